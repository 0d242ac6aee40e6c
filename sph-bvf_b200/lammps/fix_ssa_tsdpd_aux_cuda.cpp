/* ----------------------------------------------------------------------
   argument parsing of the auxiliary /cuda fixes; see fix_ssa_tsdpd_aux_cuda.h
------------------------------------------------------------------------- */

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "fix_ssa_tsdpd_aux_cuda.h"
#include "atom.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "modify.h"
#include "pair.h"
#include "update.h"

using namespace LAMMPS_NS;

enum { K_BUOYANCY = 0, K_FORCING = 1, K_BUFFER = 2, K_SETFORCE = 3, K_CHEMRXN = 4 };

void FixSphbvfRegistered::init()
{
  desc.groupbit = groupbit;
  SphbvfLmp::get(lmp)->add_fix(desc);
}

/* fix ID group ssa_tsdpd/buoyancy {boussinesq/sdpd|gravity} a dim k Cref ------------------- */

FixSsaTsdpdBuoyancyCuda::FixSsaTsdpdBuoyancyCuda(LAMMPS *lmp, int narg, char **arg) : FixSphbvfRegistered(lmp, narg, arg)
{
  if (narg != 8) error->all(FLERR, "Illegal fix ssa_tsdpd/buoyancy command");
  desc.kind = K_BUOYANCY;
  if (strcmp(arg[3], "boussinesq/sdpd") == 0) desc.ia[0] = 0;
  else if (strcmp(arg[3], "gravity") == 0) desc.ia[0] = 1;
  else error->all(FLERR, "Illegal type of force in fix ssa_tsdpd/buoyancy command. Valid options: <boussinesq/sdpd> or <gravity>");
  desc.a[0] = atof(arg[4]);       // acceleration
  desc.ia[1] = atoi(arg[5]);      // coordinate the force acts on
  desc.ia[2] = atoi(arg[6]);      // species index of the Boussinesq term
  desc.a[1] = atof(arg[7]);       // C_ref
  if (desc.ia[1] < 0 || desc.ia[1] > 2) error->all(FLERR, "Illegal fix ssa_tsdpd/buoyancy command");
  if (domain->periodicity[desc.ia[1]]) error->all(FLERR, "Cannot use buoyancy force in a periodic dimension");
  if (desc.ia[0] == 0 && (desc.ia[2] < 0 || desc.ia[2] >= atom->num_sdpd_species))
    error->all(FLERR, "Illegal fix ssa_tsdpd/buoyancy command: species index out of range");
}

/* fix ID group ssa_tsdpd/forcing {tsdpd|velocity} step idx {circle cx cy R v | rectangle cx cy Lx Ly v} */

FixSsaTsdpdForcingCuda::FixSsaTsdpdForcingCuda(LAMMPS *lmp, int narg, char **arg) : FixSphbvfRegistered(lmp, narg, arg)
{
  if (narg < 7) error->all(FLERR, "Illegal fix SsaTsdpdForcing command, first error.");
  desc.kind = K_FORCING;
  int iarg = 3;
  if (strcmp(arg[iarg], "tsdpd") == 0) desc.ia[0] = 0;
  else if (strcmp(arg[iarg], "velocity") == 0) desc.ia[0] = 1;
  else if (strcmp(arg[iarg], "ssa") == 0) error->all(FLERR, "fix ssa_tsdpd/forcing/cuda: SSA species are not supported");
  else error->all(FLERR, "Illegal argument[3]. Choose <tsdpd>, <ssa> or <velocity>");
  iarg++;
  desc.step = atoi(arg[iarg++]);
  desc.ia[1] = atoi(arg[iarg++]);   // species index / velocity component, used as given (0-based)
  if (desc.ia[0] == 0 && (desc.ia[1] < 0 || desc.ia[1] >= atom->num_sdpd_species))
    error->all(FLERR, "Illegal fix ssa_tsdpd_forcing command: species id > num_tsdpd_species.\n");
  if (desc.ia[0] == 1 && (desc.ia[1] < 0 || desc.ia[1] > 2))
    error->all(FLERR, "Illegal fix ssa_tsdpd_forcing command: velocity id out of range.\n");
  if (strcmp(arg[iarg], "circle") == 0) desc.ia[2] = 0;
  else if (strcmp(arg[iarg], "rectangle") == 0) desc.ia[2] = 1;
  else error->all(FLERR, "Illegal fix ssa_tsdpd_forcing command, symbol error.");
  iarg++;
  if (desc.ia[2] == 0) {
    if (narg != 11) error->all(FLERR, "Illegal fix ssa_tsdpd_forcing command, index0");
    desc.a[0] = atof(arg[iarg++]);  // centre
    desc.a[1] = atof(arg[iarg++]);
    desc.a[2] = atof(arg[iarg++]);  // radius
    desc.a[3] = 0.0;
    desc.a[4] = atof(arg[iarg++]);  // value
  } else {
    if (narg != 12) error->all(FLERR, "Illegal fix ssa_tsdpd_forcing command, index1");
    desc.a[0] = atof(arg[iarg++]);
    desc.a[1] = atof(arg[iarg++]);
    desc.a[2] = atof(arg[iarg++]);  // half length
    desc.a[3] = atof(arg[iarg++]);  // half width
    desc.a[4] = atof(arg[iarg++]);
  }
}

/* fix ID group ssa_tsdpd/buffer {tsdpd|velocity|density} {x|y} step idx cx cy Lx Ly v ----------- */

FixSsaTsdpdBufferCuda::FixSsaTsdpdBufferCuda(LAMMPS *lmp, int narg, char **arg) : FixSphbvfRegistered(lmp, narg, arg)
{
  if (narg != 12) error->all(FLERR, "Illegal fix ssa_tsdpd_buffer command, index1");
  desc.kind = K_BUFFER;
  int iarg = 3;
  if (strcmp(arg[iarg], "tsdpd") == 0) desc.ia[0] = 0;
  else if (strcmp(arg[iarg], "velocity") == 0) desc.ia[0] = 1;
  else if (strcmp(arg[iarg], "density") == 0) desc.ia[0] = 2;
  else error->all(FLERR, "Illegal argument[3]. Choose <tsdpd>, <velocity> or <density>");
  iarg++;
  if (strcmp(arg[iarg], "x") == 0) desc.ia[2] = 0;
  else if (strcmp(arg[iarg], "y") == 0) desc.ia[2] = 1;
  else error->all(FLERR, "Illegal argument[4]. Choose <x> or <y>");
  iarg++;
  desc.step = atoi(arg[iarg++]);
  desc.ia[1] = atoi(arg[iarg++]);
  if (desc.ia[0] == 0 && (desc.ia[1] < 0 || desc.ia[1] >= atom->num_sdpd_species))
    error->all(FLERR, "Illegal fix ssa_tsdpd_buffer command: species id > num_tsdpd_species.\n");
  if (desc.ia[0] == 1 && (desc.ia[1] < 0 || desc.ia[1] > 2))
    error->all(FLERR, "Illegal fix ssa_tsdpd_buffer command: velocity id out of range.\n");
  desc.a[0] = atof(arg[iarg++]);
  desc.a[1] = atof(arg[iarg++]);
  desc.a[2] = atof(arg[iarg++]);
  desc.a[3] = atof(arg[iarg++]);
  desc.a[4] = atof(arg[iarg++]);
}

/* fix ID group setforce fx fy fz  (numeric constants only) --------------------------------- */

FixSetForceCuda::FixSetForceCuda(LAMMPS *lmp, int narg, char **arg) : FixSphbvfRegistered(lmp, narg, arg)
{
  if (narg != 6) error->all(FLERR, "fix setforce/cuda supports the form 'setforce fx fy fz' only: other forms would act on the stale host copy of the atoms; run this deck without -sf cuda");
  desc.kind = K_SETFORCE;
  for (int k = 0; k < 3; k++) {
    if (strstr(arg[3 + k], "v_") == arg[3 + k] || strcmp(arg[3 + k], "NULL") == 0)
      error->all(FLERR, "fix setforce/cuda supports numeric constants only: a variable-style setforce would act on the stale host copy of the atoms; run this deck without -sf cuda");
    desc.a[k] = force->numeric(FLERR, arg[3 + k]);
  }
}

/* fix ID group ssa_tsdpd/chem_rxn_mass_action k nreact r0.. nprod p0.. -------------------------- */

FixSsaTsdpdChemRxnMassActionCuda::FixSsaTsdpdChemRxnMassActionCuda(LAMMPS *lmp, int narg, char **arg) :
  FixSphbvfRegistered(lmp, narg, arg)
{
  if (narg < 6) error->all(FLERR, "Illegal fix ssa_tsdpd_chem_rxn_mass_action command, first error.");
  desc.kind = K_CHEMRXN;
  int iarg = 3;
  desc.a[0] = atof(arg[iarg++]);
  const int nr = atoi(arg[iarg++]);
  if (nr > atom->num_sdpd_species)
    error->all(FLERR, "Illegal fix ssa_tsdpd_chem_rxn_mass_action command -- number of reactant species greater than number of species.\n");
  if (nr > 2 || nr < 0)
    error->all(FLERR, "Illegal fix ssa_tsdpd_chem_rxn_mass_action command -- mass action reactions can have at most 2 reactants.\n");
  if (narg < iarg + nr + 1) error->all(FLERR, "Illegal fix ssa_tsdpd_chem_rxn_mass_action command");
  for (int i = 0; i < nr; i++) {
    const int r = atoi(arg[iarg++]);
    if (r < 0 || r >= atom->num_sdpd_species) error->all(FLERR, "Illegal fix ssa_tsdpd_chem_rxn_mass_action command -- species index");
    desc.ia[1] |= r << (8 * i);
  }
  const int np = atoi(arg[iarg++]);
  if (np > atom->num_sdpd_species)
    error->all(FLERR, "Illegal fix ssa_tsdpd_chem_rxn_mass_action command -- number of product species greater than number of species.\n");
  if (np > 4 || np < 0)
    error->all(FLERR, "Illegal fix ssa_tsdpd_chem_rxn_mass_action command -- maximum number of product limited to 4 .\n");
  if (narg < iarg + np) error->all(FLERR, "Illegal fix ssa_tsdpd_chem_rxn_mass_action command");
  for (int i = 0; i < np; i++) {
    const int q = atoi(arg[iarg++]);
    if (q < 0 || q >= atom->num_sdpd_species) error->all(FLERR, "Illegal fix ssa_tsdpd_chem_rxn_mass_action command -- species index");
    desc.ia[2] |= q << (8 * i);
  }
  desc.ia[0] = nr | (np << 8);
}

/* fix ID group dt/adaptive N tmin tmax CFLmax dxAve ------------------------------------------ */

FixDtAdaptiveCuda::FixDtAdaptiveCuda(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg)
{
  if (narg < 8) error->all(FLERR, "Illegal fix dt/adaptive command");
  time_depend = 1;
  scalar_flag = 1;
  global_freq = 1;
  extscalar = 0;
  dynamic_group_allow = 1;
  nevery = force->inumeric(FLERR, arg[3]);
  if (nevery <= 0) error->all(FLERR, "Illegal fix dt/adaptive command");
  minbound = maxbound = 1;
  tmin = tmax = 0.0;
  if (strcmp(arg[4], "NULL") == 0) minbound = 0;
  else tmin = force->numeric(FLERR, arg[4]);
  if (strcmp(arg[5], "NULL") == 0) maxbound = 0;
  else tmax = force->numeric(FLERR, arg[5]);
  CFLmax = force->numeric(FLERR, arg[6]);
  dxAve = force->numeric(FLERR, arg[7]);
  if (minbound && tmin < 0.0) error->all(FLERR, "Illegal fix dt/adaptive command");
  if (maxbound && tmax < 0.0) error->all(FLERR, "Illegal fix dt/adaptive command");
  if (minbound && maxbound && tmin >= tmax) error->all(FLERR, "Illegal fix dt/adaptive command");
  if (CFLmax <= 0.0) error->all(FLERR, "Illegal fix dt/adaptive command");
  if (dxAve <= 0.0) error->all(FLERR, "Illegal fix dt/adaptive command");
  laststep = update->ntimestep;
  dt = update->dt;
}

int FixDtAdaptiveCuda::setmask() { return FixConst::END_OF_STEP; }

void FixDtAdaptiveCuda::init() { dt = update->dt; }

void FixDtAdaptiveCuda::setup(int) { end_of_step(); }

void FixDtAdaptiveCuda::end_of_step()
{
  SphbvfLmp *engine = SphbvfLmp::get(lmp);
  if (!engine->active()) error->all(FLERR, "fix dt/adaptive/cuda requires pair_style ssa_tsdpd/bvf/<style>/cuda");
  const double maxAllVsq = engine->max_vsq(groupbit);
  dt = CFLmax * dxAve / sqrt(maxAllVsq);
  if (minbound) dt = MAX(dt, tmin);
  if (maxbound) dt = MIN(dt, tmax);
  if (dt == update->dt) return;
  laststep = update->ntimestep;
  update->update_time();
  update->dt = dt;
  if (force->pair) force->pair->reset_dt();
  for (int i = 0; i < modify->nfix; i++) modify->fix[i]->reset_dt();
}

double FixDtAdaptiveCuda::compute_scalar() { return (double)laststep; }

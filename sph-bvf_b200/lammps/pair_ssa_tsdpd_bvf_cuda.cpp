/* ----------------------------------------------------------------------
   pair_style ssa_tsdpd/bvf/<style>/cuda -- host side.  The arithmetic lives in libsphbvf.so
   (csrc/kernels_pair.cu); this class keeps the reference's input-script contract:
     pair_style <name>                                   (no arguments)
     pair_coeff I J rho0 c0 eta h cutc G0 kappa[0..S-1]  (pair_..._transport_velocity.cpp:967-1026)
   per-type values (rho0, c0, G0) are keyed on the I range and later lines overwrite earlier ones,
   exactly as upstream.
------------------------------------------------------------------------- */

#include <stdlib.h>
#include <string.h>
#include "pair_ssa_tsdpd_bvf_cuda.h"
#include "sphbvf_lmp.h"
#include "atom.h"
#include "comm.h"
#include "error.h"
#include "force.h"
#include "memory.h"
#include "neighbor.h"
#include "neigh_request.h"
#include "update.h"

using namespace LAMMPS_NS;

/* ---------------------------------------------------------------------- */

PairSsaTsdpdBvfCuda::PairSsaTsdpdBvfCuda(LAMMPS *lmp, int variant_in) : Pair(lmp)
{
  restartinfo = 0;
  single_enable = 0;
  variant = variant_in;
  rho0 = soundspeed = B = G0 = NULL;
  cut = viscosity = cutc = NULL;
  kappa = NULL;
  engine = SphbvfLmp::get(lmp);
  engine->pair = this;
  engine->variant = variant;
}

/* ---------------------------------------------------------------------- */

PairSsaTsdpdBvfCuda::~PairSsaTsdpdBvfCuda()
{
  SphbvfLmp::release(lmp);   // syncs the host arrays and frees the device context
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
    memory->destroy(rho0);
    memory->destroy(soundspeed);
    memory->destroy(B);
    memory->destroy(G0);
    memory->destroy(cut);
    memory->destroy(viscosity);
    memory->destroy(cutc);
    memory->destroy(kappa);
  }
}

/* ----------------------------------------------------------------------
   force_clear + the three neighbour sweeps + reverse_comm of the reference, as ONE kernel over the
   device-resident neighbour structure.  The first call of a run (from Verlet::setup) uploads.
------------------------------------------------------------------------- */

void PairSsaTsdpdBvfCuda::compute(int eflag, int vflag)
{
  if (eflag || vflag) ev_setup(eflag, vflag);
  else evflag = vflag_fdotr = 0;

  if (!engine->active()) engine->start();                       // setup: neighbour build included
  else {
    int rebuilt = 0;
    engine->neighbor_step(&rebuilt);                                        // decide + rebuild | forward halo
  }
  engine->call(sphbvf_pair_compute);
  engine->mark_dirty();
  // thermo steps: the virial LAMMPS would get from virial_fdotr_compute() on the host arrays
  if (vflag_fdotr || vflag_global) {
    double v[6];
    engine->virial(v);
    for (int k = 0; k < 6; k++) virial[k] += v[k];
    vflag_fdotr = 0;
  }
}

/* ---------------------------------------------------------------------- */

void PairSsaTsdpdBvfCuda::allocate()
{
  allocated = 1;
  const int n = atom->ntypes, S = atom->num_sdpd_species;
  memory->create(setflag, n + 1, n + 1, "pair:setflag");
  for (int i = 1; i <= n; i++)
    for (int j = i; j <= n; j++) setflag[i][j] = 0;
  memory->create(cutsq, n + 1, n + 1, "pair:cutsq");
  memory->create(rho0, n + 1, "pair:rho0");
  memory->create(soundspeed, n + 1, "pair:soundspeed");
  memory->create(B, n + 1, "pair:B");
  memory->create(G0, n + 1, "pair:G0");
  memory->create(cut, n + 1, n + 1, "pair:cut");
  memory->create(viscosity, n + 1, n + 1, "pair:viscosity");
  memory->create(cutc, n + 1, n + 1, "pair:cutc");
  memory->create(kappa, n + 1, n + 1, S > 0 ? S : 1, "pair:kappa");
}

/* ---------------------------------------------------------------------- */

void PairSsaTsdpdBvfCuda::settings(int narg, char **)
{
  if (narg != 0) error->all(FLERR, "Illegal number of setting arguments for pair_style ssa_tsdpd/bvf");
}

/* ---------------------------------------------------------------------- */

void PairSsaTsdpdBvfCuda::coeff(int narg, char **arg)
{
  const int S = atom->num_sdpd_species;
  if (narg < 8 + S) error->all(FLERR, "Incorrect args for pair_style ssa_tsdpd/bvf coefficients");
  if (!allocated) allocate();

  int ilo, ihi, jlo, jhi;
  force->bounds(FLERR, arg[0], atom->ntypes, ilo, ihi);
  force->bounds(FLERR, arg[1], atom->ntypes, jlo, jhi);

  const double rho0_one = force->numeric(FLERR, arg[2]);
  const double c0_one = force->numeric(FLERR, arg[3]);
  const double eta_one = force->numeric(FLERR, arg[4]);
  const double cut_one = force->numeric(FLERR, arg[5]);
  const double cutc_one = force->numeric(FLERR, arg[6]);
  const double G0_one = force->numeric(FLERR, arg[7]);

  int count = 0;
  for (int i = ilo; i <= ihi; i++) {
    rho0[i] = rho0_one;
    soundspeed[i] = c0_one;
    B[i] = c0_one * c0_one * rho0_one / 7.0;
    G0[i] = G0_one;
    for (int j = MAX(jlo, i); j <= jhi; j++) {
      viscosity[i][j] = eta_one;
      cut[i][j] = cut_one;
      cutc[i][j] = cutc_one;
      for (int k = 0; k < S; k++) kappa[i][j][k] = atof(arg[8 + k]);
      setflag[i][j] = 1;
      count++;
    }
  }
  if (count == 0) error->all(FLERR, "Incorrect args for pair coefficients");
}

/* ----------------------------------------------------------------------
   no LAMMPS neighbour list is requested: the device builds its own (kernels_neigh.cu)
------------------------------------------------------------------------- */

void PairSsaTsdpdBvfCuda::init_style()
{
  engine->pair = this;
  engine->variant = variant;
  engine->rho0 = rho0;
  engine->soundspeed = soundspeed;
  engine->G0 = G0;
  engine->viscosity = viscosity;
  engine->cut = cut;
  engine->cutc = cutc;
  engine->kappa = kappa;
  engine->reset_fixes();
}

/* ---------------------------------------------------------------------- */

double PairSsaTsdpdBvfCuda::init_one(int i, int j)
{
  if (setflag[i][j] == 0) error->all(FLERR, "Not all pair ssa_tsdpd/bvf coeffs are not set");
  cut[j][i] = cut[i][j];
  viscosity[j][i] = viscosity[i][j];
  cutc[j][i] = cutc[i][j];
  for (int k = 0; k < atom->num_sdpd_species; k++) kappa[j][i][k] = kappa[i][j][k];
  return cut[i][j];
}

/* ---------------------------------------------------------------------- */

double PairSsaTsdpdBvfCuda::single(int, int, int, int, double, double, double, double &fforce)
{
  fforce = 0.0;
  return 0.0;
}

/* ----------------------------------------------------------------------
   fix ssa_tsdpd/bvf/<style>/cuda -- host side of the integrator (csrc/kernels_integrate.cu).
   Hook order per step is Verlet::run's (verlet.cpp:240-353):
     initial_integrate -> post_integrate -> [pair/cuda: neighbour + pair kernel] -> post_force
     -> final_integrate -> end_of_step
   The host arrays of class Atom are refreshed only when LAMMPS is about to read them: on
   timesteps with thermo / dump output (output->next) and at the end of the run.
------------------------------------------------------------------------- */

#include <string.h>
#include "fix_ssa_tsdpd_bvf_cuda.h"
#include "sphbvf_lmp.h"
#include "atom.h"
#include "comm.h"
#include "error.h"
#include "force.h"
#include "modify.h"
#include "output.h"
#include "pair.h"
#include "update.h"

using namespace LAMMPS_NS;
using namespace FixConst;

/* ---------------------------------------------------------------------- */

FixSsaTsdpdBvfCuda::FixSsaTsdpdBvfCuda(LAMMPS *lmp, int narg, char **arg, int variant_in) : Fix(lmp, narg, arg)
{
  if ((atom->e_flag != 1) || (atom->rho_flag != 1))
    error->all(FLERR, "fix ssa_tsdpd/bvf command requires atom_style with both energy and density");
  if (narg != 3) error->all(FLERR, "Illegal number of arguments for fix ssa_tsdpd/bvf command");
  time_integrate = 1;
  variant = variant_in;
  engine = SphbvfLmp::get(lmp);
}

/* ---------------------------------------------------------------------- */

int FixSsaTsdpdBvfCuda::setmask()
{
  int mask = 0;
  mask |= INITIAL_INTEGRATE;
  mask |= POST_INTEGRATE;
  mask |= PRE_FORCE;        // setup_pre_force only, like the reference (mask bit without a pre_force body)
  mask |= POST_FORCE;
  mask |= FINAL_INTEGRATE;
  mask |= END_OF_STEP;
  return mask;
}

/* ---------------------------------------------------------------------- */

void FixSsaTsdpdBvfCuda::init()
{
  engine = SphbvfLmp::get(lmp);
  if (!force->pair || !strstr(force->pair_style, "/cuda"))
    error->all(FLERR, "fix ssa_tsdpd/bvf/<style>/cuda requires pair_style ssa_tsdpd/bvf/<style>/cuda");
  if (engine->variant != variant)
    error->all(FLERR, "fix ssa_tsdpd/bvf/<style>/cuda and pair_style ssa_tsdpd/bvf/<style>/cuda must be the same variant");
  engine->integrate_groupbit = groupbit;

  // Between output steps the host arrays of class Atom are stale: a stock (non-/cuda) fix that integrates or
  // edits per-atom state would act on dead host data and never reach the device.  Refuse those outright;
  // stock fixes that only hook end_of_step (fix print, ave/time, ave/atom ...: readers) are served by a full
  // download before every end_of_step, which is correct but slow, so say so.
  need_host_every_step = 0;
  const int writers = INITIAL_INTEGRATE | POST_INTEGRATE | PRE_EXCHANGE | PRE_NEIGHBOR | PRE_FORCE | POST_FORCE |
                      FINAL_INTEGRATE | INITIAL_INTEGRATE_RESPA | POST_INTEGRATE_RESPA | PRE_FORCE_RESPA |
                      POST_FORCE_RESPA | FINAL_INTEGRATE_RESPA;
  for (int i = 0; i < modify->nfix; i++) {
    const char *fs = modify->fix[i]->style;
    const size_t n = strlen(fs);
    if (n >= 5 && strcmp(fs + n - 5, "/cuda") == 0) continue;
    const int m = modify->fmask[i];
    if (m & writers) {
      char msg[512];
      snprintf(msg, sizeof msg, "fix %s (style %s) has no /cuda version: it would act on the stale host copy of the atoms "
               "while the ssa_tsdpd/bvf /cuda styles keep them on the device.  Remove it or run without -sf cuda",
               modify->fix[i]->id, fs);
      error->all(FLERR, msg);
    }
    if (m & END_OF_STEP) {
      need_host_every_step = 1;
      if (comm->me == 0) {
        char msg[512];
        snprintf(msg, sizeof msg, "fix %s (style %s) reads host data in end_of_step: the /cuda styles download every "
                 "per-atom field on every step for it", modify->fix[i]->id, fs);
        error->warning(FLERR, msg);
      }
    }
  }
}

/* ----------------------------------------------------------------------
   vest = v, rhoI = rho on the host, before the atoms are uploaded by the first Pair::compute
   (fix_ssa_tsdpd_bvf_transport_velocity.cpp:76-95)
------------------------------------------------------------------------- */

void FixSsaTsdpdBvfCuda::setup_pre_force(int)
{
  engine->stop();   // a previous run's context, if any: host arrays are authoritative between runs
  double **v = atom->v;
  double **vest = atom->vest;
  double *rhoI = atom->rhoI;
  double *rho = atom->rho;
  int *mask = atom->mask;
  int nlocal = atom->nlocal;
  if (igroup == atom->firstgroup) nlocal = atom->nfirst;
  for (int i = 0; i < nlocal; i++) {
    if (mask[i] & groupbit) {
      vest[i][0] = v[i][0];
      vest[i][1] = v[i][1];
      vest[i][2] = v[i][2];
      rhoI[i] = rho[i];
    }
  }
}

/* modify->setup(): post_force of the registered fixes once (FixSetForce::setup, FixSsaTsdpdBuoyancy::setup) */

void FixSsaTsdpdBvfCuda::setup(int)
{
  engine->call(sphbvf_setup_post_force);
  engine->fetch(engine->output_fields(true));   // output of step 0 (Verlet::setup -> output->setup)
}

/* ---------------------------------------------------------------------- */

void FixSsaTsdpdBvfCuda::initial_integrate(int)
{
  engine->set_timestep(update->ntimestep);
  engine->call(sphbvf_initial_integrate);
  engine->mark_dirty();
}

void FixSsaTsdpdBvfCuda::post_integrate() { engine->call(sphbvf_post_integrate); }

void FixSsaTsdpdBvfCuda::post_force(int) { engine->call(sphbvf_post_force); }

void FixSsaTsdpdBvfCuda::final_integrate()
{
  engine->call(sphbvf_final_integrate);
  if (need_host_every_step) engine->to_host();   // a stock end_of_step fix reads the host arrays (see init)
}

void FixSsaTsdpdBvfCuda::end_of_step()
{
  engine->call(sphbvf_end_of_step);
  if (update->ntimestep == output->next) engine->fetch(engine->output_fields());
}

void FixSsaTsdpdBvfCuda::post_run() { engine->stop(); }

/* ---------------------------------------------------------------------- */

void FixSsaTsdpdBvfCuda::reset_dt()
{
  if (engine->active()) engine->set_dt(update->dt);
}

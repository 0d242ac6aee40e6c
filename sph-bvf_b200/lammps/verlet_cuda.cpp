/* ----------------------------------------------------------------------
   run_style verlet/cuda -- see verlet_cuda.h
------------------------------------------------------------------------- */

#include <string.h>
#include "verlet_cuda.h"
#include "force.h"
#include "modify.h"
#include "output.h"
#include "pair.h"
#include "timer.h"
#include "update.h"

using namespace LAMMPS_NS;

VerletCuda::VerletCuda(LAMMPS *lmp, int narg, char **arg) : Verlet(lmp, narg, arg) {}

void VerletCuda::run(int n)
{
  if (!force->pair || !strstr(force->pair_style, "ssa_tsdpd/bvf") || !strstr(force->pair_style, "/cuda")) {
    Verlet::run(n);
    return;
  }

  const int n_post_integrate = modify->n_post_integrate;
  const int n_post_force = modify->n_post_force;
  const int n_end_of_step = modify->n_end_of_step;

  for (int i = 0; i < n; i++) {
    if (timer->check_timeout(i)) {
      update->nsteps = i;
      break;
    }
    const bigint ntimestep = ++update->ntimestep;
    ev_set(ntimestep);

    timer->stamp();
    modify->initial_integrate(vflag);
    if (n_post_integrate) modify->post_integrate();
    timer->stamp(Timer::MODIFY);

    // neighbour decision, rebuild or halo, force_clear and the pair sweeps: all inside the pair style
    force->pair->compute(eflag, vflag);
    timer->stamp(Timer::PAIR);

    if (n_post_force) modify->post_force(vflag);
    modify->final_integrate();
    if (n_end_of_step) modify->end_of_step();
    timer->stamp(Timer::MODIFY);

    if (ntimestep == output->next) {
      timer->stamp();
      output->write(ntimestep);
      timer->stamp(Timer::OUTPUT);
    }
  }
}

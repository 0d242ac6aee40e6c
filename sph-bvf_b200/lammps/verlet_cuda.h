/* -*- c++ -*- ----------------------------------------------------------
   run_style verlet/cuda -- picked automatically under "-sf cuda" (update.cpp:335-350).

   Verlet::run (verlet.cpp:223-354) minus everything the device now does itself: the host-side
   Neighbor::decide / Comm::forward_comm / exchange / borders / Neighbor::build on stale host
   coordinates and the per-step memset of the host force arrays (Verlet::force_clear +
   AtomVecSsaTsdpdAtomic::force_clear, ~60 doubles per atom).  With those gone the host does O(1)
   work per step and a 64 M-atom run is not throttled by host memory bandwidth.  When the pair
   style is not one of ssa_tsdpd/bvf/<style>/cuda it behaves exactly like run_style verlet.
------------------------------------------------------------------------- */

#ifdef INTEGRATE_CLASS

IntegrateStyle(verlet/cuda,VerletCuda)

#else

#ifndef LMP_VERLET_CUDA_H
#define LMP_VERLET_CUDA_H

#include "verlet.h"

namespace LAMMPS_NS {

class VerletCuda : public Verlet {
 public:
  VerletCuda(class LAMMPS *, int, char **);
  virtual ~VerletCuda() {}
  virtual void run(int);
};

}

#endif
#endif

/* -*- c++ -*- ----------------------------------------------------------
   fix ID group ssa_tsdpd/bvf/{transportVelocity,mechanics,fsi}/cuda

   Drop-in twins of FixSsaTsdpdBvf{TransportVelocity,Mechanics,Fsi}
   (fix_ssa_tsdpd_bvf_transport_velocity.cpp:40-461 and siblings): same "fix ID group style"
   syntax (3 arguments).  The integrator also drives the hooks of the auxiliary /cuda fixes
   (buoyancy, forcing, buffer, setforce), which only register their parameters, so that the
   device executes them in Modify order inside the matching hook.
------------------------------------------------------------------------- */

#ifdef FIX_CLASS

FixStyle(ssa_tsdpd/bvf/transportVelocity/cuda,FixSsaTsdpdBvfTransportVelocityCuda)
FixStyle(ssa_tsdpd/bvf/mechanics/cuda,FixSsaTsdpdBvfMechanicsCuda)
FixStyle(ssa_tsdpd/bvf/fsi/cuda,FixSsaTsdpdBvfFsiCuda)

#else

#ifndef LMP_FIX_SSA_TSDPD_BVF_CUDA_H
#define LMP_FIX_SSA_TSDPD_BVF_CUDA_H

#include "fix.h"

namespace LAMMPS_NS {

class FixSsaTsdpdBvfCuda : public Fix {
 public:
  FixSsaTsdpdBvfCuda(class LAMMPS *, int, char **, int variant);
  int setmask();
  virtual void init();
  virtual void setup_pre_force(int);
  virtual void setup(int);
  virtual void initial_integrate(int);
  virtual void post_integrate();
  virtual void post_force(int);
  virtual void final_integrate();
  virtual void end_of_step();
  virtual void post_run();
  void reset_dt();

 protected:
  int variant;
  int need_host_every_step;   // a stock (non-/cuda) fix hooks end_of_step: refresh the host arrays before it
  class SphbvfLmp *engine;
};

class FixSsaTsdpdBvfTransportVelocityCuda : public FixSsaTsdpdBvfCuda {
 public:
  FixSsaTsdpdBvfTransportVelocityCuda(class LAMMPS *lmp, int narg, char **arg) : FixSsaTsdpdBvfCuda(lmp, narg, arg, 0) {}
};

class FixSsaTsdpdBvfMechanicsCuda : public FixSsaTsdpdBvfCuda {
 public:
  FixSsaTsdpdBvfMechanicsCuda(class LAMMPS *lmp, int narg, char **arg) : FixSsaTsdpdBvfCuda(lmp, narg, arg, 1) {}
};

class FixSsaTsdpdBvfFsiCuda : public FixSsaTsdpdBvfCuda {
 public:
  FixSsaTsdpdBvfFsiCuda(class LAMMPS *lmp, int narg, char **arg) : FixSsaTsdpdBvfCuda(lmp, narg, arg, 2) {}
};

}

#endif
#endif

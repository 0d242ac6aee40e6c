/* -*- c++ -*- ----------------------------------------------------------
   fix ssa_tsdpd/buoyancy/cuda, ssa_tsdpd/forcing/cuda, ssa_tsdpd/buffer/cuda, setforce/cuda

   Twins of FixSsaTsdpdBuoyancy (fix_ssa_tsdpd_buoyancy.cpp:28-140), FixSsaTsdpdForcing
   (fix_ssa_tsdpd_forcing.cpp:38-176), FixSsaTsdpdBuffer (fix_ssa_tsdpd_buffer.cpp:32-240) and of
   FixSetForce with constant components (fix_setforce.cpp:40-290), with the same arguments.
   They own no arithmetic: init() registers the parsed parameters with the device engine, and
   the integrator fix ssa_tsdpd/bvf/<style>/cuda executes them inside the matching hook in Modify order.
------------------------------------------------------------------------- */

#ifdef FIX_CLASS

FixStyle(ssa_tsdpd/buoyancy/cuda,FixSsaTsdpdBuoyancyCuda)
FixStyle(ssa_tsdpd/forcing/cuda,FixSsaTsdpdForcingCuda)
FixStyle(ssa_tsdpd/buffer/cuda,FixSsaTsdpdBufferCuda)
FixStyle(setforce/cuda,FixSetForceCuda)
FixStyle(ssa_tsdpd/chem_rxn_mass_action/cuda,FixSsaTsdpdChemRxnMassActionCuda)
FixStyle(dt/adaptive/cuda,FixDtAdaptiveCuda)

#else

#ifndef LMP_FIX_SSA_TSDPD_AUX_CUDA_H
#define LMP_FIX_SSA_TSDPD_AUX_CUDA_H

#include "fix.h"
#include "sphbvf_lmp.h"

namespace LAMMPS_NS {

class FixSphbvfRegistered : public Fix {
 public:
  FixSphbvfRegistered(class LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg) { memset(&desc, 0, sizeof desc); }
  int setmask() { return 0; }      // executed by the integrator fix on the device
  void init();
 protected:
  SphbvfLmp::FixDesc desc;
};

class FixSsaTsdpdBuoyancyCuda : public FixSphbvfRegistered {
 public:
  FixSsaTsdpdBuoyancyCuda(class LAMMPS *, int, char **);
};

class FixSsaTsdpdForcingCuda : public FixSphbvfRegistered {
 public:
  FixSsaTsdpdForcingCuda(class LAMMPS *, int, char **);
};

class FixSsaTsdpdBufferCuda : public FixSphbvfRegistered {
 public:
  FixSsaTsdpdBufferCuda(class LAMMPS *, int, char **);
};

class FixSetForceCuda : public FixSphbvfRegistered {
 public:
  FixSetForceCuda(class LAMMPS *, int, char **);
};

// fix ID group ssa_tsdpd/chem_rxn_mass_action k nreact r.. nprod p..  (fix_ssa_tsdpd_chem_rxn_mass_action.cpp)
class FixSsaTsdpdChemRxnMassActionCuda : public FixSphbvfRegistered {
 public:
  FixSsaTsdpdChemRxnMassActionCuda(class LAMMPS *, int, char **);
};

// fix ID group dt/adaptive N tmin tmax CFLmax dxAve (fix_dt_adaptive.cpp:39-170): the max |v|^2 comes from
// the device (sphbvf_max_vsq), the CFL arithmetic and the reset_dt() fan-out stay on the host
class FixDtAdaptiveCuda : public Fix {
 public:
  FixDtAdaptiveCuda(class LAMMPS *, int, char **);
  int setmask();
  void init();
  void setup(int);
  void end_of_step();
  double compute_scalar();
 private:
  int minbound, maxbound;
  double tmin, tmax, CFLmax, dxAve, dt;
  bigint laststep;
};

}

#endif
#endif

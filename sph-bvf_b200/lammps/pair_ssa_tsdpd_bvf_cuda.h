/* -*- c++ -*- ----------------------------------------------------------
   pair_style ssa_tsdpd/bvf/{transportVelocity,mechanics,fsi}/cuda

   Drop-in twins of PairSsaTsdpdBvf{TransportVelocity,Mechanics,Fsi}
   (pair_ssa_tsdpd_bvf_transport_velocity.h:14-59 and siblings): same pair_style / pair_coeff
   syntax, same per-type and per-type-pair tables; compute() runs on the GPU through the C ABI
   of libsphbvf.so.  Selected by "-sf cuda" / "suffix cuda" (force.cpp:230-266).
------------------------------------------------------------------------- */

#ifdef PAIR_CLASS

PairStyle(ssa_tsdpd/bvf/transportVelocity/cuda,PairSsaTsdpdBvfTransportVelocityCuda)
PairStyle(ssa_tsdpd/bvf/mechanics/cuda,PairSsaTsdpdBvfMechanicsCuda)
PairStyle(ssa_tsdpd/bvf/fsi/cuda,PairSsaTsdpdBvfFsiCuda)

#else

#ifndef LMP_PAIR_SSA_TSDPD_BVF_CUDA_H
#define LMP_PAIR_SSA_TSDPD_BVF_CUDA_H

#include "pair.h"

namespace LAMMPS_NS {

class PairSsaTsdpdBvfCuda : public Pair {
 public:
  PairSsaTsdpdBvfCuda(class LAMMPS *, int variant);
  virtual ~PairSsaTsdpdBvfCuda();
  virtual void compute(int, int);
  void settings(int, char **);
  void coeff(int, char **);
  virtual void init_style();
  virtual double init_one(int, int);
  virtual double single(int, int, int, int, double, double, double, double &);

 protected:
  int variant;
  double *rho0, *soundspeed, *B, *G0;
  double **cut, **viscosity, **cutc;
  double ***kappa;
  class SphbvfLmp *engine;
  void allocate();
};

class PairSsaTsdpdBvfTransportVelocityCuda : public PairSsaTsdpdBvfCuda {
 public:
  PairSsaTsdpdBvfTransportVelocityCuda(class LAMMPS *lmp) : PairSsaTsdpdBvfCuda(lmp, 0) {}
};

class PairSsaTsdpdBvfMechanicsCuda : public PairSsaTsdpdBvfCuda {
 public:
  PairSsaTsdpdBvfMechanicsCuda(class LAMMPS *lmp) : PairSsaTsdpdBvfCuda(lmp, 1) {}
};

class PairSsaTsdpdBvfFsiCuda : public PairSsaTsdpdBvfCuda {
 public:
  PairSsaTsdpdBvfFsiCuda(class LAMMPS *lmp) : PairSsaTsdpdBvfCuda(lmp, 2) {}
};

}

#endif
#endif

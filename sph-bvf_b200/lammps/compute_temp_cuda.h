/* -*- c++ -*- ----------------------------------------------------------
   compute temp/cuda -- picked under "-sf cuda" for every `compute ID group temp`, including the
   thermo_temp LAMMPS creates itself (output.cpp).  Same arguments and same numbers as compute temp
   (compute_temp.cpp:78-135); while a /cuda run has the atoms on the device the kinetic-energy sums
   come from sphbvf_ke_tensor instead of the (stale) host velocities, so thermo-only output steps
   need no device -> host copy.  Without an active device context it IS compute temp.
------------------------------------------------------------------------- */

#ifdef COMPUTE_CLASS

ComputeStyle(temp/cuda,ComputeTempCuda)

#else

#ifndef LMP_COMPUTE_TEMP_CUDA_H
#define LMP_COMPUTE_TEMP_CUDA_H

#include "compute_temp.h"

namespace LAMMPS_NS {

class ComputeTempCuda : public ComputeTemp {
 public:
  ComputeTempCuda(class LAMMPS *, int, char **);
  virtual ~ComputeTempCuda() {}
  virtual double compute_scalar();
  virtual void compute_vector();

 private:
  bool device_sums(double *ke6);
};

}

#endif
#endif

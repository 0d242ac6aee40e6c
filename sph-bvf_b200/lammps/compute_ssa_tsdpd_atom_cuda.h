/* -*- c++ -*- ----------------------------------------------------------
   compute ssa_tsdpd/{rho,phi,p,C,stress}/atom/cuda -- picked under "-sf cuda" for the package's per-atom computes
   (compute_ssa_tsdpd_rho_atom.cpp:61-87, ..._phi_atom.cpp:61-87, ..._p_atom.cpp:61-87, ..._C_atom.cpp:64-92,
   ..._stress_atom.cpp:40-41,66-100).  Same arguments, same numbers: each class IS the reference compute; before it
   reads its column of class Atom it asks the engine for exactly that column (SphbvfLmp::fetch), so a dump step of
   a /cuda run copies the fields its dump lists and nothing else; with atom_style
   ssa_tsdpd/atomic/cuda that request is also what allocates the host array.  Without a device context they are
   the reference computes.
------------------------------------------------------------------------- */

#ifdef COMPUTE_CLASS

ComputeStyle(ssa_tsdpd/rho/atom/cuda,ComputeSsaTsdpdRhoAtomCuda)
ComputeStyle(ssa_tsdpd/phi/atom/cuda,ComputeSsaTsdpdPhiAtomCuda)
ComputeStyle(ssa_tsdpd/p/atom/cuda,ComputeSsaTsdpdPAtomCuda)
ComputeStyle(ssa_tsdpd/C/atom/cuda,ComputeSsaTsdpdCAtomCuda)
ComputeStyle(ssa_tsdpd/stress/atom/cuda,ComputeSsaTsdpdStressAtomCuda)

#else

#ifndef LMP_COMPUTE_SSA_TSDPD_ATOM_CUDA_H
#define LMP_COMPUTE_SSA_TSDPD_ATOM_CUDA_H

#include "compute_ssa_tsdpd_rho_atom.h"
#include "compute_ssa_tsdpd_phi_atom.h"
#include "compute_ssa_tsdpd_p_atom.h"
#include "compute_ssa_tsdpd_C_atom.h"
#include "compute_ssa_tsdpd_stress_atom.h"

namespace LAMMPS_NS {


#define SPHBVF_ATOM_COMPUTE(Name, Base)                                              \
  class Name : public Base {                                                         \
   public:                                                                           \
    Name(class LAMMPS *lmp, int narg, char **arg) : Base(lmp, narg, arg) {}          \
    void compute_peratom();                                                          \
  };

SPHBVF_ATOM_COMPUTE(ComputeSsaTsdpdRhoAtomCuda, ComputeSsaTsdpdRhoAtom)
SPHBVF_ATOM_COMPUTE(ComputeSsaTsdpdPhiAtomCuda, ComputeSsaTsdpdPhiAtom)
SPHBVF_ATOM_COMPUTE(ComputeSsaTsdpdPAtomCuda, ComputeSsaTsdpdPAtom)
SPHBVF_ATOM_COMPUTE(ComputeSsaTsdpdCAtomCuda, ComputeSsaTsdpdCAtom)
SPHBVF_ATOM_COMPUTE(ComputeSsaTsdpdStressAtomCuda, ComputeSsaTsdpdStressAtom)

#undef SPHBVF_ATOM_COMPUTE

}

#endif
#endif

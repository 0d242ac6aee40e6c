/* ----------------------------------------------------------------------
   compute ssa_tsdpd/{rho,phi,p,C,stress}/atom/cuda -- see compute_ssa_tsdpd_atom_cuda.h
------------------------------------------------------------------------- */

#include "compute_ssa_tsdpd_atom_cuda.h"
#include "sphbvf_lmp.h"

using namespace LAMMPS_NS;

#include "atom.h"

/* the host array behind the column: allocate it if it is a lazy mirror, refresh it if a device context holds newer data */
#define sphbvf_fetch_for_compute(mask) SphbvfLmp::host_fields(atom, mask)

/* each compute reads ONE column of class Atom (stress: the pressure and the deviatoric tensor) */

void ComputeSsaTsdpdRhoAtomCuda::compute_peratom()
{
  sphbvf_fetch_for_compute(SphbvfLmp::HF_RHO);
  ComputeSsaTsdpdRhoAtom::compute_peratom();
}

void ComputeSsaTsdpdPhiAtomCuda::compute_peratom()
{
  sphbvf_fetch_for_compute(SphbvfLmp::HF_PHI);
  ComputeSsaTsdpdPhiAtom::compute_peratom();
}

void ComputeSsaTsdpdPAtomCuda::compute_peratom()
{
  sphbvf_fetch_for_compute(SphbvfLmp::HF_PNEW);
  ComputeSsaTsdpdPAtom::compute_peratom();
}

void ComputeSsaTsdpdCAtomCuda::compute_peratom()
{
  sphbvf_fetch_for_compute(SphbvfLmp::HF_C);
  ComputeSsaTsdpdCAtom::compute_peratom();
}

void ComputeSsaTsdpdStressAtomCuda::compute_peratom()
{
  sphbvf_fetch_for_compute(SphbvfLmp::HF_PNEW | SphbvfLmp::HF_DEV);
  ComputeSsaTsdpdStressAtom::compute_peratom();
}

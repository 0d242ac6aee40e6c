/* ----------------------------------------------------------------------
   compute temp/cuda -- see compute_temp_cuda.h
------------------------------------------------------------------------- */

#include "compute_temp_cuda.h"
#include "error.h"
#include "force.h"
#include "sphbvf_lmp.h"
#include "update.h"

using namespace LAMMPS_NS;

ComputeTempCuda::ComputeTempCuda(LAMMPS *lmp, int narg, char **arg) : ComputeTemp(lmp, narg, arg) {}

/* sum m v_a v_b over the group from the device if that is where the current velocities are */

bool ComputeTempCuda::device_sums(double *ke6)
{
  SphbvfLmp *engine = SphbvfLmp::peek();
  if (!engine || !engine->active() || engine->host_has(SphbvfLmp::HF_V)) return false;
  engine->ke_tensor(groupbit, ke6);
  engine->count_device_thermo();
  return true;
}

/* ---------------------------------------------------------------------- */

double ComputeTempCuda::compute_scalar()
{
  double t[6];
  if (!device_sums(t)) return ComputeTemp::compute_scalar();
  invoked_scalar = update->ntimestep;
  scalar = t[0] + t[1] + t[2];   // one rank drives the device(s): no MPI_Allreduce
  if (dynamic) dof_compute();
  if (dof < 0.0 && natoms_temp > 0.0) error->all(FLERR, "Temperature compute degrees of freedom < 0");
  scalar *= tfactor;
  return scalar;
}

/* ---------------------------------------------------------------------- */

void ComputeTempCuda::compute_vector()
{
  double t[6];
  if (!device_sums(t)) {
    ComputeTemp::compute_vector();
    return;
  }
  invoked_vector = update->ntimestep;
  for (int i = 0; i < 6; i++) vector[i] = t[i] * force->mvv2e;
}

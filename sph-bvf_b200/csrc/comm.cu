// comm.cu -- brick decomposition across the GPUs of one box: one process per GPU, NCCL over
// NVLink/NVSwitch.  Replaces comm.cpp:445 set_proc_grid + procmap.cpp (grid factorisation),
// CommBrick::{forward_comm,exchange,borders} (comm_brick.cpp:460-880) and the MPI_Allreduce of the
// rebuild vote (neighbor.cpp:1997).  No reverse communication exists: every rank gathers over the
// full neighbour set of its owned atoms, so ghosts are read-only.
#include <math.h>
#include <string.h>

#include "context.cuh"

using namespace sphbvf;

extern "C" {

// rank r owns brick (r % px, (r / px) % py, r / (px*py)); uniform cuts like comm->xsplit defaults
int sphbvf_brick_bounds(const sphbvf_config *cfg, int rank, double sublo[3], double subhi[3]) {
  int pg[3];
  for (int k = 0; k < 3; k++) pg[k] = cfg->procgrid[k] < 1 ? 1 : cfg->procgrid[k];
  const int loc[3] = {rank % pg[0], (rank / pg[0]) % pg[1], rank / (pg[0] * pg[1])};
  if (rank < 0 || loc[2] >= pg[2]) return SPHBVF_EINVAL;
  for (int k = 0; k < 3; k++) {
    const double prd = cfg->boxhi[k] - cfg->boxlo[k];
    // domain.cpp:308-330 set_local_box: sublo = boxlo + prd * split[loc]; the last brick ends at boxhi
    sublo[k] = cfg->boxlo[k] + prd * ((double)loc[k] / pg[k]);
    subhi[k] = loc[k] == pg[k] - 1 ? cfg->boxhi[k] : cfg->boxlo[k] + prd * ((double)(loc[k] + 1) / pg[k]);
  }
  return 0;
}

// procmap.cpp:71-134 onelevel_grid / best_factors: among all factorisations of nranks choose the
// one minimising the surface area of a brick (2D: pz = 1)
int sphbvf_proc_grid(int nranks, int dim, const double prd[3], int grid[3]) {
  if (nranks < 1) return SPHBVF_EINVAL;
  double best = 1e300;
  grid[0] = nranks; grid[1] = 1; grid[2] = 1;
  for (int px = 1; px <= nranks; px++) {
    if (nranks % px) continue;
    const int rest = nranks / px;
    for (int py = 1; py <= rest; py++) {
      if (rest % py) continue;
      const int pz = rest / py;
      if (dim == 2 && pz != 1) continue;
      const double lx = prd[0] / px, ly = prd[1] / py, lz = prd[2] / pz;
      const double surf = dim == 2 ? lx + ly : lx * ly + ly * lz + lx * lz;
      if (surf < best - 1e-12 * fabs(best)) {
        best = surf;
        grid[0] = px; grid[1] = py; grid[2] = pz;
      }
    }
  }
  return 0;
}

// CommBrick::setup (comm_brick.cpp:161-410) for a one-hop halo: for each of the 27 directions
// code = (dx+1) + 3 (dy+1) + 9 (dz+1) the rank owning the adjacent brick (-1: none, i.e. a fixed
// boundary, the centre, or dz != 0 in 2D) and the shift added to positions sent that way
// (-d_k * prd_k when the step crosses a periodic face, as pack_comm does with pbc flags).
int sphbvf_comm_plan(const sphbvf_config *cfg, int rank, int *peer, double *shift) {
  int pg[3];
  for (int k = 0; k < 3; k++) pg[k] = cfg->procgrid[k] < 1 ? 1 : cfg->procgrid[k];
  if (rank < 0 || rank >= pg[0] * pg[1] * pg[2]) return SPHBVF_EINVAL;
  const int loc[3] = {rank % pg[0], (rank / pg[0]) % pg[1], rank / (pg[0] * pg[1])};
  for (int code = 0; code < 27; code++) {
    const int s[3] = {code % 3 - 1, (code / 3) % 3 - 1, code / 9 - 1};
    int t[3];
    bool ok = code != 13 && !(cfg->dim == 2 && s[2] != 0);
    for (int k = 0; k < 3; k++) {
      shift[3 * code + k] = 0.0;
      t[k] = loc[k] + s[k];
      if (t[k] < 0 || t[k] >= pg[k]) {
        if (!cfg->periodic[k]) ok = false;
        else {
          t[k] = (t[k] + pg[k]) % pg[k];
          shift[3 * code + k] = -s[k] * (cfg->boxhi[k] - cfg->boxlo[k]);
        }
      }
    }
    peer[code] = ok ? t[0] + pg[0] * (t[1] + pg[1] * t[2]) : -1;
    if (!ok) shift[3 * code] = shift[3 * code + 1] = shift[3 * code + 2] = 0.0;
  }
  return 0;
}

}  // extern "C"

// ---- NCCL halo (implemented in comm_nccl.cu when built with NCCL) ----------------------------
#ifndef SPHBVF_WITH_NCCL
int comm_rebuild(sphbvf_ctx *ctx) { return ctx->fail(SPHBVF_ECOMM, "library built without NCCL"); }
int comm_forward(sphbvf_ctx *ctx, int) { return ctx->fail(SPHBVF_ECOMM, "library built without NCCL"); }
int comm_halo_join(sphbvf_ctx *) { return 0; }
cudaStream_t comm_halo_stream(sphbvf_ctx *) { return nullptr; }
int comm_halo_mark(sphbvf_ctx *) { return 0; }
int comm_face_has_peer(const sphbvf_ctx *, int, int) { return 0; }
int comm_vote(sphbvf_ctx *ctx, int *) { return ctx->fail(SPHBVF_ECOMM, "library built without NCCL"); }
int comm_allreduce_max(sphbvf_ctx *ctx, int *, int) { return ctx->fail(SPHBVF_ECOMM, "library built without NCCL"); }
int comm_allreduce_max_double(sphbvf_ctx *ctx, double *) { return ctx->fail(SPHBVF_ECOMM, "library built without NCCL"); }
void comm_destroy(sphbvf_ctx *) {}
extern "C" int sphbvf_comm_unique_id(void *) { return SPHBVF_ECOMM; }
extern "C" int sphbvf_comm_init(sphbvf_ctx *ctx, const void *) { return ctx->fail(SPHBVF_ECOMM, "library built without NCCL"); }
#endif

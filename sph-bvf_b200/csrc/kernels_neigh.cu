// kernels_neigh.cu -- kernel (1) of the hot path: periodic wrap, rebuild trigger, cell binning,
// deterministic counting sort of the owned atoms, periodic-image ghosts, and the Verlet list.
//
// Replaces Domain::pbc (domain.cpp:498-600), Neighbor::check_distance (neighbor.cpp:1950-2006),
// NBinStandard::{setup_bins,bin_atoms} + NBin::coord2bin (nbin_standard.cpp:53-232, nbin.cpp:116-148),
// the NStencil*Bin* offsets (nstencil.cpp:145-228), NPair{HalfBinAtomonlyNewton,FullBinAtomonly}::build
// (npair_half_bin_atomonly_newton.cpp:37-118, npair_full_bin_atomonly.cpp:34-95) and, for one rank,
// CommBrick::borders (comm_brick.cpp:709-880).
//
// The reference's linked-list bins become a counting sort that physically reorders the owned
// atoms (coalesced streaming everywhere else); its half list with Newton mirror becomes a FULL
// list per owned atom (gather form, no atomics, no reverse communication).  The SET of
// interacting pairs is the reference's: rsq <= cutneighsq[itype][jtype] at rebuild time, with rsq
// evaluated without FMA contraction in the reference's operation order, so the lists compare
// bit-exactly as sorted tag pairs.
#include <stdlib.h>
#include <string.h>

#include "sphbvf_internal.cuh"
#include "tile_common.cuh"

namespace sphbvf {

static inline int nblocks(long n, int t) { return (int)((n + t - 1) / t); }

// flags: [0] non-finite coordinate, [1] lost atom (outside the cell grid), [2] max neighbour
// count, [3] some atom moved more than skin/2
__device__ __forceinline__ double rsq_nofma(double dx, double dy, double dz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__global__ void check_distance_kernel(const DevState d, const double triggersq, int *flag_moved) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
  const size_t i3 = 3 * (size_t)i;
  const double rsq = rsq_nofma(d.x[i3] - d.xhold[i3], d.x[i3 + 1] - d.xhold[i3 + 1], d.x[i3 + 2] - d.xhold[i3 + 2]);
  if (rsq > triggersq) *flag_moved = 1;
}

void launch_check_distance(const DevState &d, double triggersq, int *flag_moved, cudaStream_t st) {
  if (!d.nlocal) return;
  check_distance_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, triggersq, flag_moved);
  SPHBVF_LAUNCHED(1);
}

__device__ __forceinline__ int cell_of(const Grid &g, double x, double y, double z, bool &ok) {
  const double fx = (x - g.lo[0]) * g.inv[0], fy = (y - g.lo[1]) * g.inv[1], fz = (z - g.lo[2]) * g.inv[2];
  int cx = (int)floor(fx), cy = (int)floor(fy), cz = (int)floor(fz);
  ok = cx >= 0 && cy >= 0 && cz >= 0 && cx < g.n[0] && cy < g.n[1] && cz < g.n[2];
  return ok ? cell_index(g, cx, cy, cz) : 0;
}

__global__ void pbc_cellid_kernel(const DevState d, const Box b, const Grid g, int *cellid, int *cell_count, int *flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
  double x[3];
  for (int k = 0; k < 3; k++) {
    double xk = d.x[3 * (size_t)i + k];
    if (!isfinite(xk)) { flags[0] = 1; xk = b.lo[k]; }
    if (b.periodic[k]) {
      if (xk < b.lo[k]) xk += b.prd[k];
      if (xk >= b.hi[k]) {
        xk -= b.prd[k];
        xk = xk > b.lo[k] ? xk : b.lo[k];
      }
      d.x[3 * (size_t)i + k] = xk;
    }
    x[k] = xk;
  }
  if (b.dim == 2) x[2] = g.lo[2];
  bool ok;
  int c = cell_of(g, x[0], x[1], x[2], ok);
  if (!ok) { flags[1] = 1; c = 0; }
  cellid[i] = c;
  atomicAdd(&cell_count[c], 1);
}

void launch_cell_ids(const DevState &d, const Grid &g, const Box &b, const NeighWork &w, cudaStream_t st) {
  cudaMemsetAsync(w.cell_count, 0, sizeof(int) * (g.ncells + 1), st);
  if (!d.nlocal) return;
  pbc_cellid_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, b, g, w.cellid, w.cell_count, w.flags);
  SPHBVF_LAUNCHED(1);
}

// ---------------------------------------------------------------- exclusive scan (int32)
constexpr int SCAN_T = 256, SCAN_E = 8, SCAN_B = SCAN_T * SCAN_E;

__global__ void scan_block_kernel(const int *in, int *out, long n, int *block_sums) {
  __shared__ int warp_tot[SCAN_T / 32];
  const long base = (long)blockIdx.x * SCAN_B + (long)threadIdx.x * SCAN_E;
  int v[SCAN_E], tsum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    tsum += v[k];
  }
  // inclusive scan of per-thread sums across the block
  int x = tsum;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int t = lane < SCAN_T / 32 ? warp_tot[lane] : 0;
#pragma unroll
    for (int o = 1; o < SCAN_T / 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    if (lane < SCAN_T / 32) warp_tot[lane] = t;
  }
  __syncthreads();
  int excl = x - tsum + (wid ? warp_tot[wid - 1] : 0);
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    if (base + k < n) out[base + k] = excl;
    excl += v[k];
  }
  if (threadIdx.x == SCAN_T - 1 && block_sums) block_sums[blockIdx.x] = excl;
}

__global__ void scan_add_kernel(int *out, long n, const int *block_offsets) {
  const long i = (long)blockIdx.x * SCAN_B + threadIdx.x;
  const int off = block_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    const long q = i + (long)k * SCAN_T;
    if (q < n) out[q] += off;
  }
}

// out[0..n) = exclusive scan of in[0..n); tmp needs >= n/SCAN_B + 2*SCAN_B ints.  `in` must have
// n+1 readable entries with in[n] arbitrary: callers pass n+1 so out[n] is the grand total.
void exclusive_scan(const int *in, int *out, long n, int *tmp, cudaStream_t st) {
  if (n <= 0) return;
  const long nb = (n + SCAN_B - 1) / SCAN_B;
  if (nb == 1) {
    scan_block_kernel<<<1, SCAN_T, 0, st>>>(in, out, n, nullptr);
    SPHBVF_LAUNCHED(1);
    return;
  }
  scan_block_kernel<<<(int)nb, SCAN_T, 0, st>>>(in, out, n, tmp);
  exclusive_scan(tmp, tmp, nb, tmp + nb, st);
  scan_add_kernel<<<(int)nb, SCAN_T, 0, st>>>(out, n, tmp);
  SPHBVF_LAUNCHED(2);
}

// ---------------------------------------------------------------- counting sort of owned atoms
__global__ void scatter_kernel(const int n, const int *cellid, const int *cell_start, int *cursor, int *perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cellid[i];
  const int r = atomicAdd(&cursor[c], 1);
  perm[cell_start[c] + r] = i;
}

// restore determinism: within each cell order by atom tag (cells hold a handful of atoms), so
// that neither the atomics above nor the arrival order of migrated / ghost atoms shows in results
__global__ void cell_order_kernel(const long ncells, const int *cell_start, int *perm, const int *key) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const int a = cell_start[c], b = cell_start[c + 1];
  for (int p = a + 1; p < b; p++) {
    const int v = perm[p], kv = key[v];
    int q = p - 1;
    while (q >= a && (key[perm[q]] > kv || (key[perm[q]] == kv && perm[q] > v))) { perm[q + 1] = perm[q]; q--; }
    perm[q + 1] = v;
  }
}

void launch_sort_owned(const DevState &d, const Grid &g, const NeighWork &w, cudaStream_t st) {
  // cell_count[ncells] = 0 sentinel so that cell_start[ncells] = nlocal
  exclusive_scan(w.cell_count, w.cell_start, g.ncells + 1, w.scan_tmp, st);
  cudaMemsetAsync(w.cell_count, 0, sizeof(int) * (g.ncells + 1), st);
  if (!d.nlocal) return;
  scatter_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d.nlocal, w.cellid, w.cell_start, w.cell_count, w.perm);
  cell_order_kernel<<<nblocks(g.ncells, 256), 256, 0, st>>>(g.ncells, w.cell_start, w.perm, d.tag);
  SPHBVF_LAUNCHED(2);
}

template <typename T>
__global__ void permute_kernel(const T *in, T *out, const int *perm, const int n, const int ncols) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long)n * ncols) return;
  const int row = (int)(q / ncols), col = (int)(q - (long)row * ncols);
  out[q] = in[(long)perm[row] * ncols + col];
}

// arr[new] = arr[perm[new]] through the staging buffer
void launch_permute(void *arr, void *tmp, const int *perm, int n, int ncols, int elem_bytes, cudaStream_t st) {
  if (!n || !ncols) return;
  const long tot = (long)n * ncols;
  if (elem_bytes == 8) permute_kernel<double><<<nblocks(tot, 256), 256, 0, st>>>((const double *)arr, (double *)tmp, perm, n, ncols);
  else permute_kernel<int><<<nblocks(tot, 256), 256, 0, st>>>((const int *)arr, (int *)tmp, perm, n, ncols);
  SPHBVF_LAUNCHED(1);
  cudaMemcpyAsync(arr, tmp, (size_t)tot * elem_bytes, cudaMemcpyDeviceToDevice, st);
}

// Every primary array in one pass: out.<f>[i] = in.<f>[perm[i]].  One thread per atom; the reads of a warp are a
// near-sequential gather (a rebuild moves atoms by a few cells), the writes are coalesced.  The caller swaps the two
// buffers afterwards, so nothing is copied back (the per-array permute above wrote a staging buffer and copied it
// back: twice the bytes and 28 launches per rebuild).
__global__ void __launch_bounds__(256)
gather_state_kernel(const StateArrays in, const StateArrays out, const int *__restrict__ perm, const int n, const int S,
                    const int with_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = perm[i];
  const size_t i3 = 3 * (size_t)i, o3 = 3 * (size_t)o;
  const int tag = in.tag[o], type = in.type[o], mask = in.mask[o], solid = in.solid[o], fixed = in.fixed[o], slot = in.slot[o];
  double x[3], v[3], ve[3];
#pragma unroll
  for (int k = 0; k < 3; k++) { x[k] = in.x[o3 + k]; v[k] = in.v[o3 + k]; ve[k] = in.vest[o3 + k]; }
  const double rho = in.rho[o], rhoI = in.rhoI[o], e = in.e[o];
  out.tag[i] = tag; out.type[i] = type; out.mask[i] = mask; out.solid[i] = solid; out.fixed[i] = fixed; out.slot[i] = slot;
#pragma unroll
  for (int k = 0; k < 3; k++) { out.x[i3 + k] = x[k]; out.v[i3 + k] = v[k]; out.vest[i3 + k] = ve[k]; }
  out.rho[i] = rho; out.rhoI[i] = rhoI; out.e[i] = e;
  for (int k = 0; k < S; k++) out.C[(size_t)i * S + k] = in.C[(size_t)o * S + k];
  if (with_dev)
    for (int k = 0; k < 9; k++) out.dev[9 * (size_t)i + k] = in.dev[9 * (size_t)o + k];
}

void launch_gather_state(const StateArrays &in, const StateArrays &out, const int *perm, int n, int S, int with_dev,
                         cudaStream_t st) {
  if (!n) return;
  gather_state_kernel<<<nblocks(n, 256), 256, 0, st>>>(in, out, perm, n, S, with_dev);
  SPHBVF_LAUNCHED(1);
}

// ---------------------------------------------------------------- periodic self-image ghosts
// An owned atom needs an image shifted by s (s_k in {-1,0,+1}, periodic dims only, s != 0) iff for
// every shifted dim it lies in the slab the reference's staged swaps send across that face:
// s_k = +1: x_k <= lo + cutghost ; s_k = -1: x_k >= hi - cutghost (comm_brick.cpp:355-395, 765-770).
__device__ __forceinline__ void image_slabs(const Box &b, const double cutghost, const double *x, int lo[3], int hi[3]) {
  for (int k = 0; k < 3; k++) {
    const bool use = b.periodic[k] && !(b.dim == 2 && k == 2) && b.sublo[k] == b.lo[k] && b.subhi[k] == b.hi[k];
    lo[k] = use && x[k] <= b.lo[k] + cutghost;
    hi[k] = use && x[k] >= b.hi[k] - cutghost;
  }
}

__global__ void count_images_kernel(const DevState d, const Box b, const double cutghost, int *nimg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > d.nlocal) return;
  if (i == d.nlocal) { nimg[i] = 0; return; }
  int lo[3], hi[3];
  image_slabs(b, cutghost, &d.x[3 * (size_t)i], lo, hi);
  nimg[i] = (1 + lo[0] + hi[0]) * (1 + lo[1] + hi[1]) * (1 + lo[2] + hi[2]) - 1;
}

__global__ void fill_images_kernel(const DevState d, const Box b, const double cutghost, const int *img_start) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
  int lo[3], hi[3];
  image_slabs(b, cutghost, &d.x[3 * (size_t)i], lo, hi);
  int g = img_start[i];
  for (int sz = -1; sz <= 1; sz++)
    for (int sy = -1; sy <= 1; sy++)
      for (int sx = -1; sx <= 1; sx++) {
        if (!sx && !sy && !sz) continue;
        const int s[3] = {sx, sy, sz};
        bool need = true;
        for (int k = 0; k < 3; k++)
          if ((s[k] == 1 && !lo[k]) || (s[k] == -1 && !hi[k])) need = false;
        if (!need) continue;
        d.gowner[g] = i;
        for (int k = 0; k < 3; k++) d.gshift[3 * (size_t)g + k] = s[k] * b.prd[k];
        d.ptag[d.nlocal + g] = d.tag[i];
        g++;
      }
}

void launch_count_images(const DevState &d, const Box &b, double cutghost, const NeighWork &w, cudaStream_t st) {
  count_images_kernel<<<nblocks(d.nlocal + 1, 256), 256, 0, st>>>(d, b, cutghost, w.nimg);
  SPHBVF_LAUNCHED(1);
  exclusive_scan(w.nimg, w.nimg, d.nlocal + 1, w.scan_tmp, st);
}

void launch_fill_images(const DevState &d, const Box &b, double cutghost, const NeighWork &w, cudaStream_t st) {
  if (!d.nlocal) return;
  fill_images_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, b, cutghost, w.nimg);
  SPHBVF_LAUNCHED(1);
}

// ---------------------------------------------------------------- ghosts into cells (index sort)
__global__ void ghost_cellid_kernel(const DevState d, const Grid g, const int dim, int *cellid, int *gcell_count, int *flags) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= d.nghost) return;
  const Rec4 A = d.prec[d.nlocal + q].A;
  bool ok;
  int c = cell_of(g, A.x, A.y, dim == 2 ? g.lo[2] : A.z, ok);
  if (!ok) c = -1;   // beyond the stencil reach of every owned atom: never a neighbour
  cellid[d.nlocal + q] = c;
  if (c >= 0) atomicAdd(&gcell_count[c], 1);
}

__global__ void ghost_scatter_kernel(const DevState d, const int *cellid, const int *gcell_start, int *cursor, int *gorder) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= d.nghost) return;
  const int c = cellid[d.nlocal + q];
  if (c < 0) return;
  const int r = atomicAdd(&cursor[c], 1);
  gorder[gcell_start[c] + r] = q;
}

void launch_bin_ghosts(const DevState &d, const Grid &g, const NeighWork &w, cudaStream_t st) {
  if (!d.nghost) return;   // the list build never opens the ghost table then
  cudaMemsetAsync(w.gcell_count, 0, sizeof(int) * (g.ncells + 1), st);
  if (d.nghost) ghost_cellid_kernel<<<nblocks(d.nghost, 256), 256, 0, st>>>(d, g, g.dim, w.cellid, w.gcell_count, w.flags);
  exclusive_scan(w.gcell_count, w.gcell_start, g.ncells + 1, w.scan_tmp, st);
  cudaMemsetAsync(w.gcell_count, 0, sizeof(int) * (g.ncells + 1), st);
  ghost_scatter_kernel<<<nblocks(d.nghost, 256), 256, 0, st>>>(d, w.cellid, w.gcell_start, w.gcell_count, w.gorder);
  cell_order_kernel<<<nblocks(g.ncells, 256), 256, 0, st>>>(g.ncells, w.gcell_start, w.gorder, d.ptag + d.nlocal);
  SPHBVF_LAUNCHED(3);
}

// ---------------------------------------------------------------- Verlet list
// Two builders produce the same list (same entries, same order inside each atom's row is not required:
// the pair SET is what is compared bit-exactly, and both orders are deterministic):
//
// (A) build_list_tile_kernel (default): one CTA per tile of the cell grid (4x4x4 cells in 3D, 8x8 in 2D: a
//     contiguous range of ~145 atoms).  The candidates of the whole tile -- the atoms of the cells within the
//     stencil reach of the tile, at most 8x8x8 cells, i.e. <= 3 contiguous index ranges per (y,z) row -- are
//     staged ONCE in shared memory as {x, y, z, packed list entry} (32 B).  Every thread (one atom) then sweeps
//     the staged candidates in lock step: no divergence, every shared-memory read is a broadcast, and the test
//     is the reference's exact FP64 criterion (rsq evaluated without FMA in the reference's operation order);
//     a warp only sweeps the z-layers of cells its own atoms can reach.  FP64-pipe bound (8 FP64 operations
//     per candidate, ~800 candidates per atom) instead of divergence bound: the thread-per-atom walk (B) spent
//     26 k warp instructions per 32 atoms at 12.6 of 32 active lanes (ncu, profiles/) in tiny per-cell loops
//     whose trip counts differ in every lane.
// (B) build_list_kernel (SPHBVF_LIST_BUILD=thread): one thread per owned atom walks its own stencil; for every
//     (dz, dy) row the cells cx-s .. cx+s are visited as at most two contiguous index ranges, pruned by the
//     atom's position.  Kept as the cross-check of (A) (tests/test_gpu_parity.py) and for A/B timing.
// Ghosts live in their own cell table (gcell_start / gorder).

#ifndef TB_THREADS
#define TB_THREADS 224
#endif
#ifndef TB_UNROLL
#define TB_UNROLL 4
#endif
constexpr int TB_U = TB_UNROLL;               // unroll factor of the sweep
constexpr int TB_T = TB_THREADS;              // threads per CTA: a bulk 3D tile holds 125..216 atoms (one batch)
constexpr int TB_CH = 1536;                   // staged candidates per chunk (48 KB, dynamic): a bulk halo is <= 11^3

struct __align__(16) Cand {   // two LDS.128 broadcasts per candidate: {x, y} and {z, entry}
  double2 xy;
  double2 ze;
};

// LIST16 = false: gather form, 32-bit entries j | type_j << 27 | solid_j << 30 written transposed (neigh[k * stride + i]).
// LIST16 = true:  tile form, 16-bit entries slot | type_j << 12 | solid_j << 15 written row-major; a thread collects
//                 four entries in a 64-bit shift register and stores 8 aligned bytes at a time (a quarter of the
//                 scattered store instructions of the 4-byte emission, which cost 3.4 of 10.3 ms per rebuild).
template <bool UNIFORM, bool LIST16>
__global__ void __launch_bounds__(TB_T)
build_list_tile_kernel(const DevState d, const __grid_constant__ Grid g, const __grid_constant__ Coeffs co,
                       const int *__restrict__ cell_start, const int *__restrict__ gcell_start,
                       const int *__restrict__ gorder, const double cutmaxsq, int *flags) {
  extern __shared__ __align__(16) unsigned char tb_smem[];
  Cand *cand = reinterpret_cast<Cand *>(tb_smem);
  __shared__ int seg_src[TB_MAXSEG];
  __shared__ int seg_off[TB_MAXSEG + 1];
  __shared__ int layer_off[TB_MAXLAY + 1];

  const int tid = threadIdx.x;
  TileGeom t;
  if (!tile_geometry(g, blockIdx.x, cell_start, t)) return;   // empty tile (whole CTA)
  const int first = t.first, last = t.last, hz0 = t.hz0, hz1 = t.hz1, nz = t.nz, nseg = t.nseg;
  tile_segments(g, t, cell_start, gcell_start, d.nghost > 0, seg_src, seg_off);
  if (tid <= nz) layer_off[tid] = seg_off[tid * t.ny * 6];
  if (tid == 0) atomicMax(&flags[4], seg_off[nseg]);
  __syncthreads();
  const int total = seg_off[nseg];
  const size_t stride = d.stride;

  for (int base = first; base < last; base += TB_T) {
    const int i = base + tid;
    const bool valid = i < last;
    Rec4 Ai = make_rec4(0, 0, 0, 0);
    int ti = 1;
    if (valid) { Ai = d.prec[i].A; ti = d.pflags[i] & 7; }
    double cut_i[MAXT];
#pragma unroll
    for (int q = 0; q < MAXT; q++) cut_i[q] = UNIFORM ? cutmaxsq : co.cutneighsq[ti][q];
    // halo z-layers this WARP can reach
    int llo = TB_MAXLAY, lhi = -1;
    if (valid) {
      int cz = g.dim == 2 ? 0 : (int)floor((Ai.z - g.lo[2]) * g.inv[2]);
      cz = min(max(cz, hz0), hz1);
      llo = max(cz - g.s[2], hz0) - hz0;
      lhi = min(cz + g.s[2], hz1) - hz0;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      llo = min(llo, __shfl_xor_sync(0xffffffffu, llo, o));
      lhi = max(lhi, __shfl_xor_sync(0xffffffffu, lhi, o));
    }
    int n = 0;
    int *out = d.neigh + (valid ? i : first);
    const int maxn = valid ? (LIST16 ? d.pitch16 : d.maxneigh) : 0;
    unsigned long long *row64 = LIST16 ? reinterpret_cast<unsigned long long *>(d.neigh16 + (size_t)(valid ? i : first) * d.pitch16) : nullptr;
    unsigned long long acc = 0;
    const int wq0 = lhi >= 0 ? layer_off[llo] : 0, wq1 = lhi >= 0 ? layer_off[lhi + 1] : 0;

    for (int clo = 0; clo < total; clo += TB_CH) {
      const int chi = min(total, clo + TB_CH);
      // ---- stage candidates clo .. chi
      for (int q = clo + tid; q < chi; q += TB_T) {
        int lo = 0, hi = nseg;   // largest sid with seg_off[sid] <= q
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (seg_off[mid] <= q) lo = mid; else hi = mid;
        }
        const int j = tile_source(seg_src[lo], q - seg_off[lo], d.nlocal, gorder);
        const Rec4 A = d.prec[j].A;
        const int fj = d.pflags[j];
        Cand c;
        c.xy = make_double2(A.x, A.y);
        c.ze = make_double2(A.z, __hiloint2double(0, j | ((fj & 7) << NEIGH_JBITS) | (((fj >> 4) & 1) << 30)));
        cand[q - clo] = c;
      }
      __syncthreads();
      // ---- converged sweep with the exact criterion
      const int qa = max(wq0, clo) - clo, qb = min(wq1, chi) - clo;
#pragma unroll TB_U
      for (int q = qa; q < qb; q++) {
        const double2 cxy = cand[q].xy, cze = cand[q].ze;
        const int ent = __double2loint(cze.y);
        const double rsq = rsq_nofma(Ai.x - cxy.x, Ai.y - cxy.y, Ai.z - cze.x);
        double cut = cut_i[0];
        if (!UNIFORM) {
          const int tj = (ent >> NEIGH_JBITS) & 7;
#pragma unroll
          for (int q2 = 1; q2 < MAXT; q2++) cut = tj == q2 ? cut_i[q2] : cut;
        }
        if (rsq <= cut && (ent & NEIGH_JMASK) != i) {
          if (LIST16) {
            const unsigned e16 = (unsigned)(q + clo) | (((unsigned)ent >> NEIGH_JBITS) & 7u) << TILE_SLOT_BITS | (((unsigned)ent >> 30) & 1u) << 15;
            acc = (acc >> 16) | ((unsigned long long)e16 << 48);
            n++;
            if ((n & 3) == 0 && n <= maxn) row64[(n >> 2) - 1] = acc;
          } else {
            if (n < maxn) out[(size_t)n * stride] = ent;
            n++;
          }
        }
      }
      // the staging area is reused only if another chunk or batch of atoms follows (CTA-uniform)
      if (chi < total || base + TB_T < last) __syncthreads();
    }
    if (valid) {
      if (LIST16 && (n & 3) && n < maxn) row64[n >> 2] = acc >> (16 * (4 - (n & 3)));   // partial last word (maxn is a multiple of 8)
      d.numneigh[i] = n < maxn ? n : maxn;
      atomicMax(&flags[2], n);
    }
  }
}

// (A') build_list_tile32_kernel (default since round 2): same tile geometry, candidate enumeration and output as (A),
//     but the lock-step sweep CLASSIFIES in FP32 and only the candidates the classification cannot decide see FP64:
//       * positions relative to the centre of the tile's halo box are rounded to float (p); a candidate is staged as
//         {-2 p.x, -2 p.y, -2 p.z, |p|^2} (16 B, ONE broadcast LDS.128) and atom i keeps p_i and c_i = |p_i|^2 - cut, so
//         t = r^2 - cut = fma(p_i.x, s.x, fma(p_i.y, s.y, fma(p_i.z, s.z, s.w))) + c_i costs 3 FFMA + 1 FADD (the
//         all-FP64 sweep: 8 FP64 instructions at a quarter of the issue rate); the sign of t is shifted into a bit word
//         (one funnel shift) and min |t| of 32 candidates is tracked (FMNMX3);
//       * with u = 2^-24, R_i = |p_i|, R_c = max |p| of the staged candidates (tracked while staging) and
//         M = (R_i + R_c)^2 + cut, the roundings of the inputs (2 r u (R_i + R_c)) and of the five float results
//         (each <= u M) keep |t - (rsq_ref - cut)| below 8 u M for EVERY candidate; `band` is twice that (about 3e-5
//         cut in the bulk).  |t| > band: the sign of t decides; otherwise (about 0.01 candidates per atom) the
//         reference's exact FP64 test (rsq_nofma on the FP64 records) does, so the pair set stays bit-exact
//         (tests: tile32 == tile64 == thread walk on every fixture, pairs placed ON the cutoff, golden pair sets);
//       * hits are compacted per 32-candidate block into a per-thread ring of 16-bit candidate indices in shared memory
//         (column of the thread: conflict-free) and written out by row flushes with the lanes in lock step on the
//         neighbour index k: a warp store covers one row of the transposed list (out[k * stride + i], consecutive i)
//         instead of one scattered 4-byte store per hit.
#ifndef TB32_CH
#define TB32_CH 1408
#endif
#ifndef TB32_KS
#define TB32_KS 64
#endif
constexpr int TB_CH32 = TB32_CH;              // staged candidates per chunk (a bulk halo holds about 1160)
constexpr int TB_KS = TB32_KS;                // ring rows per thread (power of two, >= 32)
constexpr int TB_NST = (TB_CH32 + TB_T - 1) / TB_T;   // staged candidates per thread and chunk
#ifndef TB_SB
#define TB_SB 4
#endif
constexpr size_t TB32_SMEM = (size_t)TB_CH32 * 16 + (size_t)TB_CH32 * 4 + (size_t)TB_KS * TB_T * 2;

template <bool UNIFORM, bool LIST16>
__global__ void __launch_bounds__(TB_T, 3)
build_list_tile32_kernel(const DevState d, const __grid_constant__ Grid g, const __grid_constant__ Coeffs co,
                         const int *__restrict__ cell_start, const int *__restrict__ gcell_start,
                         const int *__restrict__ gorder, const double cutmaxsq, int *flags) {
  extern __shared__ __align__(16) unsigned char tb_smem[];
  float4 *cand = reinterpret_cast<float4 *>(tb_smem);                          // {-2 p, |p|^2}
  int *ents = reinterpret_cast<int *>(tb_smem + (size_t)TB_CH32 * 16);         // global index, then the packed 32-bit list entry
  unsigned short *ring = reinterpret_cast<unsigned short *>(tb_smem + (size_t)TB_CH32 * 20);   // [row][thread]
  __shared__ int seg_src[TB_MAXSEG];
  __shared__ int seg_off[TB_MAXSEG + 1];
  __shared__ int layer_off[TB_MAXLAY + 1];
  __shared__ int qself[TB_T];
  __shared__ int s_rmax;

  const int tid = threadIdx.x;
  TileGeom t;
  if (!tile_geometry(g, blockIdx.x, cell_start, t)) return;   // empty tile (whole CTA)
  const int first = t.first, last = t.last, hz0 = t.hz0, hz1 = t.hz1, nz = t.nz, nseg = t.nseg;
  tile_segments(g, t, cell_start, gcell_start, d.nghost > 0, seg_src, seg_off);
  if (tid <= nz) layer_off[tid] = seg_off[tid * t.ny * 6];
  if (tid == 0) { atomicMax(&flags[4], seg_off[nseg]); s_rmax = 0; }
  qself[tid] = 0;
  __syncthreads();
  const int total = seg_off[nseg];
  const size_t stride = d.stride;
  // centre of the halo box: the float coordinates stay within a few cell sizes of zero
  const double ox = g.lo[0] + 0.5 * (t.hx0 + t.hx1 + 1) / g.inv[0];
  const double oy = g.lo[1] + 0.5 * (t.hy0 + t.hy1 + 1) / g.inv[1];
  const double oz = g.lo[2] + 0.5 * (t.hz0 + t.hz1 + 1) / g.inv[2];
  constexpr unsigned FULL = 0xffffffffu;

  for (int base = first; base < last; base += TB_T) {
    const int i = base + tid;
    const bool valid = i < last;
    Rec4 Ai = make_rec4(0, 0, 0, 0);
    int ti = 1;
    float pix = 0.f, piy = 0.f, piz = 0.f;
    double pi2 = 0.0;
    if (valid) {
      Ai = d.prec[i].A; ti = d.pflags[i] & 7;
      pix = __double2float_rn(Ai.x - ox); piy = __double2float_rn(Ai.y - oy); piz = __double2float_rn(Ai.z - oz);
      pi2 = (double)pix * pix + (double)piy * piy + (double)piz * piz;
    }
    // halo z-layers this WARP can reach
    int llo = TB_MAXLAY, lhi = -1;
    if (valid) {
      int cz = g.dim == 2 ? 0 : (int)floor((Ai.z - g.lo[2]) * g.inv[2]);
      cz = min(max(cz, hz0), hz1);
      llo = max(cz - g.s[2], hz0) - hz0;
      lhi = min(cz + g.s[2], hz1) - hz0;
    }
    llo = __reduce_min_sync(FULL, llo);
    lhi = __reduce_max_sync(FULL, lhi);
    int n = 0, nfl = 0;   // hits so far (= next row of this atom), rows already written out
    int *out = d.neigh + (valid ? i : first);
    const int maxn = valid ? (LIST16 ? d.pitch16 : d.maxneigh) : 0;
    unsigned long long *row64 = LIST16 ? reinterpret_cast<unsigned long long *>(d.neigh16 + (size_t)(valid ? i : first) * d.pitch16) : nullptr;
    unsigned long long acc = 0;
    const int wq0 = lhi >= 0 ? layer_off[llo] : 0, wq1 = lhi >= 0 ? layer_off[lhi + 1] : 0;

    for (int clo = 0; clo < total; clo += TB_CH32) {
      const int chi = min(total, clo + TB_CH32), cnt = chi - clo, cpad = (cnt + 31) & ~31;
      // ---- stage candidates clo .. chi.  Global indices segment by segment (a segment is a run of consecutive atoms) ...
      for (int sid = tid; sid < nseg; sid += TB_T) {
        const int off = seg_off[sid], src = seg_src[sid];
        const int k0 = max(0, clo - off), k1 = min(seg_off[sid + 1], chi) - off;
        for (int k = k0; k < k1; k++) ents[off + k - clo] = tile_source(src, k, d.nlocal, gorder);
      }
      __syncthreads();
      // ... then the records, TB_SB loads per thread in flight together
      float rmax = 0.f;
#pragma unroll
      for (int u0 = 0; u0 < TB_NST; u0 += TB_SB) {
        Rec4 A[TB_SB];
        int fl[TB_SB], jj[TB_SB];
#pragma unroll
        for (int u = 0; u < TB_SB; u++) {
          const int q = tid + (u0 + u) * TB_T;
          if (u0 + u < TB_NST && q < cnt) { jj[u] = ents[q]; A[u] = d.prec[jj[u]].A; fl[u] = d.pflags[jj[u]]; }
        }
#pragma unroll
        for (int u = 0; u < TB_SB; u++) {
          const int q = tid + (u0 + u) * TB_T;
          if (u0 + u >= TB_NST) continue;
          if (q < cnt) {
            const int j = jj[u];
            const float px = __double2float_rn(A[u].x - ox), py = __double2float_rn(A[u].y - oy), pz = __double2float_rn(A[u].z - oz);
            const float p2 = __double2float_rn((double)px * px + (double)py * py + (double)pz * pz);
            rmax = fmaxf(rmax, p2);
            cand[q] = make_float4(-2.f * px, -2.f * py, -2.f * pz, p2);
            ents[q] = j | ((fl[u] & 7) << NEIGH_JBITS) | (((fl[u] >> 4) & 1) << 30);
            if (j >= base && j < base + TB_T && j < last) qself[j - base] = q;
          } else if (q < cpad) {
            cand[q] = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));   // +inf: never inside
          }
        }
      }
      rmax = __int_as_float(__reduce_max_sync(FULL, __float_as_int(rmax)));   // non-negative floats order like their bit patterns
      if ((tid & 31) == 0) atomicMax(&s_rmax, __float_as_int(rmax));
      __syncthreads();
      // ---- error bound of the float classification (s_rmax only grows: later chunks stay covered)
      const double Rsum = sqrt(pi2) + sqrt((double)__int_as_float(s_rmax) * (1.0 + 1e-6));
      const float bandf = __double2float_ru(1.001 * (16.0 / 16777216.0) * (Rsum * Rsum + cutmaxsq) + 1e-12 * cutmaxsq);
      float ci_i[MAXT];
      double cut_i[MAXT];
#pragma unroll
      for (int q = 0; q < MAXT; q++) {
        cut_i[q] = UNIFORM ? cutmaxsq : co.cutneighsq[ti][q];
        ci_i[q] = valid ? __double2float_rn(pi2 - cut_i[q]) : __int_as_float(0x7f800000);
      }
      // this atom's own slot (its bit is cleared instead of testing every hit)
      int qs = qself[tid];
      if (!(valid && qs < cnt && (ents[qs] & NEIGH_JMASK) == i)) qs = -1;
      // ---- converged float sweep, 32 candidates per bit word
      const int qa = (max(wq0, clo) - clo) & ~31, qb = min((min(wq1, chi) - clo + 31) & ~31, cpad);
      for (int q0 = qa; q0 < qb; q0 += 32) {
        unsigned bits = 0;
        float tmin = 3.0e38f;
#pragma unroll
        for (int u = 0; u < 32; u++) {
          const float4 c = cand[q0 + u];
          float ci = ci_i[0];
          if (!UNIFORM) {
            const int tj = (ents[q0 + u] >> NEIGH_JBITS) & 7;
#pragma unroll
            for (int q2 = 1; q2 < MAXT; q2++) ci = tj == q2 ? ci_i[q2] : ci;
          }
          const float tt = fmaf(pix, c.x, fmaf(piy, c.y, fmaf(piz, c.z, c.w))) + ci;   // r^2 - cut; < 0: inside
          bits = __funnelshift_l((unsigned)__float_as_int(tt), bits, 1);   // (bits << 1) | sign(tt): candidate u ends up in bit 31 - u
          tmin = fminf(tmin, fabsf(tt));
        }
        bits = __brev(bits);
        if (tmin <= bandf) {   // rare: some candidate of the block is undecided in float -> the reference's FP64 criterion on the FP64 records
          for (int u = 0; u < 32; u++) {
            const float4 c = cand[q0 + u];
            const int ent = ents[q0 + u];
            float ci = ci_i[0];
            double cut = cut_i[0];
            if (!UNIFORM) {
              const int tj = (ent >> NEIGH_JBITS) & 7;
#pragma unroll
              for (int q2 = 1; q2 < MAXT; q2++) { ci = tj == q2 ? ci_i[q2] : ci; cut = tj == q2 ? cut_i[q2] : cut; }
            }
            const float tt = fmaf(pix, c.x, fmaf(piy, c.y, fmaf(piz, c.z, c.w))) + ci;
            if (fabsf(tt) <= bandf) {
              const Rec4 Aj = d.prec[ent & NEIGH_JMASK].A;
              const double rsq = rsq_nofma(Ai.x - Aj.x, Ai.y - Aj.y, Ai.z - Aj.z);
              if (rsq <= cut) bits |= 1u << u; else bits &= ~(1u << u);
            }
          }
        }
        if ((qs & ~31) == q0) bits &= ~(1u << (qs & 31));
        // ---- compaction into the ring; rows are written out when some lane could run out of ring
        if (__any_sync(FULL, n - nfl + __popc(bits) > TB_KS)) {
          const int k0 = __reduce_min_sync(FULL, nfl), k1 = __reduce_max_sync(FULL, n);
          for (int k = k0; k < k1; k++)
            if (k >= nfl && k < n) {
              const int q = ring[(k & (TB_KS - 1)) * TB_T + tid];
              const int ent = ents[q];
              if (LIST16) {
                const unsigned e16 = (unsigned)(q + clo) | (((unsigned)ent >> NEIGH_JBITS) & 7u) << TILE_SLOT_BITS | (((unsigned)ent >> 30) & 1u) << 15;
                acc = (acc >> 16) | ((unsigned long long)e16 << 48);
                if (((k + 1) & 3) == 0 && k < maxn) row64[k >> 2] = acc;
              } else if (k < maxn) {
                out[(size_t)k * stride] = ent;
              }
            }
          nfl = n;
        }
        while (bits) {
          const int bq = __ffs(bits) - 1;
          bits &= bits - 1;
          ring[(n & (TB_KS - 1)) * TB_T + tid] = (unsigned short)(q0 + bq);
          n++;
        }
      }
      // ---- the rest of the ring (its entries index this chunk's staging area)
      {
        const int k0 = __reduce_min_sync(FULL, nfl), k1 = __reduce_max_sync(FULL, n);
        for (int k = k0; k < k1; k++)
          if (k >= nfl && k < n) {
            const int q = ring[(k & (TB_KS - 1)) * TB_T + tid];
            const int ent = ents[q];
            if (LIST16) {
              const unsigned e16 = (unsigned)(q + clo) | (((unsigned)ent >> NEIGH_JBITS) & 7u) << TILE_SLOT_BITS | (((unsigned)ent >> 30) & 1u) << 15;
              acc = (acc >> 16) | ((unsigned long long)e16 << 48);
              if (((k + 1) & 3) == 0 && k < maxn) row64[k >> 2] = acc;
            } else if (k < maxn) {
              out[(size_t)k * stride] = ent;
            }
          }
        nfl = n;
      }
      // the staging area is reused only if another chunk or batch of atoms follows (CTA-uniform)
      if (chi < total || base + TB_T < last) __syncthreads();
    }
    if (valid) {
      if (LIST16 && (n & 3) && n < maxn) row64[n >> 2] = acc >> (16 * (4 - (n & 3)));   // partial last word (maxn is a multiple of 8)
      d.numneigh[i] = n < maxn ? n : maxn;
      atomicMax(&flags[2], n);
    }
  }
}

// tile form -> gather form: d.neigh[k * stride + i] = j | type_j << 27 | solid_j << 30 for every 16-bit entry.  One CTA
// per tile; slot -> global index through the same enumeration the builder used.  For sphbvf_get_pairs and for tests
// that run the gather kernel on a list built in tile form; never on the timestep path.
__global__ void __launch_bounds__(256)
expand_list_kernel(const DevState d, const __grid_constant__ Grid g, const int *__restrict__ cell_start,
                   const int *__restrict__ gcell_start, const int *__restrict__ gorder) {
  __shared__ int seg_src[TB_MAXSEG];
  __shared__ int seg_off[TB_MAXSEG + 1];
  __shared__ int gidx[TILE_MAX_SLOTS];
  TileGeom t;
  if (!tile_geometry(g, blockIdx.x, cell_start, t)) return;
  tile_segments(g, t, cell_start, gcell_start, d.nghost > 0, seg_src, seg_off);
  const int total = min(seg_off[t.nseg], TILE_MAX_SLOTS);
  for (int q = threadIdx.x; q < total; q += blockDim.x) {
    int lo = 0, hi = t.nseg;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (seg_off[mid] <= q) lo = mid; else hi = mid;
    }
    gidx[q] = tile_source(seg_src[lo], q - seg_off[lo], d.nlocal, gorder);
  }
  __syncthreads();
  for (int i = t.first + threadIdx.x; i < t.last; i += blockDim.x) {
    const int nn = d.numneigh[i];
    const unsigned short *row = d.neigh16 + (size_t)i * d.pitch16;
    for (int k = 0; k < nn; k++) {
      const unsigned e = row[k];
      d.neigh[(size_t)k * d.stride + i] = gidx[e & TILE_SLOT_MASK] | (int)((e >> TILE_SLOT_BITS) & 7u) << NEIGH_JBITS | (int)(e >> 15) << 30;
    }
  }
}

template <bool UNIFORM>
__global__ void __launch_bounds__(128, 6)
build_list_kernel(const DevState d, const __grid_constant__ Grid g, const __grid_constant__ Coeffs co,
                  const int *__restrict__ cell_start, const int *__restrict__ gcell_start,
                  const int *__restrict__ gorder, const double cutmaxsq, int *flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
  const Rec4 Ai = d.prec[i].A;
  const int ti = d.pflags[i] & 7;
  bool ok;
  const int ci = cell_of(g, Ai.x, Ai.y, co.dim == 2 ? g.lo[2] : Ai.z, ok);
  int cx, cy, cz;
  cell_coords(g, ci, cx, cy, cz);
  const double bsy = 1.0 / g.inv[1], bsz = 1.0 / g.inv[2];
  const int tmask = (1 << g.tb[0]) - 1;
  const bool have_ghosts = d.nghost > 0;
  int n = 0;
  int *out = d.neigh + i;
  auto emit = [&](int ent) {
    if (n < d.maxneigh) out[(size_t)n * d.stride] = ent;
    n++;
  };
  const int xlo = max(cx - g.s[0], 0), xhi = min(cx + g.s[0], g.n[0] - 1);
  const double xi = Ai.x, yi = Ai.y, zi = co.dim == 2 ? g.lo[2] : Ai.z;
  const double cutpad = cutmaxsq * (1.0 + 1e-10);
  for (int dz = -g.s[2]; dz <= g.s[2]; dz++) {
    const int z = cz + dz;
    if (z < 0 || z >= g.n[2]) continue;
    // distance from the atom to the slab of cells z (0 inside); NStencil::bin_distance
    // (nstencil.cpp:204-228) measures from the atom's CELL, this prunes from the atom itself
    const double zc = g.lo[2] + z * bsz;
    const double ddz = fmax(0.0, fmax(zc - zi, zi - (zc + bsz)));
    for (int dy = -g.s[1]; dy <= g.s[1]; dy++) {
      const int y = cy + dy;
      if (y < 0 || y >= g.n[1]) continue;
      const double yc = g.lo[1] + y * bsy;
      const double ddy = fmax(0.0, fmax(yc - yi, yi - (yc + bsy)));
      const double rem = cutpad - ddy * ddy - ddz * ddz;
      if (!(rem > 0.0)) continue;
      // cells of this row that can hold an atom within the cutoff: same floor() map as cell_of,
      // which is monotone, so the range is a superset whatever the rounding
      const double rx = sqrt(rem);
      const int xa = max(xlo, (int)floor((xi - rx - g.lo[0]) * g.inv[0]));
      const int xb = min(xhi, (int)floor((xi + rx - g.lo[0]) * g.inv[0]));
      const bool grow = have_ghosts && (y < g.glo[1] || y > g.ghi[1] || z < g.glo[2] || z > g.ghi[2]);
      for (int x = xa; x <= xb;) {
        const int xe = min(xb, x | tmask);
        const int c0 = cell_index(g, x, y, z), c1 = c0 + (xe - x);
        const bool gseg = have_ghosts && (grow || x < g.glo[0] || xe > g.ghi[0]);
        for (int pass = 0; pass < (gseg ? 2 : 1); pass++) {
          const int *start = pass ? gcell_start : cell_start;
          const int a = start[c0], b = start[c1 + 1];
          for (int p = a; p < b; p++) {
            const int j = pass ? d.nlocal + gorder[p] : p;
            if (j == i) continue;
            const Rec4 Aj = d.prec[j].A;
            const double rsq = rsq_nofma(Ai.x - Aj.x, Ai.y - Aj.y, Ai.z - Aj.z);
            if (UNIFORM) {
              if (rsq <= cutmaxsq) {
                const int fj = d.pflags[j];
                emit(j | ((fj & 7) << NEIGH_JBITS) | (((fj >> 4) & 1) << 30));
              }
            } else {
              const int fj = d.pflags[j];
              const int tj = fj & 7;
              if (rsq <= co.cutneighsq[ti][tj]) emit(j | (tj << NEIGH_JBITS) | (((fj >> 4) & 1) << 30));
            }
          }
        }
        x = xe + 1;
      }
    }
  }
  d.numneigh[i] = n < d.maxneigh ? n : d.maxneigh;
  atomicMax(&flags[2], n);
}

bool tile_form_possible(const Grid &g) {
  // the halo of a tile must fit the segment tables: stencil half-width <= 2 cells (always true for cells of
  // cutneigh/2, init_neighbor), <= 64 (y,z) rows, <= 12 z-layers
  const int ny = (1 << g.tb[1]) + 2 * g.s[1], nz = g.dim == 3 ? (1 << g.tb[2]) + 2 * g.s[2] : 1;
  return ny * nz <= TB_MAXROW && nz <= TB_MAXLAY;
}

// d.list16 selects the encoding (and with it the builder: the tile form exists only for the tile builder)
void launch_build_list(const DevState &d, const Grid &g, const Coeffs &co, const NeighWork &w, cudaStream_t st) {
  double cutmax = 0.0;
  bool uniform = true;
  for (int i = 1; i <= co.ntypes; i++)
    for (int j = 1; j <= co.ntypes; j++) {
      if (co.cutneighsq[i][j] > cutmax) cutmax = co.cutneighsq[i][j];
      if (co.cutneighsq[i][j] != co.cutneighsq[1][1]) uniform = false;
    }
  if (!d.nlocal) return;
  const char *env = getenv("SPHBVF_LIST_BUILD");   // read per rebuild so that tests can compare the builders
  const bool per_thread = env && env[0] == 't' && env[1] == 'h' && !d.list16;
  const bool sweep64 = env && !strcmp(env, "tile64");   // the all-FP64 sweep (A) instead of the float classification (A')
  if (!per_thread && tile_form_possible(g)) {
    const long ntiles = (long)g.nt[0] * g.nt[1] * g.nt[2];
    constexpr int smem64 = TB_CH * (int)sizeof(Cand);
    constexpr int smem32 = (int)TB32_SMEM;
    // the opt-in to > 48 KB of dynamic shared memory is per device (one process may drive several GPUs, each from
    // its own thread): setting it again is harmless, missing it is a launch failure
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      cudaFuncSetAttribute(build_list_tile_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64);
      cudaFuncSetAttribute(build_list_tile_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64);
      cudaFuncSetAttribute(build_list_tile_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64);
      cudaFuncSetAttribute(build_list_tile_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64);
      cudaFuncSetAttribute(build_list_tile32_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem32);
      cudaFuncSetAttribute(build_list_tile32_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem32);
      cudaFuncSetAttribute(build_list_tile32_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem32);
      cudaFuncSetAttribute(build_list_tile32_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem32);
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
#define TBK(K, SM, U, L) K<U, L><<<(int)ntiles, TB_T, SM, st>>>(d, g, co, w.cell_start, w.gcell_start, w.gorder, cutmax, w.flags)
    if (sweep64) {
      if (d.list16) { if (uniform) TBK(build_list_tile_kernel, smem64, true, true); else TBK(build_list_tile_kernel, smem64, false, true); }
      else { if (uniform) TBK(build_list_tile_kernel, smem64, true, false); else TBK(build_list_tile_kernel, smem64, false, false); }
    } else {
      if (d.list16) { if (uniform) TBK(build_list_tile32_kernel, smem32, true, true); else TBK(build_list_tile32_kernel, smem32, false, true); }
      else { if (uniform) TBK(build_list_tile32_kernel, smem32, true, false); else TBK(build_list_tile32_kernel, smem32, false, false); }
    }
#undef TBK
    SPHBVF_LAUNCHED(1);
    return;
  }
  const int nb = nblocks(d.nlocal, 128);
  if (uniform) build_list_kernel<true><<<nb, 128, 0, st>>>(d, g, co, w.cell_start, w.gcell_start, w.gorder, cutmax, w.flags);
  else build_list_kernel<false><<<nb, 128, 0, st>>>(d, g, co, w.cell_start, w.gcell_start, w.gorder, cutmax, w.flags);
  SPHBVF_LAUNCHED(1);
}

__global__ void tile_counts_kernel(const int *__restrict__ tile_order, const int ntiles, const int bits,
                                   const int *__restrict__ cell_start, int *cnt) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q > ntiles) return;
  if (q == ntiles) { cnt[q] = 0; return; }
  const long t = tile_order[q];
  cnt[q] = cell_start[(t + 1) << bits] - cell_start[t << bits];
}

__global__ void atom_order_kernel(const int *__restrict__ tile_order, const int ntiles, const int bits,
                                  const int *__restrict__ cell_start, const int *__restrict__ off, int *aorder,
                                  const int n_first, int *flags) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;   // one warp per tile
  if (q >= ntiles) return;
  const long t = tile_order[q];
  const int first = cell_start[t << bits], n = cell_start[(t + 1) << bits] - first, o = off[q];
  for (int k = lane; k < n; k += 32) aorder[o + k] = first + k;
  if (lane == 0 && q == n_first - 1) flags[5] = o + n;
  if (lane == 0 && q == 0 && n_first == 0) flags[5] = 0;
}

void launch_atom_order(const Grid &g, const NeighWork &w, const int *tile_order, int ntiles, int n_first, int *cnt, int *off,
                       int *aorder, cudaStream_t st) {
  if (ntiles <= 0) return;
  const int bits = g.tb[0] + g.tb[1] + g.tb[2];
  tile_counts_kernel<<<nblocks(ntiles + 1, 256), 256, 0, st>>>(tile_order, ntiles, bits, w.cell_start, cnt);
  SPHBVF_LAUNCHED(1);
  exclusive_scan(cnt, off, ntiles + 1, w.scan_tmp, st);
  atom_order_kernel<<<nblocks(ntiles, 8), 256, 0, st>>>(tile_order, ntiles, bits, w.cell_start, off, aorder, n_first, w.flags);
  SPHBVF_LAUNCHED(1);
}

void launch_expand_list(const DevState &d, const Grid &g, const NeighWork &w, cudaStream_t st) {
  if (!d.nlocal) return;
  const long ntiles = (long)g.nt[0] * g.nt[1] * g.nt[2];
  expand_list_kernel<<<(int)ntiles, 256, 0, st>>>(d, g, w.cell_start, w.gcell_start, w.gorder);
  SPHBVF_LAUNCHED(1);
}

void launch_copy_xhold(const DevState &d, cudaStream_t st) {
  cudaMemcpyAsync(d.xhold, d.x, sizeof(double) * 3 * (size_t)d.nlocal, cudaMemcpyDeviceToDevice, st);
}

}  // namespace sphbvf

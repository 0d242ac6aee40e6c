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

#include "sphbvf_internal.cuh"

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
  if (d.nlocal) check_distance_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, triggersq, flag_moved);
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
  if (d.nlocal) pbc_cellid_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, b, g, w.cellid, w.cell_count, w.flags);
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
    return;
  }
  scan_block_kernel<<<(int)nb, SCAN_T, 0, st>>>(in, out, n, tmp);
  exclusive_scan(tmp, tmp, nb, tmp + nb, st);
  scan_add_kernel<<<(int)nb, SCAN_T, 0, st>>>(out, n, tmp);
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
  cudaMemcpyAsync(arr, tmp, (size_t)tot * elem_bytes, cudaMemcpyDeviceToDevice, st);
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
  exclusive_scan(w.nimg, w.nimg, d.nlocal + 1, w.scan_tmp, st);
}

void launch_fill_images(const DevState &d, const Box &b, double cutghost, const NeighWork &w, cudaStream_t st) {
  if (d.nlocal) fill_images_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, b, cutghost, w.nimg);
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
constexpr int TB_NH = 1;                      // 2: stage the halo as two y-halves (24 KB chunks, 8 CTAs/SM): measured slower
constexpr int TB_MAXROW = 64;                 // (y,z) rows of the halo: 8 x 8 in 3D, 12 x 1 in 2D
constexpr int TB_MAXSEG = TB_MAXROW * 3 * 2;  // x-parts per row (<= 3 tiles) x {owned, ghost}
constexpr int TB_MAXLAY = 12;                 // z-layers of the halo (8 in 3D, 1 in 2D)

struct __align__(16) Cand {   // two LDS.128 broadcasts per candidate: {x, y} and {z, entry}
  double2 xy;
  double2 ze;
};

// TB_NH == 2 stages the halo in two halves (low-y rows, then high-y rows): every warp has work in both halves (a
// warp's atoms span the tile in y) and inside a half the rows stay z-major, so the z-layers a warp can reach are
// still one contiguous range.  With TB_NH == 1 the second half is empty.
template <bool UNIFORM>
__global__ void __launch_bounds__(TB_T)
build_list_tile_kernel(const DevState d, const __grid_constant__ Grid g, const __grid_constant__ Coeffs co,
                       const int *__restrict__ cell_start, const int *__restrict__ gcell_start,
                       const int *__restrict__ gorder, const double cutmaxsq, int *flags) {
  extern __shared__ __align__(16) unsigned char tb_smem[];
  Cand *cand = reinterpret_cast<Cand *>(tb_smem);
  __shared__ int seg_src[TB_MAXSEG];
  __shared__ int seg_off[TB_MAXSEG + 1];
  __shared__ int layer_off[2][TB_MAXLAY + 1];

  const int tid = threadIdx.x;
  const int bits = g.tb[0] + g.tb[1] + g.tb[2];
  const int tile = blockIdx.x;
  const int first = cell_start[(long)tile << bits], last = cell_start[((long)tile + 1) << bits];
  if (first == last) return;   // empty tile (whole CTA)
  const int tx = tile % g.nt[0], ty = (tile / g.nt[0]) % g.nt[1], tz = tile / (g.nt[0] * g.nt[1]);
  const int x0 = tx << g.tb[0], y0 = ty << g.tb[1], z0 = tz << g.tb[2];
  const int hx0 = max(x0 - g.s[0], 0), hx1 = min(x0 + (1 << g.tb[0]) - 1 + g.s[0], g.n[0] - 1);
  const int hy0 = max(y0 - g.s[1], 0), hy1 = min(y0 + (1 << g.tb[1]) - 1 + g.s[1], g.n[1] - 1);
  const int hz0 = max(z0 - g.s[2], 0), hz1 = min(z0 + (1 << g.tb[2]) - 1 + g.s[2], g.n[2] - 1);
  const int ny = hy1 - hy0 + 1, nz = hz1 - hz0 + 1;
  const int nya = TB_NH == 2 ? (ny + 1) >> 1 : ny;   // rows of the first half; the second has ny - nya
  const int nseg_a = nz * nya * 6, nseg = nz * ny * 6;
  const int tmask = (1 << g.tb[0]) - 1;
  const bool have_ghosts = d.nghost > 0;

  // ---- segment table: (half, z, y, x-part, owned|ghost) -> contiguous source range
  for (int sid = tid; sid < nseg; sid += TB_T) {
    const int half = sid >= nseg_a;
    const int rs = half ? sid - nseg_a : sid, nyh = half ? ny - nya : nya;
    const int pass = rs & 1, part = (rs >> 1) % 3, row = rs / 6;
    const int y = hy0 + (half ? nya : 0) + row % nyh, z = hz0 + row / nyh;
    int x = hx0, xe = min(hx1, x | tmask);
    for (int q = 0; q < part && x <= hx1; q++) { x = xe + 1; xe = min(hx1, x | tmask); }
    int a = 0, len = 0;
    if (x <= hx1 && (pass == 0 || have_ghosts)) {
      const int *start = pass ? gcell_start : cell_start;
      const int c0 = cell_index(g, x, y, z), c1 = c0 + (xe - x);
      a = start[c0];
      len = start[c1 + 1] - a;
    }
    seg_src[sid] = a | (pass << 31);
    seg_off[sid] = len;
  }
  __syncthreads();
  if (tid < 32) {   // exclusive scan of <= 384 lengths: 12 per lane
    constexpr int PER = TB_MAXSEG / 32;
    int v[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int sid = tid * PER + k;
      v[k] = sid < nseg ? seg_off[sid] : 0;
      sum += v[k];
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += y;
    }
    int run = incl - sum;
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int sid = tid * PER + k;
      if (sid <= nseg) seg_off[sid] = run;
      run += v[k];
    }
    if (tid == 31) seg_off[nseg] = run;   // nseg == TB_MAXSEG: no lane owns that slot
  }
  __syncthreads();
  if (tid <= nz) {
    layer_off[0][tid] = seg_off[tid * nya * 6];
    layer_off[1][tid] = seg_off[nseg_a + tid * (ny - nya) * 6];
  }
  __syncthreads();
  const size_t stride = d.stride;
#ifdef TB_DIAG_NOSTORE
  const bool nostore = flags[7] != 0;
#endif

  for (int base = first; base < last; base += TB_T) {
    const int i = base + tid;
    const bool valid = i < last;
    Rec4 Ai = make_rec4(0, 0, 0, 0);
    int ti = 1;
    if (valid) { Ai = d.prec[i].A; ti = d.pflags[i] & 7; }
    double cut_i[MAXT];
#pragma unroll
    for (int t = 0; t < MAXT; t++) cut_i[t] = UNIFORM ? cutmaxsq : co.cutneighsq[ti][t];
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
    const int maxn = valid ? d.maxneigh : 0;

    for (int half = 0; half < 2; half++) {
      const int h0 = half ? seg_off[nseg_a] : 0, h1 = half ? seg_off[nseg] : seg_off[nseg_a];
      const int wq0 = lhi >= 0 ? layer_off[half][llo] : 0, wq1 = lhi >= 0 ? layer_off[half][lhi + 1] : 0;
      for (int clo = h0; clo < h1; clo += TB_CH) {
        const int chi = min(h1, clo + TB_CH);
        // ---- stage candidates clo .. chi
        for (int q = clo + tid; q < chi; q += TB_T) {
          int lo = 0, hi = nseg;   // largest sid with seg_off[sid] <= q
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (seg_off[mid] <= q) lo = mid; else hi = mid;
          }
          const int src = seg_src[lo];
          const int p = (src & 0x7fffffff) + (q - seg_off[lo]);
          const int j = src < 0 ? d.nlocal + gorder[p] : p;
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
            for (int t = 1; t < MAXT; t++) cut = tj == t ? cut_i[t] : cut;
          }
          if (rsq <= cut && (ent & NEIGH_JMASK) != i) {
#ifdef TB_DIAG_NOSTORE   // tools/: measure what the scattered emission costs (the list keeps its previous content)
            if (n < maxn && !nostore) out[(size_t)n * stride] = ent;
#else
            if (n < maxn) out[(size_t)n * stride] = ent;
#endif
            n++;
          }
        }
        // the staging area is reused only if another chunk, half or batch of atoms follows (CTA-uniform)
        if (chi < h1 || (half == 0 && seg_off[nseg] > seg_off[nseg_a]) || base + TB_T < last) __syncthreads();
      }
    }
    if (valid) {
      d.numneigh[i] = n < d.maxneigh ? n : d.maxneigh;
      atomicMax(&flags[2], n);
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

void launch_build_list(const DevState &d, const Grid &g, const Coeffs &co, const NeighWork &w, cudaStream_t st) {
  double cutmax = 0.0;
  bool uniform = true;
  for (int i = 1; i <= co.ntypes; i++)
    for (int j = 1; j <= co.ntypes; j++) {
      if (co.cutneighsq[i][j] > cutmax) cutmax = co.cutneighsq[i][j];
      if (co.cutneighsq[i][j] != co.cutneighsq[1][1]) uniform = false;
    }
  if (!d.nlocal) return;
  const char *env = getenv("SPHBVF_LIST_BUILD");   // read per rebuild so that tests can compare both builders
  const bool per_thread = env && env[0] == 't';
  // the tile builder needs the halo of a tile to fit its tables: stencil half-width <= 2 cells (always true for
  // cells of cutneigh/2, init_neighbor) and <= 64 (y,z) rows
  const int ny = (1 << g.tb[1]) + 2 * g.s[1], nz = g.dim == 3 ? (1 << g.tb[2]) + 2 * g.s[2] : 1;
  if (!per_thread && ny * nz <= TB_MAXROW && nz <= TB_MAXLAY) {
    const long ntiles = (long)g.nt[0] * g.nt[1] * g.nt[2];
    constexpr int smem = TB_CH * (int)sizeof(Cand);
#ifdef TB_DIAG_NOSTORE
    { const char *e = getenv("SPHBVF_TB_NOSTORE"); const int v = e && atoi(e); cudaMemcpyAsync(w.flags + 7, &v, sizeof(int), cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st); }
#endif
    // the opt-in to > 48 KB of dynamic shared memory is per device (one process may drive several GPUs, each from
    // its own thread): setting it again is harmless, missing it is a launch failure
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      cudaFuncSetAttribute(build_list_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaFuncSetAttribute(build_list_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    if (uniform) build_list_tile_kernel<true><<<(int)ntiles, TB_T, smem, st>>>(d, g, co, w.cell_start, w.gcell_start, w.gorder, cutmax, w.flags);
    else build_list_tile_kernel<false><<<(int)ntiles, TB_T, smem, st>>>(d, g, co, w.cell_start, w.gcell_start, w.gorder, cutmax, w.flags);
    return;
  }
  const int nb = nblocks(d.nlocal, 128);
  if (uniform) build_list_kernel<true><<<nb, 128, 0, st>>>(d, g, co, w.cell_start, w.gcell_start, w.gorder, cutmax, w.flags);
  else build_list_kernel<false><<<nb, 128, 0, st>>>(d, g, co, w.cell_start, w.gcell_start, w.gorder, cutmax, w.flags);
}

void launch_copy_xhold(const DevState &d, cudaStream_t st) {
  cudaMemcpyAsync(d.xhold, d.x, sizeof(double) * 3 * (size_t)d.nlocal, cudaMemcpyDeviceToDevice, st);
}

}  // namespace sphbvf

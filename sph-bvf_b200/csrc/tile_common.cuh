// tile_common.cuh -- the candidate enumeration of one tile of the cell grid, shared by the tile list builder
// (kernels_neigh.cu), the list expansion and the tile-staged pair kernel (kernels_pair.cu).
//
// A tile is 4x4x4 cells in 3D (8x8 in 2D): a contiguous range of owned atoms.  Its CANDIDATES are the atoms of
// the cells within stencil reach of the tile (at most 8x8x8 cells, 12x12 in 2D), enumerated in a fixed order:
// (z, y) rows of cells, each row cut into the <= 3 x-parts that are contiguous in the tile-major cell numbering,
// each part first its owned atoms, then its ghosts (own cell table).  The position of a candidate in this
// enumeration is its SLOT; 16-bit neighbour-list entries of the tile form hold slots.  Everything here depends
// only on the frozen cell tables of the last rebuild, so the builder and every later pair pass see the same slots.
#pragma once

#include "sphbvf_internal.cuh"

namespace sphbvf {

constexpr int TB_MAXROW = 64;                 // (y,z) rows of the halo: 8 x 8 in 3D, 12 x 1 in 2D
constexpr int TB_MAXSEG = TB_MAXROW * 3 * 2;  // x-parts per row (<= 3 tiles) x {owned, ghost}
constexpr int TB_MAXLAY = 12;                 // z-layers of the halo (8 in 3D, 1 in 2D)
constexpr int TILE_SLOT_BITS = 12;            // 16-bit entry: slot | type_j << 12 | solid_j << 15
constexpr int TILE_SLOT_MASK = (1 << TILE_SLOT_BITS) - 1;
constexpr int TILE_MAX_SLOTS = 1 << TILE_SLOT_BITS;

struct TileGeom {
  int first, last;            // owned atoms of the tile
  int hx0, hx1, hy0, hy1, hz0, hz1;   // halo cell box
  int ny, nz, nseg;
};

// false: the tile holds no owned atom (the whole CTA can leave)
__device__ __forceinline__ bool tile_geometry(const Grid &g, const int tile, const int *__restrict__ cell_start, TileGeom &t) {
  const int bits = g.tb[0] + g.tb[1] + g.tb[2];
  t.first = cell_start[(long)tile << bits];
  t.last = cell_start[((long)tile + 1) << bits];
  if (t.first == t.last) return false;
  const int tx = tile % g.nt[0], ty = (tile / g.nt[0]) % g.nt[1], tz = tile / (g.nt[0] * g.nt[1]);
  const int x0 = tx << g.tb[0], y0 = ty << g.tb[1], z0 = tz << g.tb[2];
  t.hx0 = max(x0 - g.s[0], 0); t.hx1 = min(x0 + (1 << g.tb[0]) - 1 + g.s[0], g.n[0] - 1);
  t.hy0 = max(y0 - g.s[1], 0); t.hy1 = min(y0 + (1 << g.tb[1]) - 1 + g.s[1], g.n[1] - 1);
  t.hz0 = max(z0 - g.s[2], 0); t.hz1 = min(z0 + (1 << g.tb[2]) - 1 + g.s[2], g.n[2] - 1);
  t.ny = t.hy1 - t.hy0 + 1;
  t.nz = t.hz1 - t.hz0 + 1;
  t.nseg = t.nz * t.ny * 6;
  return true;
}

// Segment table of the tile: seg_src[sid] = first source index (bit 31: ghost table), seg_off[sid] = first slot,
// seg_off[nseg] = number of candidates; sid = ((z * ny + y) * 3 + part) * 2 + pass.  Called by every thread of the
// CTA (>= 32 threads); ends with a barrier.
__device__ __forceinline__ void tile_segments(const Grid &g, const TileGeom &t, const int *__restrict__ cell_start,
                                              const int *__restrict__ gcell_start, const bool have_ghosts,
                                              int *seg_src, int *seg_off) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int tmask = (1 << g.tb[0]) - 1;
  for (int sid = tid; sid < t.nseg; sid += nthr) {
    const int pass = sid & 1, part = (sid >> 1) % 3, row = sid / 6;
    const int y = t.hy0 + row % t.ny, z = t.hz0 + row / t.ny;
    int x = t.hx0, xe = min(t.hx1, x | tmask);
    for (int q = 0; q < part && x <= t.hx1; q++) { x = xe + 1; xe = min(t.hx1, x | tmask); }
    int a = 0, len = 0;
    if (x <= t.hx1 && (pass == 0 || have_ghosts)) {
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
      v[k] = sid < t.nseg ? seg_off[sid] : 0;
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
      if (sid <= t.nseg) seg_off[sid] = run;
      run += v[k];
    }
    if (tid == 31) seg_off[t.nseg] = run;   // nseg == TB_MAXSEG: no lane owns that slot
  }
  __syncthreads();
}

// global index of candidate `k` of segment `src`
__device__ __forceinline__ int tile_source(const int src, const int k, const int nlocal, const int *__restrict__ gorder) {
  const int p = (src & 0x7fffffff) + k;
  return src < 0 ? nlocal + gorder[p] : p;
}

}  // namespace sphbvf

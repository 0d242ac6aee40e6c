// sphbvf_internal.cuh -- shared declarations of the CUDA implementation behind include/sphbvf.h
//
// Data layout in HBM (all FP64 / int32, device resident for the whole run):
//   primary state, owned atoms only, cell-sorted order, LAMMPS-compatible row layout:
//     x v vest [n][3] | rho rhoI e [n] | C [n][S] | dev [n][9] | tag type mask solid fixed slot [n]
//   pair outputs, owned atoms (written once per step by the pair kernel, no atomics, no memset):
//     f nw ddv ddx [n][3] | drho phi nd rhoAux1 rhoAux2 Pnew [n] | ddev [n][9] | Q [n][S]
//   packed pair inputs, owned + ghost atoms: one 96-byte record (three 32-byte aligned quarters, six 16-byte
//   granules) per atom, so a neighbour visit is three 256-bit loads that touch one or two cache lines:
//     A = {x, y, z, rho}   B = {vest.x, vest.y, vest.z, V = m/rho}   C = {w.x, w.y, w.z, P/rho^2}
//     with w = vest - v (momentum minus transport velocity); pD = {rhoI, art, C0, e} is only read on
//     Shepard-filter steps and by the stochastic term; pCs [nall][S], pdev [nall][9].
//   neighbour structure: full (both directions) Verlet list of the owned atoms, frozen between
//   rebuilds exactly like the reference's list, in one of two encodings chosen at every rebuild:
//     list16 = 1 (tile form, SPHBVF_PAIR=tile): 16-bit entries = slot of j among the staged candidates of i's tile
//       (12 bits) | type_j (3 bits) | solid_tag_j (1 bit), row-major neigh16[i * pitch16 + k]: 2 B per
//       neighbour, consumed by pair_tile_kernel (one CTA per tile, candidates staged in shared memory);
//     list16 = 0 (gather form, default): 32-bit entries j (27 bits) | type_j (3) | solid_tag_j (1), stored
//       TRANSPOSED neigh[k * stride + i] so that the 32 lanes of a warp read 128 contiguous bytes per k,
//       consumed by pair_kernel (one thread per atom, records gathered through L1).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sphbvf.h"

namespace sphbvf {

constexpr int MAXT = 5;  // atom types 1..4
constexpr int MAXS = 4;  // species
constexpr int MAXFIX = 16;
constexpr int NEIGH_JBITS = 27;
constexpr int NEIGH_JMASK = (1 << NEIGH_JBITS) - 1;

// per-type / per-type-pair coefficients, passed to kernels by value (constant bank)
struct Coeffs {
  int dim, variant, nspecies, ntypes;
  double mass[MAXT], rho0[MAXT], c0[MAXT], B[MAXT], G0[MAXT];
  double eta[MAXT][MAXT], cut[MAXT][MAXT], cutsq[MAXT][MAXT], cutc[MAXT][MAXT];
  double cutneighsq[MAXT][MAXT];
  double kappa[MAXT][MAXT][MAXS];
};

// 32-byte aligned record: one LDG.E.ENL2.256 per gather on sm_100a
struct __align__(32) Rec4 {
  double x, y, z, w;
};
__host__ __device__ __forceinline__ Rec4 make_rec4(double x, double y, double z, double w) {
  Rec4 r;
  r.x = x; r.y = y; r.z = z; r.w = w;
  return r;
}

// the record the pair kernels read for every neighbour: 96 bytes, 32-byte aligned (1.33 particles per 128-byte
// line, no padding):  A = {x, y, z, rho}   B = {vest.x, vest.y, vest.z, V}   C = {w.x, w.y, w.z, P/rho^2}
struct __align__(32) Prec {
  Rec4 A, B, C;
};
constexpr int PREC_GRANULES = 6;   // 16-byte granules per record

// cell grid used for sorting and for the list build (covers sub-box + ghost shell).
// Cells are numbered TILE-major: the grid is cut into tiles of 2^tb[0] x 2^tb[1] x 2^tb[2] cells
// (4x4x4 in 3D, 8x8x1 in 2D = 64 cells, about one CTA's worth of atoms), tiles x-fastest, cells
// x-fastest inside a tile.  Consecutive atoms therefore fill compact cubes instead of pencils,
// which is what gives the pair kernel's gathers their L1 hit rate.
struct Grid {
  double lo[3];      // grid origin
  double inv[3];     // 1 / cell size
  int n[3];          // cells per dimension
  int s[3];          // stencil half-width in cells
  int tb[3];         // log2 of the tile edge per dimension
  int nt[3];         // tiles per dimension
  int glo[3], ghi[3];  // cells with coordinate < glo or > ghi (any dim) may hold ghost atoms
  int dim;
  long ncells;       // nt[0]*nt[1]*nt[2] << (tb[0]+tb[1]+tb[2])  (>= n[0]*n[1]*n[2])
};

__host__ __device__ __forceinline__ int cell_index(const Grid &g, int cx, int cy, int cz) {
  const int tx = cx >> g.tb[0], ty = cy >> g.tb[1], tz = cz >> g.tb[2];
  const int lx = cx & ((1 << g.tb[0]) - 1), ly = cy & ((1 << g.tb[1]) - 1), lz = cz & ((1 << g.tb[2]) - 1);
  const int tile = (tz * g.nt[1] + ty) * g.nt[0] + tx;
  return (tile << (g.tb[0] + g.tb[1] + g.tb[2])) | (((lz << g.tb[1]) | ly) << g.tb[0]) | lx;
}
__host__ __device__ __forceinline__ void cell_coords(const Grid &g, int c, int &cx, int &cy, int &cz) {
  const int bits = g.tb[0] + g.tb[1] + g.tb[2];
  const int tile = c >> bits, l = c & ((1 << bits) - 1);
  const int tx = tile % g.nt[0], ty = (tile / g.nt[0]) % g.nt[1], tz = tile / (g.nt[0] * g.nt[1]);
  cx = (tx << g.tb[0]) | (l & ((1 << g.tb[0]) - 1));
  cy = (ty << g.tb[1]) | ((l >> g.tb[0]) & ((1 << g.tb[1]) - 1));
  cz = (tz << g.tb[2]) | (l >> (g.tb[0] + g.tb[1]));
}

struct Box {
  double lo[3], hi[3], prd[3];   // global box
  double sublo[3], subhi[3];     // this rank's brick
  int periodic[3];
  int dim;
};

// Shepard filter of the density (fix_...transport_velocity.cpp:287, ..._mechanics.cpp:311: every 20 steps; the fsi fix
// has freqFilter = 1e16 -> INT_MAX, ..._fsi.cpp:304): ONE predicate for the integrator that consumes rhoAux1/2 and
// for the pair pass that has to compute the numerator on exactly those steps
__host__ __device__ __forceinline__ bool shepard_filter_step(int variant, long step) {
  return variant == SPHBVF_FSI ? (step % 2147483647L) == 0 : (step % 20) == 0;
}

enum FixKind { FIX_BUOYANCY = 0, FIX_FORCING = 1, FIX_BUFFER = 2, FIX_SETFORCE = 3, FIX_CHEMRXN = 4 };
struct FixDesc {
  int kind, groupbit;
  int ia[4];
  long step;
  double a[6];
};

// the primary per-atom state (the arrays a rebuild reorders); DevState holds two sets and swaps them
struct StateArrays {
  int *tag, *type, *mask, *solid, *fixed, *slot;
  double *x, *v, *vest, *rho, *rhoI, *e, *C, *dev;
};

// device pointers of one context (plain struct so kernels can take it by value)
struct DevState {
  int nlocal, nghost, nmax, nallmax;
  // primary
  int *tag, *type, *mask, *solid, *fixed, *slot;
  double *x, *v, *vest, *rho, *rhoI, *e, *C, *dev;
  StateArrays alt;      // second buffer of the primary state: a rebuild gathers into it and swaps (no copy back)
  // pair outputs
  double *f, *nw, *ddv, *ddx, *drho, *phi, *nd, *rhoAux1, *rhoAux2, *Pnew, *ddev, *Q;
  // packed pair inputs (owned + ghost)
  Prec *prec;
  Rec4 *pD;
  double *pCs, *pdev;
  int *pflags, *ptag;
  // rebuild bookkeeping
  double *xhold;
  // ghosts that image owned atoms of this rank (periodic self images)
  int *gowner;          // owner index (owned atom) of self-image ghost g
  double *gshift;       // [nghost][3]
  // neighbour list (see the header comment for the two encodings)
  int *neigh;           // [maxneigh][stride]   gather form (also the expanded copy sphbvf_get_pairs reads)
  int *numneigh;        // [nlocal]
  int stride, maxneigh;
  unsigned short *neigh16;   // [nmax][pitch16]  tile form
  int pitch16;          // entries per row, multiple of 8
  int list16;           // encoding of the CURRENT list
  int tile_cap;         // tile form: staged slots per tile (>= the largest candidate count, multiple of 8)
};

// kernels launched by the calling thread (every launch site of the library bumps it); the context attributes the
// difference between tic() and toc() to a kernel family, so sphbvf_launch_count is a count, not an estimate
extern thread_local long tl_launches;
#define SPHBVF_LAUNCHED(n) (::sphbvf::tl_launches += (n))

enum KernelFamily { K_PAIR = 0, K_INITIAL = 1, K_FINAL = 2, K_NEIGH = 3, K_PACK = 4, K_FIX = 5, K_FUSED = 6, K_NFAM = 7 };

}  // namespace sphbvf

// -------- launchers implemented in the .cu files (all asynchronous on `st`) -----------------
namespace sphbvf {

// kernels_integrate.cu
void launch_setup_pre_force(const DevState &d, int groupbit, cudaStream_t st);
void launch_initial_integrate(const DevState &d, const Coeffs &co, double dt, long ntimestep,
                              int groupbit, int with_dev, cudaStream_t st);
void launch_final_integrate(const DevState &d, const Coeffs &co, double dt, long ntimestep,
                            int groupbit, int with_dev, cudaStream_t st);
// final_integrate(step_final) + initial_integrate(step_init) [+ pack] in one pass (bit-identical to the two calls)
void launch_final_initial(const DevState &d, const Coeffs &co, double dt_final, long step_final, double dt_init,
                          long step_init, int groupbit, int do_pack, int with_dev, cudaStream_t st, const int *order = nullptr,
                          int a0 = 0, int a1 = -1);
void launch_max_vsq(const DevState &d, int groupbit, unsigned long long *out, cudaStream_t st);
// out[0] = some atom has solid_tag, out[1] = some solid can carry deviatoric stress (G0 of its type != 0) or some
// dev != 0, out[2] = some e != 0: what selects the pair-kernel instantiation.  out must be zeroed by the caller.
void launch_derive_flags(const DevState &d, const Coeffs &co, int *out, cudaStream_t st);
void launch_ke_tensor(const DevState &d, const Coeffs &co, int groupbit, double *scratch, double *out6, cudaStream_t st);
void launch_fix(const DevState &d, const Coeffs &co, const FixDesc &fx, int hook, long ntimestep,
                cudaStream_t st);   // hook: 0 post_integrate, 1 post_force, 2 end_of_step
bool fix_runs(const FixDesc &fx, int hook, long ntimestep);   // would launch_fix launch anything?
// pack owned atoms into pA..pD (+pCs, pdev) and refresh self-image ghosts from their owners
void launch_pack(const DevState &d, const Coeffs &co, int with_dev, cudaStream_t st);
void launch_ghost_refresh(const DevState &d, const Coeffs &co, int with_dev, cudaStream_t st);

// kernels_pair.cu
struct NeighWork;
struct PairFlags {
  int filter_step;   // Shepard sums rhoAux1/2 are consumed this step
  int uniform;       // every type pair shares h, eta and the masses are equal
  int with_dev;      // deviatoric tensors may be non-zero (elastic solids present): bit t = solid atoms of type t can carry stress
  int any_solid;     // some atom has solid_tag == 1
  double damp;       // density-diffusion amplitude of the fsi variant (0 otherwise)
  int random;        // stochastic stress term on (some e != 0 and sphbvf_set_random was called)
  double rand_pref;  // 4 kB / dt
  unsigned long long seed;
  long ntimestep;
};
// part of the owned atoms a pair launch covers (the overlapped halo runs the interior first): tile form -> a list of
// tiles; gather form -> positions [a0, a1) of an atom order (nullptr: the atoms themselves)
struct PairSubset {
  const int *tile_list;
  int ntiles;
  const int *aorder;
  int a0, a1;
  int *queues;   // gather form: nq + 1 chunk counters of the persistent, SM-local schedule (nullptr: one CTA per chunk)
  int nq;
};
void launch_pair(const DevState &d, const Coeffs &co, const PairFlags &pf, const Grid &g, const NeighWork &w,
                 const PairSubset *part, cudaStream_t st);
void launch_virial(const DevState &d, const Coeffs &co, const PairFlags &pf, const Grid &g, const NeighWork &w, double *out6,
                   cudaStream_t st);
// dynamic shared memory pair_tile_kernel needs for `cap` staged slots: 96 B per record and 4 B per slot for the
// slot -> global index map of the instantiations that gather extras (species, deviatoric tensors, rhoI, noise)
size_t pair_tile_smem(int cap, bool index_map);

// kernels_neigh.cu
struct NeighWork {      // scratch owned by the context
  int *cellid;          // [nallmax]
  int *perm;            // [nmax]  new position -> old index (owned)
  int *cell_count;      // [ncells+1]
  int *cell_start;      // [ncells+1]  exclusive scan over owned atoms
  int *gcell_count, *gcell_start;   // same for ghosts
  int *gorder;          // [nghost] ghost indices (relative to nlocal) sorted by cell
  int *scan_tmp;        // block sums for the scan
  long ncells_cap;
  int *nimg;            // [nmax+1] images per owned atom and its exclusive scan
  int *flags;           // device flags: [0] nonfinite, [1] lost, [2] max neighbours, [3] moved, [4] max candidates of a tile
  void *tmp_perm;       // staging buffer for the permutation of one array
  size_t tmp_perm_bytes;
};
// all return cudaError_t of the enqueue; results that the host needs are in w.flags
void launch_check_distance(const DevState &d, double triggersq, int *flag_moved, cudaStream_t st);
void launch_cell_ids(const DevState &d, const Grid &g, const Box &b, const NeighWork &w, cudaStream_t st);
void exclusive_scan(const int *in, int *out, long n, int *tmp, cudaStream_t st);  // out has n+1 entries
void launch_sort_owned(const DevState &d, const Grid &g, const NeighWork &w, cudaStream_t st);
void launch_permute(void *arr, void *tmp, const int *perm, int n, int ncols, int elem_bytes, cudaStream_t st);
// out.<field>[i] = in.<field>[perm[i]] for every primary array in ONE pass (dev only if with_dev, C only if S)
void launch_gather_state(const StateArrays &in, const StateArrays &out, const int *perm, int n, int S, int with_dev, cudaStream_t st);
void launch_count_images(const DevState &d, const Box &b, double cutghost, const NeighWork &w, cudaStream_t st);
void launch_fill_images(const DevState &d, const Box &b, double cutghost, const NeighWork &w, cudaStream_t st);
void launch_bin_ghosts(const DevState &d, const Grid &g, const NeighWork &w, cudaStream_t st);
void launch_build_list(const DevState &d, const Grid &g, const Coeffs &co, const NeighWork &w, cudaStream_t st);
bool tile_form_possible(const Grid &g);   // the halo of a tile fits the tile kernels' tables
// owned atoms in the order of `tile_order` (a permutation of the tiles): aorder[p] = atom index; flags[5] = number of
// atoms in the first `n_first` tiles.  cnt / off: scratch of ntiles + 1 ints each.
void launch_atom_order(const Grid &g, const NeighWork &w, const int *tile_order, int ntiles, int n_first, int *cnt, int *off,
                       int *aorder, cudaStream_t st);
// tile form -> gather form (d.neigh, transposed 32-bit entries) for sphbvf_get_pairs and cross-checks
void launch_expand_list(const DevState &d, const Grid &g, const NeighWork &w, cudaStream_t st);
void launch_copy_xhold(const DevState &d, cudaStream_t st);
long count_pairs_host(const DevState &d, cudaStream_t st);

}  // namespace sphbvf

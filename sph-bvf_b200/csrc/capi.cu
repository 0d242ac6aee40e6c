// capi.cu -- the extern "C" layer of include/sphbvf.h: context, device memory, and the
// timestep driver that sequences the kernels in the order Verlet::setup / Verlet::run call the
// reference's styles (verlet.cpp:88-170, 223-354).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "sphbvf_internal.cuh"
#include "tile_common.cuh"
#include "context.cuh"

using namespace sphbvf;

namespace sphbvf {
thread_local long tl_launches = 0;
}

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
int sphbvf_ctx::fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  err = buf;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return ctx->fail(SPHBVF_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define CKLAUNCH() CK(cudaGetLastError())

template <typename T>
static cudaError_t dev_realloc(T *&p, size_t old_n, size_t new_n, cudaStream_t st, bool keep, bool zero) {
  T *q = nullptr;
  if (new_n == 0) new_n = 1;
  cudaError_t e = cudaMalloc((void **)&q, new_n * sizeof(T));
  if (e != cudaSuccess) return e;
  if (zero) cudaMemsetAsync(q, 0, new_n * sizeof(T), st);
  if (p && keep && old_n) cudaMemcpyAsync(q, p, std::min(old_n, new_n) * sizeof(T), cudaMemcpyDeviceToDevice, st);
  if (p) {
    cudaStreamSynchronize(st);
    cudaFree(p);
  }
  p = q;
  return cudaSuccess;
}

// tic(fam) ... toc(): every kernel the calling thread launches in between is attributed to `fam` (the launch sites
// bump sphbvf::tl_launches), and with profiling on the span is bracketed by two CUDA events on the context's stream
void sphbvf_ctx::tic(int fam) {
  open_fam = fam;
  launch_mark = tl_launches;
  if (!profiling) return;
  cudaEvent_t a, b;
  if (ev_pool.size() >= 2) {
    a = ev_pool.back(); ev_pool.pop_back();
    b = ev_pool.back(); ev_pool.pop_back();
  } else {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
  }
  cudaEventRecord(a, st);
  ev_open = {fam, a, b};
}

void sphbvf_ctx::toc() {
  if (open_fam >= 0) launches_fam[open_fam] += tl_launches - launch_mark;
  open_fam = -1;
  if (!profiling) return;
  cudaEventRecord(ev_open.b, st);
  ev_list.push_back(ev_open);
}

void sphbvf_ctx::drain_events() {
  for (auto &e : ev_list) {
    float ms = 0.f;
    cudaEventSynchronize(e.b);
    cudaEventElapsedTime(&ms, e.a, e.b);
    ms_fam[e.fam] += ms;
    ev_pool.push_back(e.a);
    ev_pool.push_back(e.b);
  }
  ev_list.clear();
}

// staging buffer shared by the permutation of the rebuild, uploads / downloads and reductions
static int stage(sphbvf_ctx *ctx, size_t bytes) {
  if (bytes > ctx->w.tmp_perm_bytes) {
    if (ctx->w.tmp_perm) cudaFree(ctx->w.tmp_perm);
    ctx->w.tmp_perm_bytes = bytes;
    CK(cudaMalloc(&ctx->w.tmp_perm, bytes));
  }
  return 0;
}

// see context.cuh: a deferred final_integrate runs here unless sphbvf_initial_integrate absorbed it
int flush_final(sphbvf_ctx *ctx) {
  if (!ctx->final_pending) return 0;
  ctx->final_pending = 0;
  ctx->tic(K_FINAL);
  launch_final_integrate(ctx->d, ctx->co, ctx->pend_dt, ctx->pend_step, ctx->cfg.integrate_groupbit, ctx->with_dev, ctx->st);
  ctx->toc();
  CKLAUNCH();
  return 0;
}
#define FLUSH()                                  \
  do {                                           \
    int rcf_ = flush_final(ctx);                 \
    if (rcf_) return rcf_;                       \
  } while (0)

int make_tile_order(sphbvf_ctx *ctx);

static int ensure_scan(sphbvf_ctx *ctx, long n) {
  const long need = n / 1024 + 8192;
  if (need <= ctx->scan_cap) return 0;
  if (ctx->w.scan_tmp) cudaFree(ctx->w.scan_tmp);
  ctx->w.scan_tmp = nullptr;
  CK(cudaMalloc((void **)&ctx->w.scan_tmp, sizeof(int) * need));
  ctx->scan_cap = need;
  return 0;
}

// ensure capacity for `nmax` owned atoms and `nallmax` owned+ghost atoms
static int ensure_capacity(sphbvf_ctx *ctx, int nmax, int nallmax) {
  DevState &d = ctx->d;
  const int S = ctx->co.nspecies > 0 ? ctx->co.nspecies : 1;
  cudaStream_t st = ctx->st;
  if (nmax > d.nmax) {
    const size_t o = d.nmax, n = nmax;
#define RE(p, w, keep) CK(dev_realloc(p, o *(w), n *(w), st, keep, true))
    RE(d.tag, 1, true); RE(d.type, 1, true); RE(d.mask, 1, true); RE(d.solid, 1, true); RE(d.fixed, 1, true); RE(d.slot, 1, true);
    RE(d.x, 3, true); RE(d.v, 3, true); RE(d.vest, 3, true); RE(d.rho, 1, true); RE(d.rhoI, 1, true); RE(d.e, 1, true);
    RE(d.C, S, true); RE(d.dev, 9, true);
    RE(d.alt.tag, 1, false); RE(d.alt.type, 1, false); RE(d.alt.mask, 1, false); RE(d.alt.solid, 1, false);
    RE(d.alt.fixed, 1, false); RE(d.alt.slot, 1, false);
    RE(d.alt.x, 3, false); RE(d.alt.v, 3, false); RE(d.alt.vest, 3, false); RE(d.alt.rho, 1, false); RE(d.alt.rhoI, 1, false);
    RE(d.alt.e, 1, false); RE(d.alt.C, S, false); RE(d.alt.dev, 9, false);
    RE(d.f, 3, true); RE(d.nw, 3, true); RE(d.ddv, 3, true); RE(d.ddx, 3, true); RE(d.drho, 1, true); RE(d.phi, 1, true);
    RE(d.nd, 1, true); RE(d.rhoAux1, 1, true); RE(d.rhoAux2, 1, true); RE(d.Pnew, 1, true); RE(d.ddev, 9, true); RE(d.Q, S, true);
    RE(d.xhold, 3, true); RE(d.numneigh, 1, true);
    RE(ctx->w.perm, 1, false);
#undef RE
    CK(dev_realloc(ctx->w.nimg, 0, n + 1, st, false, true));   // nmax+1 entries (scan sentinel)
    { int rc_ = ensure_scan(ctx, (long)n + 2); if (rc_) return rc_; }
    if ((size_t)9 * n * sizeof(double) > ctx->w.tmp_perm_bytes) {
      if (ctx->w.tmp_perm) cudaFree(ctx->w.tmp_perm);
      ctx->w.tmp_perm_bytes = (size_t)9 * n * sizeof(double);
      CK(cudaMalloc(&ctx->w.tmp_perm, ctx->w.tmp_perm_bytes));
    }
    // neighbour list storage is tied to nmax (stride of the gather form, rows of the tile form): dropped here,
    // re-created by ensure_neigh at the next rebuild (growth only happens on the way into one)
    if (d.neigh) cudaFree(d.neigh);
    if (d.neigh16) cudaFree(d.neigh16);
    d.neigh = nullptr;
    d.neigh16 = nullptr;
    d.stride = nmax;
    d.nmax = nmax;
  }
  if (nallmax > d.nallmax) {
    const size_t o = d.nallmax, n = nallmax;
#define RE(p, w) CK(dev_realloc(p, o *(w), n *(w), st, true, true))
    RE(d.prec, 1); RE(d.pD, 1); RE(d.pCs, S); RE(d.pdev, 9); RE(d.pflags, 1); RE(d.ptag, 1);
    RE(d.gowner, 1); RE(d.gshift, 3); RE(ctx->w.cellid, 1); RE(ctx->w.gorder, 1);
#undef RE
    d.nallmax = nallmax;
  }
  return 0;
}

static int ensure_cells(sphbvf_ctx *ctx, long ncells) {
  NeighWork &w = ctx->w;
  if (ncells + 1 <= w.ncells_cap) return 0;
  const long cap = ncells + 1 + ncells / 8;
  for (int **p : {&w.cell_count, &w.cell_start, &w.gcell_count, &w.gcell_start}) {
    if (*p) cudaFree(*p);
    CK(cudaMalloc((void **)p, sizeof(int) * cap));
  }
  w.ncells_cap = cap;
  return ensure_scan(ctx, cap);
}

// storage of the Verlet list in the encoding d.list16 selects, for `maxneigh` entries per atom
static int ensure_neigh(sphbvf_ctx *ctx, int maxneigh) {
  DevState &d = ctx->d;
  if (maxneigh > d.maxneigh) {   // growth invalidates both encodings
    if (d.neigh) cudaFree(d.neigh);
    if (d.neigh16) cudaFree(d.neigh16);
    d.neigh = nullptr;
    d.neigh16 = nullptr;
    d.maxneigh = maxneigh;
  }
  d.pitch16 = (d.maxneigh + 7) & ~7;
  if (d.list16 && !d.neigh16) CK(cudaMalloc((void **)&d.neigh16, sizeof(unsigned short) * (size_t)d.pitch16 * d.nmax));
  if (!d.list16 && !d.neigh) CK(cudaMalloc((void **)&d.neigh, sizeof(int) * (size_t)d.maxneigh * d.stride));
  return 0;
}

// Neighbor::init cutoffs (neighbor.cpp:278-310) + the cell grid (nbin_standard.cpp:53-186)
static int init_neighbor(sphbvf_ctx *ctx) {
  Coeffs &co = ctx->co;
  const double skin = ctx->cfg.skin;
  ctx->triggersq = 0.25 * skin * skin;
  ctx->cutneighmax = 0.0;
  for (int i = 1; i <= co.ntypes; i++)
    for (int j = 1; j <= co.ntypes; j++) {
      if (!ctx->pairset[i][j]) return ctx->fail(SPHBVF_EINVAL, "Not all pair ssa_tsdpd/bvf coeffs are set");
      const double cutoff = sqrt(co.cutsq[i][j]);
      const double delta = cutoff > 0.0 ? skin : 0.0;
      const double cut = cutoff + delta;
      co.cutneighsq[i][j] = cut * cut;
      ctx->cutneighmax = std::max(ctx->cutneighmax, cut);
    }
  if (ctx->cutneighmax <= 0.0) return ctx->fail(SPHBVF_EINVAL, "all pair cutoffs are zero");
  Grid &g = ctx->grid;
  const Box &b = ctx->box;
  const double binsize = 0.5 * ctx->cutneighmax;
  g.dim = b.dim;
  for (int k = 0; k < 3; k++) {
    if (b.dim == 2 && k == 2) {
      g.lo[k] = b.sublo[k];
      g.n[k] = 1;
      g.inv[k] = 1.0 / std::max(b.subhi[k] - b.sublo[k], 1e-300);
      g.s[k] = 0;
      g.glo[k] = 0;
      g.ghi[k] = 0;
      continue;
    }
    const double lo = b.sublo[k] - ctx->cutneighmax, hi = b.subhi[k] + ctx->cutneighmax;
    int n = (int)((hi - lo) / binsize);
    if (n < 1) n = 1;
    const double size = (hi - lo) / n;
    g.lo[k] = lo;
    g.n[k] = n;
    g.inv[k] = 1.0 / size;
    int s = (int)(ctx->cutneighmax * g.inv[k]);
    if (s * size < ctx->cutneighmax) s++;
    g.s[k] = s;
    if (s > 2) return ctx->fail(SPHBVF_EINVAL, "internal: stencil half-width %d > 2 (cells are cutneigh/2 or larger)", s);
    // cells strictly inside the brick cannot hold ghosts (one cell of safety margin per side)
    g.glo[k] = (int)floor((b.sublo[k] - lo) * g.inv[k]) + 2;
    g.ghi[k] = (int)floor((b.subhi[k] - lo) * g.inv[k]) - 2;
  }
  // tile-major cell numbering: 4x4x4 (3D) or 8x8x1 (2D) cells per tile
  g.tb[0] = b.dim == 3 ? 2 : 3;
  g.tb[1] = b.dim == 3 ? 2 : 3;
  g.tb[2] = b.dim == 3 ? 2 : 0;
  g.ncells = 1L << (g.tb[0] + g.tb[1] + g.tb[2]);
  for (int k = 0; k < 3; k++) {
    g.nt[k] = (g.n[k] + (1 << g.tb[k]) - 1) >> g.tb[k];
    g.ncells *= g.nt[k];
  }
  if (g.ncells > 2000000000L) return ctx->fail(SPHBVF_EINVAL, "Too many neighbor bins");
  return ensure_cells(ctx, g.ncells);
}

static int fetch_flags(sphbvf_ctx *ctx) {
  CK(cudaMemcpyAsync(ctx->h_flags, ctx->w.flags, sizeof(int) * 8, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return 0;
}

// ---- pieces of the rebuild branch of verlet.cpp:268-296, shared by the single-rank path below and
// the brick-decomposed path in comm_nccl.cu -----------------------------------------------------

// every primary array gathered through w.perm (new -> old) into the second buffer, then the buffers swap
int permute_state(sphbvf_ctx *ctx, int n, bool always_dev) {
  DevState &d = ctx->d;
  StateArrays cur = {d.tag, d.type, d.mask, d.solid, d.fixed, d.slot, d.x, d.v, d.vest, d.rho, d.rhoI, d.e, d.C, d.dev};
  launch_gather_state(cur, d.alt, ctx->w.perm, n, ctx->co.nspecies, ctx->with_dev || always_dev, ctx->st);
  CKLAUNCH();
  const StateArrays nw = d.alt;
  d.alt = cur;
  d.tag = nw.tag; d.type = nw.type; d.mask = nw.mask; d.solid = nw.solid; d.fixed = nw.fixed; d.slot = nw.slot;
  d.x = nw.x; d.v = nw.v; d.vest = nw.vest; d.rho = nw.rho; d.rhoI = nw.rhoI; d.e = nw.e; d.C = nw.C; d.dev = nw.dev;
  if (!(ctx->with_dev || always_dev)) std::swap(d.dev, d.alt.dev);   // not gathered (identically zero): keep the buffer
  if (!ctx->co.nspecies) std::swap(d.C, d.alt.C);
  return 0;
}

// Domain::pbc + cell ids + counting sort + permutation of every primary array into cell order
int rebuild_sort(sphbvf_ctx *ctx) {
  DevState &d = ctx->d;
  NeighWork &w = ctx->w;
  cudaStream_t st = ctx->st;
  launch_cell_ids(d, ctx->grid, ctx->box, w, st);
  launch_sort_owned(d, ctx->grid, w, st);
  CKLAUNCH();
  return permute_state(ctx, d.nlocal, false);
}

// ghosts are in place (packed records of owned + ghost atoms written): bin ghosts, Verlet list, xhold
int rebuild_finish(sphbvf_ctx *ctx) {
  DevState &d = ctx->d;
  NeighWork &w = ctx->w;
  cudaStream_t st = ctx->st;
  const Coeffs &co = ctx->co;
  int rc;
  launch_bin_ghosts(d, ctx->grid, w, st);
  CKLAUNCH();
  // Encoding of this rebuild's list.  The gather form is the default (faster: DESIGN.md section 3); with
  // SPHBVF_PAIR=tile the tile form (16-bit slots, candidates of a tile staged in shared memory by pair_tile_kernel) is
  // tried first: the builder reports the largest candidate count of a tile, and if that does not fit 12-bit slots or
  // the shared memory of one CTA (96-byte records + index map) the list is built again in the gather form.
  // multi-rank: the order in which the overlapped pair pass takes the atoms (interior of the brick first); the atom
  // count of the interior comes back with the flag fetch of the list build below
  ctx->aorder_valid = 0;
  if (ctx->cfg.nranks > 1 && ctx->overlap_halo && tile_form_possible(ctx->grid)) {
    if ((rc = make_tile_order(ctx))) return rc;
    if (d.nmax > ctx->aorder_cap) {
      if (ctx->aorder) cudaFree(ctx->aorder);
      ctx->aorder = nullptr;
      CK(cudaMalloc((void **)&ctx->aorder, sizeof(int) * (size_t)d.nmax));
      ctx->aorder_cap = d.nmax;
    }
    launch_atom_order(ctx->grid, w, ctx->tile_order, ctx->ntiles_total, ctx->ntiles_interior, ctx->tile_cnt, ctx->tile_off,
                      ctx->aorder, st);
    CKLAUNCH();
    ctx->aorder_valid = 1;
  }
  const char *lb = getenv("SPHBVF_LIST_BUILD");
  bool want_tile = ctx->pair_pref == 1 && !(lb && lb[0] == 't' && lb[1] == 'h') && tile_form_possible(ctx->grid);
  for (int attempt = 0; attempt < 3; attempt++) {
    if (d.maxneigh == 0) {
      // first guess from the number density: neighbours within cutneighmax of a uniform fluid
      const Box &b = ctx->box;
      double vol = 1.0;
      for (int k = 0; k < b.dim; k++) vol *= (b.subhi[k] - b.sublo[k]);
      const double dens = d.nlocal / std::max(vol, 1e-300);
      const double sph = b.dim == 3 ? 4.18879 * pow(ctx->cutneighmax, 3) : 3.14159 * pow(ctx->cutneighmax, 2);
      d.maxneigh = std::max(16, (int)(1.3 * dens * sph) + 8);
    }
    d.list16 = want_tile;
    if ((rc = ensure_neigh(ctx, d.maxneigh))) return rc;
    CK(cudaMemsetAsync(w.flags + 2, 0, sizeof(int), st));
    CK(cudaMemsetAsync(w.flags + 4, 0, sizeof(int), st));
    launch_build_list(d, ctx->grid, co, w, st);
    CKLAUNCH();
    if ((rc = fetch_flags(ctx))) return rc;
    if (d.list16) {
      const int cap = (ctx->h_flags[4] + 7) & ~7;
      if (ctx->h_flags[4] > TILE_MAX_SLOTS || pair_tile_smem(cap, true) + 6144 > (size_t)ctx->smem_optin) {
        want_tile = false;   // dense tiles: gather form for this rebuild interval
        attempt--;
        continue;
      }
      d.tile_cap = std::max(cap, 8);
      if (getenv("SPHBVF_VERBOSE") && ctx->nbuilds < 2)
        fprintf(stderr, "sphbvf[%d]: tile form, %d candidates in the fullest tile, %zu B of shared memory per CTA\n",
                ctx->cfg.rank, ctx->h_flags[4], pair_tile_smem(d.tile_cap, false));
    }
    const int capacity = d.list16 ? d.pitch16 : d.maxneigh;
    if (ctx->h_flags[2] <= capacity) break;
    if (attempt == 2) return ctx->fail(SPHBVF_EOVERFLOW, "Neighbor list overflow");
    if ((rc = ensure_neigh(ctx, ctx->h_flags[2] + ctx->h_flags[2] / 8 + 4))) return rc;
  }
  ctx->maxneigh_seen = ctx->h_flags[2];
  ctx->expanded_valid = 0;
  ctx->natoms_interior = ctx->aorder_valid ? ctx->h_flags[5] : 0;
  launch_copy_xhold(d, st);
  ctx->ago = 0;
  ctx->nbuilds++;
  return 0;
}

// Tiles in the order the overlapped pair pass wants them: first the tiles whose halo box holds no cell that can
// contain a ghost (they can run while the per-step halo is still on the wire), then the ones along the brick faces
// that have a neighbour.  A function of the cell grid and the halo plan only, so it is rebuilt when those change.
int make_tile_order(sphbvf_ctx *ctx) {
  const Grid &g = ctx->grid;
  const long ntiles = (long)g.nt[0] * g.nt[1] * g.nt[2];
  static_assert(sizeof(int) == 4, "tile ids are 32-bit");
  int peer[3][2];
  long peermask = 0;
  for (int k = 0; k < 3; k++)
    for (int sd = 0; sd < 2; sd++) {
      peer[k][sd] = comm_face_has_peer(ctx, k, sd);
      peermask = 2 * peermask + peer[k][sd];
    }
  long key[11] = {g.nt[0], g.nt[1], g.nt[2], g.glo[0], g.glo[1], g.glo[2], g.ghi[0], g.ghi[1], g.ghi[2],
                  g.n[0] + 4096L * (g.n[1] + 4096L * g.n[2]), peermask};
  if (ctx->tile_order && ntiles == ctx->ntiles_total && memcmp(key, ctx->tile_key, sizeof key) == 0) return 0;
  std::vector<int> inner, outer;
  inner.reserve(ntiles);
  for (long t = 0; t < ntiles; t++) {
    const int tc[3] = {(int)(t % g.nt[0]), (int)((t / g.nt[0]) % g.nt[1]), (int)(t / ((long)g.nt[0] * g.nt[1]))};
    bool ghost = false;
    for (int k = 0; k < g.dim; k++) {
      const int c0 = tc[k] << g.tb[k];
      // The early halo also relies on "every atom a neighbour brick receives (x within cutghost of the face,
      // border_kernel) lives in a face tile": such an atom sits in a cell c <= floor(2 cutghost / cellsize) <= 4 (= 4 only
      // when cellsize is exactly cutghost / 2, and then glo = 4), its tile starts at c0 <= c, and c0 - s <= 2 < glo
      // (glo carries two cells of margin, init_neighbor); symmetrically on the high side.
      const int h0 = std::max(c0 - g.s[k], 0), h1 = std::min(c0 + (1 << g.tb[k]) - 1 + g.s[k], g.n[k] - 1);
      // cells below glo / above ghi may hold ghosts, if a neighbour brick (or this brick's own periodic image) sends any
      if ((h0 < g.glo[k] && peer[k][0]) || (h1 > g.ghi[k] && peer[k][1])) ghost = true;
    }
    (ghost ? outer : inner).push_back((int)t);
  }
  if (ntiles > ctx->tile_order_cap) {
    // the three tables sized by the tile count (the atom order and the schedule's queues have their own owners)
    for (int **q : {&ctx->tile_order, &ctx->tile_cnt, &ctx->tile_off}) {
      if (*q) cudaFree(*q);
      *q = nullptr;
    }
    CK(cudaMalloc((void **)&ctx->tile_order, sizeof(int) * (size_t)ntiles));
    CK(cudaMalloc((void **)&ctx->tile_cnt, sizeof(int) * (size_t)(ntiles + 2)));
    CK(cudaMalloc((void **)&ctx->tile_off, sizeof(int) * (size_t)(ntiles + 2)));
    ctx->tile_order_cap = ntiles;
  }
  // (A Z-order sequence of the tiles behind the persistent pair schedule -- every SM then owns a compact blob of tiles
  // instead of a slab of rows -- was measured at 8 M atoms: L2 hit rate 26 -> 36 %, DRAM reads 4.98 -> 4.85 GB per pass,
  // L1 hit rate unchanged at 87 %, and the pass 0.04 ms SLOWER because of the atom-order indirection it needs
  // (gpurun_out/r2q_*).  Rows stay.)
  ctx->ntiles_interior = (int)inner.size();
  ctx->ntiles_total = (int)ntiles;
  inner.insert(inner.end(), outer.begin(), outer.end());
  // the pair pass that reads the list is ordered after this copy on the same stream; the vector dies at return
  CK(cudaMemcpyAsync(ctx->tile_order, inner.data(), sizeof(int) * (size_t)ntiles, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  memcpy(ctx->tile_key, key, sizeof key);
  return 0;
}

int ctx_ensure_capacity(sphbvf_ctx *ctx, int nmax, int nallmax) { return ensure_capacity(ctx, nmax, nallmax); }
int ctx_fetch_flags(sphbvf_ctx *ctx) { return fetch_flags(ctx); }

// single rank: ghosts are periodic self images
static int rebuild(sphbvf_ctx *ctx) {
  DevState &d = ctx->d;
  NeighWork &w = ctx->w;
  cudaStream_t st = ctx->st;
  const Coeffs &co = ctx->co;
  int rc;
  ctx->tic(K_NEIGH);
  CK(cudaMemsetAsync(w.flags, 0, sizeof(int) * 8, st));
  if ((rc = rebuild_sort(ctx))) return rc;
  const int n = d.nlocal;
  launch_count_images(d, ctx->box, ctx->cutneighmax, w, st);
  int nghost = 0;
  CK(cudaMemcpyAsync(&ctx->h_flags[8], w.nimg + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  if ((rc = fetch_flags(ctx))) return rc;
  nghost = ctx->h_flags[8];
  if (ctx->h_flags[0]) return ctx->fail(SPHBVF_ENONFINITE, "Non-numeric positions - simulation unstable");
  if (ctx->h_flags[1]) return ctx->fail(SPHBVF_ELOST, "Lost atoms: an owned atom left the non-periodic box");
  if (d.nlocal + nghost > d.nallmax)
    if ((rc = ensure_capacity(ctx, d.nmax, d.nlocal + nghost + nghost / 4 + 1024))) return rc;
  d.nghost = nghost;
  if (nghost) launch_fill_images(d, ctx->box, ctx->cutneighmax, w, st);   // 27 shift tests per atom: skip without images
  CK(cudaMemcpyAsync(d.ptag, d.tag, sizeof(int) * n, cudaMemcpyDeviceToDevice, st));
  launch_pack(d, co, ctx->with_dev, st);
  launch_ghost_refresh(d, co, ctx->with_dev, st);
  CKLAUNCH();
  if ((rc = rebuild_finish(ctx))) return rc;
  ctx->toc();
  return 0;
}

static PairFlags pair_flags(const sphbvf_ctx *ctx) {
  PairFlags pf;
  const int var = ctx->co.variant;
  pf.filter_step = shepard_filter_step(var, ctx->ntimestep);
  pf.with_dev = ctx->with_dev;
  pf.any_solid = ctx->any_solid;
  // density diffusion of the fsi pair style: amplDamp = 0.1 while ntimestep*dt <= dt*nsteps
  // (pair_ssa_tsdpd_bvf_fsi.cpp:531-539)
  const double tnow = ctx->ntimestep * ctx->cfg.dt, tmax = ctx->cfg.dt * ctx->run_nsteps;
  pf.damp = (var == SPHBVF_FSI && tnow <= tmax) ? 0.1 : 0.0;
  pf.random = ctx->e_nonzero && ctx->random_set;
  pf.rand_pref = 4.0 * ctx->kboltz / ctx->cfg.dt;
  pf.seed = ctx->seed;
  pf.ntimestep = ctx->ntimestep;
  return pf;
}

// any_solid / with_dev / e_nonzero select the pair-kernel instantiation and the halo record width.  They are derived
// from the DEVICE state (a reduction over the owned atoms), so that fields changed through sphbvf_upload after
// sphbvf_set_atoms count: set_atoms(e = NULL) followed by upload(E != 0) must switch the stochastic term on.
// the flags must agree on every brick; with_dev is a type mask, so its bits are reduced one by one (max = or)
static int allreduce_kernel_flags(sphbvf_ctx *ctx) {
  // comm_allreduce_max carries at most 8 ints: two flags + one bit per atom type (types 1 .. MAXT-1, MAXT = 5)
  static_assert(MAXT <= 6, "with_dev bits must fit the 8-int all-reduce");
  int v[2 + 6] = {ctx->any_solid, ctx->e_nonzero};
  for (int t = 0; t < 6; t++) v[2 + t] = (ctx->with_dev >> t) & 1;
  const int rc = comm_allreduce_max(ctx, v, 8);
  if (rc) return rc;
  ctx->any_solid = v[0]; ctx->e_nonzero = v[1];
  ctx->with_dev = 0;
  for (int t = 0; t < 6; t++) ctx->with_dev |= v[2 + t] << t;
  return 0;
}

static int derive_flags(sphbvf_ctx *ctx) {
  int *out = ctx->w.flags + 5;   // flags[5..7]: scratch between rebuilds
  CK(cudaMemsetAsync(out, 0, sizeof(int) * 3, ctx->st));
  launch_derive_flags(ctx->d, ctx->co, out, ctx->st);
  CKLAUNCH();
  int rc;
  if ((rc = fetch_flags(ctx))) return rc;
  ctx->any_solid = ctx->h_flags[5];
  ctx->with_dev = ctx->h_flags[6];
  ctx->e_nonzero = ctx->h_flags[7];
  ctx->flags_dirty = 0;
  return 0;
}

template <typename T>
__global__ void scatter_rows_kernel(const T *in, T *out, const int *slot, int n, int ncols, int to_slot) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long)n * ncols) return;
  const int row = (int)(q / ncols), col = (int)(q - (long)row * ncols);
  const long s = (long)slot[row] * ncols + col;
  if (to_slot) out[s] = in[q];   // device order -> host slot order
  else out[q] = in[s];           // host slot order -> device order
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int sphbvf_version(void) { return SPHBVF_VERSION; }

int sphbvf_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int sphbvf_create(const sphbvf_config *cfg, sphbvf_ctx **out) {
  if (!cfg || !out) return SPHBVF_EINVAL;
  *out = nullptr;
  if (cfg->ntypes < 1 || cfg->ntypes + 1 > MAXT || cfg->nspecies < 0 || cfg->nspecies > MAXS ||
      (cfg->dim != 2 && cfg->dim != 3) || cfg->variant < 0 || cfg->variant > 2)
    return SPHBVF_EINVAL;
  if (sphbvf_device_count() <= cfg->device) return SPHBVF_ECUDA;   // no GPU: fail loudly, no fallback
  if (cudaSetDevice(cfg->device) != cudaSuccess) return SPHBVF_ECUDA;
  sphbvf_ctx *ctx = new sphbvf_ctx();
  ctx->cfg = *cfg;
  if (ctx->cfg.nranks < 1) { ctx->cfg.nranks = 1; ctx->cfg.rank = 0; }
  for (int k = 0; k < 3; k++) if (ctx->cfg.procgrid[k] < 1) ctx->cfg.procgrid[k] = 1;
  if (ctx->cfg.neigh_every < 1) ctx->cfg.neigh_every = 1;
  memset(&ctx->co, 0, sizeof ctx->co);
  ctx->co.dim = cfg->dim;
  ctx->co.variant = cfg->variant;
  ctx->co.nspecies = cfg->nspecies;
  ctx->co.ntypes = cfg->ntypes;
  Box &b = ctx->box;
  b.dim = cfg->dim;
  for (int k = 0; k < 3; k++) {
    b.lo[k] = cfg->boxlo[k];
    b.hi[k] = cfg->boxhi[k];
    b.prd[k] = cfg->boxhi[k] - cfg->boxlo[k];
    b.periodic[k] = cfg->periodic[k];
  }
  sphbvf_brick_bounds(&ctx->cfg, ctx->cfg.rank, b.sublo, b.subhi);
  if (cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost((void **)&ctx->h_flags, sizeof(int) * 16) != cudaSuccess ||
      cudaMalloc((void **)&ctx->w.flags, sizeof(int) * 8) != cudaSuccess) {
    delete ctx;
    return SPHBVF_ECUDA;
  }
  cudaMemset(ctx->w.flags, 0, sizeof(int) * 8);
  ctx->run_nsteps_user = -1;
  { const char *e = getenv("SPHBVF_NO_FUSE"); ctx->fuse = !(e && atoi(e)); }
  { const char *e = getenv("SPHBVF_PAIR"); ctx->pair_pref = (e && e[0] == 't') ? 1 : 0; }   // tile | gather (default)
  // SPHBVF_HALO: serial (on the compute stream) | overlap (own stream, beside the interior of the pair pass) |
  // early (default: own stream, started by the fused integrator right after it has integrated the face atoms, so the
  // exchange hides behind the integration of the interior and the pair pass stays ONE launch)
  { const char *e = getenv("SPHBVF_HALO"); ctx->overlap_halo = !(e && e[0] == 's'); ctx->halo_early = !e || e[0] == 'e'; }
  {   // gather form, launches of more than 4 x SMs chunks: SPHBVF_PAIR_SCHED = warp (default: persistent CTAs, every warp
      // draws 32-atom chunks from the queue of its SM: 5.42 vs 5.61 ms at 8 M atoms) | smid (the same queues, one 192-atom
      // chunk per CTA and draw: 5.82 ms) | grid (one CTA per chunk, the hardware's dispatch order: 5.61 ms)
    const char *e = getenv("SPHBVF_PAIR_SCHED");
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, cfg->device);
    ctx->pair_nq = nsm > 0 ? nsm : 1;
    const bool grid = e && e[0] == 'g';
    if (!grid && cudaMalloc((void **)&ctx->pair_queues, sizeof(int) * 2 * (ctx->pair_nq + 1)) != cudaSuccess)
      ctx->pair_queues = nullptr;
    ctx->pair_warp = !(e && e[0] == 's');
  }
  if (cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device) != cudaSuccess)
    ctx->smem_optin = 48 * 1024;
  *out = ctx;
  return 0;
}

void sphbvf_destroy(sphbvf_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  comm_halo_join(ctx);
  cudaStreamSynchronize(ctx->st);
  ctx->drain_events();
  for (auto e : ctx->ev_pool) cudaEventDestroy(e);
  comm_destroy(ctx);
  for (int *q : {ctx->tile_order, ctx->tile_cnt, ctx->tile_off, ctx->aorder, ctx->pair_queues}) if (q) cudaFree(q);
  DevState &d = ctx->d;
  void *ptrs[] = {d.tag, d.type, d.mask, d.solid, d.fixed, d.slot, d.x, d.v, d.vest, d.rho, d.rhoI, d.e, d.C, d.dev,
                  d.f, d.nw, d.ddv, d.ddx, d.drho, d.phi, d.nd, d.rhoAux1, d.rhoAux2, d.Pnew, d.ddev, d.Q,
                  d.alt.tag, d.alt.type, d.alt.mask, d.alt.solid, d.alt.fixed, d.alt.slot, d.alt.x, d.alt.v, d.alt.vest,
                  d.alt.rho, d.alt.rhoI, d.alt.e, d.alt.C, d.alt.dev, d.neigh16,
                  d.prec, d.pD, d.pCs, d.pdev, d.pflags, d.ptag, d.xhold, d.gowner, d.gshift, d.neigh,
                  d.numneigh, ctx->w.cellid, ctx->w.perm, ctx->w.cell_count, ctx->w.cell_start, ctx->w.gcell_count,
                  ctx->w.gcell_start, ctx->w.gorder, ctx->w.scan_tmp, ctx->w.nimg, ctx->w.flags, ctx->w.tmp_perm};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->d_virial) cudaFree(ctx->d_virial);
  cudaStreamDestroy(ctx->st);
  delete ctx;
}

const char *sphbvf_last_error(const sphbvf_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int sphbvf_set_type(sphbvf_ctx *ctx, int t, double mass, double rho0, double c0, double G0) {
  if (t < 1 || t > ctx->co.ntypes) return ctx->fail(SPHBVF_EINVAL, "type %d out of range", t);
  Coeffs &co = ctx->co;
  co.mass[t] = mass;
  co.rho0[t] = rho0;
  co.c0[t] = c0;
  co.B[t] = c0 * c0 * rho0 / 7.0;   // pair_...transport_velocity.cpp:981
  co.G0[t] = G0;
  return 0;
}

int sphbvf_set_pair(sphbvf_ctx *ctx, int i, int j, double eta, double h, double cutc, const double *kappa) {
  Coeffs &co = ctx->co;
  if (i < 1 || j < 1 || i > co.ntypes || j > co.ntypes) return ctx->fail(SPHBVF_EINVAL, "type pair %d %d out of range", i, j);
  const int a[2] = {i, j}, b[2] = {j, i};
  for (int s = 0; s < 2; s++) {   // init_one mirrors (pair_...:1040-1050); cutsq = cut*cut (pair.cpp:245)
    co.eta[a[s]][b[s]] = eta;
    co.cut[a[s]][b[s]] = h;
    co.cutsq[a[s]][b[s]] = h * h;
    co.cutc[a[s]][b[s]] = cutc;
    for (int k = 0; k < co.nspecies; k++) co.kappa[a[s]][b[s]][k] = kappa ? kappa[k] : 0.0;
    ctx->pairset[a[s]][b[s]] = 1;
  }
  return 0;
}

int sphbvf_set_dt(sphbvf_ctx *ctx, double dt) { ctx->cfg.dt = dt; return 0; }
int sphbvf_set_random(sphbvf_ctx *ctx, double kboltz, unsigned long long seed) {
  if (!(kboltz >= 0.0)) return ctx->fail(SPHBVF_EINVAL, "set_random: kboltz must be >= 0");
  ctx->kboltz = kboltz;
  ctx->seed = seed;
  ctx->random_set = 1;
  return 0;
}
int sphbvf_set_timestep(sphbvf_ctx *ctx, long n) { ctx->ntimestep = n; return 0; }
int sphbvf_set_run_length(sphbvf_ctx *ctx, long n) { ctx->run_nsteps_user = n; ctx->run_nsteps = n; return 0; }

int sphbvf_set_atoms(sphbvf_ctx *ctx, int n, const int *tag, const int *type, const int *mask, const int *solid,
                     const int *fixed, const double *x, const double *v, const double *rho, const double *e,
                     const double *C, const double *dev) {
  if (n < 0 || !tag || !type || !solid || !fixed || !x || !rho) return ctx->fail(SPHBVF_EINVAL, "set_atoms: null array");
  if (n >= NEIGH_JMASK / 2) return ctx->fail(SPHBVF_EINVAL, "too many atoms for one GPU");
  cudaSetDevice(ctx->cfg.device);
  int rc;
  if ((rc = ensure_capacity(ctx, n + n / 8 + 1024, n + n / 4 + 4096))) return rc;
  DevState &d = ctx->d;
  cudaStream_t st = ctx->st;
  const int S = ctx->co.nspecies;
  d.nlocal = n;
  d.nghost = 0;
  for (int i = 0; i < n; i++)
    if (type[i] < 1 || type[i] > ctx->co.ntypes) return ctx->fail(SPHBVF_EINVAL, "atom %d has type %d out of range", i, type[i]);
  ctx->any_solid = 0;
  int has_dev = 0;
  for (int i = 0; i < n; i++) {
    if (solid[i]) {
      ctx->any_solid = 1;
      if (ctx->co.G0[type[i]] != 0.0) has_dev |= 1 << type[i];
    }
  }
  if (dev)
    for (int i = 0; i < n; i++) {
      if ((has_dev >> type[i]) & 1) continue;
      for (int q = 0; q < 9; q++) if (dev[9 * (size_t)i + q] != 0.0) { has_dev |= 1 << type[i]; break; }
    }
  ctx->with_dev = has_dev;   // bit t: atoms of type t may carry a deviatoric stress (non-zero = elastic solids present)
#define UP(dst, src, cnt, T)                                                                         \
  do {                                                                                               \
    if (src) CK(cudaMemcpyAsync(dst, src, sizeof(T) * (size_t)(cnt), cudaMemcpyHostToDevice, st));   \
    else CK(cudaMemsetAsync(dst, 0, sizeof(T) * (size_t)(cnt), st));                                 \
  } while (0)
  UP(d.tag, tag, n, int); UP(d.type, type, n, int); UP(d.solid, solid, n, int); UP(d.fixed, fixed, n, int);
  UP(d.x, x, 3 * (size_t)n, double); UP(d.v, v, 3 * (size_t)n, double); UP(d.rho, rho, n, double); UP(d.e, e, n, double);
  if (S) UP(d.C, C, (size_t)S * n, double);
  UP(d.dev, dev, 9 * (size_t)n, double);
#undef UP
  std::vector<int> iota(n), ones;
  for (int i = 0; i < n; i++) iota[i] = i;
  CK(cudaMemcpyAsync(d.slot, iota.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  if (mask) CK(cudaMemcpyAsync(d.mask, mask, sizeof(int) * n, cudaMemcpyHostToDevice, st));
  else {
    ones.assign(n, 1);
    CK(cudaMemcpyAsync(d.mask, ones.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  }
  // create_atom defaults (atom_vec_ssa_tsdpd_atomic.cpp:1873-1936): vest = 0, rhoI = 0, outputs 0
  CK(cudaMemsetAsync(d.vest, 0, sizeof(double) * 3 * n, st));
  CK(cudaMemsetAsync(d.rhoI, 0, sizeof(double) * n, st));
  for (double *p : {d.f, d.nw, d.ddv, d.ddx}) CK(cudaMemsetAsync(p, 0, sizeof(double) * 3 * n, st));
  for (double *p : {d.drho, d.phi, d.nd, d.rhoAux1, d.rhoAux2, d.Pnew}) CK(cudaMemsetAsync(p, 0, sizeof(double) * n, st));
  CK(cudaMemsetAsync(d.ddev, 0, sizeof(double) * 9 * n, st));
  CK(cudaMemsetAsync(d.Q, 0, sizeof(double) * (S ? S : 1) * n, st));
  CK(cudaStreamSynchronize(st));
  ctx->e_nonzero = 0;
  if (e) for (int i = 0; i < n; i++) if (e[i] != 0.0) { ctx->e_nonzero = 1; break; }
  ctx->atoms_set = 1;
  ctx->setup_done = 0;
  ctx->migrated = 0;
  ctx->final_pending = 0;
  ctx->pack_valid = 0;
  return 0;
}

static int add_fix(sphbvf_ctx *ctx, const FixDesc &f) {
  if (ctx->nfix == MAXFIX) return ctx->fail(SPHBVF_EINVAL, "too many fixes");
  ctx->fixes[ctx->nfix++] = f;
  return 0;
}
int sphbvf_add_buoyancy(sphbvf_ctx *ctx, int groupbit, int gravity, double accel, int coord, int k, double Cref) {
  if (coord < 0 || coord > 2 || (!gravity && (k < 0 || k >= ctx->co.nspecies)))
    return ctx->fail(SPHBVF_EINVAL, "Illegal fix ssa_tsdpd/buoyancy command");
  FixDesc f = {FIX_BUOYANCY, groupbit, {gravity, coord, k, 0}, 0, {accel, Cref, 0, 0, 0, 0}};
  return add_fix(ctx, f);
}
int sphbvf_add_forcing(sphbvf_ctx *ctx, int groupbit, int kind, long step, int idx, int shape, double cx, double cy,
                       double a, double b, double value) {
  if ((kind == 0 && (idx < 0 || idx >= ctx->co.nspecies)) || (kind == 1 && (idx < 0 || idx > 2)) || kind < 0 || kind > 1)
    return ctx->fail(SPHBVF_EINVAL, "Illegal fix ssa_tsdpd_forcing command");
  FixDesc f = {FIX_FORCING, groupbit, {kind, idx, shape, 0}, step, {cx, cy, a, b, value, 0}};
  return add_fix(ctx, f);
}
int sphbvf_add_buffer(sphbvf_ctx *ctx, int groupbit, int kind, int axis, long step, int idx, double cx, double cy,
                      double length, double width, double value) {
  if ((kind == 0 && (idx < 0 || idx >= ctx->co.nspecies)) || (kind == 1 && (idx < 0 || idx > 2)) || kind < 0 || kind > 2)
    return ctx->fail(SPHBVF_EINVAL, "Illegal fix ssa_tsdpd_buffer command");
  FixDesc f = {FIX_BUFFER, groupbit, {kind, idx, axis, 0}, step, {cx, cy, length, width, value, 0}};
  return add_fix(ctx, f);
}
int sphbvf_add_chem_rxn(sphbvf_ctx *ctx, int groupbit, double k_rate, int nreact, const int *reactants, int nprod,
                        const int *products) {
  const int S = ctx->co.nspecies;
  if (nreact < 0 || nreact > 2 || nprod < 0 || nprod > 4 || (nreact && !reactants) || (nprod && !products))
    return ctx->fail(SPHBVF_EINVAL, "Illegal fix ssa_tsdpd_chem_rxn_mass_action command");
  int r = 0, p = 0;
  for (int j = 0; j < nreact; j++) {
    if (reactants[j] < 0 || reactants[j] >= S) return ctx->fail(SPHBVF_EINVAL, "chem_rxn: reactant species out of range");
    r |= reactants[j] << (8 * j);
  }
  for (int j = 0; j < nprod; j++) {
    if (products[j] < 0 || products[j] >= S) return ctx->fail(SPHBVF_EINVAL, "chem_rxn: product species out of range");
    p |= products[j] << (8 * j);
  }
  FixDesc f = {FIX_CHEMRXN, groupbit, {nreact | (nprod << 8), r, p, 0}, 0, {k_rate, 0, 0, 0, 0, 0}};
  return add_fix(ctx, f);
}
int sphbvf_add_setforce(sphbvf_ctx *ctx, int groupbit, double fx, double fy, double fz) {
  FixDesc f = {FIX_SETFORCE, groupbit, {0, 0, 0, 0}, 0, {fx, fy, fz, 0, 0, 0}};
  return add_fix(ctx, f);
}

// in_setup: Modify::setup calls Fix::setup, which only FixSetForce and FixSsaTsdpdBuoyancy forward to
// post_force (fix_setforce.cpp, fix_ssa_tsdpd_buoyancy.cpp:93-96); the reaction fix has no setup()
static int run_fixes(sphbvf_ctx *ctx, int hook, bool in_setup = false) {
  bool any = false;
  for (int q = 0; q < ctx->nfix; q++) {
    if (in_setup && ctx->fixes[q].kind == FIX_CHEMRXN) continue;
    if (!fix_runs(ctx->fixes[q], hook, ctx->ntimestep)) continue;
    if (!any) {
      FLUSH();
      ctx->tic(K_FIX);
      any = true;
      if (hook == 0) ctx->pack_valid = 0;   // post_integrate fixes edit vest / C after the records were written
    }
    launch_fix(ctx->d, ctx->co, ctx->fixes[q], hook, ctx->ntimestep, ctx->st);
  }
  if (any) { ctx->toc(); CKLAUNCH(); }
  return 0;
}

int sphbvf_build_neighbors(sphbvf_ctx *ctx) {
  if (!ctx->atoms_set) return ctx->fail(SPHBVF_ESTATE, "build_neighbors before set_atoms");
  cudaSetDevice(ctx->cfg.device);
  int rc;
  FLUSH();
  ctx->pack_valid = 0;
  if ((rc = init_neighbor(ctx))) return rc;
  if (ctx->cfg.nranks > 1) return comm_rebuild(ctx);
  return rebuild(ctx);
}

int sphbvf_setup(sphbvf_ctx *ctx) {
  if (!ctx->atoms_set) return ctx->fail(SPHBVF_ESTATE, "setup before set_atoms");
  cudaSetDevice(ctx->cfg.device);
  int rc;
  FLUSH();
  // Deliberate deviation: Verlet::setup creates ghosts BEFORE setup_pre_force sets
  // vest = v, rhoI = rho (verlet.cpp:118-132), so the reference's ghosts hold stale vest/rhoI at
  // step 0 -- and with its half list + Newton mirror the step-0 result is then not even
  // gather-consistent.  Here setup_pre_force runs first, so ghosts always mirror their owners.
  // Identical whenever the initial momentum velocity of atoms near a periodic face is zero (all
  // shipped decks) or the run was preceded by a `run 0`.
  launch_setup_pre_force(ctx->d, ctx->cfg.integrate_groupbit, ctx->st);
  if ((rc = derive_flags(ctx))) return rc;
  if (ctx->cfg.nranks > 1) {
    // kernel specialisation and halo record width must agree on every brick
    if ((rc = allreduce_kernel_flags(ctx))) return rc;
  }
  if ((rc = sphbvf_build_neighbors(ctx))) return rc;
  ctx->nbuilds = 0;   // neighbor->ncalls = 0 (verlet.cpp:128)
  ctx->ndanger = 0;
  ctx->setup_done = 1;
  if ((rc = sphbvf_pair_compute(ctx))) return rc;
  if ((rc = run_fixes(ctx, 1, true))) return rc;   // modify->setup(): FixSetForce::setup, FixSsaTsdpdBuoyancy::setup
  CK(cudaStreamSynchronize(ctx->st));
  return 0;
}

int sphbvf_setup_neighbors(sphbvf_ctx *ctx) {
  if (!ctx->atoms_set) return ctx->fail(SPHBVF_ESTATE, "setup before set_atoms");
  cudaSetDevice(ctx->cfg.device);
  int rc;
  FLUSH();
  if ((rc = derive_flags(ctx))) return rc;
  if (ctx->cfg.nranks > 1) {
    if ((rc = allreduce_kernel_flags(ctx))) return rc;
  }
  if ((rc = sphbvf_build_neighbors(ctx))) return rc;
  ctx->nbuilds = 0;
  ctx->ndanger = 0;
  ctx->setup_done = 1;
  return 0;
}

int sphbvf_initial_integrate(sphbvf_ctx *ctx) {
  if (ctx->final_pending) {
    // nothing touched the state since the deferred final_integrate: one pass does both halves and, unless a
    // post_integrate fix may still edit vest / C, writes the pair-input records as well
    int do_pack = 1;
    for (int q = 0; q < ctx->nfix; q++)
      if (ctx->fixes[q].kind == FIX_FORCING || (ctx->fixes[q].kind == FIX_BUFFER && ctx->fixes[q].ia[0] != 2)) do_pack = 0;
    ctx->final_pending = 0;
    ctx->tic(K_FUSED);
    ctx->halo_done_step = 0;
    // Early halo (multi-rank): the atoms of the face tiles -- a superset of everything a neighbour brick receives --
    // are integrated and packed first, the halo starts on its own stream behind them, and the interior is integrated
    // while the records travel.  If this step turns out to be a rebuild step the exchanged ghosts are simply replaced.
    // Every rank takes the same branch (halo_early_ok is agreed collectively at each rebuild), so the order of the
    // communicator's operations is the same everywhere.
    // (not while an upload of e / dev / type / solid_tag is waiting to be folded into the kernel flags: the width of the
    // halo record depends on them)
    const bool early = ctx->cfg.nranks > 1 && ctx->overlap_halo && ctx->halo_early && ctx->halo_early_ok && do_pack &&
                       ctx->aorder_valid && !ctx->flags_dirty;
    if (early) {
      int rc;
      launch_final_initial(ctx->d, ctx->co, ctx->pend_dt, ctx->pend_step, ctx->cfg.dt, ctx->ntimestep,
                           ctx->cfg.integrate_groupbit, do_pack, ctx->with_dev, ctx->st, ctx->aorder, ctx->natoms_interior,
                           ctx->d.nlocal);
      const PairFlags pf = pair_flags(ctx);
      if ((rc = comm_forward(ctx, pf.filter_step || pf.random))) return rc;
      launch_final_initial(ctx->d, ctx->co, ctx->pend_dt, ctx->pend_step, ctx->cfg.dt, ctx->ntimestep,
                           ctx->cfg.integrate_groupbit, do_pack, ctx->with_dev, ctx->st, ctx->aorder, 0, ctx->natoms_interior);
      ctx->halo_done_step = 1;
    } else
      launch_final_initial(ctx->d, ctx->co, ctx->pend_dt, ctx->pend_step, ctx->cfg.dt, ctx->ntimestep,
                           ctx->cfg.integrate_groupbit, do_pack, ctx->with_dev, ctx->st);
    ctx->toc();
    CKLAUNCH();
    ctx->pack_valid = do_pack;
    return 0;
  }
  ctx->halo_done_step = 0;
  ctx->pack_valid = 0;
  ctx->tic(K_INITIAL);
  launch_initial_integrate(ctx->d, ctx->co, ctx->cfg.dt, ctx->ntimestep, ctx->cfg.integrate_groupbit, ctx->with_dev, ctx->st);
  ctx->toc();
  CKLAUNCH();
  return 0;
}

int sphbvf_post_integrate(sphbvf_ctx *ctx) { return run_fixes(ctx, 0); }
int sphbvf_post_force(sphbvf_ctx *ctx) { return run_fixes(ctx, 1); }
int sphbvf_setup_post_force(sphbvf_ctx *ctx) { return run_fixes(ctx, 1, true); }
int sphbvf_end_of_step(sphbvf_ctx *ctx) { return run_fixes(ctx, 2); }

int sphbvf_neighbor(sphbvf_ctx *ctx, int *rebuilt) {
  if (!ctx->setup_done) return ctx->fail(SPHBVF_ESTATE, "neighbor before setup");
  int rc, flag = 0;
  FLUSH();
  const int pack_valid = ctx->pack_valid;
  ctx->pack_valid = 0;
  // e / dev / type / solid_tag were uploaded since the last step: the kernel specialisation follows the new state
  // (all ranks agree), and new types or solid tags sit in the packed list entries, so the list is rebuilt now
  // (multi-rank: such an upload is a collective act of the host code, every rank sees it in the same step)
  int force_rebuild = 0;
  if (ctx->flags_dirty) {
    force_rebuild = (ctx->flags_dirty & 2) != 0;
    if ((rc = derive_flags(ctx))) return rc;
    if (ctx->cfg.nranks > 1) {
      if ((rc = allreduce_kernel_flags(ctx))) return rc;
    }
  }
  // Neighbor::decide (neighbor.cpp:1922-1937)
  ctx->ago++;
  if (ctx->ago >= ctx->cfg.neigh_delay && ctx->ago % ctx->cfg.neigh_every == 0) {
    if (ctx->cfg.neigh_check == 0) flag = 1;
    else {
      CK(cudaMemsetAsync(ctx->w.flags + 3, 0, sizeof(int), ctx->st));
      launch_check_distance(ctx->d, ctx->triggersq, ctx->w.flags + 3, ctx->st);
      if ((rc = fetch_flags(ctx))) return rc;
      flag = ctx->h_flags[3];
      if (ctx->cfg.nranks > 1 && (rc = comm_vote(ctx, &flag))) return rc;   // MPI_Allreduce(MAX) neighbor.cpp:1997
      if (flag && ctx->ago == std::max(ctx->cfg.neigh_every, ctx->cfg.neigh_delay)) ctx->ndanger++;
    }
  }
  if (force_rebuild) flag = 1;
  if (rebuilt) *rebuilt = flag;
  if (flag) return ctx->cfg.nranks > 1 ? comm_rebuild(ctx) : rebuild(ctx);
  // Comm::forward_comm (comm_brick.cpp:460-520): refresh the packed records of owned atoms and ghosts
  if (pack_valid && ctx->cfg.nranks == 1 && !ctx->d.nghost) return 0;   // records are current, nothing to refresh
  if (ctx->halo_done_step && pack_valid) return 0;   // the integrator has packed and started this step's halo already
  ctx->tic(K_PACK);
  if (!pack_valid) launch_pack(ctx->d, ctx->co, ctx->with_dev, ctx->st);
  if (ctx->cfg.nranks > 1) {
    const PairFlags pf = pair_flags(ctx);   // pD (rhoI, e) travels only when the coming pair pass reads it
    if ((rc = comm_forward(ctx, pf.filter_step || pf.random))) return rc;
  }
  else launch_ghost_refresh(ctx->d, ctx->co, ctx->with_dev, ctx->st);
  ctx->toc();
  CKLAUNCH();
  return 0;
}

int sphbvf_pair_compute(sphbvf_ctx *ctx) {
  if (!ctx->setup_done) return ctx->fail(SPHBVF_ESTATE, "pair_compute before setup");
  FLUSH();
  int rc;
  ctx->tic(K_PAIR);
  if (ctx->halo_pending && !ctx->halo_done_step && ctx->tile_order && ctx->ntiles_interior > 0 &&
      (ctx->d.list16 || ctx->aorder_valid)) {
    // the halo of this step is still in flight on its own stream: the atoms that cannot see a ghost go first, the
    // compute stream then waits for the unpack, and the atoms along the brick faces follow
    const int nin = ctx->ntiles_interior, ntot = ctx->ntiles_total;
    const int snq = ctx->pair_warp ? -ctx->pair_nq : ctx->pair_nq;   // sign = chunk granularity (SPHBVF_PAIR_SCHED=warp)
    PairSubset in = {ctx->tile_order, nin, ctx->aorder, 0, ctx->natoms_interior, ctx->pair_queues, snq};
    PairSubset out = {ctx->tile_order + nin, ntot - nin, ctx->aorder, ctx->natoms_interior, ctx->d.nlocal,
                      ctx->pair_queues ? ctx->pair_queues + ctx->pair_nq + 1 : nullptr, snq};
    launch_pair(ctx->d, ctx->co, pair_flags(ctx), ctx->grid, ctx->w, &in, ctx->st);
    // the face atoms run on the halo stream, right behind the unpack: they fill the tail of the interior launch instead
    // of waiting for it; the compute stream then waits for both
    launch_pair(ctx->d, ctx->co, pair_flags(ctx), ctx->grid, ctx->w, &out, comm_halo_stream(ctx));
    if ((rc = comm_halo_mark(ctx))) return rc;
    if ((rc = comm_halo_join(ctx))) return rc;
  } else {
    if ((rc = comm_halo_join(ctx))) return rc;
    PairSubset all = {nullptr, (int)((long)ctx->grid.nt[0] * ctx->grid.nt[1] * ctx->grid.nt[2]), nullptr, 0, ctx->d.nlocal,
                      ctx->pair_queues, ctx->pair_warp ? -ctx->pair_nq : ctx->pair_nq};
    launch_pair(ctx->d, ctx->co, pair_flags(ctx), ctx->grid, ctx->w, &all, ctx->st);
  }
  ctx->toc();
  CKLAUNCH();
  return 0;
}

// Pair::virial_fdotr_compute (pair.cpp:1511-1560) = sum over local AND ghost atoms of x (x) f as the
// half-list pair style leaves them before reverse communication.  In the gather formulation ghosts
// carry no force; writing a ghost as x_g = x_owner + s_g and folding its force into the owner gives
//   V = sum_{owned i} x_i (x) f_i  -  1/2 sum_{owned i} sum_{ghost g in N(i)} s_g (x) F_{i<-g}
// (each periodic pair is seen twice in the full lists with opposite shifts and opposite forces, the
// half list holds it once); s_g = 0 for plain inter-brick ghosts.  Call right after
// sphbvf_pair_compute, before post_force fixes touch f.  Multi-rank: each rank returns its share.
int sphbvf_virial(sphbvf_ctx *ctx, double *virial6) {
  if (!ctx->setup_done) return ctx->fail(SPHBVF_ESTATE, "virial before setup");
  cudaSetDevice(ctx->cfg.device);
  FLUSH();
  { int rcj = comm_halo_join(ctx); if (rcj) return rcj; }
  if (!ctx->d_virial) CK(cudaMalloc((void **)&ctx->d_virial, sizeof(double) * 6));
  CK(cudaMemsetAsync(ctx->d_virial, 0, sizeof(double) * 6, ctx->st));
  launch_virial(ctx->d, ctx->co, pair_flags(ctx), ctx->grid, ctx->w, ctx->d_virial, ctx->st);
  CKLAUNCH();
  CK(cudaMemcpyAsync(virial6, ctx->d_virial, sizeof(double) * 6, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return 0;
}

int sphbvf_max_vsq(sphbvf_ctx *ctx, int groupbit, double *max_vsq) {
  if (!ctx->atoms_set) return ctx->fail(SPHBVF_ESTATE, "max_vsq before set_atoms");
  cudaSetDevice(ctx->cfg.device);
  FLUSH();
  if (!ctx->d_virial) CK(cudaMalloc((void **)&ctx->d_virial, sizeof(double) * 6));
  unsigned long long *slot = (unsigned long long *)ctx->d_virial;
  CK(cudaMemsetAsync(slot, 0, sizeof(unsigned long long), ctx->st));
  launch_max_vsq(ctx->d, groupbit, slot, ctx->st);
  CKLAUNCH();
  double v = 0.0;
  CK(cudaMemcpyAsync(&v, slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  if (ctx->cfg.nranks > 1) {   // MPI_Allreduce(MAX) of fix_dt_adaptive.cpp:148
    int rc;
    if ((rc = comm_allreduce_max_double(ctx, &v))) return rc;
  }
  *max_vsq = v;
  return 0;
}

int sphbvf_ke_tensor(sphbvf_ctx *ctx, int groupbit, double *ke6) {
  if (!ctx->atoms_set) return ctx->fail(SPHBVF_ESTATE, "ke_tensor before set_atoms");
  if (!ke6) return ctx->fail(SPHBVF_EINVAL, "ke_tensor: null output");
  cudaSetDevice(ctx->cfg.device);
  FLUSH();
  if (!ctx->d_virial) CK(cudaMalloc((void **)&ctx->d_virial, sizeof(double) * 6));
  int rc;
  const size_t need = sizeof(double) * 6 * (size_t)(ctx->d.nlocal / 256 + 2);
  if ((rc = stage(ctx, need))) return rc;
  launch_ke_tensor(ctx->d, ctx->co, groupbit, (double *)ctx->w.tmp_perm, ctx->d_virial, ctx->st);
  CKLAUNCH();
  CK(cudaMemcpyAsync(ke6, ctx->d_virial, sizeof(double) * 6, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return 0;
}

int sphbvf_final_integrate(sphbvf_ctx *ctx) {
  FLUSH();
  if (ctx->fuse) {   // deferred: see flush_final / sphbvf_initial_integrate
    ctx->final_pending = 1;
    ctx->pend_dt = ctx->cfg.dt;
    ctx->pend_step = ctx->ntimestep;
    return 0;
  }
  ctx->tic(K_FINAL);
  launch_final_integrate(ctx->d, ctx->co, ctx->cfg.dt, ctx->ntimestep, ctx->cfg.integrate_groupbit, ctx->with_dev, ctx->st);
  ctx->toc();
  CKLAUNCH();
  return 0;
}

int sphbvf_run(sphbvf_ctx *ctx, int nsteps) {
  if (!ctx->setup_done) return ctx->fail(SPHBVF_ESTATE, "run before setup");
  cudaSetDevice(ctx->cfg.device);
  if (ctx->run_nsteps_user < 0) ctx->run_nsteps = nsteps;   // update->nsteps
  int rc;
  for (int s = 0; s < nsteps; s++) {
    ctx->ntimestep++;
    if ((rc = sphbvf_initial_integrate(ctx))) return rc;
    if ((rc = sphbvf_post_integrate(ctx))) return rc;
    if ((rc = sphbvf_neighbor(ctx, nullptr))) return rc;
    if ((rc = sphbvf_pair_compute(ctx))) return rc;
    if ((rc = sphbvf_post_force(ctx))) return rc;
    if ((rc = sphbvf_final_integrate(ctx))) return rc;
    if ((rc = sphbvf_end_of_step(ctx))) return rc;
  }
  FLUSH();
  CK(cudaStreamSynchronize(ctx->st));
  return 0;
}

int sphbvf_nlocal(const sphbvf_ctx *ctx) { return ctx->d.nlocal; }
int sphbvf_nghost(const sphbvf_ctx *ctx) { return ctx->d.nghost; }
long sphbvf_ntimestep(const sphbvf_ctx *ctx) { return ctx->ntimestep; }
int sphbvf_nbuilds(const sphbvf_ctx *ctx) { return ctx->nbuilds; }
int sphbvf_pair_mode(const sphbvf_ctx *ctx) { return ctx->d.list16; }
int sphbvf_ndanger(const sphbvf_ctx *ctx) { return ctx->ndanger; }
void *sphbvf_stream(sphbvf_ctx *ctx) { return (void *)ctx->st; }

int sphbvf_sync(sphbvf_ctx *ctx) {
  FLUSH();
  { int rcj = comm_halo_join(ctx); if (rcj) return rcj; }
  CK(cudaStreamSynchronize(ctx->st));
  return 0;
}

long sphbvf_launch_count(const sphbvf_ctx *ctx) {
  long n = 0;
  for (int k = 0; k < K_NFAM; k++) n += ctx->launches_fam[k];
  return n;
}

int sphbvf_set_profiling(sphbvf_ctx *ctx, int on) {
  FLUSH();
  cudaStreamSynchronize(ctx->st);
  ctx->drain_events();
  ctx->profiling = on;
  for (int k = 0; k < K_NFAM; k++) { ctx->ms_fam[k] = 0.0; ctx->launches_fam[k] = 0; }
  return 0;
}

double sphbvf_kernel_ms(const sphbvf_ctx *cctx, int which, long *launches) {
  sphbvf_ctx *ctx = const_cast<sphbvf_ctx *>(cctx);
  if (which < 0 || which >= K_NFAM) return -1.0;
  if (flush_final(ctx)) return -1.0;
  cudaStreamSynchronize(ctx->st);
  ctx->drain_events();
  if (launches) *launches = ctx->launches_fam[which];
  return ctx->ms_fam[which];
}

// ---- field access --------------------------------------------------------------------------
static bool field_info(sphbvf_ctx *ctx, int field, void **ptr, int *ncols, int *is_int) {
  DevState &d = ctx->d;
  const int S = ctx->co.nspecies;
  *is_int = 0;
  *ncols = 1;
  switch (field) {
    case SPHBVF_F_TAG: *ptr = d.tag; *is_int = 1; break;
    case SPHBVF_F_TYPE: *ptr = d.type; *is_int = 1; break;
    case SPHBVF_F_MASK: *ptr = d.mask; *is_int = 1; break;
    case SPHBVF_F_SOLID_TAG: *ptr = d.solid; *is_int = 1; break;
    case SPHBVF_F_FIXED_TAG: *ptr = d.fixed; *is_int = 1; break;
    case SPHBVF_F_X: *ptr = d.x; *ncols = 3; break;
    case SPHBVF_F_V: *ptr = d.v; *ncols = 3; break;
    case SPHBVF_F_VEST: *ptr = d.vest; *ncols = 3; break;
    case SPHBVF_F_F: *ptr = d.f; *ncols = 3; break;
    case SPHBVF_F_RHO: *ptr = d.rho; break;
    case SPHBVF_F_RHOI: *ptr = d.rhoI; break;
    case SPHBVF_F_DRHO: *ptr = d.drho; break;
    case SPHBVF_F_E: *ptr = d.e; break;
    case SPHBVF_F_PHI: *ptr = d.phi; break;
    case SPHBVF_F_NUMBER_DENSITY: *ptr = d.nd; break;
    case SPHBVF_F_NW: *ptr = d.nw; *ncols = 3; break;
    case SPHBVF_F_DDV: *ptr = d.ddv; *ncols = 3; break;
    case SPHBVF_F_DDX: *ptr = d.ddx; *ncols = 3; break;
    case SPHBVF_F_RHOAUX1: *ptr = d.rhoAux1; break;
    case SPHBVF_F_RHOAUX2: *ptr = d.rhoAux2; break;
    case SPHBVF_F_PNEW: *ptr = d.Pnew; break;
    case SPHBVF_F_DEV: *ptr = d.dev; *ncols = 9; break;
    case SPHBVF_F_DDEV: *ptr = d.ddev; *ncols = 9; break;
    case SPHBVF_F_C: *ptr = d.C; *ncols = S; break;
    case SPHBVF_F_Q: *ptr = d.Q; *ncols = S; break;
    default: return false;
  }
  return true;
}


int sphbvf_download(sphbvf_ctx *ctx, int field, void *host) {
  void *p;
  int nc, is_int, rc;
  if (!field_info(ctx, field, &p, &nc, &is_int)) return ctx->fail(SPHBVF_EINVAL, "unknown field %d", field);
  if (ctx->migrated) return ctx->fail(SPHBVF_ESTATE, "atoms migrated between ranks: use sphbvf_download_local");
  FLUSH();
  const int n = ctx->d.nlocal;
  if (!n || !nc) return 0;
  cudaSetDevice(ctx->cfg.device);
  const size_t eb = is_int ? 4 : 8, bytes = eb * (size_t)n * nc;
  if ((rc = stage(ctx, bytes))) return rc;
  const long tot = (long)n * nc;
  const int blocks = (int)((tot + 255) / 256);
  if (is_int) scatter_rows_kernel<int><<<blocks, 256, 0, ctx->st>>>((const int *)p, (int *)ctx->w.tmp_perm, ctx->d.slot, n, nc, 1);
  else scatter_rows_kernel<double><<<blocks, 256, 0, ctx->st>>>((const double *)p, (double *)ctx->w.tmp_perm, ctx->d.slot, n, nc, 1);
  CK(cudaMemcpyAsync(host, ctx->w.tmp_perm, bytes, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return 0;
}

int sphbvf_download_local(sphbvf_ctx *ctx, int field, void *host, int cap_rows) {
  void *p;
  int nc, is_int;
  if (!field_info(ctx, field, &p, &nc, &is_int)) return ctx->fail(SPHBVF_EINVAL, "unknown field %d", field);
  const int n = ctx->d.nlocal;
  if (cap_rows < n) return ctx->fail(SPHBVF_EINVAL, "download_local: buffer too small (%d < %d)", cap_rows, n);
  if (!n || !nc) return 0;
  cudaSetDevice(ctx->cfg.device);
  FLUSH();
  CK(cudaMemcpyAsync(host, p, (is_int ? 4 : 8) * (size_t)n * nc, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return 0;
}

// fields that decide the kernel specialisation (and, for type / solid_tag, the packed list entries)
static void note_upload(sphbvf_ctx *ctx, int field) {
  if (field == SPHBVF_F_E || field == SPHBVF_F_DEV) ctx->flags_dirty |= 1;
  if (field == SPHBVF_F_TYPE || field == SPHBVF_F_SOLID_TAG || field == SPHBVF_F_FIXED_TAG) ctx->flags_dirty |= 3;   // + rebuild
}

int sphbvf_upload_local(sphbvf_ctx *ctx, int field, const void *host, int nrows) {
  void *p;
  int nc, is_int;
  if (!field_info(ctx, field, &p, &nc, &is_int)) return ctx->fail(SPHBVF_EINVAL, "unknown field %d", field);
  note_upload(ctx, field);
  const int n = ctx->d.nlocal;
  if (nrows != n) return ctx->fail(SPHBVF_EINVAL, "upload_local: %d rows given, %d atoms owned", nrows, n);
  if (!n || !nc) return 0;
  cudaSetDevice(ctx->cfg.device);
  FLUSH();
  ctx->pack_valid = 0;
  CK(cudaMemcpyAsync(p, host, (is_int ? 4 : 8) * (size_t)n * nc, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return 0;
}

int sphbvf_upload(sphbvf_ctx *ctx, int field, const void *host) {
  void *p;
  int nc, is_int, rc;
  if (!field_info(ctx, field, &p, &nc, &is_int)) return ctx->fail(SPHBVF_EINVAL, "unknown field %d", field);
  if (ctx->migrated) return ctx->fail(SPHBVF_ESTATE, "atoms migrated between ranks: upload by slot is undefined");
  note_upload(ctx, field);
  FLUSH();
  ctx->pack_valid = 0;
  const int n = ctx->d.nlocal;
  if (!n || !nc) return 0;
  cudaSetDevice(ctx->cfg.device);
  const size_t eb = is_int ? 4 : 8, bytes = eb * (size_t)n * nc;
  if ((rc = stage(ctx, bytes))) return rc;
  CK(cudaMemcpyAsync(ctx->w.tmp_perm, host, bytes, cudaMemcpyHostToDevice, ctx->st));
  const long tot = (long)n * nc;
  const int blocks = (int)((tot + 255) / 256);
  if (is_int) scatter_rows_kernel<int><<<blocks, 256, 0, ctx->st>>>((const int *)ctx->w.tmp_perm, (int *)p, ctx->d.slot, n, nc, 0);
  else scatter_rows_kernel<double><<<blocks, 256, 0, ctx->st>>>((const double *)ctx->w.tmp_perm, (double *)p, ctx->d.slot, n, nc, 0);
  CKLAUNCH();
  CK(cudaStreamSynchronize(ctx->st));
  return 0;
}

// each unordered pair once, as (tag_i, tag_j): what the reference's half list holds
long sphbvf_get_pairs(sphbvf_ctx *ctx, int *out, long cap) {
  DevState &d = ctx->d;
  const int n = d.nlocal, nall = d.nlocal + d.nghost;
  if (!n || !(d.list16 ? (void *)d.neigh16 : (void *)d.neigh)) return 0;
  cudaSetDevice(ctx->cfg.device);
  if (d.list16 && !ctx->expanded_valid) {   // tile form: expand the 16-bit slot entries into the gather form first
    if (!d.neigh && cudaMalloc((void **)&d.neigh, sizeof(int) * (size_t)d.maxneigh * d.stride) != cudaSuccess) {
      ctx->fail(SPHBVF_ECUDA, "get_pairs: out of device memory for the expanded list");
      return -1;
    }
    launch_expand_list(d, ctx->grid, ctx->w, ctx->st);
    ctx->expanded_valid = 1;
  }
  cudaStreamSynchronize(ctx->st);
  std::vector<int> nn(n), ptag(nall), gowner(std::max(d.nghost, 1));
  std::vector<double> gshift(3 * (size_t)std::max(d.nghost, 1));
  cudaMemcpy(nn.data(), d.numneigh, sizeof(int) * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(ptag.data(), d.ptag, sizeof(int) * nall, cudaMemcpyDeviceToHost);
  if (d.nghost) {
    cudaMemcpy(gowner.data(), d.gowner, sizeof(int) * d.nghost, cudaMemcpyDeviceToHost);
    cudaMemcpy(gshift.data(), d.gshift, sizeof(double) * 3 * d.nghost, cudaMemcpyDeviceToHost);
  }
  int mx = 0;
  for (int i = 0; i < n; i++) mx = std::max(mx, nn[i]);
  std::vector<int> row(n);
  long cnt = 0;
  for (int k = 0; k < mx; k++) {
    cudaMemcpy(row.data(), d.neigh + (size_t)k * d.stride, sizeof(int) * n, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; i++) {
      if (k >= nn[i]) continue;
      const int j = row[i] & NEIGH_JMASK;
      bool emit;
      if (j < n) emit = i < j;
      else {
        const int g = j - n;
        if (ptag[i] != ptag[j]) emit = ptag[i] < ptag[j];
        else {   // an atom and its own periodic image: keep the lexicographically positive shift
          const double *s = &gshift[3 * (size_t)g];
          emit = s[2] > 0 || (s[2] == 0 && (s[1] > 0 || (s[1] == 0 && s[0] > 0)));
        }
      }
      if (!emit) continue;
      if (out && cnt < cap) { out[2 * cnt] = ptag[i]; out[2 * cnt + 1] = ptag[j]; }
      cnt++;
    }
  }
  return cnt;
}

}  // extern "C"

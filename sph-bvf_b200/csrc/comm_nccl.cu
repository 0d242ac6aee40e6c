// comm_nccl.cu -- brick decomposition across the GPUs of one box, one process per GPU, NCCL
// point-to-point over NVLink 5 / NVSwitch.
//
// Replaces CommBrick::{setup,exchange,borders,forward_comm} + AtomVecSsaTsdpdAtomic::{pack,unpack}_
// {exchange,border,comm} (comm_brick.cpp:161-880, atom_vec_ssa_tsdpd_atomic.cpp:426-1638) and the
// MPI_Allreduce of the rebuild vote (neighbor.cpp:1997).  Differences that are deliberate:
//   * every peer is one NVSwitch hop away at full bandwidth, so the reference's three staged
//     sweeps (x, then y, then z, each forwarding ghosts of ghosts) become ONE exchange with up to
//     26 neighbour bricks (7 on a 2x2x2 grid), posted as a single NCCL group;
//   * no reverse communication: each rank gathers over the full neighbour set of the atoms it
//     owns, so ghosts are read-only and only the 16 (+S, +9) doubles the pair
//     kernel reads travel per ghost per step;
//   * migration moves the primary state only (the pair outputs are recomputed before use).
// A brick that is its own neighbour across a periodic face (one brick in that dimension)
// exchanges with itself through a device copy; the arithmetic x + shift is the reference's
// (atom_vec_ssa_tsdpd_atomic.cpp:487-500).
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is bound at run time, see NcclApi
#include <string.h>

#include <algorithm>
#include <vector>

#include "context.cuh"

using namespace sphbvf;

namespace {

// libnccl.so.2 is dlopen'ed on first use instead of being a link-time dependency: single-GPU and
// CPU-only processes never load it, and a process that already holds an NCCL (PyTorch bundles its
// own libnccl.so.2) gets that same copy instead of a second one with clashing symbols.
struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId;
  decltype(&ncclCommInitRank) CommInitRank;
  decltype(&ncclCommDestroy) CommDestroy;
  decltype(&ncclGetErrorString) GetErrorString;
  decltype(&ncclGroupStart) GroupStart;
  decltype(&ncclGroupEnd) GroupEnd;
  decltype(&ncclSend) Send;
  decltype(&ncclRecv) Recv;
  decltype(&ncclAllGather) AllGather;
  decltype(&ncclAllReduce) AllReduce;
  bool ok = false;
};

NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
#define BIND(name) api.name = (decltype(api.name))dlsym(h, "nccl" #name); if (!api.name) return nullptr
  BIND(GetUniqueId); BIND(CommInitRank); BIND(CommDestroy); BIND(GetErrorString); BIND(GroupStart); BIND(GroupEnd);
  BIND(Send); BIND(Recv); BIND(AllGather); BIND(AllReduce);
#undef BIND
  api.ok = true;
  return &api;
}

constexpr int ND = 27;       // directions (dx+1) + 3 (dy+1) + 9 (dz+1); 13 = stay
constexpr int NDX = 32;      // stride of a rank's count vector: ND counts + [27] non-finite flag, [28] lost flag, [29] status
constexpr int NHALO = 16;    // the 96-byte record (12) + pD (4); pD travels only when the coming pair pass reads it
struct DirTable {
  int peer[ND];
  int off[ND + 1];           // first slot of each direction in the send list / ghost slab
  double shift[ND][3];
};

}  // namespace

struct CommState {
  ncclComm_t comm = nullptr;
  DirTable send{}, recv{};
  int nsend = 0;
  int *sendidx = nullptr;    // [nsend] owned atom of each send slot, grouped by direction
  int send_cap = 0;
  double *sendbuf = nullptr, *recvbuf = nullptr;
  size_t buf_cap = 0;        // doubles
  int *d_counts = nullptr;   // [NDX] local counters + flags | [NDX] cursors | [NDX * nranks] gathered | [16] scratch
  int *h_counts = nullptr;   // pinned mirror
  int *d_dir = nullptr;      // [nmax] migration direction of each owned atom
  int dir_cap = 0;
  int *d_keep = nullptr, *d_pos = nullptr;
  cudaStream_t halo_st = nullptr;             // the per-step halo runs here, beside the interior of the pair pass
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return ctx->fail(SPHBVF_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define NK(call)                                                                                   \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess)                                                                         \
      return ctx->fail(SPHBVF_ECOMM, "%s failed: %s (%s:%d)", #call, nccl_api()->GetErrorString(r_), __FILE__, __LINE__); \
  } while (0)

static inline int nblocks(long n, int t) { return (int)((n + t - 1) / t); }

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
struct BrickGeom {
  double lo[3], hi[3], prd[3], sublo[3], subhi[3];
  int periodic[3], pg[3], loc[3], dim;
};

__device__ __forceinline__ void brick_bounds_dev(const BrickGeom &b, int k, int t, double &lo, double &hi) {
  lo = b.lo[k] + b.prd[k] * ((double)t / b.pg[k]);
  hi = t == b.pg[k] - 1 ? b.hi[k] : b.lo[k] + b.prd[k] * ((double)(t + 1) / b.pg[k]);
}

// Domain::pbc + the owner test of CommBrick::exchange (comm_brick.cpp:585-700): direction of the
// brick that owns the atom now.  flags[1] = atom more than one brick away / outside a fixed box.
__global__ void classify_kernel(const DevState d, const BrickGeom b, int *dir, int *counts, int *flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
  int code = 0, mul = 1;
  for (int k = 0; k < 3; k++, mul *= 3) {
    double xk = d.x[3 * (size_t)i + k];
    if (!isfinite(xk)) { flags[0] = 1; xk = b.sublo[k]; }
    int delta = 0;
    if (!(b.dim == 2 && k == 2)) {
      if (b.periodic[k]) {
        if (xk < b.lo[k]) xk += b.prd[k];
        if (xk >= b.hi[k]) {
          xk -= b.prd[k];
          xk = xk > b.lo[k] ? xk : b.lo[k];
        }
        d.x[3 * (size_t)i + k] = xk;
      }
      const int p = b.pg[k], loc = b.loc[k];
      if (p > 1 && (xk < b.sublo[k] || xk >= b.subhi[k])) {
        if (xk >= b.subhi[k]) delta = loc < p - 1 ? 1 : (b.periodic[k] ? -1 : 0);
        else delta = loc > 0 ? -1 : (b.periodic[k] ? 1 : 0);
        // wrapped across the periodic face: the owner is the brick at the other end
        if (b.periodic[k] && loc == p - 1 && xk < b.sublo[k]) {
          double l0, h0;
          brick_bounds_dev(b, k, 0, l0, h0);
          if (xk < h0) delta = 1;
        }
        if (b.periodic[k] && loc == 0 && xk >= b.subhi[k]) {
          double l0, h0;
          brick_bounds_dev(b, k, p - 1, l0, h0);
          if (xk >= l0) delta = -1;
        }
        if (delta) {
          const int t = (loc + delta + p) % p;
          double tl, th;
          brick_bounds_dev(b, k, t, tl, th);
          if (xk < tl || xk >= th) flags[1] = 1;   // moved further than the adjacent brick
        }
      }
    }
    code += (delta + 1) * mul;
  }
  dir[i] = code;
  if (code != 13) atomicAdd(&counts[code], 1);
}

// keep[i] = atom stays on this rank
__global__ void keep_flag_kernel(int n, const int *dir, int *keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) keep[i] = i < n ? dir[i] == 13 : 0;
}

// stayers -> perm (new -> old, order preserved); leavers -> migration records
//   record: tag type mask solid fixed | x3 v3 vest3 rho rhoI e | C[S] | dev[9]   (ints as doubles)
__global__ void pack_leavers_kernel(const DevState d, const int S, const int *dir, const int *pos, int *perm,
                                    const DirTable t, int *cursor, double *buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
  const int code = dir[i];
  if (code == 13) { perm[pos[i]] = i; return; }
  const int NM = 26 + S;
  double *r = buf + (size_t)(t.off[code] + atomicAdd(&cursor[code], 1)) * NM;
  r[0] = d.tag[i]; r[1] = d.type[i]; r[2] = d.mask[i]; r[3] = d.solid[i]; r[4] = d.fixed[i];
  const size_t i3 = 3 * (size_t)i;
  for (int k = 0; k < 3; k++) { r[5 + k] = d.x[i3 + k]; r[8 + k] = d.v[i3 + k]; r[11 + k] = d.vest[i3 + k]; }
  r[14] = d.rho[i]; r[15] = d.rhoI[i]; r[16] = d.e[i];
  for (int k = 0; k < S; k++) r[17 + k] = d.C[(size_t)i * S + k];
  for (int k = 0; k < 9; k++) r[17 + S + k] = d.dev[9 * (size_t)i + k];
}

__global__ void unpack_arrivals_kernel(const DevState d, const int S, const int first, const int n, const double *buf) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int NM = 26 + S;
  const double *r = buf + (size_t)q * NM;
  const int i = first + q;
  d.tag[i] = (int)r[0]; d.type[i] = (int)r[1]; d.mask[i] = (int)r[2]; d.solid[i] = (int)r[3]; d.fixed[i] = (int)r[4];
  d.slot[i] = -1;
  const size_t i3 = 3 * (size_t)i;
  for (int k = 0; k < 3; k++) { d.x[i3 + k] = r[5 + k]; d.v[i3 + k] = r[8 + k]; d.vest[i3 + k] = r[11 + k]; }
  d.rho[i] = r[14]; d.rhoI[i] = r[15]; d.e[i] = r[16];
  for (int k = 0; k < S; k++) d.C[(size_t)i * S + k] = r[17 + k];
  for (int k = 0; k < 9; k++) d.dev[9 * (size_t)i + k] = r[17 + S + k];
}

// CommBrick::borders slab test (comm_brick.cpp:765-770) for all 26 directions at once:
// atom i is a ghost of the brick in direction (dx,dy,dz) iff it lies within cutghost of every
// face that direction crosses.
__device__ __forceinline__ void slab_flags(const BrickGeom &b, const double cut, const double *x, int lo[3], int hi[3]) {
  for (int k = 0; k < 3; k++) {
    const bool use = !(b.dim == 2 && k == 2);
    lo[k] = use && x[k] <= b.sublo[k] + cut;
    hi[k] = use && x[k] >= b.subhi[k] - cut;
  }
}

template <bool FILL>
__global__ void border_kernel(const DevState d, const BrickGeom b, const double cut, const DirTable t, int *counts,
                              int *sendidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
  int lo[3], hi[3];
  slab_flags(b, cut, &d.x[3 * (size_t)i], lo, hi);
  if (!(lo[0] | hi[0] | lo[1] | hi[1] | lo[2] | hi[2])) return;
  for (int code = 0; code < ND; code++) {
    if (code == 13 || t.peer[code] < 0) continue;
    const int s[3] = {code % 3 - 1, (code / 3) % 3 - 1, code / 9 - 1};
    bool need = true;
    for (int k = 0; k < 3; k++)
      if ((s[k] == -1 && !lo[k]) || (s[k] == 1 && !hi[k])) need = false;
    if (!need) continue;
    const int r = atomicAdd(&counts[code], 1);
    if (FILL) sendidx[t.off[code] + r] = i;
  }
}

// pack_comm / pack_border: the records the pair kernel reads, position shifted by the periodic image
//   record: prec(12) [pD(4) if with_pd] | pCs[S] | pdev[9] (with_dev) | flags tag (border)
// pD = {rhoI, art, C0, e} is read by the Shepard-filter steps (every 20th) and by the stochastic term only: on all
// other steps a ghost costs 96 B instead of 128 B on the wire.
__host__ __device__ __forceinline__ int halo_width(int S, int with_dev, int border, int with_pd) {
  return 12 + (with_pd ? 4 : 0) + S + (with_dev ? 9 : 0) + (border ? 2 : 0);
}

__global__ void halo_pack_kernel(const DevState d, const int S, const int with_dev, const int border, const int with_pd,
                                 const int nsend, const int *sendidx, const DirTable t, double *buf) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nsend) return;
  int code = 0;
  while (s >= t.off[code + 1]) code++;
  const int i = sendidx[s];
  const int R = halo_width(S, with_dev, border, with_pd);
  double *r = buf + (size_t)s * R;
  const Prec P = d.prec[i];
  Rec4 A = P.A;
  // x + shift in the reference's order: one rounded add per shifted dimension
  A.x += t.shift[code][0]; A.y += t.shift[code][1]; A.z += t.shift[code][2];
  const Rec4 B = P.B, C = P.C;
  r[0] = A.x; r[1] = A.y; r[2] = A.z; r[3] = A.w;
  r[4] = B.x; r[5] = B.y; r[6] = B.z; r[7] = B.w;
  r[8] = C.x; r[9] = C.y; r[10] = C.z; r[11] = C.w;
  int q = 12;
  if (with_pd) {
    const Rec4 D = d.pD[i];
    r[12] = D.x; r[13] = D.y; r[14] = D.z; r[15] = D.w;
    q = 16;
  }
  for (int k = 0; k < S; k++) r[q++] = d.pCs[(size_t)i * S + k];
  if (with_dev)
    for (int k = 0; k < 9; k++) r[q++] = d.pdev[9 * (size_t)i + k];
  if (border) { r[q++] = d.pflags[i]; r[q++] = d.tag[i]; }
}

__global__ void halo_unpack_kernel(const DevState d, const int S, const int with_dev, const int border, const int with_pd,
                                   const DirTable t, const double *buf) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= d.nghost) return;
  const int R = halo_width(S, with_dev, border, with_pd);
  const double *r = buf + (size_t)g * R;
  const int j = d.nlocal + g;
  Prec P;
  P.A = make_rec4(r[0], r[1], r[2], r[3]);
  P.B = make_rec4(r[4], r[5], r[6], r[7]);
  P.C = make_rec4(r[8], r[9], r[10], r[11]);
  d.prec[j] = P;
  int q = 12;
  if (with_pd) {
    d.pD[j] = make_rec4(r[12], r[13], r[14], r[15]);
    q = 16;
  }
  for (int k = 0; k < S; k++) d.pCs[(size_t)j * S + k] = r[q++];
  if (with_dev)
    for (int k = 0; k < 9; k++) d.pdev[9 * (size_t)j + k] = r[q++];
  if (border) {
    d.pflags[j] = (int)r[q++];
    d.ptag[j] = (int)r[q++];
    int code = 0;
    while (g >= t.off[code + 1]) code++;
    d.gowner[g] = -1;
    for (int k = 0; k < 3; k++) d.gshift[3 * (size_t)g + k] = t.shift[code][k];
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
extern "C" int sphbvf_comm_plan(const sphbvf_config *cfg, int rank, int *peer, double *shift);

static BrickGeom geom(const sphbvf_ctx *ctx) {
  BrickGeom g;
  const Box &b = ctx->box;
  for (int k = 0; k < 3; k++) {
    g.lo[k] = b.lo[k]; g.hi[k] = b.hi[k]; g.prd[k] = b.prd[k];
    g.sublo[k] = b.sublo[k]; g.subhi[k] = b.subhi[k];
    g.periodic[k] = b.periodic[k];
    g.pg[k] = ctx->cfg.procgrid[k];
  }
  const int r = ctx->cfg.rank;
  g.loc[0] = r % g.pg[0];
  g.loc[1] = (r / g.pg[0]) % g.pg[1];
  g.loc[2] = r / (g.pg[0] * g.pg[1]);
  g.dim = b.dim;
  return g;
}

static int ensure_buf(sphbvf_ctx *ctx, size_t doubles) {
  CommState *c = ctx->comm;
  if (doubles <= c->buf_cap) return 0;
  CK(cudaStreamSynchronize(ctx->st));
  if (c->sendbuf) cudaFree(c->sendbuf);
  if (c->recvbuf) cudaFree(c->recvbuf);
  c->buf_cap = doubles + doubles / 4 + 4096;
  CK(cudaMalloc((void **)&c->sendbuf, sizeof(double) * c->buf_cap));
  CK(cudaMalloc((void **)&c->recvbuf, sizeof(double) * c->buf_cap));
  return 0;
}

static int ensure_dir(sphbvf_ctx *ctx, int n) {
  CommState *c = ctx->comm;
  if (n + 1 <= c->dir_cap) return 0;
  for (int **p : {&c->d_dir, &c->d_keep, &c->d_pos}) {
    if (*p) cudaFree(*p);
    CK(cudaMalloc((void **)p, sizeof(int) * (size_t)(ctx->d.nmax + 2)));
  }
  c->dir_cap = ctx->d.nmax + 2;
  return 0;
}

// all ranks learn every rank's per-direction counts: recvcnt[d'] = what peer(d') sends towards -d'.
// The error decision rides on the same all-gather (slots 27..29 of each rank's vector: non-finite positions,
// lost atoms, local status), so that EVERY rank returns the same error before any payload is posted -- a rank
// that left alone would leave its peers spinning in ncclSend/ncclRecv.
static int exchange_counts(sphbvf_ctx *ctx, int *d_local, int *sendcnt, int *recvcnt, const char *lost_msg, int rc_local) {
  CommState *c = ctx->comm;
  const int P = ctx->cfg.nranks;
  int *gathered = c->d_counts + 2 * NDX;
  CK(cudaMemcpyAsync(d_local + ND, ctx->w.flags, sizeof(int) * 2, cudaMemcpyDeviceToDevice, ctx->st));
  int *h_rc = c->h_counts + NDX * P + 8;   // pinned scratch behind the gathered counts
  *h_rc = rc_local ? -rc_local : 0;
  CK(cudaMemcpyAsync(d_local + ND + 2, h_rc, sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  NK(nccl_api()->AllGather(d_local, gathered, NDX, ncclInt, c->comm, ctx->st));
  CK(cudaMemcpyAsync(c->h_counts, gathered, sizeof(int) * NDX * P, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  for (int r = 0; r < P; r++) {
    const int *v = c->h_counts + r * NDX;
    if (v[ND]) return ctx->fail(SPHBVF_ENONFINITE, "Non-numeric positions - simulation unstable%s", r == ctx->cfg.rank ? "" : " (on another GPU)");
    if (v[ND + 1]) return ctx->fail(SPHBVF_ELOST, "%s%s", lost_msg, r == ctx->cfg.rank ? "" : " (on another GPU)");
  }
  if (rc_local) return rc_local;   // keep the local message
  for (int r = 0; r < P; r++)
    if (c->h_counts[r * NDX + ND + 2])
      return ctx->fail(-c->h_counts[r * NDX + ND + 2], "GPU %d of this run failed with status %d (see its message)", r, -c->h_counts[r * NDX + ND + 2]);
  for (int dcode = 0; dcode < ND; dcode++) {
    sendcnt[dcode] = c->send.peer[dcode] >= 0 ? c->h_counts[ctx->cfg.rank * NDX + dcode] : 0;
    const int p = c->recv.peer[dcode];
    recvcnt[dcode] = p >= 0 ? c->h_counts[p * NDX + (ND - 1 - dcode)] : 0;
  }
  sendcnt[13] = recvcnt[13] = 0;
  return 0;
}

// one NCCL group: sends in ascending direction order, receives in descending order, so that
// several messages between the same two ranks pair up in the order NCCL matches them; a brick
// that neighbours itself copies on the device.
static int exchange_payload(sphbvf_ctx *ctx, const int *sendcnt, const int *sendoff, const int *recvcnt,
                            const int *recvoff, int width, cudaStream_t st) {
  CommState *c = ctx->comm;
  const int me = ctx->cfg.rank;
  bool any = false;
  for (int dcode = 0; dcode < ND; dcode++) {
    if (sendcnt[dcode] && c->send.peer[dcode] == me) {
      const int rd = ND - 1 - dcode;   // arrives as "from direction -d"
      CK(cudaMemcpyAsync(c->recvbuf + (size_t)recvoff[rd] * width, c->sendbuf + (size_t)sendoff[dcode] * width,
                         sizeof(double) * (size_t)sendcnt[dcode] * width, cudaMemcpyDeviceToDevice, st));
    } else if (sendcnt[dcode] || (recvcnt[dcode] && c->recv.peer[dcode] != me)) any = true;
  }
  if (!any) return 0;
  NK(nccl_api()->GroupStart());
  for (int dcode = 0; dcode < ND; dcode++)
    if (sendcnt[dcode] && c->send.peer[dcode] != me)
      NK(nccl_api()->Send(c->sendbuf + (size_t)sendoff[dcode] * width, (size_t)sendcnt[dcode] * width, ncclDouble,
                  c->send.peer[dcode], c->comm, st));
  for (int dcode = ND - 1; dcode >= 0; dcode--)
    if (recvcnt[dcode] && c->recv.peer[dcode] != me)
      NK(nccl_api()->Recv(c->recvbuf + (size_t)recvoff[dcode] * width, (size_t)recvcnt[dcode] * width, ncclDouble,
                  c->recv.peer[dcode], c->comm, st));
  NK(nccl_api()->GroupEnd());
  return 0;
}

static int halo(sphbvf_ctx *ctx, int border, int with_pd, cudaStream_t st) {
  CommState *c = ctx->comm;
  DevState &d = ctx->d;
  const int S = ctx->co.nspecies;
  const int R = halo_width(S, ctx->with_dev, border, with_pd);
  int rc;
  if ((rc = ensure_buf(ctx, (size_t)std::max(c->nsend, d.nghost) * R + 64))) return rc;
  if (c->nsend) {
    halo_pack_kernel<<<nblocks(c->nsend, 256), 256, 0, st>>>(d, S, ctx->with_dev, border, with_pd, c->nsend, c->sendidx,
                                                              c->send, c->sendbuf);
    SPHBVF_LAUNCHED(1);
  }
  int sendcnt[ND], recvcnt[ND];
  for (int k = 0; k < ND; k++) {
    sendcnt[k] = c->send.off[k + 1] - c->send.off[k];
    recvcnt[k] = c->recv.off[k + 1] - c->recv.off[k];
  }
  if ((rc = exchange_payload(ctx, sendcnt, c->send.off, recvcnt, c->recv.off, R, st))) return rc;
  if (d.nghost) {
    halo_unpack_kernel<<<nblocks(d.nghost, 256), 256, 0, st>>>(d, S, ctx->with_dev, border, with_pd, c->recv, c->recvbuf);
    SPHBVF_LAUNCHED(1);
  }
  CK(cudaGetLastError());
  return 0;
}

// Per-step halo (Comm::forward_comm).  It runs on its OWN high-priority stream, ordered after the kernel that wrote
// the records (event on the compute stream), and only the tiles of the pair pass that can see a ghost wait for it
// (sphbvf_pair_compute launches the interior tiles first, then joins): pack -> NVLink -> unpack overlaps the interior
// of the pair pass instead of stretching the step.  NCCL operations of the one communicator never overlap each other:
// everything else that uses it (rebuild, votes) joins the halo stream first (comm_halo_join).
int comm_forward(sphbvf_ctx *ctx, int with_pd) {
  CommState *c = ctx->comm;
  if (!c) return ctx->fail(SPHBVF_ECOMM, "sphbvf_comm_init has not been called");
  if (!c->halo_st) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = numerically lowest = highest priority
    CK(cudaStreamCreateWithPriority(&c->halo_st, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
  }
  const bool overlap = ctx->overlap_halo != 0;
  cudaStream_t hs = overlap ? c->halo_st : ctx->st;
  if (overlap) {
    CK(cudaEventRecord(c->ev_ready, ctx->st));
    CK(cudaStreamWaitEvent(hs, c->ev_ready, 0));
  }
  const int rc = halo(ctx, 0, with_pd, hs);
  if (rc) return rc;
  if (overlap) {
    CK(cudaEventRecord(c->ev_done, hs));
    ctx->halo_pending = 1;
  }
  return 0;
}

// the stream the halo in flight runs on, and a way to extend what comm_halo_join waits for: the pair pass queues
// the atoms along the brick faces behind the unpack on this stream, beside the interior's tail on the compute stream
cudaStream_t comm_halo_stream(sphbvf_ctx *ctx) { return ctx->comm ? ctx->comm->halo_st : nullptr; }
int comm_halo_mark(sphbvf_ctx *ctx) {
  CK(cudaEventRecord(ctx->comm->ev_done, ctx->comm->halo_st));
  return 0;
}
// can ghosts arrive through the face of this brick in dimension k (side 0: low, 1: high)?
int comm_face_has_peer(const sphbvf_ctx *ctx, int k, int side) {
  if (!ctx->comm) return 0;
  static const int step[3] = {1, 3, 9};
  return ctx->comm->recv.peer[13 + (side ? step[k] : -step[k])] >= 0;
}

// make the compute stream wait for a halo in flight (no host synchronisation)
int comm_halo_join(sphbvf_ctx *ctx) {
  if (!ctx->halo_pending) return 0;
  ctx->halo_pending = 0;
  CK(cudaStreamWaitEvent(ctx->st, ctx->comm->ev_done, 0));
  return 0;
}

// max over ranks of up to 8 ints: the rebuild vote (neighbor.cpp:1997) and the per-run flags
int comm_allreduce_max(sphbvf_ctx *ctx, int *vals, int n) {
  CommState *c = ctx->comm;
  if (!c) return ctx->fail(SPHBVF_ECOMM, "sphbvf_comm_init has not been called");
  if (n > 8) return ctx->fail(SPHBVF_EINVAL, "comm_allreduce_max: n > 8");
  { int rcj = comm_halo_join(ctx); if (rcj) return rcj; }
  int *v = c->d_counts + 2 * NDX + NDX * ctx->cfg.nranks;
  for (int k = 0; k < n; k++) c->h_counts[k] = vals[k];
  CK(cudaMemcpyAsync(v, c->h_counts, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->st));
  NK(nccl_api()->AllReduce(v, v, n, ncclInt, ncclMax, c->comm, ctx->st));
  CK(cudaMemcpyAsync(c->h_counts, v, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  for (int k = 0; k < n; k++) vals[k] = c->h_counts[k];
  return 0;
}

int comm_vote(sphbvf_ctx *ctx, int *flag) { return comm_allreduce_max(ctx, flag, 1); }

// Collective status: every rank passes the code of the local work it just did and all of them return the same
// decision, so that nobody goes on to post sends / receives towards a rank that has already given up.
static int comm_agree(sphbvf_ctx *ctx, int rc_local) {
  int worst = rc_local ? -rc_local : 0;   // codes are negative
  const int rc = comm_allreduce_max(ctx, &worst, 1);
  if (rc_local) return rc_local;          // keep the local message
  if (rc) return rc;
  if (worst) return ctx->fail(-worst, "another GPU of this run failed with status %d (see its message)", -worst);
  return 0;
}

int comm_allreduce_max_double(sphbvf_ctx *ctx, double *val) {
  CommState *c = ctx->comm;
  if (!c) return ctx->fail(SPHBVF_ECOMM, "sphbvf_comm_init has not been called");
  { int rcj = comm_halo_join(ctx); if (rcj) return rcj; }
  double *v = (double *)(c->d_counts + 2 * NDX + NDX * ctx->cfg.nranks + 8);   // 8-byte aligned scratch (NDX is even)
  double *h = (double *)(c->h_counts + 2);
  *h = *val;
  CK(cudaMemcpyAsync(v, h, sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  NK(nccl_api()->AllReduce(v, v, 1, ncclDouble, ncclMax, c->comm, ctx->st));
  CK(cudaMemcpyAsync(h, v, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  *val = *h;
  return 0;
}

// the rebuild branch of verlet.cpp:268-296 on a brick: pbc, exchange, sort, borders, list.
// Error protocol: rank-local work (allocations, launches) never returns between two collectives; its status goes
// through comm_agree / the flag slots of exchange_counts, so all ranks leave with the same decision.
static int migrate_prepare(sphbvf_ctx *ctx, const int *sendoff, int nleave, int narrive, int *cur) {
  CommState *c = ctx->comm;
  DevState &d = ctx->d;
  NeighWork &w = ctx->w;
  cudaStream_t st = ctx->st;
  const int S = ctx->co.nspecies;
  const int NM = 26 + S;
  const int nstay = d.nlocal - nleave;
  int rc;
  if ((rc = ensure_buf(ctx, (size_t)std::max(nleave, narrive) * NM + 64))) return rc;
  if (nstay + narrive > d.nmax) {
    if ((rc = ctx_ensure_capacity(ctx, nstay + narrive + (nstay + narrive) / 8 + 1024, d.nallmax))) return rc;
    if ((rc = ensure_dir(ctx, d.nmax))) return rc;
  }
  if (nleave) {
    DirTable t = c->send;
    for (int k = 0; k <= ND; k++) t.off[k] = sendoff[k];
    keep_flag_kernel<<<nblocks(d.nlocal + 1, 256), 256, 0, st>>>(d.nlocal, c->d_dir, c->d_keep);
    exclusive_scan(c->d_keep, c->d_pos, d.nlocal + 1, w.scan_tmp, st);
    pack_leavers_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, S, c->d_dir, c->d_pos, w.perm, t, cur, c->sendbuf);
    SPHBVF_LAUNCHED(2);
    // compaction of the stayers, order preserved
    if ((rc = permute_state(ctx, nstay, true))) return rc;
  }
  CK(cudaGetLastError());
  return 0;
}

int comm_rebuild(sphbvf_ctx *ctx) {
  CommState *c = ctx->comm;
  if (!c) return ctx->fail(SPHBVF_ECOMM, "sphbvf_comm_init has not been called");
  DevState &d = ctx->d;
  NeighWork &w = ctx->w;
  cudaStream_t st = ctx->st;
  const int S = ctx->co.nspecies;
  const BrickGeom bg = geom(ctx);
  int rc;
  if ((rc = comm_halo_join(ctx))) return rc;
  ctx->tic(K_NEIGH);
  CK(cudaMemsetAsync(w.flags, 0, sizeof(int) * 8, st));

  // ---- migration (Comm::exchange)
  rc = ensure_dir(ctx, d.nlocal);
  int *cnt = c->d_counts, *cur = c->d_counts + NDX;
  CK(cudaMemsetAsync(cnt, 0, sizeof(int) * 2 * NDX, st));
  if (!rc && d.nlocal) {
    classify_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, bg, c->d_dir, cnt, w.flags);
    SPHBVF_LAUNCHED(1);
  }
  int sendcnt[ND], recvcnt[ND], sendoff[ND + 1], recvoff[ND + 1];
  if ((rc = exchange_counts(ctx, cnt, sendcnt, recvcnt, "Lost atoms: an atom moved further than the neighbouring brick", rc))) return rc;
  sendoff[0] = recvoff[0] = 0;
  for (int k = 0; k < ND; k++) {
    sendoff[k + 1] = sendoff[k] + sendcnt[k];
    recvoff[k + 1] = recvoff[k] + recvcnt[k];
  }
  const int nleave = sendoff[ND], narrive = recvoff[ND];
  rc = (nleave || narrive) ? migrate_prepare(ctx, sendoff, nleave, narrive, cur) : 0;
  if ((rc = comm_agree(ctx, rc))) return rc;
  if (nleave || narrive) {
    const int NM = 26 + S;
    const int nstay = d.nlocal - nleave;
    if ((rc = exchange_payload(ctx, sendcnt, sendoff, recvcnt, recvoff, NM, st))) return rc;
    d.nlocal = nstay;
    if (narrive) {
      unpack_arrivals_kernel<<<nblocks(narrive, 256), 256, 0, st>>>(d, S, nstay, narrive, c->recvbuf);
      SPHBVF_LAUNCHED(1);
    }
    d.nlocal = nstay + narrive;
    ctx->migrated = 1;
    CK(cudaGetLastError());
  }

  // ---- sort into cell order
  rc = rebuild_sort(ctx);

  // ---- borders: who is a ghost of which neighbour brick (frozen until the next rebuild)
  CK(cudaMemsetAsync(cnt, 0, sizeof(int) * 2 * NDX, st));
  if (!rc && d.nlocal) {
    border_kernel<false><<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, bg, ctx->cutneighmax, c->send, cnt, nullptr);
    SPHBVF_LAUNCHED(1);
  }
  if ((rc = exchange_counts(ctx, cnt, sendcnt, recvcnt, "Lost atoms: an owned atom left the cell grid of its brick", rc))) return rc;
  c->send.off[0] = c->recv.off[0] = 0;
  for (int k = 0; k < ND; k++) {
    c->send.off[k + 1] = c->send.off[k] + sendcnt[k];
    c->recv.off[k + 1] = c->recv.off[k] + recvcnt[k];
  }
  c->nsend = c->send.off[ND];
  const int nghost = c->recv.off[ND];
  rc = [&]() -> int {
    if (c->nsend > c->send_cap) {
      if (c->sendidx) cudaFree(c->sendidx);
      c->sendidx = nullptr;
      c->send_cap = c->nsend + c->nsend / 4 + 1024;
      CK(cudaMalloc((void **)&c->sendidx, sizeof(int) * (size_t)c->send_cap));
    }
    if (d.nlocal + nghost > d.nallmax) {
      int rc2;
      if ((rc2 = ctx_ensure_capacity(ctx, d.nmax, d.nlocal + nghost + nghost / 4 + 1024))) return rc2;
    }
    const int S2 = ctx->co.nspecies;
    const int R = halo_width(S2, ctx->with_dev, 1, 1);
    return ensure_buf(ctx, (size_t)std::max(c->nsend, nghost) * R + 64);
  }();
  if ((rc = comm_agree(ctx, rc))) return rc;
  d.nghost = nghost;
  if (c->nsend) {
    border_kernel<true><<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, bg, ctx->cutneighmax, c->send, cur, c->sendidx);
    SPHBVF_LAUNCHED(1);
  }
  CK(cudaMemcpyAsync(d.ptag, d.tag, sizeof(int) * (size_t)d.nlocal, cudaMemcpyDeviceToDevice, st));
  launch_pack(d, ctx->co, ctx->with_dev, st);
  if ((rc = halo(ctx, 1, 1, st))) return rc;

  rc = rebuild_finish(ctx);
  {
    // one all-reduce for the status (comm_agree) and for "some rank has no atom order": the early halo changes the
    // ORDER of this communicator's operations within a step, so either every rank uses it or none does
    int v[2] = {rc ? -rc : 0, ctx->aorder_valid ? 0 : 1};
    const int rc2 = comm_allreduce_max(ctx, v, 2);
    if (rc) return rc;
    if (rc2) return rc2;
    if (v[0]) return ctx->fail(-v[0], "another GPU of this run failed with status %d (see its message)", -v[0]);
    ctx->halo_early_ok = v[1] == 0;
  }
  ctx->toc();
  return 0;
}

void comm_destroy(sphbvf_ctx *ctx) {
  CommState *c = ctx->comm;
  if (!c) return;
  if (c->halo_st) {
    cudaStreamSynchronize(c->halo_st);
    cudaStreamDestroy(c->halo_st);
    cudaEventDestroy(c->ev_ready);
    cudaEventDestroy(c->ev_done);
  }
  if (c->comm && nccl_api()) nccl_api()->CommDestroy(c->comm);
  for (void *p : {(void *)c->sendidx, (void *)c->sendbuf, (void *)c->recvbuf, (void *)c->d_counts, (void *)c->d_dir,
                  (void *)c->d_keep, (void *)c->d_pos})
    if (p) cudaFree(p);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  delete c;
  ctx->comm = nullptr;
}

extern "C" int sphbvf_comm_unique_id(void *id128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NcclApi *api = nccl_api();
  if (!api) return SPHBVF_ECOMM;
  return api->GetUniqueId((ncclUniqueId *)id128) == ncclSuccess ? 0 : SPHBVF_ECOMM;
}

extern "C" int sphbvf_comm_init(sphbvf_ctx *ctx, const void *id128) {
  if (ctx->comm) return ctx->fail(SPHBVF_ESTATE, "sphbvf_comm_init called twice");
  if (!nccl_api()) return ctx->fail(SPHBVF_ECOMM, "libnccl.so.2 could not be loaded: %s", dlerror());
  const int P = ctx->cfg.nranks;
  if (ctx->cfg.procgrid[0] * ctx->cfg.procgrid[1] * ctx->cfg.procgrid[2] != P)
    return ctx->fail(SPHBVF_EINVAL, "procgrid %d x %d x %d does not match %d ranks", ctx->cfg.procgrid[0],
                     ctx->cfg.procgrid[1], ctx->cfg.procgrid[2], P);
  cudaSetDevice(ctx->cfg.device);
  CommState *c = new CommState();
  ctx->comm = c;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NK(nccl_api()->CommInitRank(&c->comm, P, id, ctx->cfg.rank));
  CK(cudaMalloc((void **)&c->d_counts, sizeof(int) * (2 * NDX + NDX * P + 16)));
  CK(cudaMallocHost((void **)&c->h_counts, sizeof(int) * (NDX * P + 16)));
  double shift[ND * 3];
  int rc = sphbvf_comm_plan(&ctx->cfg, ctx->cfg.rank, c->send.peer, shift);
  if (rc) return ctx->fail(rc, "sphbvf_comm_plan failed");
  for (int k = 0; k < ND; k++) {
    for (int q = 0; q < 3; q++) {
      c->send.shift[k][q] = shift[3 * k + q];
      // what arrives from direction k was sent by that peer towards -k with the opposite wrap
      c->recv.shift[k][q] = shift[3 * k + q] == 0.0 ? 0.0 : -shift[3 * k + q];
    }
    c->recv.peer[k] = c->send.peer[k];
  }
  return 0;
}

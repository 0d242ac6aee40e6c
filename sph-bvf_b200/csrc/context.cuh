// context.cuh -- the opaque sphbvf_ctx behind include/sphbvf.h
#pragma once

#include <string>
#include <vector>

#include "sphbvf_internal.cuh"

struct CommState;   // comm.cu

struct EvPair {
  int fam;
  cudaEvent_t a, b;
};

struct sphbvf_ctx {
  sphbvf_config cfg{};
  sphbvf::Coeffs co{};
  sphbvf::Box box{};
  sphbvf::Grid grid{};
  sphbvf::DevState d{};
  sphbvf::NeighWork w{};
  cudaStream_t st = nullptr;
  int pairset[sphbvf::MAXT][sphbvf::MAXT] = {};
  sphbvf::FixDesc fixes[sphbvf::MAXFIX];
  int nfix = 0;
  long ntimestep = 0, run_nsteps = 0, run_nsteps_user = -1;
  int ago = 0, nbuilds = 0, ndanger = 0, maxneigh_seen = 0;
  int atoms_set = 0, setup_done = 0, with_dev = 0, any_solid = 0, e_nonzero = 0, migrated = 0;
  double cutneighmax = 0.0, triggersq = 0.0;
  // lazy fusion of the integrators: sphbvf_final_integrate only records its arguments; if the next call that
  // touches the state is sphbvf_initial_integrate, ONE kernel does final(n) + initial(n+1) + pack, otherwise
  // flush_final() launches the plain final_integrate first (any other entry point calls it)
  int fuse = 1, final_pending = 0, pack_valid = 0;
  double pend_dt = 0.0;
  long pend_step = 0;
  int *pair_queues = nullptr;  // 2 x (nq + 1) chunk counters of the gather form's persistent schedule (two concurrent launches)
  int pair_nq = 1;             // queues = SMs of the device
  bool pair_warp = false;      // chunks are drawn per warp (32 atoms) instead of per CTA
  int flags_dirty = 0;         // bit 0: e / dev uploaded, bit 1: type / solid_tag / fixed_tag uploaded since the flags were derived
  int pair_pref = 0;           // 0: gather form (default), 1: tile form when it fits (SPHBVF_PAIR=tile)
  int smem_optin = 0;          // cudaDevAttrMaxSharedMemoryPerBlockOptin of the device
  int expanded_valid = 0;      // d.neigh holds the expansion of the current tile-form list (sphbvf_get_pairs)
  int overlap_halo = 1;        // multi-rank: per-step halo on its own stream beside the interior tiles (SPHBVF_HALO=serial: 0)
  int halo_pending = 0;        // a halo is in flight on the halo stream; ghost readers join it first
  int halo_early = 1;          // SPHBVF_HALO=early (default): start the halo from inside the fused integrator (faces first)
  int halo_early_ok = 0;       // every rank has an atom order for it (agreed in comm_rebuild)
  int halo_done_step = 0;      // the halo of the coming pair pass has been started already (by the integrator)
  int *tile_order = nullptr;   // [ntiles] tiles that cannot see a ghost first, then the others
  int *tile_cnt = nullptr, *tile_off = nullptr;   // [ntiles + 1] scratch of the atom order
  int *aorder = nullptr;       // [nmax] owned atoms in tile_order (gather form: the split pair pass indexes through it)
  int aorder_cap = 0, aorder_valid = 0, natoms_interior = 0;
  int ntiles_interior = 0, ntiles_total = 0;
  long tile_order_cap = 0;
  long tile_key[11] = {};
  int open_fam = -1;           // kernel family of the open tic()
  long launch_mark = 0;
  int random_set = 0;
  double kboltz = 0.0;
  unsigned long long seed = 0;
  long scan_cap = 0;
  int *h_flags = nullptr;      // pinned, 16 ints
  void *h_stage = nullptr;     // pinned staging (halo counts)
  double *d_virial = nullptr;  // [6]
  CommState *comm = nullptr;
  std::string err;
  // accounting
  int profiling = 0;
  long launches_fam[sphbvf::K_NFAM] = {};
  double ms_fam[sphbvf::K_NFAM] = {};
  std::vector<cudaEvent_t> ev_pool;
  std::vector<EvPair> ev_list;
  EvPair ev_open{};

  int fail(int code, const char *fmt, ...);
  void tic(int fam);
  void toc();
  void drain_events();
};

// capi.cu: rebuild pieces shared with the brick-decomposed path
int rebuild_sort(sphbvf_ctx *ctx);       // pbc + cell sort + permutation of the primary arrays
int permute_state(sphbvf_ctx *ctx, int n, bool always_dev);   // primary arrays gathered through w.perm, buffers swapped
int rebuild_finish(sphbvf_ctx *ctx);     // ghost binning + Verlet list + xhold
int ctx_ensure_capacity(sphbvf_ctx *ctx, int nmax, int nallmax);
int ctx_fetch_flags(sphbvf_ctx *ctx);    // w.flags -> h_flags[0..8), synchronises the stream
int flush_final(sphbvf_ctx *ctx);        // launch a final_integrate that sphbvf_final_integrate deferred

// comm.cu / comm_nccl.cu: brick decomposition over NCCL (one rank per GPU)
int comm_rebuild(sphbvf_ctx *ctx);        // pbc + migration + sort + borders + list
int comm_forward(sphbvf_ctx *ctx, int with_pd);   // per-step halo of the packed records (own stream when overlap_halo)
int comm_halo_join(sphbvf_ctx *ctx);      // compute stream waits for a halo in flight
cudaStream_t comm_halo_stream(sphbvf_ctx *ctx);
int comm_halo_mark(sphbvf_ctx *ctx);      // comm_halo_join also waits for what was queued on the halo stream since
int comm_face_has_peer(const sphbvf_ctx *ctx, int k, int side);   // ghosts can arrive through that face of the brick
int comm_vote(sphbvf_ctx *ctx, int *flag); // rebuild vote: max over ranks
int comm_allreduce_max(sphbvf_ctx *ctx, int *vals, int n);   // n <= 8
int comm_allreduce_max_double(sphbvf_ctx *ctx, double *val);
void comm_destroy(sphbvf_ctx *ctx);

/* sphbvf.h -- C ABI of the B200-native SPH-BVF timestep library (libsphbvf.so).
 *
 * This is the drop-in boundary for the hot path of the USER-SSA-TSDPD package of
 * briandrawert/SPH-BVF (a LAMMPS 22Aug2018 fork): the LAMMPS-style host classes registered under
 * the "/cuda" suffix (sph-bvf_b200/lammps/) call ONLY these functions, with plain pointers and
 * sizes.  Each entry point cites the reference interface it replaces (paths relative to the
 * reference's src/).  All functions return 0 on success or a negative SPHBVF_E* code and never
 * throw or exit; sphbvf_last_error() gives the message the host style passes to error->one().
 *
 * Conventions are the reference's: atom types 1..ntypes; `v` is the TRANSPORT velocity (atom->v),
 * `vest` the MOMENTUM velocity (atom->vest); host per-atom arrays are dense row-major
 * ([n][3], [n][S], [n][3][3]) exactly as LAMMPS' Memory::create lays out &x[0][0].
 * All arithmetic is FP64.  There is no CPU fallback: without a CUDA device create() fails.
 */
#ifndef SPHBVF_H
#define SPHBVF_H

#ifdef __cplusplus
extern "C" {
#endif

#define SPHBVF_VERSION 1

enum sphbvf_error {
  SPHBVF_OK = 0,
  SPHBVF_EINVAL = -1,     /* bad argument / missing coefficients (pair_...:1034-1036) */
  SPHBVF_ECUDA = -2,      /* CUDA runtime failure (message holds cudaGetErrorString) */
  SPHBVF_ENONFINITE = -3, /* "Non-numeric positions - simulation unstable" (nbin.cpp:120) */
  SPHBVF_ELOST = -4,      /* atom left a non-periodic box (thermo.cpp:436-450 "Lost atoms") */
  SPHBVF_EOVERFLOW = -5,  /* neighbour storage exhausted (npair_half_bin_atomonly_newton.cpp:114) */
  SPHBVF_ESTATE = -6,     /* call out of order (run before setup ...) */
  SPHBVF_ECOMM = -7       /* NCCL failure */
};

/* pair_style / fix variant: ssa_tsdpd/bvf/{transportVelocity,mechanics,fsi} */
enum sphbvf_variant { SPHBVF_TV = 0, SPHBVF_MECHANICS = 1, SPHBVF_FSI = 2 };

/* per-atom fields for upload/download; names follow atom.h:84-109 */
enum sphbvf_field {
  SPHBVF_F_TAG = 0, SPHBVF_F_TYPE, SPHBVF_F_MASK, SPHBVF_F_SOLID_TAG, SPHBVF_F_FIXED_TAG, /* int32 [n] */
  SPHBVF_F_X, SPHBVF_F_V, SPHBVF_F_VEST, SPHBVF_F_F,                                      /* [n][3] */
  SPHBVF_F_RHO, SPHBVF_F_RHOI, SPHBVF_F_DRHO, SPHBVF_F_E,                                 /* [n] */
  SPHBVF_F_PHI, SPHBVF_F_NUMBER_DENSITY, SPHBVF_F_NW, SPHBVF_F_DDV, SPHBVF_F_DDX,         /* nw,ddv,ddx [n][3] */
  SPHBVF_F_RHOAUX1, SPHBVF_F_RHOAUX2, SPHBVF_F_PNEW,                                      /* [n] */
  SPHBVF_F_DEV, SPHBVF_F_DDEV,                                                            /* [n][9] */
  SPHBVF_F_C, SPHBVF_F_Q,                                                                 /* [n][S] */
  SPHBVF_F_COUNT
};

typedef struct sphbvf_config {
  int dim;                 /* `dimension` 2|3 */
  int periodic[3];         /* `boundary` p=1 / f=0 */
  double boxlo[3], boxhi[3]; /* global box (domain.h boxlo/boxhi), orthogonal */
  int ntypes;              /* create_box N, <= 4 */
  int nspecies;            /* atom_style ssa_tsdpd/atomic S (atom_vec_ssa_tsdpd_atomic.cpp:58-108), <= 4 */
  int variant;             /* enum sphbvf_variant: selects pair AND integrator arithmetic */
  double skin;             /* `neighbor SKIN bin` */
  int neigh_every, neigh_delay, neigh_check; /* `neigh_modify` (neighbor.cpp:88-90: 1, 10, 1) */
  double dt;               /* `timestep` */
  int integrate_groupbit;  /* groupbit of the integrator fix (fix.h:26) */
  int device;              /* CUDA device ordinal */
  /* brick decomposition (comm.cpp:445 set_proc_grid / procmap.cpp): rank r owns brick
     (r % px, (r / px) % py, r / (px*py)); 1 1 1 for a single GPU */
  int procgrid[3];
  int rank, nranks;
} sphbvf_config;

typedef struct sphbvf_ctx sphbvf_ctx;

/* ---- lifetime ------------------------------------------------------------------------------ */
int sphbvf_version(void);
/* number of CUDA devices visible; 0 when there is no GPU (then sphbvf_create fails loudly) */
int sphbvf_device_count(void);
int sphbvf_create(const sphbvf_config *cfg, sphbvf_ctx **out);
void sphbvf_destroy(sphbvf_ctx *ctx);
const char *sphbvf_last_error(const sphbvf_ctx *ctx);

/* ---- coefficients: Pair*::coeff / init_one (pair_ssa_tsdpd_bvf_transport_velocity.cpp:967-1052)
 * `mass I m` + the per-type part of `pair_coeff I * rho0 c0 . . . G0` */
int sphbvf_set_type(sphbvf_ctx *ctx, int itype, double mass, double rho0, double c0, double G0);
/* the per-type-pair part of `pair_coeff I J . . eta h cutc . kappa[0..S-1]` (stored symmetric) */
int sphbvf_set_pair(sphbvf_ctx *ctx, int itype, int jtype, double eta, double h, double cutc,
                    const double *kappa);
/* reset_dt() of the integrator fixes (fix_ssa_tsdpd_bvf_transport_velocity.cpp:465-468) */
int sphbvf_set_dt(sphbvf_ctx *ctx, double dt);
/* the stochastic stress term (pair_ssa_tsdpd_bvf_transport_velocity.cpp:403-431), active when some
 * atom has ssa_tsdpd/e != 0: kboltz = force->boltz of the unit system, seed = any 64-bit number.
 * The reference seeds a sequential Marsaglia generator from clock() (:957-959) and draws d*d
 * Gaussians per half-list pair, so its random forces are neither reproducible nor independent of
 * the list orientation; here the symmetric traceless matrix of a pair is a pure function of
 * (seed, timestep, min tag, max tag) (Philox-4x32-10 + Box-Muller), hence identical for both
 * partners (momentum conserving without Newton mirroring), independent of decomposition and
 * reproducible; e_ij = (e_i + e_j)/2.  Not calling this leaves the term switched off. */
int sphbvf_set_random(sphbvf_ctx *ctx, double kboltz, unsigned long long seed);
/* update->ntimestep / update->nsteps as seen by the styles */
int sphbvf_set_timestep(sphbvf_ctx *ctx, long ntimestep);
int sphbvf_set_run_length(sphbvf_ctx *ctx, long nsteps);

/* ---- atoms: the owned atoms of class Atom (atom.h:49-109) -> device SoA.  C and dev may be NULL.
 * Replaces AtomVecSsaTsdpdAtomic::create_atom/data_atom defaults (:1851-2042) for the other fields. */
int sphbvf_set_atoms(sphbvf_ctx *ctx, int n, const int *tag, const int *type, const int *mask,
                     const int *solid_tag, const int *fixed_tag, const double *x, const double *v,
                     const double *rho, const double *e, const double *C, const double *dev);
/* overwrite / read one field for the owned atoms, in the order of the last set_atoms call
 * (single-rank) -- the host<->device transposition AtomVec pack/unpack does on the CPU side. */
int sphbvf_upload(sphbvf_ctx *ctx, int field, const void *host);
int sphbvf_download(sphbvf_ctx *ctx, int field, void *host);
/* multi-rank: atoms migrate, so rows come in device order; pair with SPHBVF_F_TAG */
int sphbvf_download_local(sphbvf_ctx *ctx, int field, void *host, int cap_rows);
/* overwrite one field with rows in device order, i.e. the order of the last download_local (valid
 * until the next neighbour rebuild reorders the atoms); nrows must equal sphbvf_nlocal */
int sphbvf_upload_local(sphbvf_ctx *ctx, int field, const void *host, int nrows);

/* ---- auxiliary fixes (registered once, executed inside the matching hook in the order added,
 * modify.cpp:385-480).  The /cuda fix classes may instead call the *_now entry points. */
/* fix ssa_tsdpd/buoyancy {boussinesq/sdpd|gravity} a dim k Cref (fix_ssa_tsdpd_buoyancy.cpp:28-140) */
int sphbvf_add_buoyancy(sphbvf_ctx *ctx, int groupbit, int gravity, double accel, int coord, int k,
                        double Cref);
/* fix ssa_tsdpd/forcing {tsdpd|velocity} step idx {circle cx cy R v | rectangle cx cy Lx Ly v}
 * (fix_ssa_tsdpd_forcing.cpp:40-176): kind 0 tsdpd, 1 velocity; shape 0 circle, 1 rectangle */
int sphbvf_add_forcing(sphbvf_ctx *ctx, int groupbit, int kind, long step, int idx, int shape,
                       double cx, double cy, double a, double b, double value);
/* fix ssa_tsdpd/buffer {tsdpd|velocity|density} {x|y} step idx cx cy Lx Ly v
 * (fix_ssa_tsdpd_buffer.cpp:36-240): kind 0 tsdpd, 1 velocity, 2 density; axis 0 x, 1 y */
int sphbvf_add_buffer(sphbvf_ctx *ctx, int groupbit, int kind, int axis, long step, int idx,
                      double cx, double cy, double length, double width, double value);
/* fix ssa_tsdpd/chem_rxn_mass_action k nreact r0.. nprod p0.. (fix_ssa_tsdpd_chem_rxn_mass_action.cpp:24-112):
 * post_force source  flux = k * prod_r C[r];  Q[r] -= flux;  Q[p] += flux;  nreact <= 2, nprod <= 4 */
int sphbvf_add_chem_rxn(sphbvf_ctx *ctx, int groupbit, double k_rate, int nreact, const int *reactants,
                        int nprod, const int *products);
/* fix setforce fx fy fz with constants (fix_setforce.cpp:222-290), used by the cavity decks */
int sphbvf_add_setforce(sphbvf_ctx *ctx, int groupbit, double fx, double fy, double fz);

/* ---- the timestep.  sphbvf_setup == Verlet::setup (verlet.cpp:88-170); sphbvf_run(n) executes n
 * iterations of Verlet::run (verlet.cpp:240-353) entirely on the device. */
int sphbvf_setup(sphbvf_ctx *ctx);
int sphbvf_run(sphbvf_ctx *ctx, int nsteps);
/* the neighbour part of Verlet::setup only (verlet.cpp:100-128: pbc, exchange, borders,
 * neighbor->build, ncalls = 0) for hosts that call the hooks below themselves: vest/rhoI must have
 * been uploaded, pair_compute and post_force follow through their own entry points */
int sphbvf_setup_neighbors(sphbvf_ctx *ctx);

/* the same step in the pieces the /cuda host classes are called with by Verlet/Modify */
int sphbvf_initial_integrate(sphbvf_ctx *ctx); /* Fix*::initial_integrate (fix_...transport_velocity.cpp:99) */
int sphbvf_post_integrate(sphbvf_ctx *ctx);    /* forcing/buffer post_integrate */
/* Neighbor::decide + (Comm::forward_comm | pbc/exchange/borders/Neighbor::build)
 * (verlet.cpp:258-296, neighbor.cpp:1922-2081, comm_brick.cpp:460-880); *rebuilt = 1 on a rebuild */
int sphbvf_neighbor(sphbvf_ctx *ctx, int *rebuilt);
/* Verlet::force_clear + Pair*::compute (+ reverse_comm, which the gather formulation makes a
 * no-op) (verlet.cpp:300-335, pair_ssa_tsdpd_bvf_*.cpp compute) */
int sphbvf_pair_compute(sphbvf_ctx *ctx);
/* Pair::virial_fdotr_compute (pair.cpp:1511-1560) for the forces just computed: virial[6] in LAMMPS
 * order xx yy zz xy xz yz (this rank's share); call between pair_compute and post_force */
int sphbvf_virial(sphbvf_ctx *ctx, double *virial6);
int sphbvf_post_force(sphbvf_ctx *ctx);        /* buoyancy / setforce / chem_rxn post_force */
/* Modify::setup -> Fix::setup at the start of a run: only the fixes whose setup() forwards to
 * post_force (setforce, buoyancy) act; follows the first pair_compute */
int sphbvf_setup_post_force(sphbvf_ctx *ctx);
/* Fix*::final_integrate (:244).  The launch is deferred: if the next call on this context is
 * sphbvf_initial_integrate (nothing scheduled between two steps, as in Verlet::run without output or
 * end_of_step fixes) ONE kernel runs final_integrate(n) + initial_integrate(n+1) and writes the pair-input
 * records, bit-identical to the separate kernels; any other entry point first runs the plain
 * final_integrate, so results are never observed out of order.  SPHBVF_NO_FUSE=1 disables the deferral. */
int sphbvf_final_integrate(sphbvf_ctx *ctx);
int sphbvf_end_of_step(sphbvf_ctx *ctx);       /* buffer(density) end_of_step */
/* force a neighbour rebuild now (pbc + ghosts + sort + list), as Neighbor::build on a rebuild step */
int sphbvf_build_neighbors(sphbvf_ctx *ctx);

/* ---- introspection */
int sphbvf_nlocal(const sphbvf_ctx *ctx);
int sphbvf_nghost(const sphbvf_ctx *ctx);
long sphbvf_ntimestep(const sphbvf_ctx *ctx);
int sphbvf_nbuilds(const sphbvf_ctx *ctx);    /* Neighbor::ncalls */
/* how the current neighbour list is stored and traversed: 1 = tile form (16-bit entries, the candidates of a
 * 4x4x4-cell tile staged in shared memory, default), 0 = gather form (32-bit entries, records gathered through L1;
 * chosen at a rebuild when a tile's candidates do not fit, or with SPHBVF_PAIR=gather).  Same pair sets and the same
 * arithmetic per pair either way; the summation order differs. */
int sphbvf_pair_mode(const sphbvf_ctx *ctx);
int sphbvf_ndanger(const sphbvf_ctx *ctx);    /* Neighbor::ndanger */
/* max over the atoms of `groupbit` (all ranks) of |v|^2 with v the transport velocity (atom->v): the
 * reduction FixDtAdaptive::end_of_step does before its MPI_Allreduce (fix_dt_adaptive.cpp:118-148) */
int sphbvf_max_vsq(sphbvf_ctx *ctx, int groupbit, double *max_vsq);
/* sum over the owned atoms of `groupbit` of m v_a v_b (v = atom->v, m = mass[type]), order xx yy zz xy xz yz: the
 * accumulation of ComputeTemp::compute_scalar / compute_vector (compute_temp.cpp:78-135) before its MPI_Allreduce
 * and unit factors; this rank's share; summed in a fixed order (reproducible) */
int sphbvf_ke_tensor(sphbvf_ctx *ctx, int groupbit, double *ke6);
/* the neighbour structure as unordered (tag_i, tag_j) rows, each pair once -- what
 * `compute property/local patom1 patom2` dumps for the reference; out=NULL sizes */
long sphbvf_get_pairs(sphbvf_ctx *ctx, int *out, long cap);
int sphbvf_sync(sphbvf_ctx *ctx);
/* kernel launches issued by this context since creation / accumulated device time of one kernel
 * family in ms (CUDA events on the context's stream; enable with sphbvf_set_profiling) */
long sphbvf_launch_count(const sphbvf_ctx *ctx);
int sphbvf_set_profiling(sphbvf_ctx *ctx, int on);
/* which: 0 pair, 1 initial_integrate, 2 final_integrate, 3 neighbor rebuild, 4 pack/halo, 5 fixes,
 * 6 fused final_integrate(n) + initial_integrate(n+1) + pack (see sphbvf_final_integrate) */
double sphbvf_kernel_ms(const sphbvf_ctx *ctx, int which, long *launches);
void *sphbvf_stream(sphbvf_ctx *ctx);         /* the cudaStream_t all work is queued on */

/* ---- multi-GPU: one process per GPU, brick decomposition, NCCL halo + migration.
 * Replaces CommBrick::{forward_comm,exchange,borders} + MPI (comm_brick.cpp:460-880). */
int sphbvf_comm_unique_id(void *id128);       /* ncclGetUniqueId; rank 0 broadcasts the 128 bytes */
int sphbvf_comm_init(sphbvf_ctx *ctx, const void *id128);
/* host-side brick plan, usable without a GPU: sub-domain of `rank` in `procgrid` */
int sphbvf_brick_bounds(const sphbvf_config *cfg, int rank, double sublo[3], double subhi[3]);
/* host-side halo plan, usable without a GPU (CommBrick::setup, comm_brick.cpp:161-410): for each
 * of the 27 directions code = (dx+1) + 3 (dy+1) + 9 (dz+1) the rank owning the adjacent brick
 * (-1 = none) and the periodic shift [27][3] added to positions sent in that direction */
int sphbvf_comm_plan(const sphbvf_config *cfg, int rank, int *peer, double *shift);
/* procmap.cpp-style factorisation of nranks into a grid minimising brick surface */
int sphbvf_proc_grid(int nranks, int dim, const double prd[3], int grid[3]);

#ifdef __cplusplus
}
#endif
#endif

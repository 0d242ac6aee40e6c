/* oracle/sphbvf_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain-C, single-threaded CPU restatement of the SPH-BVF timestep hot path of the reference
 * (briandrawert/SPH-BVF, src/USER-SSA-TSDPD + the LAMMPS core slices it traverses).  It exists
 * only to CHECK the CUDA library (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).
 * Nothing under sph-bvf_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors;
 * this restatement is pinned against outputs of the reference itself, compiled unmodified by
 * oracle/Makefile and run through oracle/ref_harness.cpp (fixtures under tests/golden/, made by
 * tests/golden/make_golden.py; see tests/test_oracle_vs_reference.py).
 *
 * Conventions follow the reference: atom types are 1..ntypes, `v` is the TRANSPORT velocity
 * (atom->v) and `vest` the MOMENTUM velocity (atom->vest), tensors are row-major [3][3].
 */
#ifndef SPHBVF_ORACLE_H
#define SPHBVF_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_TV = 0, ORC_MECHANICS = 1, ORC_FSI = 2 };

typedef struct {
  int dim;             /* 2 or 3 */
  int periodic[3];
  double boxlo[3], boxhi[3];
  int ntypes;
  int nspecies;        /* num_sdpd_species */
  int variant;         /* ORC_TV / ORC_MECHANICS / ORC_FSI (pair AND fix variant) */
  double skin;         /* neighbor skin */
  int every, delay, check; /* neigh_modify (defaults 1, 10, 1: neighbor.cpp:88-90) */
  double dt;
  int integrate_groupbit;  /* group bit of the integrator fix (1 = all) */
} orc_config;

typedef struct orc_ctx orc_ctx;

orc_ctx *orc_create(const orc_config *cfg);
void orc_destroy(orc_ctx *c);

/* mass I m ; pair_coeff I I rho0 c0 .. .. .. G0  (per-type part, pair_...:1000-1004) */
int orc_set_type(orc_ctx *c, int itype, double mass, double rho0, double c0, double G0);
/* pair_coeff I J . . eta h cutc . kappa[0..S-1]  (per-pair part, symmetric: init_one :1032) */
int orc_set_pair(orc_ctx *c, int itype, int jtype, double eta, double h, double cutc,
                 const double *kappa);

/* host arrays are dense [n][3] / [n][S] / [n][9]; C and dev may be NULL (zeros). */
int orc_set_atoms(orc_ctx *c, int n, const int *tag, const int *type, const int *mask,
                  const int *solid_tag, const int *fixed_tag, const double *x, const double *v,
                  const double *rho, const double *e, const double *C, const double *dev);

/* auxiliary fixes, executed in the order added within each hook (modify.cpp:385-480) */
int orc_add_buoyancy(orc_ctx *c, int groupbit, int gravity, double accel, int coord, int k,
                     double Cref);
/* kind: 0 tsdpd (C[idx]=value), 1 velocity (vest[idx]=value); shape: 0 circle(cx,cy,R), 1 rectangle */
int orc_add_forcing(orc_ctx *c, int groupbit, int kind, long step, int idx, int shape, double cx,
                    double cy, double a, double b, double value);
/* kind: 0 tsdpd, 1 velocity, 2 density; axis 0 x / 1 y */
int orc_add_buffer(orc_ctx *c, int groupbit, int kind, int axis, long step, int idx, double cx,
                   double cy, double length, double width, double value);
/* LAMMPS core fix setforce with three constants (fix_setforce.cpp post_force) */
int orc_add_setforce(orc_ctx *c, int groupbit, double fx, double fy, double fz);
/* fix ssa_tsdpd/chem_rxn_mass_action (fix_ssa_tsdpd_chem_rxn_mass_action.cpp:24-112): post_force source term on Q */
int orc_add_chem_rxn(orc_ctx *c, int groupbit, double k_rate, int nreact, const int *reactants, int nprod,
                     const int *products);

int orc_setup(orc_ctx *c);              /* Verlet::setup  (verlet.cpp:88-170) */
int orc_run(orc_ctx *c, int nsteps);    /* Verlet::run    (verlet.cpp:223-354), setup must precede */

/* update->nsteps of the enclosing `run N` command when the caller splits it into several
 * orc_run() calls: the fsi pair style's density-diffusion switch compares ntimestep*dt with
 * dt*nsteps (pair_ssa_tsdpd_bvf_fsi.cpp:531-539).  Default: the argument of each orc_run. */
void orc_set_run_length(orc_ctx *c, long nsteps);

/* 0 (default): ghosts are created before setup_pre_force as in Verlet::setup, i.e. with stale
 * vest/rhoI at step 0.  1: setup_pre_force first (what the CUDA library does). */
void orc_set_consistent_ghosts(orc_ctx *c, int on);
/* 0 (default): TV pressure switch on FREE solids as the reference's half list applies it
 * (orientation dependent, pair_...transport_velocity.cpp:606 vs :633).  1: orientation-free form. */
void orc_set_symmetric_switch(orc_ctx *c, int on);

/* single pieces, for kernel-level tests */
int orc_build_neighbors(orc_ctx *c);    /* pbc + ghosts + bins + list, as on a rebuild step */
int orc_pair_compute(orc_ctx *c);       /* force_clear + PairSsaTsdpdBvf*::compute */

int orc_nlocal(const orc_ctx *c);
int orc_nghost(const orc_ctx *c);
long orc_ntimestep(const orc_ctx *c);
int orc_nbuilds(const orc_ctx *c);
/* copy a per-atom field of the nlocal owned atoms (input order); returns #columns or <0 */
int orc_get(const orc_ctx *c, const char *name, double *out);
int orc_get_int(const orc_ctx *c, const char *name, int *out);
/* neighbour list as (tag_i, tag_j) rows; returns #pairs (call with out=NULL to size) */
long orc_get_pairs(const orc_ctx *c, int *out, long cap);
const char *orc_last_error(const orc_ctx *c);

#ifdef __cplusplus
}
#endif
#endif

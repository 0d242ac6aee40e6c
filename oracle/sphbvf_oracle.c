/* oracle/sphbvf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT (see sphbvf_oracle.h).
 *
 * Single-threaded CPU restatement of the reference's timestep hot path.  Every function cites
 * the reference file:line it restates (paths relative to /root/reference/src; the package files
 * live in USER-SSA-TSDPD/ with identical copies in src/).  Loop structure follows the reference
 * (half neighbour list + Newton mirror for TV/mechanics, full list for fsi) on purpose: the CUDA
 * product uses a per-particle full-neighbour gather, so comparing the two also tests the
 * gather-equivalence argument of SURVEY.md A.8.
 *
 * Deliberate omissions (all exactly zero or never consumed in the reference's decks):
 *  - random stress term (pair_...transport_velocity.cpp:407-431): requires e==0 (orc_setup errors
 *    otherwise); the reference's RNG seed is srand(clock()) (:957-959), i.e. unreproducible;
 *  - sweep C v_weighted_solid/a_weighted_solid (:815-906): outputs never consumed;
 *  - de (:558-559) never integrated; SSA species (serial-only upstream, decks set 0).
 * Compiled with -ffp-contract=off so that rsq for the neighbour criterion is evaluated exactly
 * as the x86-64 reference build evaluates it (mul, mul, add, mul, add; no FMA).
 */
#include "sphbvf_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXT 8   /* max atom types + 1 */
#define MAXS 4   /* max species */
#define MAXFIX 16
#define BIG 1.0e20
#define SMALL 1.0e-6          /* nbin_standard.cpp:29 */
#define CUT2BIN_RATIO 100     /* nbin_standard.cpp:30 */

enum { FIX_BUOYANCY, FIX_FORCING, FIX_BUFFER, FIX_SETFORCE, FIX_CHEMRXN };

typedef struct {
  int kind, groupbit;
  int a_int[4];
  long step;
  double a[6];
} orc_fix;

struct orc_ctx {
  orc_config cfg;
  double prd[3];
  /* per-type / per-pair coefficients (1-based) */
  double mass[MAXT], rho0[MAXT], c0[MAXT], B[MAXT], G0[MAXT];
  double eta[MAXT][MAXT], cut[MAXT][MAXT], cutsq[MAXT][MAXT], cutc[MAXT][MAXT];
  double kappa[MAXT][MAXT][MAXS];
  int pairset[MAXT][MAXT];
  double cutneighsq[MAXT][MAXT], cutneighmax, triggersq;
  /* atoms: [0,nlocal) owned, [nlocal,nlocal+nghost) ghosts */
  int nlocal, nghost, nmax;
  int *tag, *type, *mask, *solid, *fixed;
  double *x, *v, *vest, *f;              /* [n][3] */
  double *rho, *rhoI, *drho, *e;
  double *phi, *nd, *nw, *ddv, *ddx;     /* nw,ddv,ddx [n][3] */
  double *rhoAux1, *rhoAux2, *Pnew;
  double *dev, *ddev, *art;              /* [n][9] */
  double *C, *Q;                         /* [n][S] */
  double *xhold;
  /* ghosts: owner index and shift, in creation order (comm_brick sendlist replay) */
  int *gowner;
  double *gshift;                        /* [nghost][3] */
  /* bins (nbin_standard) */
  int nbinx, nbiny, nbinz, mbinx, mbiny, mbinz, mbinxlo, mbinylo, mbinzlo, mbins;
  double binsizex, binsizey, binsizez, bininvx, bininvy, bininvz;
  int *binhead, *bins, *atom2bin;
  int nstencil, *stencil;
  /* neighbour list */
  int *numneigh;
  long *firstneigh;
  int *neigh;
  long neighcap;
  int ago, nbuilds, ndanger;
  long ntimestep, run_nsteps, run_nsteps_user;
  int nfix;
  orc_fix fix[MAXFIX];
  int setup_done;
  int consistent_ghosts;
  int symmetric_switch;
  char err[256];
};

static int fail(orc_ctx *c, const char *msg) {
  snprintf(c->err, sizeof c->err, "%s", msg);
  return -1;
}
const char *orc_last_error(const orc_ctx *c) { return c->err; }

orc_ctx *orc_create(const orc_config *cfg) {
  orc_ctx *c = (orc_ctx *)calloc(1, sizeof *c);
  if (!c) return NULL;
  c->cfg = *cfg;
  if (cfg->ntypes + 1 > MAXT || cfg->nspecies > MAXS || (cfg->dim != 2 && cfg->dim != 3)) {
    free(c);
    return NULL;
  }
  for (int d = 0; d < 3; d++) c->prd[d] = cfg->boxhi[d] - cfg->boxlo[d];
  c->run_nsteps_user = -1;
  return c;
}

static void free_atoms(orc_ctx *c) {
  free(c->tag); free(c->type); free(c->mask); free(c->solid); free(c->fixed);
  free(c->x); free(c->v); free(c->vest); free(c->f);
  free(c->rho); free(c->rhoI); free(c->drho); free(c->e);
  free(c->phi); free(c->nd); free(c->nw); free(c->ddv); free(c->ddx);
  free(c->rhoAux1); free(c->rhoAux2); free(c->Pnew);
  free(c->dev); free(c->ddev); free(c->art); free(c->C); free(c->Q);
  free(c->xhold); free(c->gowner); free(c->gshift);
  free(c->bins); free(c->atom2bin); free(c->numneigh); free(c->firstneigh);
}

void orc_destroy(orc_ctx *c) {
  if (!c) return;
  free_atoms(c);
  free(c->binhead); free(c->stencil); free(c->neigh);
  free(c);
}

int orc_set_type(orc_ctx *c, int t, double mass, double rho0, double c0, double G0) {
  if (t < 1 || t > c->cfg.ntypes) return fail(c, "type out of range");
  c->mass[t] = mass;
  c->rho0[t] = rho0;
  c->c0[t] = c0;
  c->B[t] = c0 * c0 * rho0 / 7.0; /* pair_...transport_velocity.cpp:981 */
  c->G0[t] = G0;
  return 0;
}

int orc_set_pair(orc_ctx *c, int i, int j, double eta, double h, double cutc, const double *kappa) {
  if (i < 1 || j < 1 || i > c->cfg.ntypes || j > c->cfg.ntypes) return fail(c, "type out of range");
  /* init_one (:1032-1052) mirrors [i][j] into [j][i]; Pair::init sets cutsq = cut*cut (pair.cpp:245) */
  int a[2] = {i, j}, b[2] = {j, i};
  for (int s = 0; s < 2; s++) {
    c->eta[a[s]][b[s]] = eta;
    c->cut[a[s]][b[s]] = h;
    c->cutsq[a[s]][b[s]] = h * h;
    c->cutc[a[s]][b[s]] = cutc;
    for (int k = 0; k < c->cfg.nspecies; k++) c->kappa[a[s]][b[s]][k] = kappa ? kappa[k] : 0.0;
    c->pairset[a[s]][b[s]] = 1;
  }
  return 0;
}

#define GROW(p, n, w) p = realloc(p, (size_t)(n) * (w) * sizeof *(p))
static void grow(orc_ctx *c, int nmax) {
  if (nmax <= c->nmax) return;
  int S = c->cfg.nspecies > 0 ? c->cfg.nspecies : 1;
  GROW(c->tag, nmax, 1); GROW(c->type, nmax, 1); GROW(c->mask, nmax, 1);
  GROW(c->solid, nmax, 1); GROW(c->fixed, nmax, 1);
  GROW(c->x, nmax, 3); GROW(c->v, nmax, 3); GROW(c->vest, nmax, 3); GROW(c->f, nmax, 3);
  GROW(c->rho, nmax, 1); GROW(c->rhoI, nmax, 1); GROW(c->drho, nmax, 1); GROW(c->e, nmax, 1);
  GROW(c->phi, nmax, 1); GROW(c->nd, nmax, 1); GROW(c->nw, nmax, 3); GROW(c->ddv, nmax, 3);
  GROW(c->ddx, nmax, 3); GROW(c->rhoAux1, nmax, 1); GROW(c->rhoAux2, nmax, 1);
  GROW(c->Pnew, nmax, 1); GROW(c->dev, nmax, 9); GROW(c->ddev, nmax, 9); GROW(c->art, nmax, 9);
  GROW(c->C, nmax, S); GROW(c->Q, nmax, S); GROW(c->xhold, nmax, 3);
  GROW(c->gowner, nmax, 1); GROW(c->gshift, nmax, 3);
  GROW(c->bins, nmax, 1); GROW(c->atom2bin, nmax, 1);
  GROW(c->numneigh, nmax, 1); GROW(c->firstneigh, nmax, 1);
  c->nmax = nmax;
}

int orc_set_atoms(orc_ctx *c, int n, const int *tag, const int *type, const int *mask,
                  const int *solid, const int *fixed, const double *x, const double *v,
                  const double *rho, const double *e, const double *C, const double *dev) {
  int S = c->cfg.nspecies;
  grow(c, n + n / 4 + 1024);
  c->nlocal = n;
  c->nghost = 0;
  for (int i = 0; i < n; i++) {
    c->tag[i] = tag[i];
    c->type[i] = type[i];
    c->mask[i] = mask ? mask[i] : 1;
    c->solid[i] = solid[i];
    c->fixed[i] = fixed[i];
    if (type[i] < 1 || type[i] > c->cfg.ntypes) return fail(c, "atom type out of range");
    for (int d = 0; d < 3; d++) {
      c->x[3 * i + d] = x[3 * i + d];
      c->v[3 * i + d] = v ? v[3 * i + d] : 0.0;
      c->vest[3 * i + d] = 0.0; /* create_atom default (atom_vec...:1873-1875) */
      c->f[3 * i + d] = c->nw[3 * i + d] = c->ddv[3 * i + d] = c->ddx[3 * i + d] = 0.0;
    }
    c->rho[i] = rho[i];
    c->rhoI[i] = 0.0;           /* atom_vec...:1936 */
    c->e[i] = e ? e[i] : 0.0;
    c->drho[i] = c->phi[i] = c->nd[i] = c->rhoAux1[i] = c->rhoAux2[i] = c->Pnew[i] = 0.0;
    for (int k = 0; k < 9; k++) {
      c->dev[9 * i + k] = dev ? dev[9 * i + k] : 0.0;
      c->ddev[9 * i + k] = c->art[9 * i + k] = 0.0;
    }
    for (int k = 0; k < S; k++) {
      c->C[S * i + k] = C ? C[S * i + k] : 0.0;
      c->Q[S * i + k] = 0.0;
    }
  }
  c->setup_done = 0;
  return 0;
}

static int add_fix(orc_ctx *c, orc_fix *f) {
  if (c->nfix == MAXFIX) return fail(c, "too many fixes");
  c->fix[c->nfix++] = *f;
  return 0;
}
int orc_add_buoyancy(orc_ctx *c, int groupbit, int gravity, double accel, int coord, int k, double Cref) {
  orc_fix f = {FIX_BUOYANCY, groupbit, {gravity, coord, k, 0}, 0, {accel, Cref, 0, 0, 0, 0}};
  return add_fix(c, &f);
}
int orc_add_forcing(orc_ctx *c, int groupbit, int kind, long step, int idx, int shape, double cx,
                    double cy, double a, double b, double value) {
  orc_fix f = {FIX_FORCING, groupbit, {kind, idx, shape, 0}, step, {cx, cy, a, b, value, 0}};
  return add_fix(c, &f);
}
int orc_add_buffer(orc_ctx *c, int groupbit, int kind, int axis, long step, int idx, double cx,
                   double cy, double length, double width, double value) {
  orc_fix f = {FIX_BUFFER, groupbit, {kind, idx, axis, 0}, step, {cx, cy, length, width, value, 0}};
  return add_fix(c, &f);
}
int orc_add_setforce(orc_ctx *c, int groupbit, double fx, double fy, double fz) {
  orc_fix f = {FIX_SETFORCE, groupbit, {0, 0, 0, 0}, 0, {fx, fy, fz, 0, 0, 0}};
  return add_fix(c, &f);
}

/* fix ssa_tsdpd/chem_rxn_mass_action k nr r.. np p.. (fix_ssa_tsdpd_chem_rxn_mass_action.cpp:24-54):
 * a_int = {nreact, nprod, reactants packed one byte each, products packed one byte each} */
int orc_add_chem_rxn(orc_ctx *c, int groupbit, double k_rate, int nreact, const int *reactants, int nprod,
                     const int *products) {
  if (nreact < 0 || nreact > 2 || nprod < 0 || nprod > 4) return fail(c, "Illegal fix ssa_tsdpd_chem_rxn_mass_action command");
  int r = 0, p = 0;
  for (int j = 0; j < nreact; j++) {
    if (reactants[j] < 0 || reactants[j] >= c->cfg.nspecies) return fail(c, "chem_rxn: reactant species out of range");
    r |= reactants[j] << (8 * j);
  }
  for (int j = 0; j < nprod; j++) {
    if (products[j] < 0 || products[j] >= c->cfg.nspecies) return fail(c, "chem_rxn: product species out of range");
    p |= products[j] << (8 * j);
  }
  orc_fix f = {FIX_CHEMRXN, groupbit, {nreact, nprod, r, p}, 0, {k_rate, 0, 0, 0, 0, 0}};
  return add_fix(c, &f);
}

/* ------------------------------------------------------------------------------------------
 * Neighbor::init cutoffs (neighbor.cpp:278-310)
 * ---------------------------------------------------------------------------------------- */
static int init_cutoffs(orc_ctx *c) {
  double skin = c->cfg.skin;
  c->triggersq = 0.25 * skin * skin;
  c->cutneighmax = 0.0;
  for (int i = 1; i <= c->cfg.ntypes; i++)
    for (int j = 1; j <= c->cfg.ntypes; j++) {
      if (!c->pairset[i][j]) return fail(c, "Not all pair ssa_tsdpd/bvf coeffs are set");
      double cutoff = sqrt(c->cutsq[i][j]);
      double delta = cutoff > 0.0 ? skin : 0.0;
      double cut = cutoff + delta;
      c->cutneighsq[i][j] = cut * cut;
      if (cut > c->cutneighmax) c->cutneighmax = cut;
    }
  return 0;
}

/* Domain::pbc (domain.cpp:498-600), orthogonal box, owned atoms */
static void domain_pbc(orc_ctx *c) {
  for (int i = 0; i < c->nlocal; i++)
    for (int d = 0; d < 3; d++) {
      if (!c->cfg.periodic[d]) continue;
      double lo = c->cfg.boxlo[d], hi = c->cfg.boxhi[d];
      double *xi = &c->x[3 * i + d];
      if (*xi < lo) *xi += c->prd[d];
      if (*xi >= hi) {
        *xi -= c->prd[d];
        if (*xi < lo) *xi = lo;
      }
    }
}

/* copy the forward-communicated fields of an owner into its ghost (pack_comm/unpack_comm,
 * atom_vec_ssa_tsdpd_atomic.cpp:426-549, 679-737): x(+shift) v rho e vest C dev rhoI */
static void ghost_forward(orc_ctx *c, int g) {
  int S = c->cfg.nspecies;
  int o = c->gowner[g - c->nlocal];
  const double *sh = &c->gshift[3 * (g - c->nlocal)];
  for (int d = 0; d < 3; d++) {
    c->x[3 * g + d] = c->x[3 * o + d] + sh[d];
    c->v[3 * g + d] = c->v[3 * o + d];
    c->vest[3 * g + d] = c->vest[3 * o + d];
  }
  c->rho[g] = c->rho[o];
  c->e[g] = c->e[o];
  c->rhoI[g] = c->rhoI[o];
  for (int k = 0; k < S; k++) c->C[S * g + k] = c->C[S * o + k];
  for (int k = 0; k < 9; k++) c->dev[9 * g + k] = c->dev[9 * o + k];
}

/* CommBrick::borders for one rank (comm_brick.cpp:709-880 with setup :161-410): per dimension,
 * first the low-face slab is imaged by +prd, then the high-face slab by -prd; later dimensions
 * also image the ghosts made by earlier ones.  Non-periodic dims make no ghosts (sendneed=0 for
 * procgrid=1, :248-262); dimension 2 makes none in z (:246). */
static void borders(orc_ctx *c) {
  c->nghost = 0;
  double cutghost = c->cutneighmax;
  for (int dim = 0; dim < 3; dim++) {
    if (!c->cfg.periodic[dim]) continue;
    if (c->cfg.dim == 2 && dim == 2) continue;
    double sublo = c->cfg.boxlo[dim], subhi = c->cfg.boxhi[dim];
    /* maxneed (:241) ; multiple images are needed only when the box is thinner than cutghost */
    int maxneed = (int)(cutghost * 1 / c->prd[dim]) + 1;
    int nfirst = 0, nlast = 0;
    for (int ineed = 0; ineed < 2 * maxneed; ineed++) {
      double lo, hi, shift;
      if (ineed % 2 == 0) {
        lo = ineed < 2 ? -BIG : 0.5 * (sublo + subhi);
        hi = sublo + cutghost;
        shift = c->prd[dim];
        nfirst = nlast;
        nlast = c->nlocal + c->nghost;
      } else {
        lo = subhi - cutghost;
        hi = ineed < 2 ? BIG : 0.5 * (sublo + subhi);
        shift = -c->prd[dim];
      }
      for (int i = nfirst; i < nlast; i++) {
        double xd = c->x[3 * i + dim];
        if (xd >= lo && xd <= hi) {
          int g = c->nlocal + c->nghost;
          if (g + 1 >= c->nmax) grow(c, c->nmax + c->nmax / 2);
          int gi = c->nghost++;
          int owner = i;
          double sh[3] = {0, 0, 0};
          if (i >= c->nlocal) { /* image of a ghost: accumulate shifts, keep the real owner */
            owner = c->gowner[i - c->nlocal];
            for (int d = 0; d < 3; d++) sh[d] = c->gshift[3 * (i - c->nlocal) + d];
          }
          sh[dim] += shift;
          c->gowner[gi] = owner;
          for (int d = 0; d < 3; d++) c->gshift[3 * gi + d] = sh[d];
          /* pack_border/unpack_border (atom_vec...:936-1075, 1296-1368) */
          c->tag[g] = c->tag[owner];
          c->type[g] = c->type[owner];
          c->mask[g] = c->mask[owner];
          c->solid[g] = c->solid[owner];
          c->fixed[g] = c->fixed[owner];
          ghost_forward(c, g);
          /* x of an image of an image is (x_owner + s1) + s2 in the reference; with at most one
             shift per dimension the sum over distinct dimensions is exact either way */
        }
      }
    }
  }
}

/* NBinStandard::setup_bins (nbin_standard.cpp:53-186), orthogonal, single rank */
static int setup_bins(orc_ctx *c) {
  const double *blo = c->cfg.boxlo, *bhi = c->cfg.boxhi;
  double bbox[3], bsublo[3], bsubhi[3];
  for (int d = 0; d < 3; d++) {
    /* comm->cutghost is cutneighmax in every dimension (comm_brick.cpp:170-178) */
    bsublo[d] = blo[d] - c->cutneighmax;
    bsubhi[d] = bhi[d] + c->cutneighmax;
    bbox[d] = bhi[d] - blo[d];
  }
  double binsize_optimal = 0.5 * c->cutneighmax;
  if (binsize_optimal == 0.0) binsize_optimal = bbox[0];
  double binsizeinv = 1.0 / binsize_optimal;
  c->nbinx = (int)(bbox[0] * binsizeinv);
  c->nbiny = (int)(bbox[1] * binsizeinv);
  c->nbinz = c->cfg.dim == 3 ? (int)(bbox[2] * binsizeinv) : 1;
  if (c->nbinx == 0) c->nbinx = 1;
  if (c->nbiny == 0) c->nbiny = 1;
  if (c->nbinz == 0) c->nbinz = 1;
  c->binsizex = bbox[0] / c->nbinx;
  c->binsizey = bbox[1] / c->nbiny;
  c->binsizez = bbox[2] / c->nbinz;
  c->bininvx = 1.0 / c->binsizex;
  c->bininvy = 1.0 / c->binsizey;
  c->bininvz = 1.0 / c->binsizez;
  if (binsize_optimal * c->bininvx > CUT2BIN_RATIO || binsize_optimal * c->bininvy > CUT2BIN_RATIO ||
      binsize_optimal * c->bininvz > CUT2BIN_RATIO)
    return fail(c, "Cannot use neighbor bins - box size << cutoff");
  int mhi[3], mlo[3];
  double inv[3] = {c->bininvx, c->bininvy, c->bininvz};
  for (int d = 0; d < 3; d++) {
    double coord = bsublo[d] - SMALL * bbox[d];
    mlo[d] = (int)((coord - blo[d]) * inv[d]);
    if (coord < blo[d]) mlo[d] -= 1;
    coord = bsubhi[d] + SMALL * bbox[d];
    mhi[d] = (int)((coord - blo[d]) * inv[d]);
    mlo[d] -= 1;
    mhi[d] += 1;
  }
  if (c->cfg.dim == 2) mlo[2] = mhi[2] = 0;
  c->mbinxlo = mlo[0]; c->mbinylo = mlo[1]; c->mbinzlo = mlo[2];
  c->mbinx = mhi[0] - mlo[0] + 1;
  c->mbiny = mhi[1] - mlo[1] + 1;
  c->mbinz = mhi[2] - mlo[2] + 1;
  double nb = (double)c->mbinx * c->mbiny * c->mbinz + 1;
  if (nb > 2147483647.0) return fail(c, "Too many neighbor bins");
  c->mbins = (int)nb;
  c->binhead = realloc(c->binhead, sizeof(int) * c->mbins);

  /* NStencil::create_setup (nstencil.cpp:145-158) + half/full create (nstencil_*_bin_*.cpp:28-40) */
  int sx = (int)(c->cutneighmax * c->bininvx);
  if (sx * c->binsizex < c->cutneighmax) sx++;
  int sy = (int)(c->cutneighmax * c->bininvy);
  if (sy * c->binsizey < c->cutneighmax) sy++;
  int sz = (int)(c->cutneighmax * c->bininvz);
  if (sz * c->binsizez < c->cutneighmax) sz++;
  if (c->cfg.dim == 2) sz = 0;
  c->stencil = realloc(c->stencil, sizeof(int) * (2 * sx + 1) * (2 * sy + 1) * (2 * sz + 1));
  c->nstencil = 0;
  double cutmaxsq = c->cutneighmax * c->cutneighmax;
  int full = c->cfg.variant == ORC_FSI; /* pair_ssa_tsdpd_bvf_fsi.cpp:72-77 requests a full list */
  for (int k = -sz; k <= sz; k++)
    for (int j = -sy; j <= sy; j++)
      for (int i = -sx; i <= sx; i++) {
        if (!full) {
          if (c->cfg.dim == 3) { if (!(k > 0 || (k == 0 && (j > 0 || (j == 0 && i > 0))))) continue; }
          else { if (!(j > 0 || (j == 0 && i > 0))) continue; }
        }
        /* NStencil::bin_distance (nstencil.cpp:204-228) */
        double delx = i > 0 ? (i - 1) * c->binsizex : (i == 0 ? 0.0 : (i + 1) * c->binsizex);
        double dely = j > 0 ? (j - 1) * c->binsizey : (j == 0 ? 0.0 : (j + 1) * c->binsizey);
        double delz = k > 0 ? (k - 1) * c->binsizez : (k == 0 ? 0.0 : (k + 1) * c->binsizez);
        if (delx * delx + dely * dely + delz * delz < cutmaxsq)
          c->stencil[c->nstencil++] = k * c->mbiny * c->mbinx + j * c->mbinx + i;
      }
  return 0;
}

/* NBin::coord2bin (nbin.cpp:116-148) */
static int coord2bin(const orc_ctx *c, const double *x, int *bad) {
  int ix, iy, iz;
  const double *lo = c->cfg.boxlo, *hi = c->cfg.boxhi;
  if (!isfinite(x[0]) || !isfinite(x[1]) || !isfinite(x[2])) { *bad = 1; return 0; }
  if (x[0] >= hi[0]) ix = (int)((x[0] - hi[0]) * c->bininvx) + c->nbinx;
  else if (x[0] >= lo[0]) { ix = (int)((x[0] - lo[0]) * c->bininvx); if (ix > c->nbinx - 1) ix = c->nbinx - 1; }
  else ix = (int)((x[0] - lo[0]) * c->bininvx) - 1;
  if (x[1] >= hi[1]) iy = (int)((x[1] - hi[1]) * c->bininvy) + c->nbiny;
  else if (x[1] >= lo[1]) { iy = (int)((x[1] - lo[1]) * c->bininvy); if (iy > c->nbiny - 1) iy = c->nbiny - 1; }
  else iy = (int)((x[1] - lo[1]) * c->bininvy) - 1;
  if (x[2] >= hi[2]) iz = (int)((x[2] - hi[2]) * c->bininvz) + c->nbinz;
  else if (x[2] >= lo[2]) { iz = (int)((x[2] - lo[2]) * c->bininvz); if (iz > c->nbinz - 1) iz = c->nbinz - 1; }
  else iz = (int)((x[2] - lo[2]) * c->bininvz) - 1;
  ix -= c->mbinxlo; iy -= c->mbinylo; iz -= c->mbinzlo;
  if (ix < 0 || iy < 0 || iz < 0 || ix >= c->mbinx || iy >= c->mbiny || iz >= c->mbinz) { *bad = 2; return 0; }
  return iz * c->mbiny * c->mbinx + iy * c->mbinx + ix;
}

static void push_neigh(orc_ctx *c, long *n, int j) {
  if (*n >= c->neighcap) {
    c->neighcap = c->neighcap ? c->neighcap * 2 : (1L << 20);
    c->neigh = realloc(c->neigh, sizeof(int) * c->neighcap);
  }
  c->neigh[(*n)++] = j;
}

/* Neighbor::build (neighbor.cpp:2008-2081): xhold, bin_atoms (nbin_standard.cpp:192-232) and
 * NPairHalfBinAtomonlyNewton::build (npair_half_bin_atomonly_newton.cpp:37-118) or
 * NPairFullBinAtomonly::build (npair_full_bin_atomonly.cpp:34-95) */
static int neighbor_build(orc_ctx *c) {
  int nlocal = c->nlocal, nall = c->nlocal + c->nghost;
  c->ago = 0;
  c->nbuilds++;
  memcpy(c->xhold, c->x, sizeof(double) * 3 * nlocal);
  for (int i = 0; i < c->mbins; i++) c->binhead[i] = -1;
  int bad = 0;
  for (int i = nall - 1; i >= 0; i--) {
    int ibin = coord2bin(c, &c->x[3 * i], &bad);
    if (bad == 1) return fail(c, "Non-numeric positions - simulation unstable");
    if (bad == 2) return fail(c, "atom outside the bin grid (lost atom)");
    c->atom2bin[i] = ibin;
    c->bins[i] = c->binhead[ibin];
    c->binhead[ibin] = i;
  }
  int full = c->cfg.variant == ORC_FSI;
  long n = 0;
  const double *x = c->x;
  for (int i = 0; i < nlocal; i++) {
    c->firstneigh[i] = n;
    int itype = c->type[i];
    double xtmp = x[3 * i], ytmp = x[3 * i + 1], ztmp = x[3 * i + 2];
    if (!full) {
      for (int j = c->bins[i]; j >= 0; j = c->bins[j]) {
        if (j >= nlocal) {
          if (x[3 * j + 2] < ztmp) continue;
          if (x[3 * j + 2] == ztmp) {
            if (x[3 * j + 1] < ytmp) continue;
            if (x[3 * j + 1] == ytmp && x[3 * j] < xtmp) continue;
          }
        }
        double delx = xtmp - x[3 * j], dely = ytmp - x[3 * j + 1], delz = ztmp - x[3 * j + 2];
        double rsq = delx * delx + dely * dely + delz * delz;
        if (rsq <= c->cutneighsq[itype][c->type[j]]) push_neigh(c, &n, j);
      }
    }
    int ibin = c->atom2bin[i];
    for (int k = 0; k < c->nstencil; k++) {
      int b = ibin + c->stencil[k];
      if (b < 0 || b >= c->mbins - 1) continue; /* cannot happen for in-box owned atoms */
      for (int j = c->binhead[b]; j >= 0; j = c->bins[j]) {
        if (full && i == j) continue;
        double delx = xtmp - x[3 * j], dely = ytmp - x[3 * j + 1], delz = ztmp - x[3 * j + 2];
        double rsq = delx * delx + dely * dely + delz * delz;
        if (rsq <= c->cutneighsq[itype][c->type[j]]) push_neigh(c, &n, j);
      }
    }
    c->numneigh[i] = (int)(n - c->firstneigh[i]);
  }
  return 0;
}

int orc_build_neighbors(orc_ctx *c) {
  if (init_cutoffs(c)) return -1;
  domain_pbc(c);
  if (setup_bins(c)) return -1;
  borders(c);
  return neighbor_build(c);
}

/* Neighbor::decide + check_distance (neighbor.cpp:1922-2006), no box change, no must_check */
static int neighbor_decide(orc_ctx *c) {
  c->ago++;
  if (c->ago >= c->cfg.delay && c->ago % c->cfg.every == 0) {
    if (c->cfg.check == 0) return 1;
    int flag = 0;
    for (int i = 0; i < c->nlocal; i++) {
      double delx = c->x[3 * i] - c->xhold[3 * i];
      double dely = c->x[3 * i + 1] - c->xhold[3 * i + 1];
      double delz = c->x[3 * i + 2] - c->xhold[3 * i + 2];
      double rsq = delx * delx + dely * dely + delz * delz;
      if (rsq > c->triggersq) flag = 1;
    }
    int mx = c->cfg.every > c->cfg.delay ? c->cfg.every : c->cfg.delay;
    if (flag && c->ago == mx) c->ndanger++;
    return flag;
  }
  return 0;
}

/* Verlet::force_clear (verlet.cpp:370-415) + AtomVecSsaTsdpdAtomic::force_clear
 * (atom_vec_ssa_tsdpd_atomic.cpp:391-422), owned + ghost (newton on) */
static void force_clear(orc_ctx *c) {
  int nall = c->nlocal + c->nghost, S = c->cfg.nspecies;
  memset(c->f, 0, sizeof(double) * 3 * nall);
  memset(c->drho, 0, sizeof(double) * nall);
  memset(c->Q, 0, sizeof(double) * (S ? S : 1) * nall);
  memset(c->ddev, 0, sizeof(double) * 9 * nall);
  memset(c->art, 0, sizeof(double) * 9 * nall);
  memset(c->phi, 0, sizeof(double) * nall);
  memset(c->nd, 0, sizeof(double) * nall);
  memset(c->nw, 0, sizeof(double) * 3 * nall);
  memset(c->ddx, 0, sizeof(double) * 3 * nall);
  memset(c->ddv, 0, sizeof(double) * 3 * nall);
  memset(c->Pnew, 0, sizeof(double) * nall);
  memset(c->rhoAux1, 0, sizeof(double) * nall);
  memset(c->rhoAux2, 0, sizeof(double) * nall);
}

/* Lucy kernel and (1/r) dW/dr as the reference evaluates them
 * (pair_ssa_tsdpd_bvf_transport_velocity.cpp:204-241 and :318-355) */
static void lucy(int dim, double h, double r, double *wf, double *wfd) {
  double ih = 1.0 / h, ihsq = ih * ih, t = h - r;
  if (dim == 3) {
    *wfd = -25.066903536973515383e0 * t * t * ihsq * ihsq * ihsq * ih;
    *wf = 2.088908628081126 * t * t * t * ihsq * ihsq * ihsq * ih * (h + 3. * r);
  } else {
    *wfd = -19.098593171027440292e0 * t * t * ihsq * ihsq * ihsq;
    *wf = 1.591549430918954 * t * t * t * ihsq * ihsq * ihsq * (h + 3. * r);
  }
}

/* reverse_comm (atom_vec...:870-932): ghost accumulators are summed into their owners.  Only
 * the consumed, additive fields are folded. */
static void reverse_comm(orc_ctx *c) {
  int S = c->cfg.nspecies;
  for (int g = c->nlocal + c->nghost - 1; g >= c->nlocal; g--) {
    int o = c->gowner[g - c->nlocal];
    for (int d = 0; d < 3; d++) {
      c->f[3 * o + d] += c->f[3 * g + d];
      c->nw[3 * o + d] += c->nw[3 * g + d];
      c->ddv[3 * o + d] += c->ddv[3 * g + d];
      c->ddx[3 * o + d] += c->ddx[3 * g + d];
    }
    c->drho[o] += c->drho[g];
    c->phi[o] += c->phi[g];
    c->nd[o] += c->nd[g];
    c->rhoAux1[o] += c->rhoAux1[g];
    c->rhoAux2[o] += c->rhoAux2[g];
    for (int k = 0; k < S; k++) c->Q[S * o + k] += c->Q[S * g + k];
    for (int k = 0; k < 9; k++) c->ddev[9 * o + k] += c->ddev[9 * g + k];
  }
}

/* PairSsaTsdpdBvf{TransportVelocity,Mechanics,Fsi}::compute
 * (pair_ssa_tsdpd_bvf_transport_velocity.cpp:68-910, ..._mechanics.cpp:68-948, ..._fsi.cpp:81-797) */
static void pair_compute(orc_ctx *c) {
  const int var = c->cfg.variant, dim = c->cfg.dim, S = c->cfg.nspecies;
  const int newton = var != ORC_FSI; /* fsi: full list, i side only */
  double *x = c->x, *v = c->vest, *vt = c->v; /* naming swap of :84-85 */
  double *rho = c->rho, *f = c->f;
  const double kron[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  const double c_art = var == ORC_FSI ? 0.1 : 0.35;
  const double delta_fac = var == ORC_TV ? (1.0 / 2.6) : (1.0 / 3.0);
  /* density-diffusion amplitude (:532-539): amplDamp while tnow <= tmax = dt*nsteps of the run */
  const double amplDamp = var == ORC_FSI ? 0.1 : 0.0;
  double tnow = c->ntimestep * c->cfg.dt, tmax = c->cfg.dt * c->run_nsteps;
  const double damp = tnow <= tmax ? amplDamp : 0.0;

  /* ---- sweep A (:170-275; mechanics :172-290; fsi :184-278) */
  for (int i = 0; i < c->nlocal; i++) {
    int itype = c->type[i];
    double imass = c->mass[itype];
    double Pi = 7.0 * c->B[itype] * (rho[i] / c->rho0[itype] - 1.0);
    if (var != ORC_TV) c->Pnew[i] = Pi;
    for (int jj = 0; jj < c->numneigh[i]; jj++) {
      int j = c->neigh[c->firstneigh[i] + jj];
      double delx = x[3 * i] - x[3 * j], dely = x[3 * i + 1] - x[3 * j + 1], delz = x[3 * i + 2] - x[3 * j + 2];
      double rsq = delx * delx + dely * dely + delz * delz;
      int jtype = c->type[j];
      double jmass = c->mass[jtype];
      if (var != ORC_TV) c->Pnew[j] = 7.0 * c->B[jtype] * (rho[j] / c->rho0[jtype] - 1.0);
      if (rsq < c->cutsq[itype][jtype]) {
        double h = c->cut[itype][jtype], r = sqrt(rsq), wf, wfd;
        lucy(dim, h, r, &wf, &wfd);
        double wfd2 = wfd; /* hRatio = 1 (:186, :214-215) */
        double Vi = imass / rho[i], Vj = jmass / rho[j];
        double del[3] = {delx, dely, delz};
        c->nd[i] += pow(Vj, 2) * wf;
        c->rhoAux1[i] += c->rhoI[j] * wf;
        c->rhoAux2[i] += wf;
        for (int d = 0; d < 3; d++) {
          if (var != ORC_TV) c->ddx[3 * i + d] += pow(Vj, 2) * (v[3 * j + d] - v[3 * i + d]) * wf;
          c->ddv[3 * i + d] += 10.0 * 7.0 * c->B[itype] * (Vi * Vi + Vj * Vj) * wfd2 * del[d];
        }
        if (newton) {
          c->nd[j] += pow(Vi, 2) * wf;
          c->rhoAux1[j] += c->rhoI[i] * wf;
          c->rhoAux2[j] += wf;
          for (int d = 0; d < 3; d++) {
            if (var != ORC_TV) c->ddx[3 * j + d] += pow(Vi, 2) * (v[3 * i + d] - v[3 * j + d]) * wf;
            c->ddv[3 * j + d] += 10.0 * 7.0 * c->B[jtype] * (Vj * Vj + Vi * Vi) * wfd2 * (-del[d]);
          }
        }
      }
    }
  }

  /* ---- sweep B (:281-737) */
  for (int i = 0; i < c->nlocal; i++) {
    int itype = c->type[i];
    double imass = c->mass[itype];
    double fi = 7.0 * c->B[itype] * (rho[i] / c->rho0[itype] - 1.0);
    const double *vi = &v[3 * i], *vti = &vt[3 * i];
    for (int jj = 0; jj < c->numneigh[i]; jj++) {
      int j = c->neigh[c->firstneigh[i] + jj];
      double del[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
      double rsq = del[0] * del[0] + del[1] * del[1] + del[2] * del[2];
      int jtype = c->type[j];
      double jmass = c->mass[jtype];
      if (!(rsq < c->cutsq[itype][jtype])) continue;
      double h = c->cut[itype][jtype], r = sqrt(rsq), wf, wfd, wdelta, dummy;
      double delta = delta_fac * h;
      lucy(dim, h, r, &wf, &wfd);
      lucy(dim, h, delta, &wdelta, &dummy);
      double fj = 7.0 * c->B[jtype] * (rho[j] / c->rho0[jtype] - 1.0);
      const double *vj = &v[3 * j], *vtj = &vt[3 * j];
      double vel[3] = {vi[0] - vj[0], vi[1] - vj[1], vi[2] - vj[2]};
      double delVdotDelR = del[0] * vel[0] + del[1] * vel[1] + del[2] * vel[2];
      double Vi = imass / rho[i], Vj = jmass / rho[j];
      double S2 = pow(Vi, 2) + pow(Vj, 2);
      /* transport tensor and force (:370-377) */
      double ftr[3];
      for (int m = 0; m < 3; m++) {
        double acc = 0.0;
        for (int n = 0; n < 3; n++) {
          double T = 0.5 * ((rho[i] * vi[m] * (vti[n] - vi[n])) + (rho[j] * vj[m] * (vtj[n] - vj[n])));
          acc += T * del[n];
        }
        ftr[m] = S2 * acc * wfd;
      }
      double fvisc = S2 * (c->eta[itype][jtype] * wfd); /* :387 */
      /* pressure force (:396-399; mechanics :408; fsi :390) */
      double pij = (fj / (rho[j] * rho[j])) + (fi / (rho[i] * rho[i]));
      double fpair;
      if (var == ORC_TV) {
        if (pij >= 0.) fpair = imass * jmass * ((fj / (rho[j] * rho[j])) + (fi / (rho[i] * rho[i]))) * wfd;
        else fpair = imass * jmass * ((fj / (rho[j] * rho[j])) - (fi / (rho[i] * rho[i]))) * wfd;
        if (c->solid[i] == 1 && c->solid[j] == 1)
          fpair = imass * jmass * ((fj / (rho[j] * rho[j])) + (fi / (rho[i] * rho[i]))) * wfd;
      } else {
        fpair = imass * jmass * ((fj / (rho[j] * rho[j])) + (fi / (rho[i] * rho[i]))) * wfd;
      }
      /* strain / rotation of i (:435-440) and Jaumann rate (:443-451) */
      const double *devi = &c->dev[9 * i], *devj = &c->dev[9 * j];
      double G0i = c->G0[itype], G0j = c->G0[jtype];
      if (var == ORC_FSI && S > 0) { /* pair_ssa_tsdpd_bvf_fsi.cpp:441-442 */
        G0i = c->G0[itype] * (1.0 - 0.99 * c->C[S * i]);
        G0j = c->G0[jtype] * (1.0 - 0.99 * c->C[S * j]);
      }
      if (c->solid[i] == 1) {
        double eps[3][3], om[3][3];
        for (int m = 0; m < 3; m++)
          for (int n = 0; n < 3; n++) {
            double a = (vj[m] - vi[m]) * del[n], b = (vj[n] - vi[n]) * del[m];
            eps[m][n] = 0.5 * Vj * wfd * (a + b);
            om[m][n] = 0.5 * Vj * wfd * (a - b);
          }
        for (int m = 0; m < 3; m++)
          for (int n = 0; n < 3; n++) {
            double dDotR = devi[3 * m] * om[n][0] + devi[3 * m + 1] * om[n][1] + devi[3 * m + 2] * om[n][2];
            double rDotD = om[m][0] * devi[n] + om[m][1] * devi[3 + n] + om[m][2] * devi[6 + n];
            c->ddev[9 * i + 3 * m + n] += 2.0 * ((2.0 * G0i * G0j) / (G0i + G0j + 1e-12)) *
                                              (eps[m][n] - (1. / 3.) * kron[m][n] * eps[m][n]) + dDotR + rDotD;
          }
      }
      /* artificial stress of i and j (:454-483), assignment semantics */
      double Ri[3][3] = {{0}}, Rj[3][3] = {{0}};
      if (c->solid[i] == 1)
        for (int m = 0; m < 3; m++)
          for (int n = 0; n < 3; n++) {
            double Ps = var == ORC_MECHANICS ? fabs(fi) : fi;
            double ts = devi[3 * m + n] - Ps * kron[m][n];
            Ri[m][n] = ts > 0.0 ? -c_art * ts / (rho[i] * rho[i]) : 0.0;
          }
      if (c->solid[j] == 1)
        for (int m = 0; m < 3; m++)
          for (int n = 0; n < 3; n++) {
            double Ps = var == ORC_MECHANICS ? fabs(fj) : fj;
            double ts = devj[3 * m + n] - Ps * kron[m][n];
            Rj[m][n] = ts > 0.0 ? -c_art * ts / (rho[j] * rho[j]) : 0.0;
          }
      double fart[3];
      for (int n = 0; n < 3; n++)
        fart[n] = imass * jmass * wfd * pow(wf / wdelta, 4) *
                  (del[0] * (Ri[0][n] + Rj[0][n]) + del[1] * (Ri[1][n] + Rj[1][n]) + del[2] * (Ri[2][n] + Rj[2][n]));
      /* momentum of i (:497-529) */
      if (c->solid[i] == 0) {
        for (int d = 0; d < 3; d++) f[3 * i + d] += -del[d] * fpair + fvisc * vel[d] + ftr[d] + fart[d];
      } else {
        double fdev[3];
        for (int n = 0; n < 3; n++)
          fdev[n] = imass * jmass * wfd *
                    (del[0] * (devi[n] / (rho[i] * rho[i]) + devj[n] / (rho[j] * rho[j])) +
                     del[1] * (devi[3 + n] / (rho[i] * rho[i]) + devj[3 + n] / (rho[j] * rho[j])) +
                     del[2] * (devi[6 + n] / (rho[i] * rho[i]) + devj[6 + n] / (rho[j] * rho[j])));
        double fviscs = 0.;
        if (delVdotDelR < 0.) {
          double mu = h * delVdotDelR / (rsq + 0.01 * h * h);
          fviscs = imass * jmass * wfd * (-(c->c0[itype] + c->c0[jtype]) * mu + 2.0 * mu * mu) / (rho[i] + rho[j]);
        }
        for (int d = 0; d < 3; d++) f[3 * i + d] += -del[d] * fpair - del[d] * fviscs + fdev[d] + fart[d];
      }
      /* density rate (:548-555) */
      double velt[3] = {vti[0] - vtj[0], vti[1] - vtj[1], vti[2] - vtj[2]};
      double delVtdotDelR = del[0] * velt[0] + del[1] * velt[1] + del[2] * velt[2];
      double ai = (vi[0] - vti[0]) * del[0] + (vi[1] - vti[1]) * del[1] + (vi[2] - vti[2]) * del[2];
      double aj = (vj[0] - vtj[0]) * del[0] + (vj[1] - vtj[1]) * del[1] + (vj[2] - vtj[2]) * del[2];
      c->drho[i] += (rho[i] * jmass * delVtdotDelR * wfd / rho[j]) -
                    damp * h * rho[i] * c->c0[itype] * jmass * 2.0 * (rho[j] / rho[i] - 1.0) *
                        (rsq / (rsq + 0.01 * h * h)) * wfd / rho[j] -
                    (jmass / rho[j]) * (rho[i] * ai + rho[j] * aj) * wfd;
      /* BVF phi and wall normal (:563-576) */
      if (c->solid[i] == 0 && c->solid[j] == 1) {
        c->phi[i] += pow(Vj, 2) * wf;
        for (int d = 0; d < 3; d++) c->nw[3 * i + d] += del[d] * wfd * pow(Vj, 2);
      }
      /* Newton mirror (:579-675) */
      if (newton) {
        if (c->solid[j] == 1) {
          double eps[3][3], om[3][3];
          for (int m = 0; m < 3; m++)
            for (int n = 0; n < 3; n++) {
              double a = (vi[m] - vj[m]) * (-del[n]), b = (vi[n] - vj[n]) * (-del[m]);
              eps[m][n] = 0.5 * Vi * wfd * (a + b);
              om[m][n] = 0.5 * Vi * wfd * (a - b);
            }
          for (int m = 0; m < 3; m++)
            for (int n = 0; n < 3; n++) {
              double dDotR = devj[3 * m] * om[n][0] + devj[3 * m + 1] * om[n][1] + devj[3 * m + 2] * om[n][2];
              double rDotD = om[m][0] * devj[n] + om[m][1] * devj[3 + n] + om[m][2] * devj[6 + n];
              c->ddev[9 * j + 3 * m + n] += 2.0 * ((2.0 * G0i * G0j) / (G0i + G0j + 1e-14)) *
                                                (eps[m][n] - (1. / 3.) * kron[m][n] * eps[m][n]) + dDotR + rDotD;
            }
        }
        if (c->solid[j] == 0) {
          double fp = fpair;
          if (var == ORC_TV && pij < 0.) fp = -fp; /* :606 */
          for (int d = 0; d < 3; d++) f[3 * j + d] -= (-del[d] * fp + fvisc * vel[d] + ftr[d] + fart[d]);
        } else {
          double fdev[3];
          for (int n = 0; n < 3; n++)
            fdev[n] = imass * jmass * wfd *
                      (del[0] * (devj[n] / (rho[j] * rho[j]) + devi[n] / (rho[i] * rho[i])) +
                       del[1] * (devj[3 + n] / (rho[j] * rho[j]) + devi[3 + n] / (rho[i] * rho[i])) +
                       del[2] * (devj[6 + n] / (rho[j] * rho[j]) + devi[6 + n] / (rho[i] * rho[i])));
          double fviscs = 0.;
          if (delVdotDelR < 0.) {
            double mu = h * delVdotDelR / (rsq + 0.01 * h * h);
            fviscs = imass * jmass * wfd * (-(c->c0[itype] + c->c0[jtype]) * mu + 2.0 * mu * mu) / (rho[i] + rho[j]);
          }
          /* reference: fpair is NOT sign-flipped for a solid j (:633), so the force on a free
             solid from a fluid neighbour with pij < 0 depends on which of the two is `i` in the half
             list.  symmetric_switch = 1 applies the flip, i.e. what a gather over
             j's own neighbours computes (the CUDA library); default 0 = reference. */
          double fpj = fpair;
          if (c->symmetric_switch && var == ORC_TV && pij < 0. && !(c->solid[i] == 1 && c->solid[j] == 1)) fpj = -fpair;
          for (int d = 0; d < 3; d++) f[3 * j + d] -= (-del[d] * fpj - del[d] * fviscs + fdev[d] + fart[d]);
        }
        c->drho[j] += (rho[j] * imass * delVtdotDelR * wfd / rho[i]) -
                      damp * h * rho[j] * c->c0[jtype] * imass * 2.0 * (rho[i] / rho[j] - 1.0) *
                          (rsq / (rsq + 0.01 * h * h)) * wfd / rho[i] +
                      (imass / rho[i]) * (rho[j] * aj + rho[i] * ai) * wfd;
        if (c->solid[j] == 0 && c->solid[i] == 1) {
          c->phi[j] += pow(Vi, 2) * wf;
          for (int d = 0; d < 3; d++) c->nw[3 * j + d] += -del[d] * wfd * pow(Vi, 2);
        }
      }
      /* species transport (:678-720; mechanics :728-731; fsi :612-615) */
      if (r < c->cutc[itype][jtype]) {
        double hc = c->cutc[itype][jtype], wfc, wfdc;
        lucy(dim, hc, r, &wfc, &wfdc);
        double q0 = 2.0 * ((imass * jmass) / (imass + jmass)) * ((rho[i] + rho[j]) / (rho[i] * rho[j])) *
                    (del[0] * del[0] + del[1] * del[1] + del[2] * del[2]) * wfdc / (rsq + 0.01 * hc * hc);
        for (int k = 0; k < S; k++) {
          double Ci = c->C[S * i + k], Cj = c->C[S * j + k];
          double kap = c->kappa[itype][jtype][k];
          if (var == ORC_TV) {
            c->Q[S * i + k] += kap * (Ci - Cj) * q0 - (jmass / rho[j]) * (Ci * ai + Cj * aj) * wfdc;
            if (newton) c->Q[S * j + k] -= (kap * (Ci - Cj) * q0 - (imass / rho[i]) * (Ci * ai + Cj * aj) * wfdc);
          } else {
            double dQc = kap * (Ci - Cj) * q0;
            c->Q[S * i + k] += dQc;
            if (newton) c->Q[S * j + k] -= dQc;
          }
        }
      }
    }
  }
}

int orc_pair_compute(orc_ctx *c) {
  force_clear(c);
  pair_compute(c);
  reverse_comm(c);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Integrator fixes (fix_ssa_tsdpd_bvf_transport_velocity.cpp:76-461, ..._mechanics.cpp:77-500,
 * ..._fsi.cpp:77-470)
 * ---------------------------------------------------------------------------------------- */
static void setup_pre_force(orc_ctx *c) {
  for (int i = 0; i < c->nlocal; i++)
    if (c->mask[i] & c->cfg.integrate_groupbit) {
      for (int d = 0; d < 3; d++) c->vest[3 * i + d] = c->v[3 * i + d];
      c->rhoI[i] = c->rho[i];
    }
}

static void damp_factors(const orc_ctx *c, double *damp, double *dampSolid) {
  double tnow = (double)c->ntimestep;
  *damp = tnow <= 1.0 ? tnow / 1.0 : 1.0;
  if (c->cfg.variant == ORC_MECHANICS) *dampSolid = tnow < 1e6 ? 0.0 : 1.0; /* ..._mechanics.cpp:151-153 */
  else *dampSolid = tnow <= 1.0 ? 0.0 : 1.0;                                  /* ..._fsi.cpp:150-152 */
}

static void initial_integrate(orc_ctx *c) {
  const int var = c->cfg.variant, S = c->cfg.nspecies;
  const double dtv = c->cfg.dt, dtf = 0.5 * c->cfg.dt;
  double damp, dampSolid;
  damp_factors(c, &damp, &dampSolid);
  for (int i = 0; i < c->nlocal; i++) {
    if (!(c->mask[i] & c->cfg.integrate_groupbit)) continue;
    double dtfm = dtf / c->mass[c->type[i]];
    double *x = &c->x[3 * i], *v = &c->v[3 * i], *vest = &c->vest[3 * i], *f = &c->f[3 * i];
    double *ddv = &c->ddv[3 * i], *ddx = &c->ddx[3 * i];
    if (c->fixed[i] == 0) {
      if (c->solid[i] == 0) {
        for (int d = 0; d < 3; d++) {
          if (var == ORC_TV) vest[d] = v[d] + dtfm * f[d];
          else vest[d] = v[d] + dtfm * f[d] * damp + 0.001 * ddx[d] / c->nd[i];
          v[d] = vest[d] - dtfm * ddv[d];
          x[d] += dtv * v[d];
        }
      } else {
        for (int d = 0; d < 3; d++) {
          if (var == ORC_TV) vest[d] = v[d] + 2.0 * dtfm * f[d];
          else vest[d] = v[d] + 2.0 * dtfm * f[d] + 0.001 * ddx[d] / c->nd[i];
          v[d] += dtfm * f[d];
          if (var != ORC_TV) { vest[d] *= dampSolid; v[d] *= dampSolid; }
          x[d] += dtf * v[d];
        }
        for (int k = 0; k < 9; k++)
          c->dev[9 * i + k] += (var == ORC_TV ? 0.5 * dtv : dtf) * c->ddev[9 * i + k];
      }
      c->rhoI[i] = c->rho[i];
      c->rho[i] += dtf * c->drho[i];
    } else {
      if (c->solid[i] == 0) {
        c->rhoI[i] = c->rho[i];
        c->rho[i] += dtf * c->drho[i];
      } else {
        for (int k = 0; k < 9; k++) c->dev[9 * i + k] += dtf * c->ddev[9 * i + k];
        c->rhoI[i] = c->rho[i];
      }
    }
    for (int k = 0; k < S; k++) {
      c->C[S * i + k] += c->Q[S * i + k] * dtf;
      c->C[S * i + k] = c->C[S * i + k] > 0 ? c->C[S * i + k] : 0.0;
    }
  }
}

static void final_integrate(orc_ctx *c) {
  const int var = c->cfg.variant, S = c->cfg.nspecies;
  const double dtv = c->cfg.dt, dtf = 0.5 * c->cfg.dt;
  /* freqFilter: 20 (TV :287, mechanics :311); fsi "1e16" -> INT_MAX, never fires (..._fsi.cpp:304) */
  const long freqFilter = var == ORC_FSI ? 2147483647L : 20;
  const int filter = (c->ntimestep % freqFilter) == 0;
  double damp, dampSolid;
  damp_factors(c, &damp, &dampSolid);
  for (int i = 0; i < c->nlocal; i++) {
    if (!(c->mask[i] & c->cfg.integrate_groupbit)) continue;
    double dtfm = dtf / c->mass[c->type[i]];
    double *x = &c->x[3 * i], *v = &c->v[3 * i], *vest = &c->vest[3 * i], *f = &c->f[3 * i];
    double *nw = &c->nw[3 * i], *ddx = &c->ddx[3 * i];
    double nd = c->nd[i];
    c->phi[i] = c->phi[i] / nd;
    for (int d = 0; d < 3; d++) nw[d] = nw[d] / nd;
    if (c->fixed[i] == 0) {
      if (c->solid[i] == 0) {
        if (c->phi[i] > 0.5) { /* BVF reflection (:310-342) */
          for (int d = 0; d < 3; d++) x[d] -= dtv * v[d];
          double norm = sqrt(nw[0] * nw[0] + nw[1] * nw[1] + nw[2] * nw[2]);
          double en[3] = {-nw[0] / norm, -nw[1] / norm, -nw[2] / norm};
          double vdot = v[0] * en[0] + v[1] * en[1] + v[2] * en[2];
          double mx = vdot > 0.0 ? vdot : 0.0; /* std::max(0.0, v_dot_en) */
          for (int d = 0; d < 3; d++) v[d] = -v[d] + 2.0 * mx * en[d];
          for (int d = 0; d < 3; d++) x[d] += dtv * v[d];
        }
        for (int d = 0; d < 3; d++) {
          if (var == ORC_TV) v[d] = vest[d] + dtfm * f[d];
          else v[d] = vest[d] + dtfm * f[d] * damp + 0.001 * ddx[d] / nd;
        }
        if (var == ORC_TV) c->rho[i] = filter ? c->rhoAux1[i] / c->rhoAux2[i] + dtf * c->drho[i] : c->rhoI[i] + dtf * c->drho[i];
        else c->rho[i] = filter ? c->rhoAux1[i] / c->rhoAux2[i] + dtf * c->drho[i] : c->rhoI[i] + dtv * c->drho[i];
      } else {
        for (int d = 0; d < 3; d++) {
          if (var == ORC_TV) v[d] += dtfm * f[d];
          else { v[d] += dtfm * f[d] + 0.001 * ddx[d] / nd; v[d] *= dampSolid; }
        }
        for (int k = 0; k < 9; k++)
          c->dev[9 * i + k] += (var == ORC_TV ? 0.5 * dtv : dtf) * c->ddev[9 * i + k];
        if (var == ORC_TV) c->rho[i] = filter ? c->rhoAux1[i] / c->rhoAux2[i] + dtf * c->drho[i] : c->rhoI[i] + dtf * c->drho[i];
        else c->rho[i] = c->rhoI[i] + dtv * c->drho[i];
      }
    } else {
      if (c->solid[i] == 0) {
        c->rho[i] = filter ? c->rhoAux1[i] / c->rhoAux2[i] + dtv * c->drho[i] : c->rhoI[i] + dtv * c->drho[i];
      } else {
        for (int k = 0; k < 9; k++) c->dev[9 * i + k] += dtf * c->ddev[9 * i + k];
        c->rho[i] = filter ? c->rhoAux1[i] / c->rhoAux2[i] : c->rhoI[i];
      }
    }
    for (int k = 0; k < S; k++) {
      c->C[S * i + k] += c->Q[S * i + k] * dtf;
      c->C[S * i + k] = c->C[S * i + k] > 0 ? c->C[S * i + k] : 0.0;
    }
  }
}

/* FixSsaTsdpdForcing::post_integrate (fix_ssa_tsdpd_forcing.cpp:133-176),
 * FixSsaTsdpdBuffer::post_integrate (fix_ssa_tsdpd_buffer.cpp:124-178) */
static void post_integrate(orc_ctx *c) {
  int S = c->cfg.nspecies;
  for (int q = 0; q < c->nfix; q++) {
    orc_fix *fx = &c->fix[q];
    if (fx->kind == FIX_FORCING) {
      if (!(c->ntimestep > fx->step)) continue;
      for (int i = 0; i < c->nlocal; i++) {
        if (!(c->mask[i] & fx->groupbit)) continue;
        double drx = c->x[3 * i] - fx->a[0], dry = c->x[3 * i + 1] - fx->a[1];
        int inside;
        if (fx->a_int[2] == 0) inside = (drx * drx + dry * dry) < fx->a[2] * fx->a[2];
        else inside = fabs(drx) < fx->a[2] && fabs(dry) < fx->a[3];
        if (!inside) continue;
        if (fx->a_int[0] == 0) c->C[S * i + fx->a_int[1]] = fx->a[4];
        else c->vest[3 * i + fx->a_int[1]] = fx->a[4];
      }
    } else if (fx->kind == FIX_BUFFER) {
      if (!(c->ntimestep > fx->step) || fx->a_int[0] == 2) continue;
      for (int i = 0; i < c->nlocal; i++) {
        if (!(c->mask[i] & fx->groupbit)) continue;
        double drx = c->x[3 * i] - fx->a[0], dry = c->x[3 * i + 1] - fx->a[1];
        if (!(fabs(drx) < fx->a[2] && fabs(dry) < fx->a[3])) continue;
        double phi;
        if (fx->a_int[2] == 0) {
          double xo = fx->a[0] - fx->a[2], xL = fx->a[0] + fx->a[2];
          phi = (c->x[3 * i] - xo) / (xL - xo);
          phi = phi * phi * phi;
        } else {
          double yo = fx->a[1] - fx->a[3], yL = fx->a[1] + fx->a[3];
          phi = (c->x[3 * i + 1] - yo) / (yL - yo);
          phi = 0.5 * (1.0 - tanh(8.0 - 16.0 * phi));
        }
        double *t = fx->a_int[0] == 0 ? &c->C[S * i + fx->a_int[1]] : &c->vest[3 * i + fx->a_int[1]];
        *t = *t - phi * (*t - fx->a[4]);
      }
    }
  }
}

/* FixSsaTsdpdBuoyancy::post_force (fix_ssa_tsdpd_buoyancy.cpp:113-140), FixSetForce::post_force,
 * FixSsaTsdpdChemRxnMassAction::post_force (fix_ssa_tsdpd_chem_rxn_mass_action.cpp:76-112).  in_setup: Modify::setup
 * calls Fix::setup, which only setforce and buoyancy forward to post_force; the reaction fix has no setup(). */
static void post_force(orc_ctx *c, int in_setup) {
  int S = c->cfg.nspecies;
  for (int q = 0; q < c->nfix; q++) {
    orc_fix *fx = &c->fix[q];
    if (fx->kind == FIX_CHEMRXN) {
      if (in_setup) continue;
      const int nr = fx->a_int[0], np = fx->a_int[1];
      for (int i = 0; i < c->nlocal; i++) {
        if (!(c->mask[i] & fx->groupbit)) continue;
        double *C = c->C + (size_t)S * i, *Q = c->Q + (size_t)S * i;
        double flux;
        if (nr == 2) flux = fx->a[0] * C[fx->a_int[2] & 255] * C[(fx->a_int[2] >> 8) & 255];
        else if (nr == 1) flux = fx->a[0] * C[fx->a_int[2] & 255];
        else flux = fx->a[0];
        for (int j = 0; j < nr; j++) Q[(fx->a_int[2] >> (8 * j)) & 255] -= flux;
        for (int j = 0; j < np; j++) Q[(fx->a_int[3] >> (8 * j)) & 255] += flux;
      }
    } else if (fx->kind == FIX_BUOYANCY) {
      for (int i = 0; i < c->nlocal; i++) {
        if (!(c->mask[i] & fx->groupbit)) continue;
        double m = c->mass[c->type[i]];
        if (fx->a_int[0]) c->f[3 * i + fx->a_int[1]] += m * fx->a[0];
        else c->f[3 * i + fx->a_int[1]] += m * fx->a[0] * (c->C[S * i + fx->a_int[2]] - fx->a[1]);
      }
    } else if (fx->kind == FIX_SETFORCE) {
      for (int i = 0; i < c->nlocal; i++)
        if (c->mask[i] & fx->groupbit)
          for (int d = 0; d < 3; d++) c->f[3 * i + d] = fx->a[d];
    }
  }
}

/* FixSsaTsdpdBuffer::end_of_step (fix_ssa_tsdpd_buffer.cpp:182-240) */
static void end_of_step(orc_ctx *c) {
  for (int q = 0; q < c->nfix; q++) {
    orc_fix *fx = &c->fix[q];
    if (fx->kind != FIX_BUFFER || fx->a_int[0] != 2) continue;
    if (!(c->ntimestep > fx->step)) continue;
    for (int i = 0; i < c->nlocal; i++) {
      if (!(c->mask[i] & fx->groupbit)) continue;
      double drx = c->x[3 * i] - fx->a[0], dry = c->x[3 * i + 1] - fx->a[1];
      if (!(fabs(drx) < fx->a[2] && fabs(dry) < fx->a[3])) continue;
      double phi;
      if (fx->a_int[2] == 0) {
        double xo = fx->a[0] - fx->a[2], xL = fx->a[0] + fx->a[2];
        phi = (c->x[3 * i] - xo) / (xL - xo);
        phi = phi * phi * phi;
      } else {
        double yo = fx->a[1] - fx->a[3], yL = fx->a[1] + fx->a[3];
        phi = (c->x[3 * i + 1] - yo) / (yL - yo);
        phi = 0.5 * (1.0 - tanh(8.0 - 16.0 * phi));
      }
      c->rho[i] = c->rho[i] - phi * (c->rho[i] - fx->a[4]);
    }
  }
}

/* Verlet::setup (verlet.cpp:88-170) */
int orc_setup(orc_ctx *c) {
  for (int i = 0; i < c->nlocal; i++)
    if (c->e[i] != 0.0) return fail(c, "oracle requires e == 0 (random stress term not restated)");
  if (init_cutoffs(c)) return -1;
  domain_pbc(c);
  if (setup_bins(c)) return -1;
  /* the reference makes ghosts BEFORE setup_pre_force (verlet.cpp:118-132), so at step 0 they
     carry stale vest/rhoI.  With half lists + Newton that is not even
     gather-consistent; the CUDA library deliberately does not reproduce it, and
     orc_set_consistent_ghosts(1) gives the oracle the library's order for direct comparison. */
  if (c->consistent_ghosts) setup_pre_force(c);
  borders(c);
  if (neighbor_build(c)) return -1;
  c->nbuilds = 0;      /* neighbor->ncalls = 0 (verlet.cpp:128) */
  force_clear(c);
  setup_pre_force(c);
  pair_compute(c);
  reverse_comm(c);
  post_force(c, 1);    /* modify->setup -> FixSetForce::setup / FixSsaTsdpdBuoyancy::setup */
  c->setup_done = 1;
  return 0;
}

/* Verlet::run (verlet.cpp:223-354) */
int orc_run(orc_ctx *c, int nsteps) {
  if (!c->setup_done) return fail(c, "orc_run before orc_setup");
  if (c->run_nsteps_user < 0) c->run_nsteps = nsteps; /* update->nsteps (pair :534) */
  for (int s = 0; s < nsteps; s++) {
    c->ntimestep++;
    initial_integrate(c);
    post_integrate(c);
    if (neighbor_decide(c)) {
      domain_pbc(c);
      borders(c);
      if (neighbor_build(c)) return -1;
    } else {
      for (int g = c->nlocal; g < c->nlocal + c->nghost; g++) ghost_forward(c, g);
    }
    force_clear(c);
    pair_compute(c);
    reverse_comm(c);
    post_force(c, 0);
    final_integrate(c);
    end_of_step(c);
  }
  return 0;
}

void orc_set_run_length(orc_ctx *c, long nsteps) {
  c->run_nsteps_user = nsteps;
  c->run_nsteps = nsteps;
}

void orc_set_consistent_ghosts(orc_ctx *c, int on) { c->consistent_ghosts = on; }
void orc_set_symmetric_switch(orc_ctx *c, int on) { c->symmetric_switch = on; }

int orc_nlocal(const orc_ctx *c) { return c->nlocal; }
int orc_nghost(const orc_ctx *c) { return c->nghost; }
long orc_ntimestep(const orc_ctx *c) { return c->ntimestep; }
int orc_nbuilds(const orc_ctx *c) { return c->nbuilds; }

int orc_get(const orc_ctx *c, const char *name, double *out) {
  const double *src = NULL;
  int nc = 1, S = c->cfg.nspecies;
  if (!strcmp(name, "x")) { src = c->x; nc = 3; }
  else if (!strcmp(name, "v")) { src = c->v; nc = 3; }
  else if (!strcmp(name, "vest")) { src = c->vest; nc = 3; }
  else if (!strcmp(name, "f")) { src = c->f; nc = 3; }
  else if (!strcmp(name, "nw")) { src = c->nw; nc = 3; }
  else if (!strcmp(name, "ddv")) { src = c->ddv; nc = 3; }
  else if (!strcmp(name, "ddx")) { src = c->ddx; nc = 3; }
  else if (!strcmp(name, "rho")) src = c->rho;
  else if (!strcmp(name, "rhoI")) src = c->rhoI;
  else if (!strcmp(name, "drho")) src = c->drho;
  else if (!strcmp(name, "e")) src = c->e;
  else if (!strcmp(name, "phi")) src = c->phi;
  else if (!strcmp(name, "number_density")) src = c->nd;
  else if (!strcmp(name, "rhoAux1")) src = c->rhoAux1;
  else if (!strcmp(name, "rhoAux2")) src = c->rhoAux2;
  else if (!strcmp(name, "Pnew")) src = c->Pnew;
  else if (!strcmp(name, "dev")) { src = c->dev; nc = 9; }
  else if (!strcmp(name, "ddev")) { src = c->ddev; nc = 9; }
  else if (!strcmp(name, "C")) { src = c->C; nc = S; }
  else if (!strcmp(name, "Q")) { src = c->Q; nc = S; }
  else return -1;
  if (out && nc) memcpy(out, src, sizeof(double) * nc * c->nlocal);
  return nc;
}

int orc_get_int(const orc_ctx *c, const char *name, int *out) {
  const int *src = NULL;
  if (!strcmp(name, "tag")) src = c->tag;
  else if (!strcmp(name, "type")) src = c->type;
  else if (!strcmp(name, "mask")) src = c->mask;
  else if (!strcmp(name, "solid_tag")) src = c->solid;
  else if (!strcmp(name, "fixed_tag")) src = c->fixed;
  else return -1;
  if (out) memcpy(out, src, sizeof(int) * c->nlocal);
  return 1;
}

long orc_get_pairs(const orc_ctx *c, int *out, long cap) {
  long n = 0;
  for (int i = 0; i < c->nlocal; i++)
    for (int jj = 0; jj < c->numneigh[i]; jj++) {
      if (out && n < cap) {
        out[2 * n] = c->tag[i];
        out[2 * n + 1] = c->tag[c->neigh[c->firstneigh[i] + jj]];
      }
      n++;
    }
  return n;
}

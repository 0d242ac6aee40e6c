/* ----------------------------------------------------------------------
   atom_style ssa_tsdpd/atomic/cuda -- see atom_vec_ssa_tsdpd_atomic_cuda.h.
   Replaces AtomVecSsaTsdpdAtomic (atom_vec_ssa_tsdpd_atomic.cpp:29-2300) for /cuda runs: same dictionary and
   message layouts, table-driven, derived arrays allocated on first host use.
------------------------------------------------------------------------- */

#include <stdlib.h>
#include <string.h>
#include "atom_vec_ssa_tsdpd_atomic_cuda.h"
#include "sphbvf_lmp.h"
#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "fix.h"
#include "force.h"
#include "memory.h"
#include "modify.h"

using namespace LAMMPS_NS;

#if defined(LAMMPS_BIGBIG)
#error "atom_style ssa_tsdpd/atomic/cuda: the GPU library keeps 32-bit tags and image flags (build without -DLAMMPS_BIGBIG)"
#endif

// indices of the state fields in fld[] (the derived ones follow in the order of enum Derived)
enum { S_X, S_V, S_F, S_TAG, S_TYPE, S_MASK, S_IMAGE, S_RHO, S_E, S_CV, S_VEST, S_C, S_SOLID, S_FIXED, S_DEV, S_RHOI };

/* ---------------------------------------------------------------------- */

AtomVecSsaTsdpdAtomicCuda::AtomVecSsaTsdpdAtomicCuda(LAMMPS *lmp) : AtomVec(lmp)
{
  molecular = 0;
  mass_type = 1;
  forceclearflag = 1;

  comm_x_only = 0;   // forward: v, rho, e, vest, C, stress and rhoI travel with x
  comm_f_only = 0;   // reverse: every pair-sweep output travels with f
  // per-atom message sizes of the reference (atom_vec_ssa_tsdpd_atomic.cpp:36-38); + S in process_args
  size_forward = 22;
  size_reverse = 51;
  size_border = 27;
  size_velocity = 3;
  size_data_atom = 8;   // id solid_tag type rho x y z <ignored>
  size_data_vel = 3;
  xcol_data = 5;

  atom->e_flag = 1;
  atom->rho_flag = 1;
  atom->cv_flag = 1;
  atom->vest_flag = 1;
  atom->sdpd_flag = 1;
  nfld = 0;
}

/* ----------------------------------------------------------------------
   atom_style ssa_tsdpd/atomic S [Nssa Nrxn]   (atom_vec_ssa_tsdpd_atomic.cpp:58-108)
------------------------------------------------------------------------- */

void AtomVecSsaTsdpdAtomicCuda::process_args(int narg, char **arg)
{
  if (narg < 1) error->all(FLERR, "Must provide NUMBER OF DETERMINISTIC SPECIES (int).");
  if (narg != 1 && narg != 3)
    error->all(FLERR, "Must provide NUMBER OF STOCHASTIC SPECIES (int) and NUMBER OF STOCHASTIC REACTIONS (int).");
  const int S = atoi(arg[0]);
  if (S < 0) error->all(FLERR, "Number of SDPD species must be greater than or equal to 0");
  atom->num_sdpd_species = S;
  if (comm->me == 0) printf("num_sdpd_species = %d \n", S);
  if (narg == 3) {
    const int nssa = atoi(arg[1]), nrxn = atoi(arg[2]);
    if (nssa < 0) error->all(FLERR, "Number of SSA species must be greater than or equal to 0");
    if (nrxn < 0) error->all(FLERR, "Number of SSA reactions must be greater than or equal to 0");
    if (nssa > 0 || nrxn > 0)
      error->all(FLERR, "atom_style ssa_tsdpd/atomic/cuda holds no SSA species (serial-only upstream, not part of the "
                        "GPU path): put 'suffix off' before and 'suffix on' after the atom_style command");
    atom->num_ssa_species = 0;
    atom->num_ssa_reactions = 0;
    if (comm->me == 0) printf("num_ssa_species = %d \nnum_ssa_reactions = %d \n", 0, 0);
  }
  size_forward += S;
  size_reverse += S;
  size_border += S;
  build_table();
}

/* ----------------------------------------------------------------------
   the dictionary: one row per per-atom array of the package (atom_vec_ssa_tsdpd_atomic.h:58-85)
------------------------------------------------------------------------- */

void AtomVecSsaTsdpdAtomicCuda::build_table()
{
  const int S = atom->num_sdpd_species;
  Atom *a = atom;
  nfld = 0;
  auto add = [&](const char *name, void *slot, int shape, int cols, unsigned in, double init = 0.0) {
    Field f = {name, slot, shape, cols, in, init};
    fld[nfld++] = f;
  };
  // ---- state (allocated with the atoms)
  add("x", &a->x, D2, 3, STATE | FWD | BRD | EXC);
  add("v", &a->v, D2, 3, STATE | FWD | BRD | EXC);
  add("f", &a->f, D2, 3, STATE | REV);
  add("tag", &a->tag, I1, 1, STATE | BRD | EXC);
  add("type", &a->type, I1, 1, STATE | BRD | EXC);
  add("mask", &a->mask, I1, 1, STATE | BRD | EXC, 1.0);
  add("image", &a->image, I1, 1, STATE | EXC);
  add("rho", &a->rho, D1, 1, STATE | FWD | BRD | EXC);
  add("e", &a->e, D1, 1, STATE | FWD | BRD | EXC);
  add("cv", &a->cv, D1, 1, STATE | BRD | EXC, 1.0);
  add("vest", &a->vest, D2, 3, STATE | FWD | BRD | EXC);
  add("C", &a->C, D2, S, STATE | FWD | BRD | EXC);
  add("solid_tag", &a->solid_tag, I1, 1, STATE | BRD | EXC);
  add("fixed_tag", &a->fixed_tag, I1, 1, STATE | BRD | EXC);
  add("deviatoricTensor", &a->deviatoricTensor, D33, 9, STATE | FWD | BRD | EXC);
  add("rhoI", &a->rhoI, D1, 1, STATE | FWD | BRD | EXC);
  // ---- derived (pair-sweep outputs: lazy), in the reference's reverse-communication order
  add("drho", &a->drho, D1, 1, REV | CLR);
  add("de", &a->de, D1, 1, REV | CLR);
  add("Q", &a->Q, D2, S, REV | CLR);
  add("ddeviatoricTensor", &a->ddeviatoricTensor, D33, 9, REV | CLR);
  add("artificialStressTensor", &a->artificialStressTensor, D33, 9, REV | CLR);
  add("phi", &a->phi, D1, 1, REV | CLR);
  add("number_density", &a->number_density, D1, 1, REV | CLR);
  add("nw", &a->nw, D2, 3, REV | CLR);
  add("v_weighted_solid", &a->v_weighted_solid, D2, 3, REV | CLR);
  add("a_weighted_solid", &a->a_weighted_solid, D2, 3, REV | CLR);
  add("ddx", &a->ddx, D2, 3, REV | CLR);
  add("ddv", &a->ddv, D2, 3, REV | CLR);
  add("Pold", &a->Pold, D1, 1, REV | CLR);
  add("Pnew", &a->Pnew, D1, 1, REV | CLR);
  add("Aaux", &a->Aaux, D1, 1, REV | CLR);
  add("Baux", &a->Baux, D1, 1, REV | CLR);
  add("APaux", &a->APaux, D1, 1, REV | CLR);
  add("fP", &a->fP, D2, 3, REV | CLR);
  add("rhoAux1", &a->rhoAux1, D1, 1, REV | CLR);
  add("rhoAux2", &a->rhoAux2, D1, 1, REV | CLR);
  add("rhoAux3", &a->rhoAux3, D1, 1, REV | CLR);

  static const int fwd[] = {S_X, S_V, S_RHO, S_E, S_VEST, S_C, S_DEV, S_RHOI, -1};
  static const int brd[] = {S_X, S_V, S_TAG, S_TYPE, S_MASK, S_RHO, S_E, S_CV, S_VEST, S_C, S_SOLID, S_FIXED, S_DEV, S_RHOI, -1};
  static const int exc[] = {S_X, S_V, S_TAG, S_TYPE, S_MASK, S_IMAGE, S_RHO, S_E, S_CV, S_VEST, S_C, S_SOLID, S_FIXED, S_DEV, S_RHOI, -1};
  static const int rst[] = {S_X, S_TAG, S_TYPE, S_MASK, S_IMAGE, S_V, S_RHO, S_E, S_CV, S_VEST, S_C, S_SOLID, S_FIXED, S_DEV, S_RHOI, -1};
  static_assert(sizeof fwd <= sizeof order_fwd && sizeof brd <= sizeof order_brd && sizeof exc <= sizeof order_exc &&
                    sizeof rst <= sizeof order_rst, "message layouts must fit their order arrays");
  memcpy(order_fwd, fwd, sizeof fwd);
  memcpy(order_brd, brd, sizeof brd);
  memcpy(order_exc, exc, sizeof exc);
  memcpy(order_rst, rst, sizeof rst);
}

/* ---------------------------------------------------------------------- */

double *AtomVecSsaTsdpdAtomicCuda::row(const Field &f, int i) const
{
  switch (f.shape) {
    case D1: return *(double *const *)f.slot + i;
    case D2: return (*(double **const *)f.slot)[i];
    case D33: return &(*(double ***const *)f.slot)[i][0][0];
  }
  return NULL;
}

void AtomVecSsaTsdpdAtomicCuda::alloc(Field &f, int n)
{
  char name[64];
  snprintf(name, sizeof name, "atom:%s", f.name);
  const int nt = (f.in & STATE) && strcmp(f.name, "f") != 0 ? n : n * comm->nthreads;   // accumulators: one copy per thread
  switch (f.shape) {
    case I1: memory->grow(*(int **)f.slot, nt, name); break;
    case D1: memory->grow(*(double **)f.slot, nt, name); break;
    case D2: memory->grow(*(double ***)f.slot, nt, f.cols, name); break;
    case D33: memory->grow(*(double ****)f.slot, nt, 3, 3, name); break;
  }
}

void AtomVecSsaTsdpdAtomicCuda::set_default(const Field &f, int i)
{
  if (f.shape == I1) {
    ival(f, i) = (int)f.init;
    return;
  }
  if (f.cols == 0) return;
  double *p = row(f, i);
  for (int k = 0; k < f.cols; k++) p[k] = f.init;
}

/* ----------------------------------------------------------------------
   n = 0 grows by a chunk, n > 0 allocates to size n (atom_vec_ssa_tsdpd_atomic.cpp:116-189): state fields always,
   derived fields only once they have been materialised
------------------------------------------------------------------------- */

void AtomVecSsaTsdpdAtomicCuda::grow(int n)
{
  if (!nfld) build_table();   // atom_style given without arguments cannot happen (process_args errors), but replicate / restart re-create the class
  const int old = nmax;
  if (n == 0) {
    // AtomVec::grow_nmax adds 16384 rows per call; every call re-points the row tables of all [n][3] arrays, so filling
    // 64 M atoms that way costs N^2 / 32768 = 1.2e11 pointer stores per array (a quarter of an hour of create_atoms).
    // Past a million rows grow by an eighth instead: ~ 35 calls from 1 M to 64 M.
    grow_nmax();
    if (old >= (1 << 20)) {
      const bigint want = (bigint)old + old / 8;
      if (want > nmax) nmax = (int)MIN(want, (bigint)MAXSMALLINT);
    }
  } else nmax = n;
  atom->nmax = nmax;
  if (nmax < 0 || nmax > MAXSMALLINT) error->one(FLERR, "Per-processor system is too big");
  for (int q = 0; q < nfld; q++) {
    Field &f = fld[q];
    if (!(f.in & STATE) && !have(f)) continue;
    if (f.shape == D2 && f.cols == 0 && !(f.in & STATE)) continue;
    alloc(f, nmax);
    // memory->grow leaves the new rows uninitialised; the derived mirrors are read as "last pair sweep or zero"
    if (!(f.in & STATE))
      for (int i = old > 0 ? old : 0; i < nmax; i++) set_default(f, i);
  }
  if (atom->nextra_grow)
    for (int iextra = 0; iextra < atom->nextra_grow; iextra++) modify->fix[atom->extra_grow[iextra]]->grow_arrays(nmax);
}

/* ---------------------------------------------------------------------- */

bool AtomVecSsaTsdpdAtomicCuda::materialized(int which) const { return have(fld[NSTATE + which]); }

int AtomVecSsaTsdpdAtomicCuda::nlazy() const
{
  int n = 0;
  for (int q = NSTATE; q < nfld; q++)
    if (!have(fld[q])) n++;
  return n;
}

void AtomVecSsaTsdpdAtomicCuda::materialize(int which)
{
  Field &f = fld[NSTATE + which];
  if (have(f) || nmax <= 0) return;
  if (f.shape == D2 && f.cols == 0) {
    // S == 0: upstream holds a table of row pointers to nothing; code that is correct upstream never dereferences it
    alloc(f, nmax);
    return;
  }
  alloc(f, nmax);
  for (int i = 0; i < nmax; i++) set_default(f, i);
}

void AtomVecSsaTsdpdAtomicCuda::materialize_all()
{
  for (int w = 0; w < NDERIVED; w++) materialize(w);
}

/* ----------------------------------------------------------------------
   start of every run: a pair style that is not a /cuda style computes on the host and writes every derived array
------------------------------------------------------------------------- */

void AtomVecSsaTsdpdAtomicCuda::init()
{
  AtomVec::init();
  const char *ps = force->pair_style;
  const size_t n = ps ? strlen(ps) : 0;
  const bool device_pair = force->pair && n >= 5 && strcmp(ps + n - 5, "/cuda") == 0 && strstr(ps, "ssa_tsdpd/bvf");
  if (!device_pair) materialize_all();
}

/* ---------------------------------------------------------------------- */

void AtomVecSsaTsdpdAtomicCuda::copy(int i, int j, int delflag)
{
  for (int q = 0; q < nfld; q++) {
    const Field &f = fld[q];
    if (!have(f)) continue;
    if (f.shape == I1) ival(f, j) = ival(f, i);
    else if (f.cols) memcpy(row(f, j), row(f, i), f.cols * sizeof(double));
  }
  if (atom->nextra_grow)
    for (int iextra = 0; iextra < atom->nextra_grow; iextra++) modify->fix[atom->extra_grow[iextra]]->copy_arrays(i, j, delflag);
}

/* nbytes = sizeof(double) * (number of atoms to clear), starting at atom n (atom_vec_ssa_tsdpd_atomic.cpp:391-422) */

void AtomVecSsaTsdpdAtomicCuda::force_clear(int n, size_t nbytes)
{
  if (!nbytes) return;
  for (int q = 0; q < nfld; q++) {
    const Field &f = fld[q];
    if (!(f.in & CLR) || !have(f) || f.cols == 0) continue;
    memset(row(f, n), 0, f.cols * nbytes);
  }
}

/* ----------------------------------------------------------------------
   one atom -> buffer and back, in the order of a message layout
------------------------------------------------------------------------- */

int AtomVecSsaTsdpdAtomicCuda::put(const int *order, int j, double *buf, const double *shift) const
{
  int m = 0;
  for (; *order >= 0; order++) {
    const Field &f = fld[*order];
    if (f.shape == I1) {
      buf[m++] = ubuf(ival(f, j)).d;
      continue;
    }
    if (!f.cols) continue;
    const double *p = row(f, j);
    if (*order == S_X && shift) {
      buf[m++] = p[0] + shift[0];
      buf[m++] = p[1] + shift[1];
      buf[m++] = p[2] + shift[2];
    } else
      for (int k = 0; k < f.cols; k++) buf[m++] = p[k];
  }
  return m;
}

int AtomVecSsaTsdpdAtomicCuda::take(const int *order, int i, const double *buf)
{
  int m = 0;
  for (; *order >= 0; order++) {
    const Field &f = fld[*order];
    if (f.shape == I1) {
      ival(f, i) = (int)ubuf(buf[m++]).i;
      continue;
    }
    if (!f.cols) continue;
    double *p = row(f, i);
    for (int k = 0; k < f.cols; k++) p[k] = buf[m++];
  }
  return m;
}

static void pbc_shift(LAMMPS_NS::Domain *domain, const int *pbc, bool border, double *s)
{
  // comm_brick.cpp hands over the image counts; orthogonal boxes shift by whole box lengths
  if (domain->triclinic == 0) {
    s[0] = pbc[0] * domain->xprd;
    s[1] = pbc[1] * domain->yprd;
    s[2] = pbc[2] * domain->zprd;
  } else if (border) {
    s[0] = pbc[0];
    s[1] = pbc[1];
    s[2] = pbc[2];
  } else {
    s[0] = pbc[0] * domain->xprd + pbc[5] * domain->xy + pbc[4] * domain->xz;
    s[1] = pbc[1] * domain->yprd + pbc[3] * domain->yz;
    s[2] = pbc[2] * domain->zprd;
  }
}

/* ---------------------------------------------------------------------- forward */

int AtomVecSsaTsdpdAtomicCuda::pack_comm(int n, int *list, double *buf, int pbc_flag, int *pbc)
{
  double s[3];
  if (pbc_flag) pbc_shift(domain, pbc, false, s);
  int m = 0;
  for (int i = 0; i < n; i++) m += put(order_fwd, list[i], buf + m, pbc_flag ? s : NULL);
  return m;
}

/* with `comm_modify vel yes` the same record travels (v and vest are part of it); atoms of the fix deform group
   crossing a periodic face get the box velocity added to v and vest (atom_vec_ssa_tsdpd_atomic.cpp:1209-1240) */

void AtomVecSsaTsdpdAtomicCuda::vel_remap(int n, int *list, double *buf, int stride, int voff, int vestoff, const int *pbc) const
{
  if (!deform_vremap) return;
  const double dv[3] = {pbc[0] * h_rate[0] + pbc[5] * h_rate[5] + pbc[4] * h_rate[4], pbc[1] * h_rate[1] + pbc[3] * h_rate[3],
                        pbc[2] * h_rate[2]};
  const int *mask = atom->mask;
  for (int i = 0; i < n; i++) {
    if (!(mask[list[i]] & deform_groupbit)) continue;
    double *b = buf + (size_t)i * stride;
    for (int k = 0; k < 3; k++) {
      b[voff + k] += dv[k];
      b[vestoff + k] += dv[k];
    }
  }
}

int AtomVecSsaTsdpdAtomicCuda::pack_comm_vel(int n, int *list, double *buf, int pbc_flag, int *pbc)
{
  const int m = pack_comm(n, list, buf, pbc_flag, pbc);
  if (pbc_flag && n) vel_remap(n, list, buf, m / n, 3, 8, pbc);   // x3 v3 rho e vest3 ...
  return m;
}

void AtomVecSsaTsdpdAtomicCuda::unpack_comm(int n, int first, double *buf)
{
  int m = 0;
  for (int i = first; i < first + n; i++) m += take(order_fwd, i, buf + m);
}

void AtomVecSsaTsdpdAtomicCuda::unpack_comm_vel(int n, int first, double *buf) { unpack_comm(n, first, buf); }

/* ---------------------------------------------------------------------- reverse: f + every derived field that exists */

int AtomVecSsaTsdpdAtomicCuda::pack_reverse(int n, int first, double *buf)
{
  int m = 0;
  for (int i = first; i < first + n; i++)
    for (int q = 0; q < nfld; q++) {
      const Field &f = fld[q];
      if (!(f.in & REV) || !have(f) || !f.cols) continue;
      const double *p = row(f, i);
      for (int k = 0; k < f.cols; k++) buf[m++] = p[k];
    }
  return m;
}

void AtomVecSsaTsdpdAtomicCuda::unpack_reverse(int n, int *list, double *buf)
{
  int m = 0;
  for (int i = 0; i < n; i++)
    for (int q = 0; q < nfld; q++) {
      const Field &f = fld[q];
      if (!(f.in & REV) || !have(f) || !f.cols) continue;
      double *p = row(f, list[i]);
      for (int k = 0; k < f.cols; k++) p[k] += buf[m++];
    }
}

/* ---------------------------------------------------------------------- border */

int AtomVecSsaTsdpdAtomicCuda::pack_border(int n, int *list, double *buf, int pbc_flag, int *pbc)
{
  double s[3];
  if (pbc_flag) pbc_shift(domain, pbc, true, s);
  int m = 0;
  for (int i = 0; i < n; i++) m += put(order_brd, list[i], buf + m, pbc_flag ? s : NULL);
  if (atom->nextra_border)
    for (int iextra = 0; iextra < atom->nextra_border; iextra++) m += modify->fix[atom->extra_border[iextra]]->pack_border(n, list, &buf[m]);
  return m;
}

int AtomVecSsaTsdpdAtomicCuda::pack_border_vel(int n, int *list, double *buf, int pbc_flag, int *pbc)
{
  double s[3];
  if (pbc_flag) pbc_shift(domain, pbc, true, s);
  int m = 0;
  for (int i = 0; i < n; i++) m += put(order_brd, list[i], buf + m, pbc_flag ? s : NULL);
  if (pbc_flag && n) vel_remap(n, list, buf, m / n, 3, 12, pbc);   // x3 v3 tag type mask rho e cv vest3 ...
  if (atom->nextra_border)
    for (int iextra = 0; iextra < atom->nextra_border; iextra++) m += modify->fix[atom->extra_border[iextra]]->pack_border(n, list, &buf[m]);
  return m;
}

void AtomVecSsaTsdpdAtomicCuda::unpack_border(int n, int first, double *buf)
{
  int m = 0;
  for (int i = first; i < first + n; i++) {
    if (i == nmax) grow(0);
    m += take(order_brd, i, buf + m);
  }
  if (atom->nextra_border)
    for (int iextra = 0; iextra < atom->nextra_border; iextra++) m += modify->fix[atom->extra_border[iextra]]->unpack_border(n, first, &buf[m]);
}

void AtomVecSsaTsdpdAtomicCuda::unpack_border_vel(int n, int first, double *buf) { unpack_border(n, first, buf); }

/* ---------------------------------------------------------------------- exchange: buf[0] = record length */

int AtomVecSsaTsdpdAtomicCuda::pack_exchange(int i, double *buf)
{
  int m = 1;
  m += put(order_exc, i, buf + m, NULL);
  if (atom->nextra_grow)
    for (int iextra = 0; iextra < atom->nextra_grow; iextra++) m += modify->fix[atom->extra_grow[iextra]]->pack_exchange(i, &buf[m]);
  buf[0] = m;
  return m;
}

int AtomVecSsaTsdpdAtomicCuda::unpack_exchange(double *buf)
{
  const int nlocal = atom->nlocal;
  if (nlocal == nmax) grow(0);
  for (int q = 0; q < nfld; q++)
    if (have(fld[q])) set_default(fld[q], nlocal);
  int m = 1;
  m += take(order_exc, nlocal, buf + m);
  if (atom->nextra_grow)
    for (int iextra = 0; iextra < atom->nextra_grow; iextra++) m += modify->fix[atom->extra_grow[iextra]]->unpack_exchange(nlocal, &buf[m]);
  atom->nlocal++;
  return m;
}

/* ---------------------------------------------------------------------- restart: xyz first (read_restart tests on them) */

int AtomVecSsaTsdpdAtomicCuda::size_restart()
{
  const int nlocal = atom->nlocal;
  int per = 1;
  for (const int *o = order_rst; *o >= 0; o++) per += fld[*o].shape == I1 ? 1 : fld[*o].cols;
  int n = per * nlocal;
  if (atom->nextra_restart)
    for (int iextra = 0; iextra < atom->nextra_restart; iextra++)
      for (int i = 0; i < nlocal; i++) n += modify->fix[atom->extra_restart[iextra]]->size_restart(i);
  return n;
}

int AtomVecSsaTsdpdAtomicCuda::pack_restart(int i, double *buf)
{
  int m = 1;
  m += put(order_rst, i, buf + m, NULL);
  if (atom->nextra_restart)
    for (int iextra = 0; iextra < atom->nextra_restart; iextra++) m += modify->fix[atom->extra_restart[iextra]]->pack_restart(i, &buf[m]);
  buf[0] = m;
  return m;
}

int AtomVecSsaTsdpdAtomicCuda::unpack_restart(double *buf)
{
  const int nlocal = atom->nlocal;
  if (nlocal == nmax) {
    grow(0);
    if (atom->nextra_store) memory->grow(atom->extra, nmax, atom->nextra_store, "atom:extra");
  }
  for (int q = 0; q < nfld; q++)
    if (have(fld[q])) set_default(fld[q], nlocal);
  int m = 1;
  m += take(order_rst, nlocal, buf + m);
  double **extra = atom->extra;
  if (atom->nextra_store) {
    const int size = static_cast<int>(buf[0]) - m;
    for (int i = 0; i < size; i++) extra[nlocal][i] = buf[m++];
  }
  atom->nlocal++;
  return m;
}

/* ----------------------------------------------------------------------
   new atom of type itype at coord (create_atoms; atom_vec_ssa_tsdpd_atomic.cpp:1851-1941): everything zero, cv = 1
------------------------------------------------------------------------- */

void AtomVecSsaTsdpdAtomicCuda::create_atom(int itype, double *coord)
{
  const int nlocal = atom->nlocal;
  if (nlocal == nmax) grow(0);
  for (int q = 0; q < nfld; q++)
    if (have(fld[q])) set_default(fld[q], nlocal);
  atom->tag[nlocal] = 0;
  atom->type[nlocal] = itype;
  atom->x[nlocal][0] = coord[0];
  atom->x[nlocal][1] = coord[1];
  atom->x[nlocal][2] = coord[2];
  atom->image[nlocal] = ((imageint)IMGMAX << IMG2BITS) | ((imageint)IMGMAX << IMGBITS) | IMGMAX;
  atom->nlocal++;
}

/* one line of the Atoms section: id solid_tag type rho x y z <ignored>  (atom_vec_ssa_tsdpd_atomic.cpp:1949-2040) */

void AtomVecSsaTsdpdAtomicCuda::data_atom(double *coord, imageint imagetmp, char **values)
{
  const int nlocal = atom->nlocal;
  if (nlocal == nmax) grow(0);
  for (int q = 0; q < nfld; q++)
    if (have(fld[q])) set_default(fld[q], nlocal);
  atom->tag[nlocal] = ATOTAGINT(values[0]);
  atom->solid_tag[nlocal] = atoi(values[1]);
  atom->type[nlocal] = atoi(values[2]);
  if (atom->type[nlocal] <= 0 || atom->type[nlocal] > atom->ntypes)
    error->one(FLERR, "Invalid atom type in Atoms section of data file");
  atom->rho[nlocal] = atof(values[3]);
  atom->x[nlocal][0] = coord[0];
  atom->x[nlocal][1] = coord[1];
  atom->x[nlocal][2] = coord[2];
  atom->image[nlocal] = imagetmp;
  atom->nlocal++;
}

/* write_data: the columns data_atom reads back (the 8th, ignored on input, carries e) + image flags */

void AtomVecSsaTsdpdAtomicCuda::pack_data(double **buf)
{
  const int nlocal = atom->nlocal;
  for (int i = 0; i < nlocal; i++) {
    buf[i][0] = ubuf(atom->tag[i]).d;
    buf[i][1] = ubuf(atom->solid_tag[i]).d;
    buf[i][2] = ubuf(atom->type[i]).d;
    buf[i][3] = atom->rho[i];
    buf[i][4] = atom->x[i][0];
    buf[i][5] = atom->x[i][1];
    buf[i][6] = atom->x[i][2];
    buf[i][7] = atom->e[i];
    buf[i][8] = ubuf((atom->image[i] & IMGMASK) - IMGMAX).d;
    buf[i][9] = ubuf((atom->image[i] >> IMGBITS & IMGMASK) - IMGMAX).d;
    buf[i][10] = ubuf((atom->image[i] >> IMG2BITS) - IMGMAX).d;
  }
}

void AtomVecSsaTsdpdAtomicCuda::write_data(FILE *fp, int n, double **buf)
{
  for (int i = 0; i < n; i++)
    fprintf(fp, TAGINT_FORMAT " %d %d %-1.16e %-1.16e %-1.16e %-1.16e %-1.16e %d %d %d\n", (tagint)ubuf(buf[i][0]).i,
            (int)ubuf(buf[i][1]).i, (int)ubuf(buf[i][2]).i, buf[i][3], buf[i][4], buf[i][5], buf[i][6], buf[i][7],
            (int)ubuf(buf[i][8]).i, (int)ubuf(buf[i][9]).i, (int)ubuf(buf[i][10]).i);
}

/* ----------------------------------------------------------------------
   compute property/atom names of the package (atom_vec_ssa_tsdpd_atomic.cpp:2093-2105)
------------------------------------------------------------------------- */

int AtomVecSsaTsdpdAtomicCuda::property_atom(char *name)
{
  static const char *names[] = {"rho", "drho", "e", "de", "cv", "phi", "solid_tag", "Pnew", "deviatoricTensor"};
  for (int k = 0; k < 9; k++)
    if (strcmp(name, names[k]) == 0) return k;
  return -1;
}

void AtomVecSsaTsdpdAtomicCuda::pack_property_atom(int index, double *buf, int nvalues, int groupbit)
{
  // the mirror must exist and be current before it is read
  static const int lazy[9] = {-1, DRHO, -1, DE, -1, PHI, -1, PNEW, -1};
  static const unsigned bits[9] = {SphbvfLmp::HF_RHO, SphbvfLmp::HF_DRHO, 0, 0, 0, SphbvfLmp::HF_PHI, 0, SphbvfLmp::HF_PNEW, SphbvfLmp::HF_DEV};
  if (index < 0 || index > 8) return;
  if (lazy[index] >= 0) materialize(lazy[index]);
  if (bits[index]) SphbvfLmp::host_fields(atom, bits[index]);

  const int *mask = atom->mask;
  const int nlocal = atom->nlocal;
  int n = 0;
  for (int i = 0; i < nlocal; i++, n += nvalues) {
    double val = 0.0;
    if (mask[i] & groupbit) switch (index) {
        case 0: val = atom->rho[i]; break;
        case 1: val = atom->drho[i]; break;
        case 2: val = atom->e[i]; break;
        case 3: val = atom->de[i]; break;
        case 4: val = atom->cv[i]; break;
        case 5: val = atom->phi[i]; break;
        case 6: val = atom->solid_tag[i]; break;
        case 7: val = atom->Pnew[i]; break;
        case 8: val = atom->deviatoricTensor[i][2][2]; break;   // upstream's loop leaves the last component in the column
      }
    buf[n] = val;
  }
}

/* ---------------------------------------------------------------------- */

bigint AtomVecSsaTsdpdAtomicCuda::memory_usage() { return host_bytes(true); }

/* Atom::memcheck is only valid inside Atom::memory_usage() (it reads a scratch string that exists only there) */

bigint AtomVecSsaTsdpdAtomicCuda::host_bytes(bool from_atom_memory_usage)
{
  bigint bytes = 0;
  for (int q = 0; q < nfld; q++) {
    const Field &f = fld[q];
    if (!have(f) || (from_atom_memory_usage && !atom->memcheck(f.name))) continue;
    switch (f.shape) {
      case I1: bytes += memory->usage(*(int **)f.slot, nmax); break;
      case D1: bytes += memory->usage(*(double **)f.slot, nmax); break;
      case D2: bytes += memory->usage(*(double ***)f.slot, nmax, f.cols); break;
      case D33: bytes += memory->usage(*(double ****)f.slot, nmax, 3, 3); break;
    }
  }
  return bytes;
}

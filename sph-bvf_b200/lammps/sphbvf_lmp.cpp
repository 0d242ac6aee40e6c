/* ----------------------------------------------------------------------
   sphbvf_lmp.cpp -- see sphbvf_lmp.h.  Replaces, for the /cuda styles, what Verlet::setup does on
   the host between Atom and the styles (verlet.cpp:88-170): here the atoms go to the device once
   and stay there.
------------------------------------------------------------------------- */

#include <stdlib.h>
#include <string.h>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include "sphbvf_lmp.h"
#include "atom_vec_ssa_tsdpd_atomic_cuda.h"
#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "lammps.h"
#include "memory.h"
#include "compute.h"
#include "dump.h"
#include "dump_custom.h"
#include "fix.h"
#include "modify.h"
#include "neighbor.h"
#include "output.h"
#include "pair.h"
#include "thermo.h"
#include "update.h"

using namespace LAMMPS_NS;

static SphbvfLmp *the_engine = NULL;   // one LAMMPS instance per process in this build

namespace LAMMPS_NS {
/* one persistent thread per GPU: run(f) executes f(rank) on every worker and returns the status codes.
   The library's multi-rank entry points contain NCCL collectives and stream synchronisations, so the ranks
   of one process must be inside the same call at the same time -- exactly what MPI ranks do upstream. */
class SphbvfWorkers {
 public:
  explicit SphbvfWorkers(int n) : n_(n), gen_(0), pending_(0), quit_(false), rc_(n, 0)
  {
    for (int r = 0; r < n; r++) th_.emplace_back([this, r] { loop(r); });
  }
  ~SphbvfWorkers()
  {
    {
      std::lock_guard<std::mutex> lk(m_);
      quit_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (size_t r = 0; r < th_.size(); r++) th_[r].join();
  }
  int size() const { return n_; }
  const std::vector<int> &run(const std::function<int(int)> &f)
  {
    std::unique_lock<std::mutex> lk(m_);
    task_ = &f;
    pending_ = n_;
    gen_++;
    cv_.notify_all();
    done_.wait(lk, [this] { return pending_ == 0; });
    return rc_;
  }

 private:
  void loop(int r)
  {
    unsigned long seen = 0;
    for (;;) {
      const std::function<int(int)> *f;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (quit_) return;
        f = task_;
      }
      const int rc = (*f)(r);
      {
        std::lock_guard<std::mutex> lk(m_);
        rc_[r] = rc;
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  int n_;
  unsigned long gen_;
  int pending_;
  bool quit_;
  std::vector<int> rc_;
  const std::function<int(int)> *task_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> th_;
};
}

SphbvfLmp *SphbvfLmp::get(LAMMPS *lmp)
{
  if (!the_engine) the_engine = new SphbvfLmp(lmp);
  return the_engine;
}

SphbvfLmp *SphbvfLmp::peek() { return the_engine; }

void SphbvfLmp::release(LAMMPS *)
{
  delete the_engine;
  the_engine = NULL;
}

SphbvfLmp::SphbvfLmp(LAMMPS *lmp) : Pointers(lmp)
{
  variant = SPHBVF_TV;
  pair = NULL;
  rho0 = soundspeed = G0 = NULL;
  viscosity = cut = cutc = NULL;
  kappa = NULL;
  integrate_groupbit = 1;
  nfixdesc = 0;
  ctx = NULL;
  host_mask = HF_ALL;
  nlocal_uploaded = 0;
  ndownloads = ndevice_thermo = nskipped = 0;
  nbytes_down = 0;
  nranks = 1;
  workers = NULL;
}

SphbvfLmp::~SphbvfLmp()
{
  destroy_contexts();
  delete workers;
}

void SphbvfLmp::destroy_contexts()
{
  if (ctxs.size() > 1 && workers) workers->run([&](int r) { sphbvf_destroy(ctxs[r]); return 0; });   // on the thread that owns the device
  else if (ctxs.size() == 1) sphbvf_destroy(ctxs[0]);
  ctxs.clear();
  ctx = NULL;
}

/* f(ctx_r, r) on every rank at the same time; the first failing rank's message becomes the LAMMPS error */

void SphbvfLmp::all(const std::function<int(sphbvf_ctx *, int)> &f)
{
  if (nranks == 1) {
    check(f(ctx, 0));
    return;
  }
  const std::vector<int> &rc = workers->run([&](int r) { return f(ctxs[r], r); });
  for (int r = 0; r < nranks; r++)
    if (rc[r]) {
      char msg[600];
      snprintf(msg, sizeof msg, "sphbvf (%d) on GPU %d: %s", rc[r], r, sphbvf_last_error(ctxs[r]));
      error->one(FLERR, msg);
    }
}

void SphbvfLmp::neighbor_step(int *rebuilt)
{
  std::vector<int> rb(nranks, 0);
  all([&](sphbvf_ctx *c, int r) { return sphbvf_neighbor(c, &rb[r]); });
  if (rebuilt) *rebuilt = rb[0];   // the rebuild vote is global
}

void SphbvfLmp::virial(double *v6)
{
  std::vector<double> v(6 * (size_t)nranks, 0.0);
  all([&](sphbvf_ctx *c, int r) { return sphbvf_virial(c, &v[6 * (size_t)r]); });
  for (int k = 0; k < 6; k++) {
    v6[k] = 0.0;
    for (int r = 0; r < nranks; r++) v6[k] += v[6 * (size_t)r + k];
  }
}

void SphbvfLmp::ke_tensor(int groupbit, double *t6)
{
  std::vector<double> v(6 * (size_t)nranks, 0.0);
  all([&](sphbvf_ctx *c, int r) { return sphbvf_ke_tensor(c, groupbit, &v[6 * (size_t)r]); });
  for (int k = 0; k < 6; k++) {
    t6[k] = 0.0;
    for (int r = 0; r < nranks; r++) t6[k] += v[6 * (size_t)r + k];
  }
}

double SphbvfLmp::max_vsq(int groupbit)
{
  std::vector<double> v(nranks, 0.0);
  all([&](sphbvf_ctx *c, int r) { return sphbvf_max_vsq(c, groupbit, &v[r]); });   // all-reduced inside the library
  return v[0];
}

void SphbvfLmp::check(int rc)
{
  if (rc == 0) return;
  char msg[600];
  snprintf(msg, sizeof msg, "sphbvf (%d): %s", rc, ctx ? sphbvf_last_error(ctx) : "no context");
  error->one(FLERR, msg);
}

void SphbvfLmp::add_fix(const FixDesc &f)
{
  if (nfixdesc == 16) error->all(FLERR, "Too many /cuda fixes");
  fixdesc[nfixdesc++] = f;
}

/* ----------------------------------------------------------------------
   create the device context from the LAMMPS state and upload the owned atoms
------------------------------------------------------------------------- */

void SphbvfLmp::start()
{
  if (comm->nprocs != 1)
    error->all(FLERR, "The /cuda styles drive all GPUs from one process: run LAMMPS on 1 MPI rank");
  if (atom->num_ssa_species > 0)
    error->all(FLERR, "SSA species are not supported by the /cuda styles");   // serial-only upstream as well
  if (sphbvf_device_count() < 1)
    error->all(FLERR, "No CUDA device: the /cuda styles have no CPU fallback (drop -sf cuda)");
  if (ctx) stop();

  const int S = atom->num_sdpd_species, ntypes = atom->ntypes;
  sphbvf_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.dim = domain->dimension;
  cfg.periodic[0] = domain->xperiodic;
  cfg.periodic[1] = domain->yperiodic;
  cfg.periodic[2] = domain->zperiodic;
  for (int k = 0; k < 3; k++) {
    cfg.boxlo[k] = domain->boxlo[k];
    cfg.boxhi[k] = domain->boxhi[k];
    cfg.procgrid[k] = 1;
  }
  cfg.ntypes = ntypes;
  cfg.nspecies = S;
  cfg.variant = variant;
  cfg.skin = neighbor->skin;
  cfg.neigh_every = neighbor->every;
  cfg.neigh_delay = neighbor->delay;
  cfg.neigh_check = neighbor->dist_check;
  cfg.dt = update->dt;
  cfg.integrate_groupbit = integrate_groupbit;

  // ---- how many GPUs: SPHBVF_NGPU, else every visible GPU once there are >= 10^6 atoms for each
  const int n = atom->nlocal;
  int want = sphbvf_device_count();
  {
    const char *env = getenv("SPHBVF_NGPU");
    if (env) want = atoi(env);
    else want = MIN(want, MAX(1, n / 1000000));
  }
  nranks = MAX(1, MIN(want, sphbvf_device_count()));
  {
    // brick grid: procmap.cpp's surface-minimising factorisation, or SPHBVF_GRID=PxxPyxPz.  (The `processors`
    // command cannot carry it: with one MPI rank comm.cpp:325-327 rejects any grid whose product is not 1.)
    double prd[3] = {domain->xprd, domain->yprd, domain->zprd};
    int grid[3] = {1, 1, 1};
    const char *genv = getenv("SPHBVF_GRID");
    if (genv) {
      if (sscanf(genv, "%dx%dx%d", &grid[0], &grid[1], &grid[2]) != 3 || grid[0] < 1 || grid[1] < 1 || grid[2] < 1 ||
          (cfg.dim == 2 && grid[2] != 1) || grid[0] * grid[1] * grid[2] > sphbvf_device_count())
        error->all(FLERR, "SPHBVF_GRID must be PxxPyxPz with Px*Py*Pz <= number of GPUs (Pz = 1 in 2d)");
      nranks = grid[0] * grid[1] * grid[2];
    } else
      check(sphbvf_proc_grid(nranks, cfg.dim, prd, grid));
    for (int k = 0; k < 3; k++) cfg.procgrid[k] = grid[k];
  }
  cfg.nranks = nranks;
  if (workers && workers->size() != nranks) { delete workers; workers = NULL; }   // GPU count changed between runs
  if (nranks > 1 && !workers) workers = new SphbvfWorkers(nranks);

  // ---- one context per GPU, created on the worker thread that will drive it (cudaSetDevice is per thread)
  ctxs.assign(nranks, (sphbvf_ctx *)NULL);
  std::vector<int> rc_create(nranks, 0);
  {
    auto create = [&](int r) {
      sphbvf_config c = cfg;
      c.device = r;
      c.rank = r;
      rc_create[r] = sphbvf_create(&c, &ctxs[r]);
      return 0;
    };
    if (nranks == 1) create(0);
    else workers->run(create);
  }
  for (int r = 0; r < nranks; r++)
    if (rc_create[r]) {
      for (int q = 0; q < nranks; q++) if (ctxs[q]) sphbvf_destroy(ctxs[q]);
      ctxs.clear();
      ctx = NULL;
      error->all(FLERR, "sphbvf_create failed (more than 4 atom types / species, or no usable GPU)");
    }
  ctx = ctxs[0];
  if (nranks > 1) {
    char id[128];
    check(sphbvf_comm_unique_id(id));
    all([&](sphbvf_ctx *c, int) { return sphbvf_comm_init(c, id); });
  }

  // Pair::coeff / init_one already mirrored [i][j] -> [j][i] (pair_ssa_tsdpd_bvf_<style>.cpp init_one)
  all([&](sphbvf_ctx *c, int) {
    int rc;
    for (int i = 1; i <= ntypes; i++)
      if ((rc = sphbvf_set_type(c, i, atom->mass[i], rho0[i], soundspeed[i], G0[i]))) return rc;
    for (int i = 1; i <= ntypes; i++)
      for (int j = i; j <= ntypes; j++)
        if ((rc = sphbvf_set_pair(c, i, j, viscosity[i][j], cut[i][j], cutc[i][j], S ? kappa[i][j] : NULL))) return rc;
    for (int q = 0; q < nfixdesc; q++) {
      const FixDesc &f = fixdesc[q];
      rc = 0;
      switch (f.kind) {
        case 0: rc = sphbvf_add_buoyancy(c, f.groupbit, f.ia[0], f.a[0], f.ia[1], f.ia[2], f.a[1]); break;
        case 1: rc = sphbvf_add_forcing(c, f.groupbit, f.ia[0], (long)f.step, f.ia[1], f.ia[2], f.a[0], f.a[1], f.a[2], f.a[3], f.a[4]); break;
        case 2: rc = sphbvf_add_buffer(c, f.groupbit, f.ia[0], f.ia[2], (long)f.step, f.ia[1], f.a[0], f.a[1], f.a[2], f.a[3], f.a[4]); break;
        case 3: rc = sphbvf_add_setforce(c, f.groupbit, f.a[0], f.a[1], f.a[2]); break;
        case 4: {
          int r[2] = {f.ia[1] & 255, (f.ia[1] >> 8) & 255};
          int p[4] = {f.ia[2] & 255, (f.ia[2] >> 8) & 255, (f.ia[2] >> 16) & 255, (f.ia[2] >> 24) & 255};
          rc = sphbvf_add_chem_rxn(c, f.groupbit, f.a[0], f.ia[0] & 255, r, f.ia[0] >> 8, p);
          break;
        }
      }
      if (rc) return rc;
    }
    return 0;
  });

  // FixSsaTsdpdBvf*Cuda::setup_pre_force has just set vest = v and rhoI = rho on the host
  if (nranks == 1) {
    check(sphbvf_set_atoms(ctx, n, atom->tag, atom->type, atom->mask, atom->solid_tag, atom->fixed_tag,
                           n ? &atom->x[0][0] : NULL, n ? &atom->v[0][0] : NULL, atom->rho, atom->e,
                           (S && n) ? &atom->C[0][0] : NULL, n ? &atom->deviatoricTensor[0][0][0] : NULL));
    if (n) {
      check(sphbvf_upload(ctx, SPHBVF_F_VEST, &atom->vest[0][0]));
      check(sphbvf_upload(ctx, SPHBVF_F_RHOI, atom->rhoI));
    }
  } else {
    // the atoms of each brick (domain.cpp:308-330 sub-box rule: lo <= x < hi, the last brick also takes x == boxhi)
    std::vector<std::vector<int> > mine(nranks);
    std::vector<double> lo(3 * (size_t)nranks), hi(3 * (size_t)nranks);
    for (int r = 0; r < nranks; r++) check(sphbvf_brick_bounds(&cfg, r, &lo[3 * r], &hi[3 * r]));
    const int P[3] = {cfg.procgrid[0], cfg.procgrid[1], cfg.procgrid[2]};
    int maxtag = 0;
    for (int i = 0; i < n; i++) {
      int idx[3] = {0, 0, 0};
      for (int k = 0; k < cfg.dim; k++) {
        // bricks along dimension k: rank index stride follows sphbvf_brick_bounds (x fastest)
        const int stride = k == 0 ? 1 : (k == 1 ? P[0] : P[0] * P[1]);
        int q = 0;
        while (q + 1 < P[k] && atom->x[i][k] >= hi[3 * (size_t)(q * stride) + k]) q++;
        idx[k] = q;
      }
      mine[idx[0] + P[0] * (idx[1] + P[1] * idx[2])].push_back(i);
      if (atom->tag[i] > maxtag) maxtag = atom->tag[i];
    }
    tag2idx.assign((size_t)maxtag + 1, -1);
    for (int i = 0; i < n; i++) tag2idx[atom->tag[i]] = i;
    all([&](sphbvf_ctx *c, int r) {
      const std::vector<int> &ix = mine[r];
      const int m = (int)ix.size();
      std::vector<int> tag(m), type(m), mask(m), solid(m), fixed(m);
      std::vector<double> x(3 * (size_t)m), v(3 * (size_t)m), vest(3 * (size_t)m), rho(m), rhoI(m), e(m),
          C((size_t)(S ? S : 1) * m), dev(9 * (size_t)m);
      for (int q = 0; q < m; q++) {
        const int i = ix[q];
        tag[q] = atom->tag[i]; type[q] = atom->type[i]; mask[q] = atom->mask[i];
        solid[q] = atom->solid_tag[i]; fixed[q] = atom->fixed_tag[i];
        for (int k = 0; k < 3; k++) { x[3 * (size_t)q + k] = atom->x[i][k]; v[3 * (size_t)q + k] = atom->v[i][k]; vest[3 * (size_t)q + k] = atom->vest[i][k]; }
        rho[q] = atom->rho[i]; rhoI[q] = atom->rhoI[i]; e[q] = atom->e[i];
        for (int k = 0; k < S; k++) C[(size_t)q * S + k] = atom->C[i][k];
        for (int k = 0; k < 9; k++) dev[9 * (size_t)q + k] = (&atom->deviatoricTensor[i][0][0])[k];
      }
      int rc = sphbvf_set_atoms(c, m, tag.data(), type.data(), mask.data(), solid.data(), fixed.data(), x.data(), v.data(),
                                rho.data(), e.data(), S ? C.data() : NULL, dev.data());
      if (rc || !m) return rc;
      if ((rc = sphbvf_upload(c, SPHBVF_F_VEST, vest.data()))) return rc;
      return sphbvf_upload(c, SPHBVF_F_RHOI, rhoI.data());
    });
  }
  // stochastic stress (active only if some ssa_tsdpd/e != 0): kB of the unit system and a seed; upstream
  // seeds from clock() (pair_ssa_tsdpd_bvf_transport_velocity.cpp:957-959), here runs are reproducible
  // unless SPHBVF_SEED is changed
  {
    const char *env = getenv("SPHBVF_SEED");
    const unsigned long long seed = env ? strtoull(env, NULL, 10) : 20261018ULL;
    const double kb = force->boltz;
    const long step = (long)update->ntimestep, nsteps = (long)update->nsteps;
    all([&](sphbvf_ctx *c, int) {
      int rc;
      if ((rc = sphbvf_set_random(c, kb, seed))) return rc;
      if ((rc = sphbvf_set_timestep(c, step))) return rc;
      if ((rc = sphbvf_set_run_length(c, nsteps))) return rc;
      return sphbvf_setup_neighbors(c);
    });
  }
  if (getenv("SPHBVF_PROFILE")) all([&](sphbvf_ctx *c, int) { return sphbvf_set_profiling(c, 1); });
  nlocal_uploaded = n;
  host_mask = 0;
  ndownloads = ndevice_thermo = nskipped = 0;
  nbytes_down = 0;
  // the host copy is stale from now on: LAMMPS must not reorder it behind the device's back
  atom->sortfreq = 0;
  if (nranks > 1 && comm->me == 0) {
    char msg[128];
    snprintf(msg, sizeof msg, "sphbvf: %d GPUs, brick grid %d x %d x %d\n", nranks, cfg.procgrid[0], cfg.procgrid[1], cfg.procgrid[2]);
    if (screen) fputs(msg, screen);
    if (logfile) fputs(msg, logfile);
  }
}

/* ---------------------------------------------------------------------- */

namespace {
struct HostFieldDesc { unsigned bit; int id, ncols; };
}

/* the host array of one field (NULL: not owned by this atom style / variant / species count) */

static double *host_array(Atom *atom, unsigned bit)
{
  switch (bit) {
    case SphbvfLmp::HF_X: return &atom->x[0][0];
    case SphbvfLmp::HF_V: return &atom->v[0][0];
    case SphbvfLmp::HF_VEST: return &atom->vest[0][0];
    case SphbvfLmp::HF_F: return &atom->f[0][0];
    case SphbvfLmp::HF_RHO: return atom->rho;
    case SphbvfLmp::HF_RHOI: return atom->rhoI;
    case SphbvfLmp::HF_DRHO: return atom->drho;
    case SphbvfLmp::HF_PHI: return atom->phi;
    case SphbvfLmp::HF_ND: return atom->number_density;
    case SphbvfLmp::HF_NW: return &atom->nw[0][0];
    case SphbvfLmp::HF_DDV: return &atom->ddv[0][0];
    case SphbvfLmp::HF_RAUX1: return atom->rhoAux1;
    case SphbvfLmp::HF_RAUX2: return atom->rhoAux2;
    case SphbvfLmp::HF_DEV: return &atom->deviatoricTensor[0][0][0];
    case SphbvfLmp::HF_DDEV: return &atom->ddeviatoricTensor[0][0][0];
    case SphbvfLmp::HF_DDX: return &atom->ddx[0][0];
    case SphbvfLmp::HF_PNEW: return atom->Pnew;
    case SphbvfLmp::HF_C: return &atom->C[0][0];
    case SphbvfLmp::HF_Q: return &atom->Q[0][0];
  }
  return NULL;
}

/* lazy host mirrors (atom_style ssa_tsdpd/atomic/cuda): the arrays of the pair-sweep outputs named in mask */

namespace {
const struct { unsigned bit; int derived; } lazy_map[] = {
  {SphbvfLmp::HF_DRHO, AtomVecSsaTsdpdAtomicCuda::DRHO}, {SphbvfLmp::HF_PHI, AtomVecSsaTsdpdAtomicCuda::PHI},
  {SphbvfLmp::HF_ND, AtomVecSsaTsdpdAtomicCuda::NUMBER_DENSITY}, {SphbvfLmp::HF_NW, AtomVecSsaTsdpdAtomicCuda::NW},
  {SphbvfLmp::HF_DDV, AtomVecSsaTsdpdAtomicCuda::DDV}, {SphbvfLmp::HF_RAUX1, AtomVecSsaTsdpdAtomicCuda::RHOAUX1},
  {SphbvfLmp::HF_RAUX2, AtomVecSsaTsdpdAtomicCuda::RHOAUX2}, {SphbvfLmp::HF_DDEV, AtomVecSsaTsdpdAtomicCuda::DDEV},
  {SphbvfLmp::HF_DDX, AtomVecSsaTsdpdAtomicCuda::DDX}, {SphbvfLmp::HF_PNEW, AtomVecSsaTsdpdAtomicCuda::PNEW},
  {SphbvfLmp::HF_Q, AtomVecSsaTsdpdAtomicCuda::Q}};
}

void SphbvfLmp::materialize(unsigned mask)
{
  AtomVecSsaTsdpdAtomicCuda *av = dynamic_cast<AtomVecSsaTsdpdAtomicCuda *>(atom->avec);
  if (!av) return;   // stock atom style: everything is allocated
  for (size_t q = 0; q < sizeof lazy_map / sizeof lazy_map[0]; q++)
    if (mask & lazy_map[q].bit) av->materialize(lazy_map[q].derived);
}

void SphbvfLmp::host_fields(Atom *atom, unsigned mask)
{
  if (the_engine) {
    the_engine->materialize(mask);
    the_engine->fetch(mask);
    return;
  }
  AtomVecSsaTsdpdAtomicCuda *av = dynamic_cast<AtomVecSsaTsdpdAtomicCuda *>(atom->avec);
  if (!av) return;
  for (size_t q = 0; q < sizeof lazy_map / sizeof lazy_map[0]; q++)
    if (mask & lazy_map[q].bit) av->materialize(lazy_map[q].derived);
}

void SphbvfLmp::fetch(unsigned want)
{
  if (!ctx) return;
  const int n = atom->nlocal, S = atom->num_sdpd_species;
  // fields this run does not have: they count as current (nothing to copy)
  unsigned absent = 0;
  if (variant == SPHBVF_TV) absent |= HF_DDX | HF_PNEW;
  if (!S) absent |= HF_C | HF_Q;
  host_mask |= absent;
  unsigned need = want & HF_ALL & ~host_mask;
  if (!need) return;
  materialize(need);
  if (n != nlocal_uploaded) error->one(FLERR, "Atom count changed during a /cuda run");
  if (n) {
    if (nranks > 1) fetch_multi(need);
    else {
      static const HostFieldDesc tab[] = {
        {HF_X, SPHBVF_F_X, 3}, {HF_V, SPHBVF_F_V, 3}, {HF_VEST, SPHBVF_F_VEST, 3}, {HF_F, SPHBVF_F_F, 3}, {HF_RHO, SPHBVF_F_RHO, 1},
        {HF_RHOI, SPHBVF_F_RHOI, 1}, {HF_DRHO, SPHBVF_F_DRHO, 1}, {HF_PHI, SPHBVF_F_PHI, 1}, {HF_ND, SPHBVF_F_NUMBER_DENSITY, 1},
        {HF_NW, SPHBVF_F_NW, 3}, {HF_DDV, SPHBVF_F_DDV, 3}, {HF_RAUX1, SPHBVF_F_RHOAUX1, 1}, {HF_RAUX2, SPHBVF_F_RHOAUX2, 1},
        {HF_DEV, SPHBVF_F_DEV, 9}, {HF_DDEV, SPHBVF_F_DDEV, 9}, {HF_DDX, SPHBVF_F_DDX, 3}, {HF_PNEW, SPHBVF_F_PNEW, 1},
        {HF_C, SPHBVF_F_C, -1}, {HF_Q, SPHBVF_F_Q, -1}};
      for (size_t q = 0; q < sizeof tab / sizeof tab[0]; q++) {
        if (!(need & tab[q].bit)) continue;
        check(sphbvf_download(ctx, tab[q].id, host_array(atom, tab[q].bit)));
        nbytes_down += (bigint)8 * n * (tab[q].ncols < 0 ? S : tab[q].ncols);
      }
    }
  }
  host_mask |= need;
  if (need == (HF_ALL & ~absent) || want == HF_ALL) ndownloads++;
}

/* ----------------------------------------------------------------------
   several GPUs: atoms have migrated between bricks, so every rank hands back its rows in device order together
   with their tags and the rows are scattered into the host arrays by tag
------------------------------------------------------------------------- */

void SphbvfLmp::fetch_multi(unsigned need)
{
  const int S = atom->num_sdpd_species;
  struct Field { int id, ncols; double *host; };
  static const HostFieldDesc tab[] = {
    {HF_X, SPHBVF_F_X, 3}, {HF_V, SPHBVF_F_V, 3}, {HF_VEST, SPHBVF_F_VEST, 3}, {HF_F, SPHBVF_F_F, 3}, {HF_RHO, SPHBVF_F_RHO, 1},
    {HF_RHOI, SPHBVF_F_RHOI, 1}, {HF_DRHO, SPHBVF_F_DRHO, 1}, {HF_PHI, SPHBVF_F_PHI, 1}, {HF_ND, SPHBVF_F_NUMBER_DENSITY, 1},
    {HF_NW, SPHBVF_F_NW, 3}, {HF_DDV, SPHBVF_F_DDV, 3}, {HF_RAUX1, SPHBVF_F_RHOAUX1, 1}, {HF_RAUX2, SPHBVF_F_RHOAUX2, 1},
    {HF_DEV, SPHBVF_F_DEV, 9}, {HF_DDEV, SPHBVF_F_DDEV, 9}, {HF_DDX, SPHBVF_F_DDX, 3}, {HF_PNEW, SPHBVF_F_PNEW, 1},
    {HF_C, SPHBVF_F_C, -1}, {HF_Q, SPHBVF_F_Q, -1}};
  std::vector<Field> fields;
  for (size_t q = 0; q < sizeof tab / sizeof tab[0]; q++)
    if (need & tab[q].bit) fields.push_back({tab[q].id, tab[q].ncols < 0 ? S : tab[q].ncols, host_array(atom, tab[q].bit)});
  std::vector<int> count(nranks, 0);
  all([&](sphbvf_ctx *c, int r) {
    // each rank scatters its own rows: distinct tags, so the writes of different ranks never overlap
    const int m = sphbvf_nlocal(c);
    count[r] = m;
    if (!m) return 0;
    std::vector<int> tag(m);
    int rc = sphbvf_download_local(c, SPHBVF_F_TAG, tag.data(), m);
    if (rc) return rc;
    std::vector<double> buf(9 * (size_t)m);
    for (size_t q = 0; q < fields.size(); q++) {
      const int nc = fields[q].ncols;
      if ((size_t)nc * m > buf.size()) buf.resize((size_t)nc * m);
      if ((rc = sphbvf_download_local(c, fields[q].id, buf.data(), m))) return rc;
      double *host = fields[q].host;
      for (int a = 0; a < m; a++) {
        const int i = tag2idx[tag[a]];
        for (int k = 0; k < nc; k++) host[(size_t)i * nc + k] = buf[(size_t)a * nc + k];
      }
    }
    return 0;
  });
  bigint tot = 0;
  for (int r = 0; r < nranks; r++) tot += count[r];
  if (tot != atom->nlocal) error->one(FLERR, "Atoms lost or duplicated between the GPUs of a /cuda run");
  for (size_t q = 0; q < fields.size(); q++) nbytes_down += (bigint)8 * atom->nlocal * fields[q].ncols;
}

/* ----------------------------------------------------------------------
   Which host per-atom arrays can anything that fires on this output step read?  Conservative:
   (a) a restart due now, a thermo style other than one / multi (`custom` may reference variables and arbitrary
       computes, which cannot be inspected from here), a global compute other than temp/cuda, pressure, pe, or a
       fix that is not one of the /cuda fixes -> everything;
   (b) every dump that is due now contributes the columns it writes: dump custom / cfg-less styles are parsed
       keyword by keyword (id type mass: static data; x y z xs .. : positions; vx .. : velocities; fx .. : forces;
       c_ID: a /cuda per-atom compute fetches its own field when the dump invokes it, static per-atom computes
       such as ssa_tsdpd/solid_tag/atom need nothing, property/atom reads x / v / f and -- with atom_style
       ssa_tsdpd/atomic/cuda -- fetches the package's columns itself, any other compute -> everything;
       v_ f_ d_ i_ -> everything);
       dump atom / xyz read positions; any other dump style -> everything;
   (c) otherwise nothing: the thermo line is served from the device reductions.
   SPHBVF_OUTPUT=full restores the unconditional full download.
------------------------------------------------------------------------- */

namespace {
// legal access to two protected members of DumpCustom: &Peek::member has type `T DumpCustom::*`
struct DumpCustomPeek : public DumpCustom {
  static char *DumpCustom::*columns_ptr() { return &DumpCustomPeek::columns; }
  static int DumpCustom::*nthresh_ptr() { return &DumpCustomPeek::nthresh; }
};
}

unsigned SphbvfLmp::output_fields(bool at_setup)
{
  static const int force_full = [] { const char *e = getenv("SPHBVF_OUTPUT"); return e && strcmp(e, "full") == 0; }();
  if (force_full) return HF_ALL;
  const bigint now = update->ntimestep;
  if (output->restart_flag && (at_setup || output->next_restart == now)) return HF_ALL;
  if (!output->thermo) return HF_ALL;
  const char *ts = output->thermo->style;
  if (strcmp(ts, "one") != 0 && strcmp(ts, "multi") != 0) return HF_ALL;
  for (int i = 0; i < modify->ncompute; i++) {
    Compute *c = modify->compute[i];
    if (c->peratom_flag || c->local_flag) continue;
    if (strcmp(c->style, "temp/cuda") == 0 || strcmp(c->style, "pressure") == 0 || strcmp(c->style, "pe") == 0) continue;
    return HF_ALL;
  }
  for (int i = 0; i < modify->nfix; i++) {
    const char *fs = modify->fix[i]->style;
    const size_t n = strlen(fs);
    if (n < 5 || strcmp(fs + n - 5, "/cuda") != 0) return HF_ALL;
  }
  unsigned mask = 0;
  for (int i = 0; i < output->ndump; i++) {
    // Fix::setup runs before Output::setup has scheduled this run's dumps: any dump may fire at the first step
    if (!at_setup && output->next_dump[i] != now) continue;
    Dump *dp = output->dump[i];
    if (strcmp(dp->style, "atom") == 0 || strcmp(dp->style, "xyz") == 0) { mask |= HF_X; continue; }
    if (strcmp(dp->style, "custom") != 0) return HF_ALL;
    DumpCustom *dc = (DumpCustom *)dp;
    if (dc->*DumpCustomPeek::nthresh_ptr() > 0) return HF_ALL;
    const char *cols = dc->*DumpCustomPeek::columns_ptr();
    if (!cols) return HF_ALL;
    std::string all(cols);
    size_t pos = 0;
    while (pos < all.size()) {
      size_t e = all.find(' ', pos);
      if (e == std::string::npos) e = all.size();
      const std::string w = all.substr(pos, e - pos);
      pos = e + 1;
      if (w.empty()) continue;
      if (w == "id" || w == "type" || w == "mass" || w == "mol" || w == "proc" || w == "procp1" || w == "element") continue;
      if (w == "x" || w == "y" || w == "z" || w == "xs" || w == "ys" || w == "zs" || w == "xu" || w == "yu" || w == "zu" ||
          w == "xsu" || w == "ysu" || w == "zsu" || w == "ix" || w == "iy" || w == "iz") { mask |= HF_X; continue; }
      if (w == "vx" || w == "vy" || w == "vz") { mask |= HF_V; continue; }
      if (w == "fx" || w == "fy" || w == "fz") { mask |= HF_F; continue; }
      if (w.compare(0, 2, "c_") == 0) {
        std::string id = w.substr(2);
        const size_t br = id.find('[');
        if (br != std::string::npos) id = id.substr(0, br);
        const int ic = modify->find_compute(id.c_str());
        if (ic < 0) return HF_ALL;
        const char *cs = modify->compute[ic]->style;
        const size_t n = strlen(cs);
        if (n >= 5 && strcmp(cs + n - 5, "/cuda") == 0) continue;              // fetches its own field
        if (strcmp(cs, "ssa_tsdpd/solid_tag/atom") == 0) continue;               // static per-atom data
        if (strcmp(cs, "property/atom") == 0 && dynamic_cast<AtomVecSsaTsdpdAtomicCuda *>(atom->avec)) {
          // core attributes read x / v / f; the package's own (rho drho e de cv phi Pnew ...) go through
          // AtomVecSsaTsdpdAtomicCuda::pack_property_atom, which fetches its column itself
          mask |= HF_X | HF_V | HF_F;
          continue;
        }
        return HF_ALL;
      }
      return HF_ALL;   // v_ f_ d_ i_ q mux ... : cannot be inspected
    }
  }
  if (!mask) nskipped++;
  return mask;
}

void SphbvfLmp::stop()
{
  if (!ctx) return;
  if (getenv("SPHBVF_VERBOSE") && comm->me == 0) {
    char msg[256];
    snprintf(msg, sizeof msg, "sphbvf: " BIGINT_FORMAT " full downloads, " BIGINT_FORMAT " output steps served from the device, "
             BIGINT_FORMAT " device kinetic-energy reductions, " BIGINT_FORMAT " bytes device->host before the final sync\n",
             ndownloads, nskipped, ndevice_thermo, nbytes_down);
    if (screen) fputs(msg, screen);
    if (logfile) fputs(msg, logfile);
  }
  if (getenv("SPHBVF_PROFILE") && comm->me == 0) {
    // per-kernel-family device time of GPU 0 (CUDA events on the context's stream, sphbvf_set_profiling)
    static const char *fam[] = {"pair", "initial_integrate", "final_integrate", "neighbor_rebuild", "pack_halo", "fixes",
                                "final_initial_pack_fused"};
    sphbvf_ctx *c0 = nranks == 1 ? ctx : ctxs[0];
    std::string msg = "sphbvf profile (GPU 0, ms / launches):";
    for (int k = 0; k < 7; k++) {
      long nl = 0;
      const double ms = sphbvf_kernel_ms(c0, k, &nl);
      char one[96];
      snprintf(one, sizeof one, " %s %.3f / %ld;", fam[k], ms, nl);
      msg += one;
    }
    msg += "\n";
    if (screen) fputs(msg.c_str(), screen);
    if (logfile) fputs(msg.c_str(), logfile);
  }
  // lazy host mirrors: copy back the state and the outputs that are mirrored; nothing ever asked for the rest
  AtomVecSsaTsdpdAtomicCuda *av = dynamic_cast<AtomVecSsaTsdpdAtomicCuda *>(atom->avec);
  if (av && getenv("SPHBVF_VERBOSE") && comm->me == 0) {
    char msg[256];
    snprintf(msg, sizeof msg, "sphbvf: host mirrors of %d of the %d pair-sweep output arrays were allocated "
             "(atom_style ssa_tsdpd/atomic/cuda), %d bytes of host arrays per atom slot\n",
             (int)AtomVecSsaTsdpdAtomicCuda::NDERIVED - av->nlazy(), (int)AtomVecSsaTsdpdAtomicCuda::NDERIVED,
             (int)(av->host_bytes() / MAX(1, atom->nmax)));
    if (screen) fputs(msg, screen);
    if (logfile) fputs(msg, logfile);
  }
  unsigned want = HF_ALL;
  if (av)
    for (size_t q = 0; q < sizeof lazy_map / sizeof lazy_map[0]; q++)
      if (!av->materialized(lazy_map[q].derived)) want &= ~lazy_map[q].bit;
  fetch(want);
  destroy_contexts();
}

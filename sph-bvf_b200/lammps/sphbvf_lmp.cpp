/* ----------------------------------------------------------------------
   sphbvf_lmp.cpp -- see sphbvf_lmp.h.  Replaces, for the /cuda styles, what Verlet::setup does on
   the host between Atom and the styles (verlet.cpp:88-170): here the atoms go to the device once
   and stay there.
------------------------------------------------------------------------- */

#include <stdlib.h>
#include <string.h>
#include "sphbvf_lmp.h"
#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "lammps.h"
#include "memory.h"
#include "compute.h"
#include "fix.h"
#include "modify.h"
#include "neighbor.h"
#include "output.h"
#include "pair.h"
#include "thermo.h"
#include "update.h"

using namespace LAMMPS_NS;

static SphbvfLmp *the_engine = NULL;   // one LAMMPS instance per process in this build

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
  host_current = 1;
  nlocal_uploaded = 0;
  ndownloads = ndevice_thermo = nskipped = 0;
}

SphbvfLmp::~SphbvfLmp()
{
  if (ctx) sphbvf_destroy(ctx);
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
  cfg.device = 0;
  cfg.rank = 0;
  cfg.nranks = 1;
  int rc = sphbvf_create(&cfg, &ctx);
  if (rc) {
    ctx = NULL;
    error->all(FLERR, "sphbvf_create failed (more than 4 atom types / species, or no usable GPU)");
  }

  // Pair::coeff / init_one already mirrored [i][j] -> [j][i] (pair_ssa_tsdpd_bvf_<style>.cpp init_one)
  for (int i = 1; i <= ntypes; i++) check(sphbvf_set_type(ctx, i, atom->mass[i], rho0[i], soundspeed[i], G0[i]));
  for (int i = 1; i <= ntypes; i++)
    for (int j = i; j <= ntypes; j++)
      check(sphbvf_set_pair(ctx, i, j, viscosity[i][j], cut[i][j], cutc[i][j], S ? kappa[i][j] : NULL));
  for (int q = 0; q < nfixdesc; q++) {
    const FixDesc &f = fixdesc[q];
    switch (f.kind) {
      case 0: check(sphbvf_add_buoyancy(ctx, f.groupbit, f.ia[0], f.a[0], f.ia[1], f.ia[2], f.a[1])); break;
      case 1: check(sphbvf_add_forcing(ctx, f.groupbit, f.ia[0], (long)f.step, f.ia[1], f.ia[2], f.a[0], f.a[1], f.a[2], f.a[3], f.a[4])); break;
      case 2: check(sphbvf_add_buffer(ctx, f.groupbit, f.ia[0], f.ia[2], (long)f.step, f.ia[1], f.a[0], f.a[1], f.a[2], f.a[3], f.a[4])); break;
      case 3: check(sphbvf_add_setforce(ctx, f.groupbit, f.a[0], f.a[1], f.a[2])); break;
      case 4: {
        int r[2] = {f.ia[1] & 255, (f.ia[1] >> 8) & 255};
        int p[4] = {f.ia[2] & 255, (f.ia[2] >> 8) & 255, (f.ia[2] >> 16) & 255, (f.ia[2] >> 24) & 255};
        check(sphbvf_add_chem_rxn(ctx, f.groupbit, f.a[0], f.ia[0] & 255, r, f.ia[0] >> 8, p));
        break;
      }
    }
  }

  const int n = atom->nlocal;
  check(sphbvf_set_atoms(ctx, n, atom->tag, atom->type, atom->mask, atom->solid_tag, atom->fixed_tag,
                         n ? &atom->x[0][0] : NULL, n ? &atom->v[0][0] : NULL, atom->rho, atom->e,
                         (S && n) ? &atom->C[0][0] : NULL, n ? &atom->deviatoricTensor[0][0][0] : NULL));
  if (n) {
    // FixSsaTsdpdBvf*Cuda::setup_pre_force has just set vest = v and rhoI = rho on the host
    check(sphbvf_upload(ctx, SPHBVF_F_VEST, &atom->vest[0][0]));
    check(sphbvf_upload(ctx, SPHBVF_F_RHOI, atom->rhoI));
  }
  // stochastic stress (active only if some ssa_tsdpd/e != 0): kB of the unit system and a seed; upstream
  // seeds from clock() (pair_ssa_tsdpd_bvf_transport_velocity.cpp:957-959), here runs are reproducible
  // unless SPHBVF_SEED is changed
  {
    const char *env = getenv("SPHBVF_SEED");
    const unsigned long long seed = env ? strtoull(env, NULL, 10) : 20261018ULL;
    check(sphbvf_set_random(ctx, force->boltz, seed));
  }
  check(sphbvf_set_timestep(ctx, (long)update->ntimestep));
  check(sphbvf_set_run_length(ctx, (long)update->nsteps));
  check(sphbvf_setup_neighbors(ctx));
  nlocal_uploaded = n;
  host_current = 0;
  // the host copy is stale from now on: LAMMPS must not reorder it behind the device's back
  atom->sortfreq = 0;
}

/* ---------------------------------------------------------------------- */

void SphbvfLmp::to_host()
{
  if (!ctx || host_current) return;
  const int n = atom->nlocal, S = atom->num_sdpd_species;
  if (n != nlocal_uploaded) error->one(FLERR, "Atom count changed during a /cuda run");
  if (n) {
    check(sphbvf_download(ctx, SPHBVF_F_X, &atom->x[0][0]));
    check(sphbvf_download(ctx, SPHBVF_F_V, &atom->v[0][0]));
    check(sphbvf_download(ctx, SPHBVF_F_VEST, &atom->vest[0][0]));
    check(sphbvf_download(ctx, SPHBVF_F_F, &atom->f[0][0]));
    check(sphbvf_download(ctx, SPHBVF_F_RHO, atom->rho));
    check(sphbvf_download(ctx, SPHBVF_F_RHOI, atom->rhoI));
    check(sphbvf_download(ctx, SPHBVF_F_DRHO, atom->drho));
    check(sphbvf_download(ctx, SPHBVF_F_PHI, atom->phi));
    check(sphbvf_download(ctx, SPHBVF_F_NUMBER_DENSITY, atom->number_density));
    check(sphbvf_download(ctx, SPHBVF_F_NW, &atom->nw[0][0]));
    check(sphbvf_download(ctx, SPHBVF_F_DDV, &atom->ddv[0][0]));
    check(sphbvf_download(ctx, SPHBVF_F_RHOAUX1, atom->rhoAux1));
    check(sphbvf_download(ctx, SPHBVF_F_RHOAUX2, atom->rhoAux2));
    check(sphbvf_download(ctx, SPHBVF_F_DEV, &atom->deviatoricTensor[0][0][0]));
    check(sphbvf_download(ctx, SPHBVF_F_DDEV, &atom->ddeviatoricTensor[0][0][0]));
    if (variant != SPHBVF_TV) {
      check(sphbvf_download(ctx, SPHBVF_F_DDX, &atom->ddx[0][0]));
      check(sphbvf_download(ctx, SPHBVF_F_PNEW, atom->Pnew));
    }
    if (S) {
      check(sphbvf_download(ctx, SPHBVF_F_C, &atom->C[0][0]));
      check(sphbvf_download(ctx, SPHBVF_F_Q, &atom->Q[0][0]));
    }
  }
  host_current = 1;
  ndownloads++;
}

/* ----------------------------------------------------------------------
   Does anything that fires on this output step read the host per-atom arrays?  Conservative: the download is
   skipped only when (a) no dump and no restart is due now, (b) thermo has one of the fixed keyword sets
   (style one / multi: step temp epair emol etotal press ..., all fed by temp, pe and pressure computes;
   `custom` may reference variables and arbitrary computes, which cannot be inspected from here), (c) every
   compute LAMMPS knows is temp/cuda, pressure or pe, or a per-atom / local compute (only dumps, fixes and
   variables invoke those), and (d) every fix is one of the /cuda fixes (none of them reads host arrays).
   SPHBVF_OUTPUT=full restores the unconditional download.
------------------------------------------------------------------------- */

bool SphbvfLmp::output_needs_host(bool at_setup)
{
  static const int force_full = [] { const char *e = getenv("SPHBVF_OUTPUT"); return e && strcmp(e, "full") == 0; }();
  if (force_full) return true;
  const bigint now = update->ntimestep;
  // Fix::setup runs before Output::setup has scheduled this run's dumps: any dump may fire at the first step
  if (at_setup && (output->ndump || output->restart_flag)) return true;
  if (output->ndump && output->next_dump_any == now) return true;
  if (output->restart_flag && output->next_restart == now) return true;
  if (!output->thermo) return true;
  const char *ts = output->thermo->style;
  if (strcmp(ts, "one") != 0 && strcmp(ts, "multi") != 0) return true;
  for (int i = 0; i < modify->ncompute; i++) {
    Compute *c = modify->compute[i];
    if (c->peratom_flag || c->local_flag) continue;
    if (strcmp(c->style, "temp/cuda") == 0 || strcmp(c->style, "pressure") == 0 || strcmp(c->style, "pe") == 0) continue;
    return true;
  }
  for (int i = 0; i < modify->nfix; i++) {
    const char *fs = modify->fix[i]->style;
    const size_t n = strlen(fs);
    if (n < 5 || strcmp(fs + n - 5, "/cuda") != 0) return true;
  }
  nskipped++;
  return false;
}

void SphbvfLmp::stop()
{
  if (!ctx) return;
  if (getenv("SPHBVF_VERBOSE") && comm->me == 0) {
    char msg[256];
    snprintf(msg, sizeof msg, "sphbvf: " BIGINT_FORMAT " full downloads, " BIGINT_FORMAT " output steps served from the device, "
             BIGINT_FORMAT " device kinetic-energy reductions\n", ndownloads, nskipped, ndevice_thermo);
    if (screen) fputs(msg, screen);
    if (logfile) fputs(msg, logfile);
  }
  to_host();
  sphbvf_destroy(ctx);
  ctx = NULL;
}

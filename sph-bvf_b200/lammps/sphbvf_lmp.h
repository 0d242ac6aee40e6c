/* -*- c++ -*- ----------------------------------------------------------
   sphbvf_lmp.h -- glue between the LAMMPS-style "/cuda" classes of this directory and the C ABI
   of libsphbvf.so (include/sphbvf.h).  One engine per LAMMPS instance: it owns the sphbvf_ctx,
   uploads the atoms of class Atom when a run starts, and copies device state back into the
   host arrays only when LAMMPS is about to read them (thermo / dump steps, end of run).

   One LAMMPS process (comm->nprocs == 1) drives ALL GPUs of the box: the engine owns
   one sphbvf_ctx per GPU, each a rank of the library's brick decomposition (NCCL halo + migration
   inside libsphbvf.so), and one worker thread per context; every hook fans out to the workers and
   joins.  SPHBVF_NGPU sets the count (default: all visible GPUs if the system has at least 10^6
   atoms per GPU, else 1).

   These files are meant to be dropped into the reference's src/ (see INTEGRATION.md): they use
   only LAMMPS' public class interfaces of the 22Aug2018 fork and the C ABI -- no CUDA headers.
------------------------------------------------------------------------- */

#ifndef LMP_SPHBVF_LMP_H
#define LMP_SPHBVF_LMP_H

#include <functional>
#include <vector>
#include "pointers.h"
#include "sphbvf.h"

namespace LAMMPS_NS {

class SphbvfLmp : protected Pointers {
 public:
  static SphbvfLmp *get(class LAMMPS *);   // created on first use, destroyed with the pair style
  static SphbvfLmp *peek();                // the engine if one exists (computes must not create it)
  static void release(class LAMMPS *);

  SphbvfLmp(class LAMMPS *);
  ~SphbvfLmp();

  // ---- filled by the style classes before a run starts
  int variant;                       // enum sphbvf_variant, set by the pair style
  class Pair *pair;                  // the /cuda pair style (owner of the coefficient arrays)
  double *rho0, *soundspeed, *G0;    // [ntypes+1], owned by the pair style
  double **viscosity, **cut, **cutc; // [ntypes+1][ntypes+1]
  double ***kappa;                   // [ntypes+1][ntypes+1][S]
  int integrate_groupbit;            // set by the integrator fix
  int nfixdesc;
  struct FixDesc { int kind, groupbit, ia[4]; bigint step; double a[6]; } fixdesc[16];

  void reset_fixes() { nfixdesc = 0; }         // Pair::init_style (force->init precedes modify->init)
  void add_fix(const FixDesc &);               // Fix::init of the auxiliary /cuda fixes, in Modify order

  // ---- the hooks, called by the pair style and the integrator fix
  void start();                 // first Pair::compute of a run: create ctx, upload, neighbour setup
  void stop();                  // end of run (Fix::post_run) or destruction: download, destroy ctx
  bool active() const { return ctx != NULL; }
  void check(int rc);           // rc != 0 -> error->one(FLERR, sphbvf_last_error())
  // ---- fan-out over the GPUs: f(ctx_r, r) on every rank concurrently (worker threads), status codes checked
  int nranks;
  std::vector<sphbvf_ctx *> ctxs;
  void all(const std::function<int(sphbvf_ctx *, int)> &f);
  void call(int (*fn)(sphbvf_ctx *)) { all([fn](sphbvf_ctx *c, int) { return fn(c); }); }
  void set_timestep(bigint n) { all([n](sphbvf_ctx *c, int) { return sphbvf_set_timestep(c, (long)n); }); }
  void set_dt(double dt) { all([dt](sphbvf_ctx *c, int) { return sphbvf_set_dt(c, dt); }); }
  void neighbor_step(int *rebuilt);
  void virial(double *v6);                    // summed over the ranks
  void ke_tensor(int groupbit, double *t6);   // summed over the ranks
  double max_vsq(int groupbit);
  // ---- host mirrors.  The device owns the state during a run; a host array of class Atom is refreshed only when
  // something is about to read it.  Fields are tracked one by one (HF_* bits), so a dump step copies the columns
  // the dump writes and nothing else.
  enum HostField {
    HF_X = 1 << 0, HF_V = 1 << 1, HF_VEST = 1 << 2, HF_F = 1 << 3, HF_RHO = 1 << 4, HF_RHOI = 1 << 5, HF_DRHO = 1 << 6,
    HF_PHI = 1 << 7, HF_ND = 1 << 8, HF_NW = 1 << 9, HF_DDV = 1 << 10, HF_RAUX1 = 1 << 11, HF_RAUX2 = 1 << 12,
    HF_DEV = 1 << 13, HF_DDEV = 1 << 14, HF_DDX = 1 << 15, HF_PNEW = 1 << 16, HF_C = 1 << 17, HF_Q = 1 << 18,
    HF_ALL = (1 << 19) - 1
  };
  void fetch(unsigned mask);    // device -> class Atom arrays for the fields of `mask` that are not current yet
  // With atom_style ssa_tsdpd/atomic/cuda the host arrays of the pair-sweep outputs (drho, phi, nw, ...) exist only once
  // something asked for them (AtomVecSsaTsdpdAtomicCuda::materialize); stop() copies back the state fields and the
  // outputs that are mirrored.  (LAMMPS refuses to evaluate a compute between runs unless it was invoked on the run's
  // last step -- "Compute used in dump between runs is not current" -- so nothing can ask for a new column afterwards.)
  static void host_fields(class Atom *, unsigned mask);   // materialise + fetch: what a reader of host arrays calls first
  void to_host() { fetch(HF_ALL); }   // every field the package owns
  void mark_dirty() { host_mask = 0; }
  bool host_is_current() const { return (host_mask & HF_ALL) == HF_ALL; }
  bool host_has(unsigned mask) const { return (host_mask & mask) == mask; }
  // output step: the fields anything scheduled now may read from the host arrays.  0 when thermo of style one /
  // multi is fed by compute temp/cuda + the device virial and no dump or restart is due; the columns of the dumps
  // that are due (dump custom / atom / xyz; /cuda per-atom computes fetch their own field) otherwise; HF_ALL when
  // something that cannot be inspected is involved (stock fixes, variables, other dump styles, restarts ...)
  unsigned output_fields(bool at_setup = false);
  bigint nbytes_down;           // bytes copied device -> host so far in this run (SPHBVF_VERBOSE)
  void count_device_thermo() { ndevice_thermo++; }
  sphbvf_ctx *ctx;

 private:
  unsigned host_mask;           // HF_* bits of the fields that are current on the host
  void destroy_contexts();
  void materialize(unsigned mask);   // lazy host mirrors: allocate the arrays of the derived fields in mask
  int nlocal_uploaded;
  class SphbvfWorkers *workers;
  std::vector<int> tag2idx;   // atom tag -> row of the host arrays (multi-GPU: atoms migrate between ranks)
  void fetch_multi(unsigned mask);
  bigint ndownloads, ndevice_thermo, nskipped;   // statistics printed at the end of a run (SPHBVF_VERBOSE)
};

}

#endif

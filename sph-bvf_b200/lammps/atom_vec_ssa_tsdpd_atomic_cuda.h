/* -*- c++ -*- ----------------------------------------------------------
   atom_style ssa_tsdpd/atomic/cuda -- picked under "-sf cuda" for `atom_style ssa_tsdpd/atomic S [Nssa Nrxn]`
   (atom.cpp:549-568 tries <style>/<suffix> first).

   Same per-atom dictionary as AtomVecSsaTsdpdAtomic (atom_vec_ssa_tsdpd_atomic.h:58-85, .cpp:116-189), same
   script-level contract (arguments, `set ssa_tsdpd/...`, property/atom names, data-file columns), but the host
   arrays are LAZY MIRRORS of what lives on the GPU:

     state    tag type mask image x v f vest rho e cv rhoI C deviatoricTensor solid_tag fixed_tag
              -- allocated with the atoms, like upstream (what create_atoms / set / velocity / displace_atoms write
              and what a run uploads);
     derived  drho de Q phi number_density nw v_weighted_solid a_weighted_solid ddeviatoricTensor
              artificialStressTensor ddx ddv Pold Pnew Aaux Baux APaux fP rhoAux1..3
              -- outputs of the pair sweeps.  Upstream allocates all of them for every atom (~ 830 B/atom with the row
              pointer tables); here each one is allocated the first time something on the HOST is about to read it:
              SphbvfLmp::fetch (a dump column, a /cuda per-atom compute, an uninspectable consumer), property/atom,
              or AtomVec::init() of a run whose pair style is not a /cuda style (then everything is materialised
              and the class behaves exactly like the reference one, so the suffix fallback keeps working).

   A /cuda run of 64 M atoms therefore needs ~ 290 B/atom of host memory (18 GB) instead of ~ 53 GB.

   The pack / unpack / copy / clear functions are generic loops over one field table (name, Atom member, shape, which
   messages carry it), in the reference's message layouts: forward x v rho e vest C dev rhoI; border adds tag type mask
   cv solid_tag fixed_tag; exchange adds image; reverse = f + every derived field that is allocated.  SSA species
   (serial-only upstream, out of scope of the GPU path) are not held: use `suffix off` around atom_style for those decks.
------------------------------------------------------------------------- */

#ifdef ATOM_CLASS

AtomStyle(ssa_tsdpd/atomic/cuda,AtomVecSsaTsdpdAtomicCuda)

#else

#ifndef LMP_ATOM_VEC_SSA_TSDPD_ATOMIC_CUDA_H
#define LMP_ATOM_VEC_SSA_TSDPD_ATOMIC_CUDA_H

#include "atom_vec.h"

namespace LAMMPS_NS {

class AtomVecSsaTsdpdAtomicCuda : public AtomVec {
 public:
  AtomVecSsaTsdpdAtomicCuda(class LAMMPS *);
  ~AtomVecSsaTsdpdAtomicCuda() {}
  void process_args(int, char **);
  void init();
  void grow(int);
  void grow_reset() {}
  void copy(int, int, int);
  void force_clear(int, size_t);
  int pack_comm(int, int *, double *, int, int *);
  int pack_comm_vel(int, int *, double *, int, int *);
  void unpack_comm(int, int, double *);
  void unpack_comm_vel(int, int, double *);
  int pack_reverse(int, int, double *);
  void unpack_reverse(int, int *, double *);
  int pack_border(int, int *, double *, int, int *);
  int pack_border_vel(int, int *, double *, int, int *);
  void unpack_border(int, int, double *);
  void unpack_border_vel(int, int, double *);
  int pack_exchange(int, double *);
  int unpack_exchange(double *);
  int size_restart();
  int pack_restart(int, double *);
  int unpack_restart(double *);
  void create_atom(int, double *);
  void data_atom(double *, imageint, char **);
  void pack_data(double **);
  void write_data(FILE *, int, double **);
  int property_atom(char *);
  void pack_property_atom(int, double *, int, int);
  bigint memory_usage();

  // ---- lazy mirrors
  // field ids of the derived group, for materialize(); the order is the reference's reverse-communication order
  enum Derived { DRHO, DE, Q, DDEV, ARTSTRESS, PHI, NUMBER_DENSITY, NW, VWS, AWS, DDX, DDV, POLD, PNEW, AAUX, BAUX,
                 APAUX, FP, RHOAUX1, RHOAUX2, RHOAUX3, NDERIVED };
  void materialize(int which);      // allocate (zero-filled) the host array of one derived field, if it is not there
  void materialize_all();
  bool materialized(int which) const;
  int nlazy() const;                // derived fields that are still unallocated
  bigint host_bytes(bool from_atom_memory_usage = false);   // bytes of the per-atom arrays that are allocated

 private:
  enum Shape { I1, D1, D2, D33 };
  enum Msg { FWD = 1, REV = 2, BRD = 4, EXC = 8, CLR = 16, STATE = 32 };
  struct Field {
    const char *name;
    void *slot;        // address of the Atom member (int *, double *, double ** or double ***)
    int shape, cols;   // cols: doubles per atom
    unsigned in;       // Msg bits
    double init;       // value a new atom gets
  };
  enum { NSTATE = 16, NFIELD = NSTATE + NDERIVED };
  Field fld[NFIELD];
  int nfld;
  int order_fwd[12], order_brd[16], order_exc[16], order_rst[16];   // field indices, -1 terminated
  bool have(const Field &f) const { return *(void *const *)f.slot != NULL; }
  double *row(const Field &f, int i) const;    // the cols doubles of atom i (D1 / D2 / D33)
  int &ival(const Field &f, int i) const { return (*(int *const *)f.slot)[i]; }
  void alloc(Field &f, int n);
  void set_default(const Field &f, int i);
  int put(const int *order, int j, double *buf, const double *shift) const;
  int take(const int *order, int i, const double *buf);
  void vel_remap(int n, int *list, double *buf, int stride, int voff, int vestoff, const int *pbc) const;
  void build_table();
};

}

#endif
#endif

// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product.
//
// Driver for the UNMODIFIED reference library (oracle/_ref/liblammps_serial.a, built from
// /root/reference by oracle/Makefile).  It adds one fix style of our own, "sphbvf/snapshot", to
// the running LAMMPS instance through the public creator map (modify.h:172 fix_map) and then
// executes the input script exactly as the reference's main.cpp:54-56 does.  The fix writes, at
// setup (step 0, after the first force evaluation) and every N steps at end_of_step, a binary
// snapshot of every per-atom array of the USER-SSA-TSDPD package (atom.h:84-109) ordered by atom
// tag, plus the pair style's neighbour list as (tag_i, tag_j) pairs.  Nothing in the reference
// is patched; fields no compute exposes (ddv, number_density, nw, rhoI, vest, ddev ...) are read
// from the public members of class Atom.
//
//   usage:  ref_harness -in deck.lmp [-log none ...]      deck contains
//           fix <id> all sphbvf/snapshot <every> <prefix> [pairs]
//
// Snapshot file <prefix>.<step>.bin layout (little endian):
//   char magic[8]="SPHBVF01"; int64 step; int32 natoms, S, dim, ntypes, nfields, ago;
//   double boxlo[3], boxhi[3]; int32 periodic[3]; double dt; double mass[ntypes];
//   nfields x { char name[24]; int32 ncols; int32 is_int; data[natoms*ncols] (double|int32) }
//   int64 npairs; int32 pairs[npairs][2]   (only with the "pairs" keyword, else npairs=0)

#include <mpi.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <stdint.h>

#include "lammps.h"
#include "input.h"
#include "atom.h"
#include "modify.h"
#include "fix.h"
#include "force.h"
#include "pair.h"
#include "neighbor.h"
#include "neigh_list.h"
#include "domain.h"
#include "update.h"
#include "error.h"

using namespace LAMMPS_NS;

namespace {

class FixSnapshot : public Fix {
 public:
  FixSnapshot(LAMMPS *lmp, int narg, char **arg) : Fix(lmp, narg, arg), with_pairs(0) {
    if (narg < 5) error->all(FLERR, "fix sphbvf/snapshot: every prefix [pairs]");
    nevery = atoi(arg[3]);
    prefix = arg[4];
    for (int i = 5; i < narg; i++)
      if (strcmp(arg[i], "pairs") == 0) with_pairs = 1;
  }
  int setmask() { return FixConst::END_OF_STEP; }
  void setup(int) {
    write_meta();
    write();
  }
  void end_of_step() { write(); }

 private:
  std::string prefix;
  int with_pairs;

  void put(FILE *fp, const char *name, int ncols, int is_int) {
    char nm[24];
    memset(nm, 0, sizeof nm);
    strncpy(nm, name, 23);
    fwrite(nm, 1, 24, fp);
    int32_t h[2] = {ncols, is_int};
    fwrite(h, 4, 2, fp);
  }
  // rows are emitted in tag order through `order`
  void put_d1(FILE *fp, const char *name, double *a, const std::vector<int> &order) {
    put(fp, name, 1, 0);
    for (size_t k = 0; k < order.size(); k++) {
      double v = a ? a[order[k]] : 0.0;
      fwrite(&v, 8, 1, fp);
    }
  }
  void put_dn(FILE *fp, const char *name, double **a, int n, const std::vector<int> &order) {
    put(fp, name, n, 0);
    std::vector<double> z(n, 0.0);
    for (size_t k = 0; k < order.size(); k++) fwrite((a && n) ? a[order[k]] : z.data(), 8, n, fp);
  }
  void put_t33(FILE *fp, const char *name, double ***a, const std::vector<int> &order) {
    put(fp, name, 9, 0);
    for (size_t k = 0; k < order.size(); k++)
      for (int m = 0; m < 3; m++) fwrite(a[order[k]][m], 8, 3, fp);
  }
  void put_i1(FILE *fp, const char *name, int *a, const std::vector<int> &order) {
    put(fp, name, 1, 1);
    for (size_t k = 0; k < order.size(); k++) {
      int32_t v = a[order[k]];
      fwrite(&v, 4, 1, fp);
    }
  }

  // text sidecar: neighbour settings, pair style and (id, style, groupbit) of every fix, so the
  // fixture generator need not guess group bit assignments
  void write_meta() {
    std::string fn = prefix + ".meta.txt";
    FILE *fp = fopen(fn.c_str(), "w");
    if (!fp) error->one(FLERR, "cannot open snapshot meta file");
    fprintf(fp, "neighbor %.17g %d %d %d\n", neighbor->skin, neighbor->every, neighbor->delay,
            neighbor->dist_check);
    fprintf(fp, "cutneighmax %.17g\n", neighbor->cutneighmax);
    fprintf(fp, "pair_style %s\n", force->pair_style ? force->pair_style : "none");
    for (int i = 0; i < modify->nfix; i++)
      fprintf(fp, "fix %s %s %d\n", modify->fix[i]->id, modify->fix[i]->style, modify->fix[i]->groupbit);
    fclose(fp);
  }

  void write() {
    char fn[1024];
    snprintf(fn, sizeof fn, "%s.%ld.bin", prefix.c_str(), (long)update->ntimestep);
    FILE *fp = fopen(fn, "wb");
    if (!fp) error->one(FLERR, "cannot open snapshot file");
    int n = atom->nlocal;
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    tagint *tag = atom->tag;
    std::sort(order.begin(), order.end(), [tag](int a, int b) { return tag[a] < tag[b]; });

    const int S = atom->num_sdpd_species;
    const int nfields = 28;
    fwrite("SPHBVF01", 1, 8, fp);
    int64_t step = update->ntimestep;
    fwrite(&step, 8, 1, fp);
    int32_t hdr[6] = {n, S, domain->dimension, atom->ntypes, nfields, neighbor->ago};
    fwrite(hdr, 4, 6, fp);
    fwrite(domain->boxlo, 8, 3, fp);
    fwrite(domain->boxhi, 8, 3, fp);
    int32_t per[3] = {domain->xperiodic, domain->yperiodic, domain->zperiodic};
    fwrite(per, 4, 3, fp);
    fwrite(&update->dt, 8, 1, fp);
    fwrite(atom->mass + 1, 8, atom->ntypes, fp);

    std::vector<int> tags(n);
    for (int i = 0; i < n; i++) tags[i] = (int)tag[i];
    put_i1(fp, "tag", tags.data(), order);
    put_i1(fp, "type", atom->type, order);
    put_i1(fp, "mask", atom->mask, order);
    put_i1(fp, "solid_tag", atom->solid_tag, order);
    put_i1(fp, "fixed_tag", atom->fixed_tag, order);
    put_dn(fp, "x", atom->x, 3, order);
    put_dn(fp, "v", atom->v, 3, order);
    put_dn(fp, "vest", atom->vest, 3, order);
    put_dn(fp, "f", atom->f, 3, order);
    put_d1(fp, "rho", atom->rho, order);
    put_d1(fp, "rhoI", atom->rhoI, order);
    put_d1(fp, "drho", atom->drho, order);
    put_d1(fp, "e", atom->e, order);
    put_d1(fp, "de", atom->de, order);
    put_d1(fp, "phi", atom->phi, order);
    put_d1(fp, "number_density", atom->number_density, order);
    put_dn(fp, "nw", atom->nw, 3, order);
    put_dn(fp, "ddv", atom->ddv, 3, order);
    put_dn(fp, "ddx", atom->ddx, 3, order);
    put_d1(fp, "rhoAux1", atom->rhoAux1, order);
    put_d1(fp, "rhoAux2", atom->rhoAux2, order);
    put_d1(fp, "Pnew", atom->Pnew, order);
    put_t33(fp, "dev", atom->deviatoricTensor, order);
    put_t33(fp, "ddev", atom->ddeviatoricTensor, order);
    put_dn(fp, "C", atom->C, S, order);
    put_dn(fp, "Q", atom->Q, S, order);
    put_dn(fp, "v_weighted_solid", atom->v_weighted_solid, 3, order);
    put_dn(fp, "a_weighted_solid", atom->a_weighted_solid, 3, order);

    int64_t npairs = 0;
    NeighList *list = (with_pairs && force->pair) ? force->pair->list : NULL;
    if (list)
      for (int ii = 0; ii < list->inum; ii++) npairs += list->numneigh[list->ilist[ii]];
    fwrite(&npairs, 8, 1, fp);
    if (list) {
      // ghosts carry the tag of the atom they image (atom_vec_ssa_tsdpd_atomic.cpp:1296-1368)
      for (int ii = 0; ii < list->inum; ii++) {
        int i = list->ilist[ii];
        int *jl = list->firstneigh[i];
        for (int jj = 0; jj < list->numneigh[i]; jj++) {
          int j = jl[jj] & NEIGHMASK;
          int32_t p[2] = {(int32_t)tag[i], (int32_t)tag[j]};
          fwrite(p, 4, 2, fp);
        }
      }
    }
    fclose(fp);
  }
};

Fix *snapshot_creator(LAMMPS *lmp, int narg, char **arg) { return new FixSnapshot(lmp, narg, arg); }

}  // namespace

int main(int argc, char **argv) {
  MPI_Init(&argc, &argv);
  LAMMPS *lammps = new LAMMPS(argc, argv, MPI_COMM_WORLD);
  (*lammps->modify->fix_map)["sphbvf/snapshot"] = &snapshot_creator;
  lammps->input->file();
  delete lammps;
  MPI_Finalize();
  return 0;
}

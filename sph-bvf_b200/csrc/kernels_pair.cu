// kernels_pair.cu -- the pair pass of the SPH-BVF step: density/BVF sums and the force pass fused
// into ONE traversal of the frozen neighbour structure.
//
// Replaces PairSsaTsdpdBvf{TransportVelocity,Mechanics,Fsi}::compute
// (pair_ssa_tsdpd_bvf_transport_velocity.cpp:68-910, ..._mechanics.cpp:68-948, ..._fsi.cpp:81-797)
// plus Verlet::force_clear / AtomVec::force_clear (every output is written, not accumulated) and
// the reverse communication (gather form: atom i sums over its FULL neighbour set).
//
// The reference's sweep A (number density, Shepard sums, transport-velocity correction ddv, ddx)
// and sweep B (forces, drho, phi, wall normal, Jaumann rate, species flux) read only particle
// state and write disjoint outputs, so they share one loop, one sqrt and one set of neighbour
// loads.  Sweep C (v_weighted_solid / a_weighted_solid) is never consumed.  The
// stochastic stress (active when some ssa_tsdpd/e != 0) is a counter-based pair noise (Philox), see RANDOM below.
//
// Three traversals share ONE arithmetic body (PairAcc::visit), so they differ only in summation order:
//
//  * pair_kernel (gather form, default): one thread per owned atom, 192-thread CTAs, 96-byte records gathered
//    through L1 with a hand software pipeline (records of neighbour k+1 in flight while k is evaluated, list entries
//    through a cp.async ring): 5.6 ms per pass at 8 M atoms.
//
//  * pair_tile_kernel (tile form, SPHBVF_PAIR=tile): the design BASELINE.json's north star names -- one CTA per tile
//    of the cell grid (4x4x4 cells, ~145 atoms); the records of every candidate of the tile (the atoms of the
//    <= 8x8x8 cells within stencil reach, ~1160) are staged ONCE in shared memory with cp.async (six 16-byte granule
//    arrays, bank group = slot mod 8), the Verlet list holds 16-bit slot numbers (2 B instead of 4 B per neighbour),
//    and EIGHT lanes share one atom: at every step they take eight consecutive entries of its sorted list and their
//    partial sums are combined once per atom by a halving butterfly.  No L2 round trip remains inside the neighbour
//    loop.  Built, bit-checked against the gather form and measured (profiles/r02a_*): 8.1 ms per pass with 80-byte
//    records at two CTAs per SM -- it executes 1.7 x the instructions of the gather form (staging 16 %, per-atom
//    prologue / butterfly / stores 21 %), the shared-memory gathers cost as many data-pipe wavefronts as the L1
//    gathers they replace (2.5 lanes of a quarter-warp collide on a bank group: runs of consecutive slots are only
//    ~4 long), and with 16 warps per SM issue stays at 50 %.  Kept selectable and tested; DESIGN.md section 3.
//
//  * pair_split_kernel (gather form for launches of <= 65536 atoms, i.e. the reference's shipped decks): four adjacent
//    lanes share one atom, each walks a quarter of its list with the same pipeline, the sums meet in a butterfly.  A
//    3 k - 42 k atom launch leaves most of the GPU empty, so its duration is the latency of one thread's serial walk
//    over its neighbours; four lanes cut that walk to a quarter.  SPHBVF_PAIR_LANES=1 switches it off.
//
// Common to all: every per-particle division lives in the pack kernel (V = m/rho, P/rho^2),
// sqrt is a branch-free Goldschmidt iteration on MUFU.RSQ64H (7 FP64 ops, < 1 ulp), per-type-pair coefficients are
// kernel-argument constants when every type pair shares them (all cavity decks and the synthetic lattice) and
// rows of a small shared-memory table otherwise.
#include <stdlib.h>
#include <string.h>

#include "sphbvf_internal.cuh"
#include "tile_common.cuh"

namespace sphbvf {

// ---- compile-time tuning switches (tools/build_variant.sh; DESIGN.md section 3 has the measurements)
//   PAIR_T / PAIR_MINB   gather form: threads per CTA / resident CTAs per SM it is compiled for
//   PT_T / PT_MINB       tile form: threads per CTA (multiple of 32) / resident CTAs per SM
#ifndef PAIR_T
#define PAIR_T 192   // 12 warps as 2 x 192 threads: 5.63 ms (3 x 128: 5.71, 6 x 64: 5.73, 1 x 384: 5.66)
#endif
#ifndef PAIR_MINB
#define PAIR_MINB 2
#endif
#ifndef PAIR_BC_LATE
#define PAIR_BC_LATE 0
#endif
#ifndef PT_T
#define PT_T 256
#endif
#ifndef PT_MINB
#define PT_MINB 2
#endif

// one row per (type_i, type_j); 8 doubles = 64 B so a row is two LDS.128 x2
struct __align__(16) PairRow {
  double cutsq, h, cwfd, cwf;      // h^2, h, (1/r)dW/dr = cwfd (h-r)^2, W = cwf (h-r)^3 (h+3r)
  double mimj, eta, iwdelta, h2eps;  // m_i m_j, eta, 1/W(delta), 0.01 h^2
};

// per type: what the body needs about atom j's type beyond the pair row
struct __align__(16) TypeRow {
  double imass, c0, G0, pad;    // 1/m, c0, G0
};

// species transport of one (type_i, type_j): 64 B, read from shared memory (type pairs differ: lanes read different
// rows) or straight from the constant bank with a compile-time index (UNIFORM: every pair has the same row)
struct __align__(16) SpecRow {
  double cutc, cwfdc, mred2, hc2eps;   // concentration cutoff, (1/r)dW/dr coefficient with h = cutc, 2 mi mj / (mi + mj), 0.01 cutc^2
  double kappa[MAXS];
};

struct PairTables {
  PairRow row[MAXT * MAXT];
  TypeRow type[MAXT];
  SpecRow spec[MAXT * MAXT];
  double geff[MAXT][MAXT];   // 2 Gi Gj / (Gi + Gj + 1e-12)
};

static bool make_tables(const Coeffs &co, PairTables &t) {
  const double delta_fac = co.variant == SPHBVF_TV ? (1.0 / 2.6) : (1.0 / 3.0);
  memset(&t, 0, sizeof t);
  bool uniform = true;
  for (int i = 1; i <= co.ntypes; i++) {
    t.type[i].imass = 1.0 / co.mass[i];
    t.type[i].c0 = co.c0[i];
    t.type[i].G0 = co.G0[i];
    for (int j = 1; j <= co.ntypes; j++) {
      auto coef = [&](double h, double &cwfd, double &cwf) {
        double ih = 1.0 / h, ihsq = ih * ih;
        if (co.dim == 3) {
          cwfd = -25.066903536973515383e0 * ihsq * ihsq * ihsq * ih;
          cwf = 2.088908628081126 * ihsq * ihsq * ihsq * ih;
        } else {
          cwfd = -19.098593171027440292e0 * ihsq * ihsq * ihsq;
          cwf = 1.591549430918954 * ihsq * ihsq * ihsq;
        }
      };
      PairRow &r = t.row[i * MAXT + j];
      double h = co.cut[i][j], hc = co.cutc[i][j], dummy;
      coef(h, r.cwfd, r.cwf);
      SpecRow &sr = t.spec[i * MAXT + j];
      coef(hc > 0 ? hc : h, sr.cwfdc, dummy);
      double delta = delta_fac * h, td = h - delta;
      double wdelta = r.cwf * td * td * td * (h + 3. * delta);
      r.cutsq = co.cutsq[i][j];
      r.h = h;
      r.iwdelta = 1.0 / wdelta;
      r.h2eps = 0.01 * h * h;
      r.mimj = co.mass[i] * co.mass[j];
      r.eta = co.eta[i][j];
      sr.cutc = hc;
      sr.hc2eps = 0.01 * hc * hc;
      sr.mred2 = 2.0 * ((co.mass[i] * co.mass[j]) / (co.mass[i] + co.mass[j]));
      for (int k = 0; k < MAXS; k++) sr.kappa[k] = k < co.nspecies ? co.kappa[i][j][k] : 0.0;
      t.geff[i][j] = (2.0 * co.G0[i] * co.G0[j]) / (co.G0[i] + co.G0[j] + 1e-12);
      const PairRow &r11 = t.row[MAXT + 1];
      if (r.cutsq != r11.cutsq || r.h != r11.h || r.mimj != r11.mimj || r.eta != r11.eta) uniform = false;
      // the UNIFORM instantiations also read the species row (and 1/m) of pair (1,1) for every pair
      if (co.nspecies > 0 && memcmp(&sr, &t.spec[MAXT + 1], sizeof sr)) uniform = false;
      if (co.mass[i] != co.mass[1]) uniform = false;
    }
  }
  return uniform;
}

// sqrt(x) for x > 0: MUFU.RSQ64H seed (2^-22.9) + two coupled Goldschmidt steps, branch free
__device__ __forceinline__ double fast_sqrt(double x) {
  // MUFU.RSQ64H reads the high word only; clamping it away from 0 (an integer op) keeps x == 0
  // (coincident atoms) from producing inf * 0: then g = 0 * y = 0 and the result is exactly 0
  const double xs = __hiloint2double(max(__double2hiint(x), 0x00200000), 0);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xs));
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  return fma(g, r, g);
}

// 1/x for normal x: MUFU.RCP64H seed + two Newton steps (no slow-path call, ~1 ulp)
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// ---- counter-based randomness for the stochastic stress term: Philox-4x32-10 (Salmon et al. 2011)
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// two independent standard normals from 64 random bits (Box-Muller; u1 in (0, 1])
__device__ __forceinline__ void gauss2(unsigned a, unsigned b, double &g0, double &g1) {
  const double u1 = ((double)a + 1.0) * (1.0 / 4294967296.0), u2 = (double)b * (1.0 / 4294967296.0);
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  g0 = rad * cs;
  g1 = rad * sn;
}

// kernel-wide scalars of one pair pass
struct PairConsts {
  double damp, rand_pref;
  unsigned long long seed;
  long ntimestep;
  int elmask;   // bit t: atoms of type t with solid_tag may carry a deviatoric stress (G0[t] != 0 or some dev != 0)
};

// ------------------------------------------------------------------------------------------
// The arithmetic of one atom i and of one neighbour visit, shared by both traversals.
//   SOLIDS: 0 = no atom has solid_tag, 1 = solids whose deviatoric stress is identically zero
//   (rigid walls: G0 == 0, dev == 0), 2 = elastic solids (deviatoric tensors gathered)
// A neighbour is handed over as the fields of its record; `j` (its global index) is only used by the
// instantiations that gather extras (species, deviatoric tensors, stochastic term).
// ------------------------------------------------------------------------------------------
// Everything a pair adds because atom i or atom j carries a deviatoric stress (SOLIDS == 2; reference:
// pair_ssa_tsdpd_bvf_mechanics.cpp:435-494, ..._fsi.cpp same blocks): artificial stress, the stress-divergence force on
// a solid i, and the Jaumann rate of i's stress.  Only called when atom i or atom j is ELASTIC (solid_tag and a type that
// can carry stress: PairConsts::elmask); rigid walls (dev == 0 for ever) take the scalar shortcut of SOLIDS == 1.
// Measured as a __noinline__ call (registers saved only around the call): fsi deck 1.06 -> 1.68 ms, so it is inlined.
// sol: dev_i = sol[k * stride], d(dev_i)/dt accumulator = sol[(9 + k) * stride].
struct SolidIn {
  double delx, dely, delz, velx, vely, velz;
  double wf, wfd, mmw, iwdelta, Vj, rhoj, Prrj, irhoi, irhoj, Pi, geff;
};
template <int VARIANT>
__device__ __forceinline__ double3 solid2_visit(const SolidIn in, const bool si, const bool sj, const double *__restrict__ devj_g,
                                             double *sol, const int stride) {
  constexpr double c_art = VARIANT == SPHBVF_FSI ? 0.1 : 0.35;
  double devj[9];
#pragma unroll
  for (int k = 0; k < 9; k++) devj[k] = sj ? devj_g[k] : 0.0;
  const double Pj = in.Prrj * in.rhoj * in.rhoj;
  const double Psj = VARIANT == SPHBVF_MECHANICS ? fabs(Pj) : Pj;
  const double Psi = VARIANT == SPHBVF_MECHANICS ? fabs(in.Pi) : in.Pi;
  const double q = in.wf * in.iwdelta, q2 = q * q;
  const double pre = in.mmw * q2 * q2;
  double R[9];
#pragma unroll
  for (int m = 0; m < 3; m++)
#pragma unroll
    for (int n = 0; n < 3; n++) {
      const double tsj = devj[3 * m + n] - (m == n ? Psj : 0.0);
      const double Rj = (sj && tsj > 0.0) ? -c_art * tsj * in.irhoj * in.irhoj : 0.0;
      double Ri = 0.0;
      if (si) {
        const double tsi = sol[(3 * m + n) * stride] - (m == n ? Psi : 0.0);
        Ri = tsi > 0.0 ? -c_art * tsi * in.irhoi * in.irhoi : 0.0;
      }
      R[3 * m + n] = Ri + Rj;
    }
  double3 F;
  F.x = pre * (in.delx * R[0] + in.dely * R[3] + in.delz * R[6]);
  F.y = pre * (in.delx * R[1] + in.dely * R[4] + in.delz * R[7]);
  F.z = pre * (in.delx * R[2] + in.dely * R[5] + in.delz * R[8]);
  if (si) {
    // ---- Jaumann rate for solid i (:435-451)
    const double hw = -0.5 * in.Vj * in.wfd;   // 0.5 * Vj * wfd * (v_j - v_i) = hw * vel
    const double vel[3] = {in.velx, in.vely, in.velz}, del[3] = {in.delx, in.dely, in.delz};
    double devi[9], eps[9], om[9];
#pragma unroll
    for (int k = 0; k < 9; k++) devi[k] = sol[k * stride];
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
      for (int n = 0; n < 3; n++) {
        const double a = vel[m] * del[n], b = vel[n] * del[m];
        eps[3 * m + n] = hw * (a + b);
        om[3 * m + n] = hw * (a - b);
      }
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
      for (int n = 0; n < 3; n++) {
        const double dDotR = devi[3 * m] * om[3 * n] + devi[3 * m + 1] * om[3 * n + 1] + devi[3 * m + 2] * om[3 * n + 2];
        const double rDotD = om[3 * m] * devi[n] + om[3 * m + 1] * devi[3 + n] + om[3 * m + 2] * devi[6 + n];
        const double e = eps[3 * m + n];
        sol[(9 + 3 * m + n) * stride] += 2.0 * in.geff * (m == n ? e - (1. / 3.) * e : e) + dDotR + rDotD;
      }
    // ---- stress divergence on a solid i (:511-522)
    const double ii = in.irhoi * in.irhoi, jj = in.irhoj * in.irhoj;
    F.x += in.mmw * (in.delx * (devi[0] * ii + devj[0] * jj) + in.dely * (devi[3] * ii + devj[3] * jj) + in.delz * (devi[6] * ii + devj[6] * jj));
    F.y += in.mmw * (in.delx * (devi[1] * ii + devj[1] * jj) + in.dely * (devi[4] * ii + devj[4] * jj) + in.delz * (devi[7] * ii + devj[7] * jj));
    F.z += in.mmw * (in.delx * (devi[2] * ii + devj[2] * jj) + in.dely * (devi[5] * ii + devj[5] * jj) + in.delz * (devi[8] * ii + devj[8] * jj));
  }
  return F;
}

// SOLSMEM (gather form, SOLIDS == 2): the deviatoric stress of atom i and its rate accumulator -- 18 doubles that only
// solid atoms ever touch -- live in a shared-memory column of the thread (sol[k * SOLSTRIDE]) instead of 36 registers:
// at 168 registers the SOLIDS == 2 instantiations spilled 1.9 KB per thread into the neighbour loop of EVERY atom.
template <int VARIANT, bool SPECIES, int SOLIDS, bool UNIFORM, bool FILTER, bool RANDOM, bool SOLSMEM = false, int SOLSTRIDE = 1>
struct PairAcc {
  // atom i
  double xi, yi, zi, rhoi, vxi, vyi, vzi, Vi, wxi, wyi, wzi, Prri;   // record: A = {x,y,z,rho} B = {vest,V} C = {w,P/rho^2}
  double Vi2, c0i, Pi, irhoi, G0i, ei, arti;
  int ti, tagi;
  bool si;
  double solr[(SOLIDS == 2 && !SOLSMEM) ? 18 : 1], Cspec_i[MAXS];
  double *sol;   // devi = sol[0..8], ddev = sol[9..17] (times SOLSTRIDE): shared-memory column (SOLSMEM) or solr
  __device__ __forceinline__ double &devi(int k) { return sol[k * SOLSTRIDE]; }
  __device__ __forceinline__ double &ddev(int k) { return sol[(9 + k) * SOLSTRIDE]; }
  const PairRow *myrow;
  const SpecRow *myspec;   // row ti of a shared-memory copy of tb.spec, or nullptr: read tb.spec (constant bank)
  // sums
  double fx, fy, fz, drho, nd, rA1, rA2, phi, ddvx, ddvy, ddvz, nwx, nwy, nwz, ddxx, ddxy, ddxz, spi;
  double Qs[MAXS];

  __host__ __device__ static constexpr double c_art() { return VARIANT == SPHBVF_FSI ? 0.1 : 0.35; }

  __device__ __forceinline__ void init(const DevState &d, const Coeffs &co, const PairTables &tb, const PairRow *srow,
                                       const int i, const int fl, const Rec4 &A, const Rec4 &B, const Rec4 &C) {
    ti = fl & 7;
    si = SOLIDS && ((fl >> 4) & 1);
    xi = A.x; yi = A.y; zi = A.z; rhoi = A.w;
    vxi = B.x; vyi = B.y; vzi = B.z; Vi = B.w;
    wxi = C.x; wyi = C.y; wzi = C.z; Prri = C.w;
    Vi2 = Vi * Vi;
    c0i = co.c0[ti];
    Pi = Prri * rhoi * rhoi;
    irhoi = Vi * tb.type[ti].imass;
    myrow = srow + (UNIFORM ? 0 : ti * MAXT);
    myspec = nullptr;
    arti = 0.0;
    // stress-free solid (dev == 0): R = -c_art max(0, -Psigma)/rho^2 = c_art P/rho^2 where P < 0 (TV, fsi);
    // mechanics uses |P| (:471,487) so the bracket is never positive
    if (SOLIDS) arti = (si && VARIANT != SPHBVF_MECHANICS && Prri < 0.0) ? c_art() * Prri : 0.0;   // SOLIDS == 2: used when neither atom is elastic
    if (SOLIDS == 2) {
      if (!SOLSMEM) sol = solr;
#pragma unroll
      for (int k = 0; k < 9; k++) devi(k) = si ? d.pdev[9 * (size_t)i + k] : 0.0;
    }
    if (SPECIES) {   // fully unrolled with a predicate: the arrays live in registers, not in local memory
#pragma unroll
      for (int k = 0; k < MAXS; k++) Cspec_i[k] = k < co.nspecies ? d.pCs[(size_t)i * co.nspecies + k] : 0.0;
    }
    ei = RANDOM ? d.pD[i].w : 0.0;
    tagi = RANDOM ? d.ptag[i] : 0;
    G0i = co.G0[ti];
    if (VARIANT == SPHBVF_FSI && SPECIES) G0i = co.G0[ti] * (1.0 - 0.99 * Cspec_i[0]);
    fx = fy = fz = drho = nd = rA1 = rA2 = phi = 0.0;
    ddvx = ddvy = ddvz = nwx = nwy = nwz = 0.0;
    ddxx = ddxy = ddxz = 0.0;
    spi = 0.0;   // sum_j s_ij rho_i a_i: the i-side transport term is vest_i times this
    if (SOLIDS == 2)
#pragma unroll
      for (int k = 0; k < 9; k++) ddev(k) = 0.0;
    if (SPECIES)
#pragma unroll
      for (int k = 0; k < MAXS; k++) Qs[k] = 0.0;
  }

  // one neighbour, handed over as the fields of its record; `j` (its global index) is only used by the
  // instantiations that gather extras
  __device__ __forceinline__ void visit(const DevState &d, const Coeffs &co, const PairTables &tb, const PairConsts &pc,
                                        const int tj, const bool sj_in, const int j, const double xj, const double yj,
                                        const double zj, const double rhoj, const double vxj, const double vyj,
                                        const double vzj, const double Vj, const double wxj, const double wyj,
                                        const double wzj, const double Prrj, const double rhoIj, const double Cj0) {
    const bool sj = SOLIDS && sj_in;
    const double delx = xi - xj, dely = yi - yj, delz = zi - zj;
    const double rsq = delx * delx + dely * dely + delz * delz;
    double cutsq, h, cwfd, cwf, mm, eta;
    if (UNIFORM) {
      const PairRow &r = tb.row[MAXT + 1];
      cutsq = r.cutsq; h = r.h; cwfd = r.cwfd; cwf = r.cwf; mm = r.mimj; eta = r.eta;
    } else {
      const PairRow &r = myrow[tj];
      cutsq = r.cutsq; h = r.h; cwfd = r.cwfd; cwf = r.cwf; mm = r.mimj; eta = r.eta;
    }
    if (!(rsq < cutsq)) return;
    const double iwdelta = UNIFORM ? tb.row[MAXT + 1].iwdelta : myrow[tj].iwdelta;
    const double h2eps = UNIFORM ? tb.row[MAXT + 1].h2eps : myrow[tj].h2eps;

    const double r = fast_sqrt(rsq);
    const double t = h - r, t2 = t * t;
    const double wfd = cwfd * t2;
    const double wf = cwf * t2 * t * (h + 3. * r);
    const double Vj2 = Vj * Vj;
    const double velx = vxi - vxj, vely = vyi - vyj, velz = vzi - vzj;
    const double dvr = delx * velx + dely * vely + delz * velz;
    const double ai = wxi * delx + wyi * dely + wzi * delz;   // (v_i - vt_i) . del
    const double aj = wxj * delx + wyj * dely + wzj * delz;
    const double qi = rhoi * ai, qj = rhoj * aj;
    const double S2 = Vi2 + Vj2;
    const double S2w = S2 * wfd;

    // ---- sweep A (pair_...transport_velocity.cpp:243-254); ddv is scaled by 70 B_i at the end
    nd = fma(Vj2, wf, nd);
    rA2 += wf;
    if (FILTER) rA1 = fma(rhoIj, wf, rA1);
    ddvx = fma(S2w, delx, ddvx); ddvy = fma(S2w, dely, ddvy); ddvz = fma(S2w, delz, ddvz);
    if (VARIANT != SPHBVF_TV) {   // ..._mechanics.cpp:250-252
      const double Vj2w = Vj2 * wf;
      ddxx = fma(-Vj2w, velx, ddxx); ddxy = fma(-Vj2w, vely, ddxy); ddxz = fma(-Vj2w, velz, ddxz);
    }

    // ---- pressure force (:396-399 / mechanics :408)
    const double mmw = mm * wfd;
    const double pij = Prrj + Prri;
    double fpair;
    if (VARIANT == SPHBVF_TV) fpair = (pij >= 0. || (si && sj)) ? mmw * pij : mmw * (Prrj - Prri);
    else fpair = mmw * pij;

    // ---- artificial stress (:454-494)
    double fartx = 0, farty = 0, fartz = 0;
    const bool eli = SOLIDS == 2 && si && ((pc.elmask >> ti) & 1), elj = SOLIDS == 2 && sj && ((pc.elmask >> tj) & 1);
    if (SOLIDS) {
      if ((si || sj) && !(eli || elj)) {
        const double q = wf * iwdelta, q2 = q * q;
        const double artj = (sj && VARIANT != SPHBVF_MECHANICS && Prrj < 0.0) ? c_art() * Prrj : 0.0;
        const double cc = mmw * q2 * q2 * (arti + artj);
        fartx = cc * delx; farty = cc * dely; fartz = cc * delz;
      }
    }

    if (SOLIDS == 2) {
      if (eli || elj) {
        double geff = 0.0;
        if (si) {
          if (VARIANT == SPHBVF_FSI && SPECIES) {
            const double G0j = tb.type[tj].G0 * (1.0 - 0.99 * Cj0);
            geff = (2.0 * G0i * G0j) / (G0i + G0j + 1e-12);
          } else geff = tb.geff[ti][tj];
        }
        const SolidIn in = {delx, dely, delz, velx, vely, velz, wf, wfd, mmw, iwdelta, Vj, rhoj, Prrj, irhoi,
                            Vj * tb.type[tj].imass, Pi, geff};
        const double3 F = solid2_visit<VARIANT>(in, si, sj, d.pdev + 9 * (size_t)j, sol, SOLSTRIDE);
        fartx = F.x; farty = F.y; fartz = F.z;
      }
    }

    // ---- momentum (:497-529)
    if (!si) {
      // chained FMAs into the accumulators (4 per component instead of 7 separately rounded ops);
      // the i-side transport term s rho_i a_i vest_i has a per-atom constant vector: summed as a scalar
      const double fvisc = S2w * eta;
      const double s = -0.5 * S2w;
      const double pj_ = s * qj;
      spi = fma(s, qi, spi);
      fx = fma(fvisc, velx, fx); fy = fma(fvisc, vely, fy); fz = fma(fvisc, velz, fz);
      fx = fma(-fpair, delx, fx); fy = fma(-fpair, dely, fy); fz = fma(-fpair, delz, fz);
      fx = fma(pj_, vxj, fx); fy = fma(pj_, vyj, fy); fz = fma(pj_, vzj, fz);
      if (SOLIDS) { fx += fartx; fy += farty; fz += fartz; }
      if (RANDOM) {
        // f_rand = sqrt(-4 kB e m_i m_j wfd / (rho_i rho_j dt)) / (r + 0.01 h) * (Wn . del)   (:403-431), with
        // Wn the symmetric traceless part of a d x d Gaussian matrix: off-diagonals N(0, 1/2), diagonal
        // g_ll - mean(g).  The matrix depends on (seed, step, min tag, max tag) only, so the partner
        // computes the same one with del -> -del: equal and opposite forces.
        const double eij = 0.5 * (ei + d.pD[j].w);
        const int tagj = d.ptag[j];
        const double pref = sqrt(fmax(-pc.rand_pref * eij * (Vi * Vj) * wfd, 0.0)) * fast_rcp(r + 0.01 * h);
        const uint2 key = make_uint2((unsigned)pc.seed, (unsigned)(pc.seed >> 32));
        const unsigned tlo = (unsigned)min(tagi, tagj), thi = (unsigned)max(tagi, tagj);
        const long ns = pc.ntimestep;
        const uint4 r0 = philox4x32(make_uint4(tlo, thi, (unsigned)ns, (unsigned)(ns >> 32) << 1), key);
        double g0, g1, g2, g3;
        gauss2(r0.x, r0.y, g0, g1);
        gauss2(r0.z, r0.w, g2, g3);
        double wxx, wyy, wzz, wxy, wxz = 0.0, wyz = 0.0;
        if (co.dim == 2) {
          // trace / dimension with the zz entry zero: Wxx = (gxx - gyy)/2 = -Wyy, Wzz irrelevant (delz = 0)
          wxx = 0.5 * (g0 - g1); wyy = -wxx; wzz = 0.0;
          wxy = 0.5 * (g2 + g3);
        } else {
          const uint4 r1 = philox4x32(make_uint4(tlo, thi, (unsigned)ns, ((unsigned)(ns >> 32) << 1) | 1u), key);
          double g4, g5, g6, g7;
          gauss2(r1.x, r1.y, g4, g5);
          gauss2(r1.z, r1.w, g6, g7);
          const double mean = (g0 + g1 + g2) * (1.0 / 3.0);
          wxx = g0 - mean; wyy = g1 - mean; wzz = g2 - mean;
          const double isq2 = 0.70710678118654752440;   // (g_lm + g_ml)/2 ~ N(0, 1/2)
          wxy = isq2 * g3; wxz = isq2 * g4; wyz = isq2 * g5;
          (void)g6; (void)g7;
        }
        fx += pref * (wxx * delx + wxy * dely + wxz * delz);
        fy += pref * (wxy * delx + wyy * dely + wyz * delz);
        fz += pref * (wxz * delx + wyz * dely + wzz * delz);
      }
    } else {
      double fviscs = 0.;
      if (dvr < 0.) {
        const double mu = h * dvr * fast_rcp(rsq + h2eps);
        fviscs = mmw * (-(c0i + tb.type[tj].c0) * mu + 2.0 * mu * mu) * fast_rcp(rhoi + rhoj);
      }
      const double cc = -(fpair + fviscs);
      fx += cc * delx + fartx;
      fy += cc * dely + farty;
      fz += cc * delz + fartz;
    }

    // ---- density rate (:548-555); (vt_i - vt_j).del = dvr - a_i + a_j
    {
      // rho_i (dvr - a_i + a_j) - (rho_i a_i + rho_j a_j) = rho_i (dvr + a_j) - 2 q_i - q_j
      double inner = fma(-2.0, qi, fma(rhoi, dvr + aj, -qj));
      if (VARIANT == SPHBVF_FSI)
        inner -= pc.damp * 2.0 * h * c0i * (rhoj - rhoi) * (rsq * fast_rcp(rsq + h2eps));
      drho = fma(wfd * Vj, inner, drho);
    }

    // ---- BVF (:563-576)
    if (SOLIDS && !si && sj) {
      phi = fma(Vj2, wf, phi);
      const double cc = wfd * Vj2;
      nwx = fma(cc, delx, nwx); nwy = fma(cc, dely, nwy); nwz = fma(cc, delz, nwz);
    }

    // ---- species (:678-720)
    // Cj0 = C_j[0] arrives with the record (requested one visit ahead); further species are gathered here
    if (SPECIES) {
      const SpecRow &sr = UNIFORM ? tb.spec[MAXT + 1] : (myspec ? myspec[tj] : tb.spec[ti * MAXT + tj]);
      const double hc = sr.cutc;
      if (r < hc) {
        const double tc = hc - r;
        const double wfdc = sr.cwfdc * tc * tc;
        const double irhoj = Vj * (UNIFORM ? tb.type[1].imass : tb.type[tj].imass);
        const double q0 = sr.mred2 * (irhoi + irhoj) * rsq * wfdc * fast_rcp(rsq + sr.hc2eps);
#pragma unroll
        for (int k = 0; k < MAXS; k++)
          if (k < co.nspecies) {
            const double Cjk = k == 0 ? Cj0 : d.pCs[(size_t)j * co.nspecies + k];
            double dq = sr.kappa[k] * (Cspec_i[k] - Cjk) * q0;
            if (VARIANT == SPHBVF_TV) dq -= Vj * (Cspec_i[k] * ai + Cjk * aj) * wfdc;
            Qs[k] += dq;
          }
      }
    }
  }

  // force on atom i so far (the i-side transport term folded in): used by the virial pass
  __device__ __forceinline__ void force_now(double &Fx, double &Fy, double &Fz) const {
    Fx = fx + spi * vxi; Fy = fy + spi * vyi; Fz = fz + spi * vzi;
  }
};

// does this instantiation read per-neighbour data that is not in the record (through the neighbour's global index)?
template <bool SPECIES, int SOLIDS, bool FILTER, bool RANDOM, bool VIRIAL>
struct NeedsJ { static constexpr bool value = SPECIES || SOLIDS == 2 || FILTER || RANDOM || VIRIAL; };

// ==========================================================================================
// gather form
// ==========================================================================================
constexpr int RING = 16;  // list-entry prefetch depth (rows); power of two

__device__ __forceinline__ void cp_async4(int *smem_dst, const int *gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}

// VIRIAL = true turns a pair kernel into the ghost part of Pair::virial_fdotr_compute (pair.cpp:1511): no
// per-atom output is written; for every neighbour that is a periodic image (shift s != 0) the force
// F it exerts on atom i is obtained as the change of the force accumulator and -1/2 s (x) F is summed
// into virial_out[6] (see sphbvf_virial in capi.cu for the derivation).
// threads per CTA of the gather form: the elastic-solid instantiations (SOLIDS == 2) need more than the 168 registers
// that 2 x 192 threads leave per thread (they spilled > 1 KB per thread into the neighbour loop): 2 x 128 threads at
// up to 255 registers
__host__ __device__ constexpr int pair_threads(int solids) { return solids == 2 ? 128 : PAIR_T; }

// one owned atom of the gather form: the whole neighbour loop and the stores (myring = this thread's column of the
// CTA's list-entry ring)
// LANES > 1 (small systems, pair_split_kernel): LANES adjacent lanes share atom i; lane `sub` takes the entries
// sub, sub + LANES, ... of its list and the partial sums meet in an xor butterfly before lane 0 stores.  `valid` = false
// pads the last warp: such a lane visits nothing and stores nothing but takes part in the shuffles.
template <int LANES>
__device__ __forceinline__ double lane_sum(double x) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

template <int VARIANT, bool SPECIES, int SOLIDS, bool UNIFORM, bool FILTER, bool RANDOM, bool VIRIAL, int LANES = 1>
__device__ __forceinline__ void pair_atom(const DevState &d, const Coeffs &co, const PairTables &tb, const PairConsts &pc,
                                          const PairRow *srow, const SpecRow *sspec, int *myring, double *mysol,
                                          const int i, double *virial_out, const int sub = 0, const bool valid = true) {
  constexpr int PTH = pair_threads(SOLIDS);
  PairAcc<VARIANT, SPECIES, SOLIDS, UNIFORM, FILTER, RANDOM, SOLIDS == 2, PTH> acc;
  acc.sol = mysol;
  acc.init(d, co, tb, srow, i, d.pflags[i], d.prec[i].A, d.prec[i].B, d.prec[i].C);
  if (SPECIES && !UNIFORM) acc.myspec = sspec + acc.ti * MAXT;

  double vir[6] = {0, 0, 0, 0, 0, 0};
  auto visit = [&](const int ent, const Rec4 &Aj, const Rec4 &Bj, const Rec4 &Cj, const double rhoIj, const double Cj0) {
    const int j = ent & NEIGH_JMASK;
    const int tj = (ent >> NEIGH_JBITS) & 7;
    const bool sj = (ent >> 30) & 1;
    if (!VIRIAL) {
      acc.visit(d, co, tb, pc, tj, sj, j, Aj.x, Aj.y, Aj.z, Aj.w, Bj.x, Bj.y, Bj.z, Bj.w, Cj.x, Cj.y, Cj.z, Cj.w, rhoIj, Cj0);
      return;
    }
    const int g = j - d.nlocal;
    if (g < 0) return;
    const double sx = d.gshift[3 * (size_t)g], sy = d.gshift[3 * (size_t)g + 1], sz = d.gshift[3 * (size_t)g + 2];
    if (sx == 0.0 && sy == 0.0 && sz == 0.0) return;
    double f0x, f0y, f0z, f1x, f1y, f1z;
    acc.force_now(f0x, f0y, f0z);
    acc.visit(d, co, tb, pc, tj, sj, j, Aj.x, Aj.y, Aj.z, Aj.w, Bj.x, Bj.y, Bj.z, Bj.w, Cj.x, Cj.y, Cj.z, Cj.w, rhoIj, Cj0);
    acc.force_now(f1x, f1y, f1z);
    const double Fx = f1x - f0x, Fy = f1y - f0y, Fz = f1z - f0z;
    vir[0] -= 0.5 * sx * Fx; vir[1] -= 0.5 * sy * Fy; vir[2] -= 0.5 * sz * Fz;
    vir[3] -= 0.5 * sx * Fy; vir[4] -= 0.5 * sx * Fz; vir[5] -= 0.5 * sy * Fz;
  };

  // ------------------------------------------------------------------ pipelined traversal
  // List entries stream from DRAM (4 B per neighbour, read once): each thread copies its own
  // entries RING rows ahead into a shared-memory ring with cp.async (no registers, no barrier:
  // a thread only ever reads the slots it wrote).  The records of neighbour k+1 are requested
  // before neighbour k is evaluated.
  const int nn_all = valid ? d.numneigh[i] : 0;
  const int nn = LANES == 1 ? nn_all : (nn_all > sub ? (nn_all - sub + LANES - 1) / LANES : 0);
  const int *np = d.neigh + i + (LANES == 1 ? (size_t)0 : (size_t)sub * d.stride);
  const size_t stride = d.stride * LANES;
  auto fetch2 = [&](int k) {   // entries k, k+1 -> ring slots k % RING, (k+1) % RING; one group
    if (k < nn) cp_async4(myring + (k % RING) * PTH, np + (size_t)k * stride);
    if (k + 1 < nn) cp_async4(myring + ((k + 1) % RING) * PTH, np + (size_t)(k + 1) * stride);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // groups allowed in flight after a wait: entries <= kk + 17 - 2 * PEND have landed
  constexpr int PEND = RING / 2 - 2;
#pragma unroll
  for (int k = 0; k < RING; k += 2) fetch2(k);
  asm volatile("cp.async.wait_group %0;" ::"n"(PEND) : "memory");   // entries 0..3 landed
  int e0 = nn > 0 ? myring[0] : 0;
  int e1 = nn > 1 ? myring[PTH] : 0;
  // Register pipeline: while the record of one neighbour is evaluated, the record of the next one is in flight.
  // ptxas puts all six record loads of the loop on ONE scoreboard, and a scoreboard is a counter: the first use of a
  // record waits until EVERY load issued before it has returned.  Issued in source order (next record's loads, then
  // the visit), the wait for record k therefore also waited for record k+1, requested five instructions earlier -- a
  // full memory latency exposed on every second visit (ncu source page: one DADD held 33 % of all stall samples).
  // The loads of record k+1 are therefore made DATA dependent on the first use of record k (`gate`: a NaN test the
  // compiler cannot fold), so they are issued right after record k has arrived and have a whole visit to complete.
  const int nsp = SPECIES ? co.nspecies : 0;
  auto gate = [&](const Rec4 &A) { const double g = acc.xi - A.x; return g != g ? 1 : 0; };
#if PAIR_BC_LATE
  // -DPAIR_BC_LATE=1 (experiment): only the position quarter A of the next record is requested one visit ahead; B and C
  // (velocities, V, P/rho^2: the other 64 bytes of the same one or two lines) are requested when the visit starts and
  // first used after the distance test and the square root.  16 registers less.
  Rec4 A0, A1, B, C;
  double D0 = 0.0, D1 = 0.0, S0 = 0.0, S1 = 0.0;
  {
    const Prec *p = d.prec + (e0 & NEIGH_JMASK);
    A0 = p->A;
    if (FILTER) D0 = d.pD[e0 & NEIGH_JMASK].x;
    if (SPECIES) S0 = d.pCs[(size_t)(e0 & NEIGH_JMASK) * nsp];
  }
  for (int kk = 0; kk < nn; kk += 2) {
    {
      const Prec *p0 = d.prec + (e0 & NEIGH_JMASK);
      B = p0->B; C = p0->C;
      const int j1 = (e1 & NEIGH_JMASK) + gate(A0);
      const Prec *p = d.prec + j1;
      A1 = p->A;
      if (FILTER) D1 = d.pD[j1].x;
      if (SPECIES) S1 = d.pCs[(size_t)j1 * nsp];
    }
    fetch2(kk + RING);
    asm volatile("cp.async.wait_group %0;" ::"n"(PEND) : "memory");
    const int e2 = kk + 2 < nn ? myring[((kk + 2) % RING) * PTH] : 0;
    const int e3 = kk + 3 < nn ? myring[((kk + 3) % RING) * PTH] : 0;
    visit(e0, A0, B, C, D0, S0);
    {
      const Prec *p1 = d.prec + (e1 & NEIGH_JMASK);
      B = p1->B; C = p1->C;
      const int j2 = (e2 & NEIGH_JMASK) + gate(A1);
      const Prec *p = d.prec + j2;
      A0 = p->A;
      if (FILTER) D0 = d.pD[j2].x;
      if (SPECIES) S0 = d.pCs[(size_t)j2 * nsp];
    }
    if (kk + 1 < nn) visit(e1, A1, B, C, D1, S1);
    e0 = e2;
    e1 = e3;
  }
#else
  Rec4 A0, B0, C0, A1, B1, C1;
  double D0 = 0.0, D1 = 0.0;   // rhoI_j of the Shepard numerator, part of the pipeline on filter steps
  double S0 = 0.0, S1 = 0.0;   // C_j[0], part of the pipeline when species are transported
  {
    const Prec *p = d.prec + (e0 & NEIGH_JMASK);
    A0 = p->A; B0 = p->B; C0 = p->C;
    if (FILTER) D0 = d.pD[e0 & NEIGH_JMASK].x;
    if (SPECIES) S0 = d.pCs[(size_t)(e0 & NEIGH_JMASK) * nsp];
  }
  for (int kk = 0; kk < nn; kk += 2) {
    {
      const int j1 = (e1 & NEIGH_JMASK) + gate(A0);
      const Prec *p = d.prec + j1;
      A1 = p->A; B1 = p->B; C1 = p->C;
      if (FILTER) D1 = d.pD[j1].x;
      if (SPECIES) S1 = d.pCs[(size_t)j1 * nsp];
    }
    fetch2(kk + RING);   // slots of entries kk, kk+1: already in e0, e1
    asm volatile("cp.async.wait_group %0;" ::"n"(PEND) : "memory");   // entries <= kk+5 landed
    const int e2 = kk + 2 < nn ? myring[((kk + 2) % RING) * PTH] : 0;
    const int e3 = kk + 3 < nn ? myring[((kk + 3) % RING) * PTH] : 0;
    visit(e0, A0, B0, C0, D0, S0);
    {
      const int j2 = (e2 & NEIGH_JMASK) + gate(A1);
      const Prec *p = d.prec + j2;
      A0 = p->A; B0 = p->B; C0 = p->C;
      if (FILTER) D0 = d.pD[j2].x;
      if (SPECIES) S0 = d.pCs[(size_t)j2 * nsp];
    }
    if (kk + 1 < nn) visit(e1, A1, B1, C1, D1, S1);
    e0 = e2;
    e1 = e3;
  }
#endif

  if (VIRIAL) {
    // only atoms next to a periodic face have anything to add: plain atomics
#pragma unroll
    for (int q = 0; q < 6; q++)
      if (vir[q] != 0.0) atomicAdd(virial_out + q, vir[q]);
    return;
  }
  if (LANES > 1) {
    acc.fx = lane_sum<LANES>(acc.fx); acc.fy = lane_sum<LANES>(acc.fy); acc.fz = lane_sum<LANES>(acc.fz);
    acc.spi = lane_sum<LANES>(acc.spi);
    acc.drho = lane_sum<LANES>(acc.drho); acc.nd = lane_sum<LANES>(acc.nd); acc.rA2 = lane_sum<LANES>(acc.rA2);
    if (FILTER) acc.rA1 = lane_sum<LANES>(acc.rA1);
    acc.ddvx = lane_sum<LANES>(acc.ddvx); acc.ddvy = lane_sum<LANES>(acc.ddvy); acc.ddvz = lane_sum<LANES>(acc.ddvz);
    if (SOLIDS) {
      acc.phi = lane_sum<LANES>(acc.phi);
      acc.nwx = lane_sum<LANES>(acc.nwx); acc.nwy = lane_sum<LANES>(acc.nwy); acc.nwz = lane_sum<LANES>(acc.nwz);
    }
    if (VARIANT != SPHBVF_TV) {
      acc.ddxx = lane_sum<LANES>(acc.ddxx); acc.ddxy = lane_sum<LANES>(acc.ddxy); acc.ddxz = lane_sum<LANES>(acc.ddxz);
    }
    if (SOLIDS == 2)
#pragma unroll
      for (int k = 0; k < 9; k++) acc.ddev(k) = lane_sum<LANES>(acc.ddev(k));
    if (SPECIES)
#pragma unroll
      for (int k = 0; k < MAXS; k++) acc.Qs[k] = lane_sum<LANES>(acc.Qs[k]);
    if (sub != 0 || !valid) return;
  }
  const double ddvc = 10.0 * 7.0 * co.B[acc.ti];
  const size_t i3 = 3 * (size_t)i;
  d.f[i3] = fma(acc.spi, acc.vxi, acc.fx); d.f[i3 + 1] = fma(acc.spi, acc.vyi, acc.fy); d.f[i3 + 2] = fma(acc.spi, acc.vzi, acc.fz);
  d.drho[i] = acc.drho;
  d.nd[i] = acc.nd;
  d.rhoAux1[i] = acc.rA1;
  d.rhoAux2[i] = acc.rA2;
  d.phi[i] = acc.phi;
  d.nw[i3] = acc.nwx; d.nw[i3 + 1] = acc.nwy; d.nw[i3 + 2] = acc.nwz;
  d.ddv[i3] = ddvc * acc.ddvx; d.ddv[i3 + 1] = ddvc * acc.ddvy; d.ddv[i3 + 2] = ddvc * acc.ddvz;
  if (VARIANT != SPHBVF_TV) {
    d.ddx[i3] = acc.ddxx; d.ddx[i3 + 1] = acc.ddxy; d.ddx[i3 + 2] = acc.ddxz;
    d.Pnew[i] = acc.Pi;   // pair_ssa_tsdpd_bvf_mechanics.cpp:188
  }
  if (SOLIDS == 2)
#pragma unroll
    for (int k = 0; k < 9; k++) d.ddev[9 * (size_t)i + k] = acc.si ? acc.ddev(k) : 0.0;
  if (SPECIES) {
#pragma unroll
    for (int k = 0; k < MAXS; k++)
      if (k < co.nspecies) d.Q[(size_t)i * co.nspecies + k] = acc.Qs[k];
  }
}

// Gather-form kernel.  Two schedules of the chunks of PTH = pair_threads(SOLIDS) consecutive atoms:
//   queues == nullptr (default): one CTA per chunk (chunk = blockIdx.x), the hardware hands CTAs to SMs as they free up, so the
//     CTAs that share an SM work on unrelated parts of the brick and every CTA starts on a cold L1;
//   queues != nullptr (SPHBVF_PAIR_SCHED=smid): PERSISTENT CTAs, two per SM, that pull chunks from a queue chosen by the SM they run
//     on (%smid): SM s owns the chunks [s cpq, (s + 1) cpq) -- a contiguous run of tiles -- and its two CTAs take them
//     in order, so consecutive and concurrent chunks of an SM are neighbouring tiles whose candidate records (half
//     of them shared) are already in that SM's L1; the last tenth of the chunks sits in one shared queue that
//     evens out the tail.  The 16 % of record gathers that missed L1 (first touch of a line by an SM) put an L2
//     round trip on nearly every warp-visit: this is the cheapest way to make fewer of them.  MEASURED (8 M atoms,
//     gpurun_out/r2g_*): L1 sector hit rate 84 -> 89 %, L1 data pipe 80 -> 71 %, but 6.08 instead of 5.66 ms -- the
//     launch loses the hardware's dynamic balance (296 CTAs whose chunk costs differ by the wall / bulk mix) and pays a
//     block-wide barrier pair plus an atomic round trip per chunk; kept as a switch, not the default.
// PERSIST is a template parameter so that the default launch carries none of the queue code (three inlined copies of
// the neighbour loop in one kernel cost the species instantiations up to 150 bytes of extra spills); the persistent
// schedules exist for the instantiations without species, elastic solids, noise or virial (can_persist).
template <bool SPECIES, int SOLIDS, bool RANDOM, bool VIRIAL>
struct CanPersist { static constexpr bool value = !SPECIES && SOLIDS < 2 && !RANDOM && !VIRIAL; };

template <int VARIANT, bool SPECIES, int SOLIDS, bool UNIFORM, bool FILTER, bool RANDOM, bool VIRIAL = false, bool PERSIST = false>
__global__ void __launch_bounds__(pair_threads(SOLIDS), PAIR_MINB)
pair_kernel(const DevState d, const __grid_constant__ Coeffs co, const __grid_constant__ PairTables tb,
            const PairConsts pc, const int *__restrict__ aorder, const int a0, const int a1, int *queues, const int nq,
            const int cpq, double *virial_out = nullptr) {
  constexpr int PTH = pair_threads(SOLIDS);
  __shared__ PairRow srow[UNIFORM ? 1 : MAXT * MAXT];
  __shared__ SpecRow sspec[(SPECIES && !UNIFORM) ? MAXT * MAXT : 1];
  __shared__ int ring[RING][PTH];
  __shared__ double ssol[SOLIDS == 2 ? 18 : 1][SOLIDS == 2 ? PTH : 1];   // devi / ddev columns of the threads
  __shared__ int s_chunk;
  double *mysol = &ssol[0][SOLIDS == 2 ? threadIdx.x : 0];
  if (!UNIFORM) {
    for (int q = threadIdx.x; q < MAXT * MAXT; q += blockDim.x) {
      srow[q] = tb.row[q];
      if (SPECIES) sspec[q] = tb.spec[q];
    }
    __syncthreads();
  }
  int *myring = &ring[0][threadIdx.x];
  // atoms [a0, a1) of the launch, through the atom order when the pass is split (interior of the brick while the
  // halo is in flight, then the atoms that can see a ghost): consecutive positions stay consecutive atoms of a tile
  if (!PERSIST) {
    const int p = a0 + blockIdx.x * PTH + threadIdx.x;
    if (p < a1) pair_atom<VARIANT, SPECIES, SOLIDS, UNIFORM, FILTER, RANDOM, VIRIAL>(d, co, tb, pc, srow, sspec, myring, mysol, aorder ? aorder[p] : p, virial_out);
    return;
  }
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const int q = (int)(smid % (unsigned)nq);
  if (cpq < 0) {
    // SPHBVF_PAIR_SCHED=warp: the same SM-local queues, but every WARP pulls its own chunks of 32 consecutive atoms
    // (-cpq per queue).  No block-wide barrier, no idle warps behind the slowest one of a 192-atom chunk, and the
    // ticket for the next chunk is drawn BEFORE the current chunk is worked on, so the atomic's round trip hides
    // behind ~40 us of work; the warps of an SM walk through one run of tiles together (shared L1 lines).
    const int wcpq = -cpq, lane = threadIdx.x & 31;
    const int nw = (a1 - a0 + 31) / 32, ntail = nw - nq * wcpq;
    auto draw = [&]() { return lane == 0 ? atomicAdd(&queues[q], 1) : 0; };   // may overshoot wcpq: only compared
    int ticket = draw();
    for (;;) {
      int c = __shfl_sync(0xffffffffu, ticket, 0), chunk = nw;
      if (c < wcpq) chunk = q * wcpq + c;
      else {
        if (lane == 0) {   // own queue empty: the shared tail, then other SMs' queues
          if (queues[nq] < ntail && (c = atomicAdd(&queues[nq], 1)) < ntail) chunk = nq * wcpq + c;
          else
            for (int k = 1; k < nq; k++) {
              const int qq = (q + k) % nq;
              if (queues[qq] < wcpq && (c = atomicAdd(&queues[qq], 1)) < wcpq) { chunk = qq * wcpq + c; break; }
            }
        }
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
      }
      if (chunk >= nw) break;
      ticket = draw();
      const int p = a0 + chunk * 32 + lane;
      if (p < a1) pair_atom<VARIANT, SPECIES, SOLIDS, UNIFORM, FILTER, RANDOM, VIRIAL>(d, co, tb, pc, srow, sspec, myring, mysol, aorder ? aorder[p] : p, virial_out);
      asm volatile("cp.async.wait_all;" ::: "memory");   // the ring columns are reused by the next chunk
      __syncwarp();
    }
    return;
  }
  const int nchunks = (a1 - a0 + PTH - 1) / PTH;
  for (;;) {
    __syncthreads();   // s_chunk of the previous round has been read by everyone
    if (threadIdx.x == 0) {
      // own queue, then the shared tail, then whatever another SM's queue still holds (%smid values need not cover
      // 0 .. nq-1: a queue no SM maps to is emptied this way)
      int chunk = nchunks;
      const int ntail = nchunks - nq * cpq;
      int c = queues[q] < cpq ? atomicAdd(&queues[q], 1) : cpq;
      if (c < cpq) chunk = q * cpq + c;
      else if (queues[nq] < ntail && (c = atomicAdd(&queues[nq], 1)) < ntail) chunk = nq * cpq + c;
      else
        for (int k = 1; k < nq; k++) {
          const int qq = (q + k) % nq;
          if (queues[qq] < cpq && (c = atomicAdd(&queues[qq], 1)) < cpq) { chunk = qq * cpq + c; break; }
        }
      s_chunk = chunk;
    }
    __syncthreads();
    const int chunk = s_chunk;
    if (chunk >= nchunks) break;
    const int p = a0 + chunk * PTH + threadIdx.x;
    if (p < a1) pair_atom<VARIANT, SPECIES, SOLIDS, UNIFORM, FILTER, RANDOM, VIRIAL>(d, co, tb, pc, srow, sspec, myring, mysol, aorder ? aorder[p] : p, virial_out);
    asm volatile("cp.async.wait_all;" ::: "memory");   // the ring is reused by the next chunk
  }
}

// Gather form for SMALL systems (the reference's shipped decks: 3 k - 42 k atoms).  With one thread per atom such a
// launch leaves most of the GPU empty and its duration is the latency of ONE thread's serial walk over its neighbours
// (16.6 us for the 3 136-atom cavity: ~20 visits x ~1 600 cycles of dependent entry -> record -> arithmetic).  Here
// PAIR_SPLIT adjacent lanes share an atom and each walks a quarter of its list (same pipeline, same arithmetic body);
// the sums meet in a butterfly.  Only the summation order differs from pair_kernel.
#ifndef PAIR_SPLIT
#define PAIR_SPLIT 4
#endif
template <int VARIANT, bool SPECIES, int SOLIDS, bool UNIFORM, bool FILTER, bool RANDOM>
__global__ void __launch_bounds__(pair_threads(SOLIDS), PAIR_MINB)
pair_split_kernel(const DevState d, const __grid_constant__ Coeffs co, const __grid_constant__ PairTables tb,
                  const PairConsts pc, const int *__restrict__ aorder, const int a0, const int a1) {
  constexpr int PTH = pair_threads(SOLIDS);
  __shared__ PairRow srow[UNIFORM ? 1 : MAXT * MAXT];
  __shared__ SpecRow sspec[(SPECIES && !UNIFORM) ? MAXT * MAXT : 1];
  __shared__ int ring[RING][PTH];
  __shared__ double ssol[SOLIDS == 2 ? 18 : 1][SOLIDS == 2 ? PTH : 1];
  double *mysol = &ssol[0][SOLIDS == 2 ? threadIdx.x : 0];
  if (!UNIFORM) {
    for (int q = threadIdx.x; q < MAXT * MAXT; q += blockDim.x) {
      srow[q] = tb.row[q];
      if (SPECIES) sspec[q] = tb.spec[q];
    }
    __syncthreads();
  }
  const int t = blockIdx.x * PTH + threadIdx.x;
  const int p = a0 + t / PAIR_SPLIT;
  const bool valid = p < a1;
  const int pp = valid ? p : a0;
  pair_atom<VARIANT, SPECIES, SOLIDS, UNIFORM, FILTER, RANDOM, false, PAIR_SPLIT>(
      d, co, tb, pc, srow, sspec, &ring[0][threadIdx.x], mysol, aorder ? aorder[pp] : pp, nullptr, t % PAIR_SPLIT, valid);
}

// ==========================================================================================
// tile form
// ==========================================================================================
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}

// sum of x over the 8 lanes of an octet (xor butterfly): for the rarely used sums
__device__ __forceinline__ double octet_sum(double x) {
  x += __shfl_xor_sync(0xffffffffu, x, 4);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  return x;
}

// Halving butterfly over the 8 lanes of an octet: every lane enters with its partial sums v[0..15] and leaves with
// the octet totals of TWO of them in v[0], v[1]: lane l (0..7 within the octet) holds the totals of the
// original indices 8 b2 + 4 b1 + 2 b0 and that + 1, with b2 b1 b0 the bits of l.  14 shuffle+add pairs instead of 48.
__device__ __forceinline__ void octet_reduce16(double (&v)[16], const int lane) {
  {
    const bool up = lane & 4;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const double send = up ? v[k] : v[k + 8], keep = up ? v[k + 8] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = lane & 2;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double send = up ? v[k] : v[k + 4], keep = up ? v[k + 4] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
  }
  {
    const bool up = lane & 1;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const double send = up ? v[k] : v[k + 2], keep = up ? v[k + 2] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
  }
}

size_t pair_tile_smem(int cap, bool index_map) { return (size_t)cap * (PREC_GRANULES * 16 + (index_map ? 4 : 0)); }

template <int VARIANT, bool SPECIES, int SOLIDS, bool UNIFORM, bool FILTER, bool RANDOM, bool VIRIAL = false>
__global__ void __launch_bounds__(PT_T, (SOLIDS == 2 || SPECIES || RANDOM) ? 1 : PT_MINB)   // the heavy instantiations get 255 registers
pair_tile_kernel(const DevState d, const __grid_constant__ Grid g, const __grid_constant__ Coeffs co,
                 const __grid_constant__ PairTables tb, const int *__restrict__ cell_start,
                 const int *__restrict__ gcell_start, const int *__restrict__ gorder, const int *__restrict__ tile_list,
                 const PairConsts pc, double *virial_out = nullptr) {
  constexpr int NG = PREC_GRANULES;
  constexpr bool NEEDJ = NeedsJ<SPECIES, SOLIDS, FILTER, RANDOM, VIRIAL>::value;
  extern __shared__ __align__(16) unsigned char pt_smem[];
  __shared__ PairRow srow[UNIFORM ? 1 : MAXT * MAXT];
  __shared__ int seg_src[TB_MAXSEG];
  __shared__ int seg_off[TB_MAXSEG + 1];
  __shared__ int next_group;

  TileGeom t;
  // tile_list: the tiles of this launch (interior tiles while the halo is in flight, then the ones that see ghosts)
  if (!tile_geometry(g, tile_list ? tile_list[blockIdx.x] : blockIdx.x, cell_start, t)) return;   // no owned atom (whole CTA)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cap = d.tile_cap;
  double2 *rec = reinterpret_cast<double2 *>(pt_smem);        // rec[gr * cap + slot], cap a multiple of 8: bank group = slot mod 8
  int *gidx = reinterpret_cast<int *>(pt_smem + (size_t)NG * 16 * cap);   // [cap] global index of a slot (NEEDJ only)
  if (!UNIFORM)
    for (int q = tid; q < MAXT * MAXT; q += PT_T) srow[q] = tb.row[q];
  if (tid == 0) next_group = 0;
  tile_segments(g, t, cell_start, gcell_start, d.nghost > 0, seg_src, seg_off);   // ends with a barrier

  // ---- stage the candidates' records: one warp per segment (a contiguous index range, or ghosts through gorder),
  // lanes over (atom, granule) so that consecutive lanes copy consecutive 16-byte pieces of global memory
  for (int sid = warp; sid < t.nseg; sid += PT_T / 32) {
    const int s0 = seg_off[sid], len = seg_off[sid + 1] - s0;
    if (len <= 0) continue;
    const int src = seg_src[sid];
    for (int idx = lane; idx < len * NG; idx += 32) {
      const int k = idx / NG, gr = idx - k * NG;
      const int j = tile_source(src, k, d.nlocal, gorder);
      cp_async16(&rec[(size_t)gr * cap + s0 + k], reinterpret_cast<const char *>(d.prec + j) + 16 * gr);
      if (NEEDJ && gr == 0) gidx[s0 + k] = j;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- octets: 8 lanes per atom, 4 atoms per warp at a time; groups of 4 atoms are handed out dynamically
  const int sub = lane & 7, oct = lane >> 3;
  const int natoms = t.last - t.first;
  for (;;) {
    int gbase = 0;
    if (lane == 0) gbase = atomicAdd(&next_group, 4);
    gbase = __shfl_sync(0xffffffffu, gbase, 0);
    if (gbase >= natoms) break;
    const bool valid = gbase + oct < natoms;
    const int i = t.first + (valid ? gbase + oct : 0);
    const int nn = valid ? d.numneigh[i] : 0;
    int nnmax = nn;
    nnmax = max(nnmax, __shfl_xor_sync(0xffffffffu, nnmax, 8));
    nnmax = max(nnmax, __shfl_xor_sync(0xffffffffu, nnmax, 16));
    const unsigned short *row = d.neigh16 + (size_t)i * d.pitch16 + sub;   // entry k of my lane at step s: k = 8 s + sub
    // The entries of a block of 8 steps (64 entries of the row) travel as eight 16-bit loads packed into two 64-bit
    // registers; the NEXT block is requested before the current one is evaluated (the list is the only DRAM stream
    // of the loop), and a step shifts its entry out -- no register array, so the visit loop stays ONE copy of the
    // body in the instruction cache (unrolled eight times it was 77 KB of code).
    auto load_block = [&](const int sb, unsigned long long &lo, unsigned long long &hi) {
      unsigned e[8];
#pragma unroll
      for (int q = 0; q < 8; q++) e[q] = (8 * (sb + q) + sub < nn) ? row[8 * (sb + q)] : 0u;
      lo = (unsigned long long)e[0] | (unsigned long long)e[1] << 16 | (unsigned long long)e[2] << 32 | (unsigned long long)e[3] << 48;
      hi = (unsigned long long)e[4] | (unsigned long long)e[5] << 16 | (unsigned long long)e[6] << 32 | (unsigned long long)e[7] << 48;
    };
    unsigned long long cur_lo, cur_hi, nxt_lo, nxt_hi;
    load_block(0, cur_lo, cur_hi);

    PairAcc<VARIANT, SPECIES, SOLIDS, UNIFORM, FILTER, RANDOM> acc;
    acc.init(d, co, tb, srow, i, d.pflags[i], d.prec[i].A, d.prec[i].B, d.prec[i].C);
    double vir[6] = {0, 0, 0, 0, 0, 0};

    const int nsteps = (nnmax + 7) >> 3;   // warp-uniform
    for (int s0 = 0; s0 < nsteps; s0 += 8) {
      load_block(s0 + 8, nxt_lo, nxt_hi);
      const int send = min(s0 + 8, nsteps);
#pragma unroll 1
      for (int st = s0; st < send; st++) {
        const unsigned e = (unsigned)cur_lo & 0xffffu;
        cur_lo = (cur_lo >> 16) | (cur_hi << 48);
        cur_hi >>= 16;
        const int k = 8 * st + sub;
        const int slot = e & TILE_SLOT_MASK;
        // granules of the record: {x, y} {z, rho} {vest.x, vest.y} {vest.z, V} {w.x, w.y} {w.z, P/rho^2}
        const double2 g0 = rec[slot], g1 = rec[cap + slot], g2 = rec[2 * cap + slot], g3 = rec[3 * cap + slot],
                      g4 = rec[4 * cap + slot], g5 = rec[5 * cap + slot];
        if (k < nn) {
          const int tj = (e >> TILE_SLOT_BITS) & 7;
          const bool sj = (e >> 15) & 1;
          const int j = NEEDJ ? gidx[slot] : 0;
          const double rhoIj = FILTER ? d.pD[j].x : 0.0;
          const double Cj0 = SPECIES ? d.pCs[(size_t)j * co.nspecies] : 0.0;
          if (!VIRIAL) {
            acc.visit(d, co, tb, pc, tj, sj, j, g0.x, g0.y, g1.x, g1.y, g2.x, g2.y, g3.x, g3.y, g4.x, g4.y, g5.x, g5.y, rhoIj, Cj0);
          } else {
            const int gh = j - d.nlocal;
            if (gh >= 0) {
              const double sx = d.gshift[3 * (size_t)gh], sy = d.gshift[3 * (size_t)gh + 1], sz = d.gshift[3 * (size_t)gh + 2];
              if (sx != 0.0 || sy != 0.0 || sz != 0.0) {
                double f0x, f0y, f0z, f1x, f1y, f1z;
                acc.force_now(f0x, f0y, f0z);
                acc.visit(d, co, tb, pc, tj, sj, j, g0.x, g0.y, g1.x, g1.y, g2.x, g2.y, g3.x, g3.y, g4.x, g4.y, g5.x, g5.y, rhoIj, Cj0);
                acc.force_now(f1x, f1y, f1z);
                const double Fx = f1x - f0x, Fy = f1y - f0y, Fz = f1z - f0z;
                vir[0] -= 0.5 * sx * Fx; vir[1] -= 0.5 * sy * Fy; vir[2] -= 0.5 * sz * Fz;
                vir[3] -= 0.5 * sx * Fy; vir[4] -= 0.5 * sx * Fz; vir[5] -= 0.5 * sy * Fz;
              }
            }
          }
        }
      }
      cur_lo = nxt_lo;
      cur_hi = nxt_hi;
    }

    if (VIRIAL) {
#pragma unroll
      for (int q = 0; q < 6; q++)
        if (vir[q] != 0.0) atomicAdd(virial_out + q, vir[q]);
      continue;
    }

    // ---- combine the 8 partial sums of every atom and write its outputs (each written once, no atomics)
    const double ddvc = 10.0 * 7.0 * co.B[acc.ti];
    double v[16] = {fma(acc.spi, acc.vxi, acc.fx), fma(acc.spi, acc.vyi, acc.fy), fma(acc.spi, acc.vzi, acc.fz), acc.drho,
                    acc.nd, acc.rA1, acc.rA2, acc.phi,
                    acc.nwx, acc.nwy, acc.nwz, 0.0,
                    ddvc * acc.ddvx, ddvc * acc.ddvy, ddvc * acc.ddvz, 0.0};
    octet_reduce16(v, lane);
    if (valid) {
      const size_t i3 = 3 * (size_t)i;
      switch (sub) {   // lane b2 b1 b0 holds the totals of indices 8 b2 + 4 b1 + 2 b0 (+1)
        case 0: d.f[i3] = v[0]; d.f[i3 + 1] = v[1]; break;
        case 1: d.f[i3 + 2] = v[0]; d.drho[i] = v[1]; break;
        case 2: d.nd[i] = v[0]; d.rhoAux1[i] = v[1]; break;
        case 3: d.rhoAux2[i] = v[0]; d.phi[i] = v[1]; break;
        case 4: d.nw[i3] = v[0]; d.nw[i3 + 1] = v[1]; break;
        case 5: d.nw[i3 + 2] = v[0]; break;
        case 6: d.ddv[i3] = v[0]; d.ddv[i3 + 1] = v[1]; break;
        default: d.ddv[i3 + 2] = v[0]; break;
      }
    }
    if (VARIANT != SPHBVF_TV) {
      const double sx = octet_sum(acc.ddxx), sy = octet_sum(acc.ddxy), sz = octet_sum(acc.ddxz);
      if (valid && sub == 0) {
        const size_t i3 = 3 * (size_t)i;
        d.ddx[i3] = sx; d.ddx[i3 + 1] = sy; d.ddx[i3 + 2] = sz;
        d.Pnew[i] = acc.Pi;   // pair_ssa_tsdpd_bvf_mechanics.cpp:188
      }
    }
    if (SOLIDS == 2) {
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const double sk = octet_sum(acc.ddev(k));
        if (valid && sub == 0) d.ddev[9 * (size_t)i + k] = acc.si ? sk : 0.0;
      }
    }
    if (SPECIES) {
#pragma unroll
      for (int k = 0; k < MAXS; k++) {
        const double sk = octet_sum(acc.Qs[k]);
        if (valid && sub == 0 && k < co.nspecies) d.Q[(size_t)i * co.nspecies + k] = sk;
      }
    }
  }
}

// ==========================================================================================
// dispatch
// ==========================================================================================
struct TileArgs {
  const Grid *g;
  const NeighWork *w;
  const int *tile_list;   // tile form: nullptr = every tile of the grid
  int ntiles;
  const int *aorder;      // gather form: positions [a0, a1) of this atom order (nullptr: atoms a0 .. a1)
  int a0, a1;
  int *queues;            // gather form: per-SM chunk queues of the persistent schedule (nullptr: one CTA per chunk)
  int nq;
};

template <int VARIANT, bool SPECIES, int SOLIDS, bool UNIFORM, bool FILTER, bool RANDOM, bool VIRIAL>
static void launch_one(const DevState &d, const Coeffs &co, const PairTables &tb, const PairConsts &pc, const TileArgs &ta,
                       double *vout, cudaStream_t st) {
  if (d.list16) {
    const size_t smem = pair_tile_smem(d.tile_cap, NeedsJ<SPECIES, SOLIDS, FILTER, RANDOM, VIRIAL>::value);
    auto kern = pair_tile_kernel<VARIANT, SPECIES, SOLIDS, UNIFORM, FILTER, RANDOM, VIRIAL>;
    // opt-in to > 48 KB of dynamic shared memory: per function and per device; repeated only when the size grows
    static size_t opted[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || opted[dev] < smem) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      // two CTAs of ~110 KB per SM need the whole shared-memory carve-out
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (dev >= 0 && dev < 64) opted[dev] = smem;
    }
    if (ta.ntiles <= 0) return;
    kern<<<ta.ntiles, PT_T, smem, st>>>(d, *ta.g, co, tb, ta.w->cell_start, ta.w->gcell_start, ta.w->gorder, ta.tile_list, pc, vout);
  } else {
    if (ta.a1 <= ta.a0) return;
    constexpr int PTH = pair_threads(SOLIDS);
    const int nchunks = (ta.a1 - ta.a0 + PTH - 1) / PTH;
    const int nq = ta.nq < 0 ? -ta.nq : ta.nq;
    // small launches: PAIR_SPLIT lanes per atom (SPHBVF_PAIR_LANES=1 never, =4 always; default: up to split_max atoms)
    static const int split_max = [] { const char *e = getenv("SPHBVF_PAIR_LANES"); return !e ? 65536 : (atoi(e) > 1 ? 0x7fffffff : 0); }();
    if (!VIRIAL && ta.a1 - ta.a0 <= split_max) {
      if constexpr (!VIRIAL) {
        const long nthreads = (long)(ta.a1 - ta.a0) * PAIR_SPLIT;
        pair_split_kernel<VARIANT, SPECIES, SOLIDS, UNIFORM, FILTER, RANDOM><<<(int)((nthreads + PTH - 1) / PTH), PTH, 0, st>>>(
            d, co, tb, pc, ta.aorder, ta.a0, ta.a1);
      }
    } else if (CanPersist<SPECIES, SOLIDS, RANDOM, VIRIAL>::value && ta.queues && nchunks > 4 * nq) {
      // persistent schedule: 9/10 of the chunks in per-SM queues (contiguous runs of tiles), the rest in a shared one;
      // nq < 0 (SPHBVF_PAIR_SCHED=warp): chunks of one warp (32 atoms) drawn per warp, passed as a negative count
      static const double own = [] { const char *e = getenv("SPHBVF_PAIR_TAIL"); const double t = e ? atof(e) : 0.1; return 1.0 - (t > 0.0 && t < 1.0 ? t : 0.1); }();
      const int cpq = ta.nq < 0 ? -(int)(own * ((ta.a1 - ta.a0 + 31) / 32) / nq) : (int)(own * nchunks / nq);
      cudaMemsetAsync(ta.queues, 0, sizeof(int) * (nq + 1), st);
      if constexpr (CanPersist<SPECIES, SOLIDS, RANDOM, VIRIAL>::value)
        pair_kernel<VARIANT, SPECIES, SOLIDS, UNIFORM, FILTER, RANDOM, VIRIAL, true><<<PAIR_MINB * nq, PTH, 0, st>>>(
            d, co, tb, pc, ta.aorder, ta.a0, ta.a1, ta.queues, nq, cpq, vout);
    } else {
      pair_kernel<VARIANT, SPECIES, SOLIDS, UNIFORM, FILTER, RANDOM, VIRIAL><<<nchunks, PTH, 0, st>>>(
          d, co, tb, pc, ta.aorder, ta.a0, ta.a1, nullptr, 1, 0, vout);
    }
  }
  SPHBVF_LAUNCHED(1);
}

template <int VARIANT, bool SPECIES, int SOLIDS, bool UNIFORM>
static void launch_filter(const DevState &d, const Coeffs &co, const PairTables &tb, const PairFlags &pf, const PairConsts &pc,
                          const TileArgs &ta, cudaStream_t st) {
  // the stochastic variant always carries the Shepard numerator (one instantiation less per case)
  if (pf.random) launch_one<VARIANT, SPECIES, SOLIDS, UNIFORM, true, true, false>(d, co, tb, pc, ta, nullptr, st);
  else if (pf.filter_step) launch_one<VARIANT, SPECIES, SOLIDS, UNIFORM, true, false, false>(d, co, tb, pc, ta, nullptr, st);
  else launch_one<VARIANT, SPECIES, SOLIDS, UNIFORM, false, false, false>(d, co, tb, pc, ta, nullptr, st);
}

template <int VARIANT, bool SPECIES>
static void launch_solids(const DevState &d, const Coeffs &co, const PairTables &tb, const PairFlags &pf, const PairConsts &pc,
                          const TileArgs &ta, bool uniform, cudaStream_t st) {
  if (d.nlocal == 0) return;
  const int solids = !pf.any_solid ? 0 : (pf.with_dev ? 2 : 1);
  // the constant-coefficient fast path is instantiated for the rigid-wall / no-solid cases only
  if (solids == 0) {
    if (uniform) launch_filter<VARIANT, SPECIES, 0, true>(d, co, tb, pf, pc, ta, st);
    else launch_filter<VARIANT, SPECIES, 0, false>(d, co, tb, pf, pc, ta, st);
  } else if (solids == 1) {
    if (uniform) launch_filter<VARIANT, SPECIES, 1, true>(d, co, tb, pf, pc, ta, st);
    else launch_filter<VARIANT, SPECIES, 1, false>(d, co, tb, pf, pc, ta, st);
  } else launch_filter<VARIANT, SPECIES, 2, false>(d, co, tb, pf, pc, ta, st);
}

template <int VARIANT>
static void launch_species(const DevState &d, const Coeffs &co, const PairTables &tb, const PairFlags &pf, const PairConsts &pc,
                           const TileArgs &ta, bool uniform, cudaStream_t st) {
  if (co.nspecies > 0) launch_solids<VARIANT, true>(d, co, tb, pf, pc, ta, uniform, st);
  else launch_solids<VARIANT, false>(d, co, tb, pf, pc, ta, uniform, st);
}

// sum_i x_i (x) f_i over the owned atoms, LAMMPS order xx yy zz xy xz yz with virial[ab] = x_a f_b
__global__ void virial_fdotr_kernel(const DevState d, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[6] = {0, 0, 0, 0, 0, 0};
  if (i < d.nlocal) {
    const size_t i3 = 3 * (size_t)i;
    const double x = d.x[i3], y = d.x[i3 + 1], z = d.x[i3 + 2], fx = d.f[i3], fy = d.f[i3 + 1], fz = d.f[i3 + 2];
    v[0] = fx * x; v[1] = fy * y; v[2] = fz * z; v[3] = fy * x; v[4] = fz * x; v[5] = fz * y;
  }
  __shared__ double sh[6][8];
#pragma unroll
  for (int q = 0; q < 6; q++) {
    double t = v[q];
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = t;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh[threadIdx.x][w];
    atomicAdd(out + threadIdx.x, t);
  }
}

template <int VARIANT, bool SPECIES>
static void launch_virial_solids(const DevState &d, const Coeffs &co, const PairTables &tb, const PairFlags &pf,
                                 const PairConsts &pc, const TileArgs &ta, double *out, cudaStream_t st) {
  const int solids = !pf.any_solid ? 0 : (pf.with_dev ? 2 : 1);
  if (solids == 0) launch_one<VARIANT, SPECIES, 0, false, false, false, true>(d, co, tb, pc, ta, out, st);
  else if (solids == 1) launch_one<VARIANT, SPECIES, 1, false, false, false, true>(d, co, tb, pc, ta, out, st);
  else launch_one<VARIANT, SPECIES, 2, false, false, false, true>(d, co, tb, pc, ta, out, st);
}

static PairConsts consts_of(const PairFlags &pf) {
  PairConsts pc;
  pc.damp = pf.damp;
  pc.rand_pref = pf.rand_pref;
  pc.seed = pf.seed;
  pc.ntimestep = pf.ntimestep;
  pc.elmask = pf.with_dev;
  return pc;
}

// Pair::virial_fdotr_compute for the gather formulation; out[6] must be zeroed by the caller
static TileArgs all_atoms(const DevState &d, const Grid &g, const NeighWork &w) {
  TileArgs ta = {&g, &w, nullptr, (int)((long)g.nt[0] * g.nt[1] * g.nt[2]), nullptr, 0, d.nlocal, nullptr, 1};
  return ta;
}

void launch_virial(const DevState &d, const Coeffs &co, const PairFlags &pf, const Grid &g, const NeighWork &w, double *out,
                   cudaStream_t st) {
  if (!d.nlocal) return;
  virial_fdotr_kernel<<<(d.nlocal + 255) / 256, 256, 0, st>>>(d, out);
  SPHBVF_LAUNCHED(1);
  if (!d.nghost) return;
#ifndef SPHBVF_HOT_ONLY
  PairTables tb;
  make_tables(co, tb);
  PairConsts pc = consts_of(pf);
  pc.rand_pref = 0.0;
  pc.seed = 0ULL;
  const TileArgs ta = all_atoms(d, g, w);
  const bool sp = co.nspecies > 0;
  switch (co.variant) {
    case SPHBVF_TV: sp ? launch_virial_solids<SPHBVF_TV, true>(d, co, tb, pf, pc, ta, out, st) : launch_virial_solids<SPHBVF_TV, false>(d, co, tb, pf, pc, ta, out, st); break;
    case SPHBVF_MECHANICS: sp ? launch_virial_solids<SPHBVF_MECHANICS, true>(d, co, tb, pf, pc, ta, out, st) : launch_virial_solids<SPHBVF_MECHANICS, false>(d, co, tb, pf, pc, ta, out, st); break;
    default: sp ? launch_virial_solids<SPHBVF_FSI, true>(d, co, tb, pf, pc, ta, out, st) : launch_virial_solids<SPHBVF_FSI, false>(d, co, tb, pf, pc, ta, out, st); break;
  }
#endif
}

// part: the subset of the owned atoms to process (nullptr: all of them)
void launch_pair(const DevState &d, const Coeffs &co, const PairFlags &pf, const Grid &g, const NeighWork &w,
                 const PairSubset *part, cudaStream_t st) {
  PairTables tb;
  const bool uniform = make_tables(co, tb);
  const PairConsts pc = consts_of(pf);
  TileArgs ta = all_atoms(d, g, w);
  if (part) {
    ta.tile_list = part->tile_list; ta.ntiles = part->ntiles;
    ta.aorder = part->aorder; ta.a0 = part->a0; ta.a1 = part->a1;
    ta.queues = part->queues; ta.nq = part->nq;
  }
#ifdef SPHBVF_HOT_ONLY   // tuning builds (tools/build_variant.sh): only the benchmark's instantiations, seconds to compile
  (void)uniform;
  if (d.nlocal) launch_filter<SPHBVF_TV, false, 1, true>(d, co, tb, pf, pc, ta, st);
  return;
#else
  switch (co.variant) {
    case SPHBVF_TV: launch_species<SPHBVF_TV>(d, co, tb, pf, pc, ta, uniform, st); break;
    case SPHBVF_MECHANICS: launch_species<SPHBVF_MECHANICS>(d, co, tb, pf, pc, ta, uniform, st); break;
    default: launch_species<SPHBVF_FSI>(d, co, tb, pf, pc, ta, uniform, st); break;
  }
#endif
}

}  // namespace sphbvf

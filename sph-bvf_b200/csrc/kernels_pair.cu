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
// loads.  Sweep C (v_weighted_solid / a_weighted_solid) is never consumed and the
// random stress term is omitted (reference seed is clock(); exactly zero in all decks with e=0).
//
// Mapping: one thread per owned atom, 128-thread CTAs, atoms in tile-major cell order so a CTA
// covers a compact cube.  The kernel is FP64-pipe / latency bound, not HBM bound (ncu, profiles/):
//   * the neighbour loop is software pipelined by hand: while neighbour k is evaluated, the three
//     32-byte records of neighbour k+1 (one 256-bit LDG each) and the list entry k+2 are in
//     flight, so an L2 round trip is overlapped with ~75 FP64 instructions of useful work;
//   * list entries are read transposed (coalesced) and carry type_j and solid_tag_j, so the hot
//     loop has no flag gather;
//   * every per-particle division lives in the pack kernel (V = m/rho, P/rho^2), sqrt is a
//     branch-free Goldschmidt iteration on MUFU.RSQ64H (7 FP64 ops, < 1 ulp);
//   * per-type-pair coefficients: when every type pair shares h, eta and mass (all cavity decks
//     and the synthetic lattice) they are kernel-argument constants; otherwise rows of a small
//     shared-memory table (conflict-free broadcast) instead of divergent constant-bank reads.
#include <string.h>

#include "sphbvf_internal.cuh"

namespace sphbvf {

// ---- compile-time tuning switches.  The defaults are what ships; every other setting is a measured and rejected
// experiment kept buildable so that it can be re-measured (tools/build_variant.sh; pair kernel ms
// at 8 M atoms on one B200, default 5.71):
//   PAIR_PIPE      records in flight per warp in the register pipeline: 0 = none (one buffer, 128 registers, 16 warps
//                  per SM: 7.01), 1 = default (two buffers, 160 registers, 12 warps), 2 = two (four buffers, 204
//                  registers, 8 warps: 6.34)
//   PAIR_TMA = D   records through the bulk-copy engine into a shared-memory ring D deep (16.3 with D = 4)
//   PAIR_L2HINT    evict-first on the list / output streams (1) and evict-last on the record gathers (2): 5.84 / 5.90
//   PAIR_PFL2 = D  prefetch.global.L2 of the records 2 D entries ahead: 7.33 with D = 2
//   PAIR_DIAG_SMEM ballast dynamic shared memory (occupancy probe: 7.70 at 2 CTAs, 13.4 at 1 CTA per SM)
//   PAIR_MINB / PAIR_T  resident CTAs per SM the kernel is compiled for / threads per CTA (MINB x T x registers <= 64 K,
//                  split over four register files: only 8 / 12 / 16 warps at <= 255 / 168 / 128 registers exist)
#ifndef PAIR_PIPE
#define PAIR_PIPE 1
#endif
#ifndef PAIR_TMA
#define PAIR_TMA 0
#endif
#ifndef PAIR_T
#define PAIR_T (PAIR_PIPE == 1 ? 192 : 128)   // 12 warps as 2 x 192 threads: 5.63 (3 x 128: 5.71, 6 x 64: 5.73, 1 x 384: 5.66)
#endif
#ifndef PAIR_MINB
#define PAIR_MINB (PAIR_PIPE == 2 ? 2 : (PAIR_PIPE == 0 ? 4 : (PAIR_T == 192 ? 2 : 3)))
#endif

// one row per (type_i, type_j); 8 doubles = 64 B so a row is two LDS.128 x2
struct __align__(16) PairRow {
  double cutsq, h, cwfd, cwf;      // h^2, h, (1/r)dW/dr = cwfd (h-r)^2, W = cwf (h-r)^3 (h+3r)
  double mimj, eta, iwdelta, h2eps;  // m_i m_j, eta, 1/W(delta), 0.01 h^2
};

struct PairTables {
  PairRow row[MAXT * MAXT];
  double cwfdc[MAXT][MAXT];  // same as cwfd with h = cutc
  double hc2eps[MAXT][MAXT]; // 0.01 cutc^2
  double mred2[MAXT][MAXT];  // 2 mi mj / (mi + mj)
  double geff[MAXT][MAXT];   // 2 Gi Gj / (Gi + Gj + 1e-12)
  double imass[MAXT];
};

static bool make_tables(const Coeffs &co, PairTables &t) {
  const double delta_fac = co.variant == SPHBVF_TV ? (1.0 / 2.6) : (1.0 / 3.0);
  memset(&t, 0, sizeof t);
  bool uniform = true;
  for (int i = 1; i <= co.ntypes; i++) {
    t.imass[i] = 1.0 / co.mass[i];
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
      coef(hc > 0 ? hc : h, t.cwfdc[i][j], dummy);
      double delta = delta_fac * h, td = h - delta;
      double wdelta = r.cwf * td * td * td * (h + 3. * delta);
      r.cutsq = co.cutsq[i][j];
      r.h = h;
      r.iwdelta = 1.0 / wdelta;
      r.h2eps = 0.01 * h * h;
      r.mimj = co.mass[i] * co.mass[j];
      r.eta = co.eta[i][j];
      t.hc2eps[i][j] = 0.01 * hc * hc;
      t.mred2[i][j] = 2.0 * ((co.mass[i] * co.mass[j]) / (co.mass[i] + co.mass[j]));
      t.geff[i][j] = (2.0 * co.G0[i] * co.G0[j]) / (co.G0[i] + co.G0[j] + 1e-12);
      const PairRow &r11 = t.row[MAXT + 1];
      if (r.cutsq != r11.cutsq || r.h != r11.h || r.mimj != r11.mimj || r.eta != r11.eta) uniform = false;
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

constexpr int RING = 16;  // list-entry prefetch depth (rows); power of two
constexpr int TMA_RSTRIDE = 112;   // bytes per staged record: 96 + 16 pad, so that the 16-byte LDS of 8 lanes hit 32 banks

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
// one bulk copy per calling lane: `bytes` (multiple of 16) from global to this CTA's shared memory, completion on `b`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// L2 residency (tuning switches):
//   PAIR_L2HINT >= 1: the streams that are touched once per launch (list entries in, per-atom outputs out) are
//                     marked evict-first so they do not push the gathered records out of L2;
//   PAIR_L2HINT >= 2: the record gathers are marked evict-last on top of that;
//   PAIR_PFL2 = D > 0: the records of the entries 2*D ahead of the register pipeline are prefetched into L2.
#ifndef PAIR_L2HINT
#define PAIR_L2HINT 0
#endif
#ifndef PAIR_PFL2
#define PAIR_PFL2 0
#endif

__device__ __forceinline__ void cp_async4(int *smem_dst, const int *gsrc, unsigned long long pol) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
#if PAIR_L2HINT >= 1
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(sa), "l"(gsrc), "l"(pol) : "memory");
#else
  (void)pol;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
#endif
}

__device__ __forceinline__ Rec4 ld_rec(const Rec4 *p) {
#if PAIR_L2HINT >= 2
  Rec4 r;
  asm("ld.global.L2::evict_last.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
#else
  return *p;
#endif
}

__device__ __forceinline__ void st_out(double *p, double v) {
#if PAIR_L2HINT >= 1
  __stcs(p, v);
#else
  *p = v;
#endif
}

__device__ __forceinline__ void prefetch_rec_l2(const Prec *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  asm volatile("prefetch.global.L2 [%0];" ::"l"((const char *)p + 64));
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

// SOLIDS: 0 = no atom has solid_tag, 1 = solids whose deviatoric stress is identically zero
// (rigid walls: G0 == 0, dev == 0), 2 = elastic solids (deviatoric tensors gathered)
// VIRIAL = true turns the kernel into the ghost part of Pair::virial_fdotr_compute (pair.cpp:1511): no
// per-atom output is written; for every neighbour that is a periodic image (shift s != 0) the force
// F it exerts on atom i is obtained as the change of the force accumulator and -1/2 s (x) F is summed
// into virial_out[6] (see sphbvf_virial in capi.cu for the derivation).
template <int VARIANT, bool SPECIES, int SOLIDS, bool UNIFORM, bool FILTER, bool RANDOM, bool VIRIAL = false>
__global__ void __launch_bounds__(PAIR_T, PAIR_MINB)
pair_kernel(const DevState d, const __grid_constant__ Coeffs co, const __grid_constant__ PairTables tb,
            const double damp, const double rand_pref, const unsigned long long seed, const long ntimestep,
            double *virial_out = nullptr) {
  __shared__ PairRow srow[UNIFORM ? 1 : MAXT * MAXT];
  __shared__ int ring[RING][PAIR_T];
  if (!UNIFORM) {
    for (int q = threadIdx.x; q < MAXT * MAXT; q += blockDim.x) srow[q] = tb.row[q];
    __syncthreads();
  }
#if PAIR_TMA
  extern __shared__ __align__(128) unsigned char pair_dyn[];
  unsigned long long *mbar_all = reinterpret_cast<unsigned long long *>(pair_dyn + (size_t)PAIR_TMA * PAIR_T * TMA_RSTRIDE);
  if (threadIdx.x < (PAIR_T / 32) * PAIR_TMA) mbar_init(mbar_all + threadIdx.x, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  // every lane stays alive (warp-wide mbarrier protocol): lanes past the last atom shadow atom 0 with an empty list
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = gi < d.nlocal;
  const int i = live ? gi : 0;
#else
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
#endif

  const int fl = d.pflags[i];
  const int ti = fl & 7;
  const bool si = SOLIDS && ((fl >> 4) & 1);
  const Rec4 Ai = d.prec[i].A, Bi = d.prec[i].B, Ci = d.prec[i].C;
  const double rhoi = Ai.w, Vi = Bi.w, Vi2 = Vi * Vi, Prri = Ci.w;
  const double c0i = co.c0[ti];
  const double Pi = Prri * rhoi * rhoi;
  const double irhoi = Vi * tb.imass[ti];
  const PairRow *myrow = srow + (UNIFORM ? 0 : ti * MAXT);

  double devi[9];
  double arti = 0.0;        // scalar artificial stress when dev == 0
  double Ri[9];             // artificial stress tensor when dev != 0
  const double c_art = VARIANT == SPHBVF_FSI ? 0.1 : 0.35;
  // stress-free solid (dev == 0): R = -c_art max(0, -Psigma)/rho^2 = c_art P/rho^2 where P < 0 (TV, fsi);
  // mechanics uses |P| (:471,487) so the bracket is never positive
  if (SOLIDS == 1) arti = (si && VARIANT != SPHBVF_MECHANICS && Prri < 0.0) ? c_art * Prri : 0.0;
  if (SOLIDS == 2) {
#pragma unroll
    for (int k = 0; k < 9; k++) devi[k] = si ? d.pdev[9 * (size_t)i + k] : 0.0;
    const double Ps = VARIANT == SPHBVF_MECHANICS ? fabs(Pi) : Pi;
#pragma unroll
    for (int m = 0; m < 3; m++)
#pragma unroll
      for (int n = 0; n < 3; n++) {
        double ts = devi[3 * m + n] - (m == n ? Ps : 0.0);
        Ri[3 * m + n] = (si && ts > 0.0) ? -c_art * ts * irhoi * irhoi : 0.0;
      }
  }
  double Cspec_i[MAXS];
  if (SPECIES)
    for (int k = 0; k < co.nspecies; k++) Cspec_i[k] = d.pCs[(size_t)i * co.nspecies + k];
  const double ei = RANDOM ? d.pD[i].w : 0.0;
  const int tagi = RANDOM ? d.ptag[i] : 0;
  double G0i = co.G0[ti];
  if (VARIANT == SPHBVF_FSI && SPECIES) G0i = co.G0[ti] * (1.0 - 0.99 * Cspec_i[0]);

  double fx = 0, fy = 0, fz = 0, drho = 0, nd = 0, rA1 = 0, rA2 = 0, phi = 0;
  double ddvx = 0, ddvy = 0, ddvz = 0, nwx = 0, nwy = 0, nwz = 0;
  double ddxx = 0, ddxy = 0, ddxz = 0;
  double spi = 0;           // sum_j s_ij rho_i a_i: the i-side transport term is vest_i times this
  double ddev[9];
  double Qs[MAXS];
  if (SOLIDS == 2)
#pragma unroll
    for (int k = 0; k < 9; k++) ddev[k] = 0.0;
  if (SPECIES)
#pragma unroll
    for (int k = 0; k < MAXS; k++) Qs[k] = 0.0;

  // ------------------------------------------------------------------ one neighbour
  auto body = [&](const int ent, const Rec4 &Aj, const Rec4 &Bj, const Rec4 &Cj, const double rhoIj) {
    const int j = ent & NEIGH_JMASK;
    const int tj = (ent >> NEIGH_JBITS) & 7;
    const bool sj = SOLIDS && ((ent >> 30) & 1);
    const double delx = Ai.x - Aj.x, dely = Ai.y - Aj.y, delz = Ai.z - Aj.z;
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
    const double rhoj = Aj.w, Vj = Bj.w, Vj2 = Vj * Vj, Prrj = Cj.w;
    const double velx = Bi.x - Bj.x, vely = Bi.y - Bj.y, velz = Bi.z - Bj.z;
    const double dvr = delx * velx + dely * vely + delz * velz;
    const double ai = Ci.x * delx + Ci.y * dely + Ci.z * delz;   // (v_i - vt_i) . del
    const double aj = Cj.x * delx + Cj.y * dely + Cj.z * delz;
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
    if (SOLIDS == 1) {
      if (si || sj) {
        const double q = wf * iwdelta, q2 = q * q;
        const double artj = (sj && VARIANT != SPHBVF_MECHANICS && Prrj < 0.0) ? c_art * Prrj : 0.0;
        const double cc = mmw * q2 * q2 * (arti + artj);
        fartx = cc * delx; farty = cc * dely; fartz = cc * delz;
      }
    }

    double devj[9];
    if (SOLIDS == 2) {
      const double irhoj = Vj * tb.imass[tj];
      if (si || sj) {
#pragma unroll
        for (int k = 0; k < 9; k++) devj[k] = sj ? d.pdev[9 * (size_t)j + k] : 0.0;
        const double Pj = Prrj * rhoj * rhoj;
        const double Psj = VARIANT == SPHBVF_MECHANICS ? fabs(Pj) : Pj;
        const double q = wf * iwdelta, q2 = q * q;
        const double pre = mmw * q2 * q2;
        double R[9];
#pragma unroll
        for (int m = 0; m < 3; m++)
#pragma unroll
          for (int n = 0; n < 3; n++) {
            double ts = devj[3 * m + n] - (m == n ? Psj : 0.0);
            double Rj = (sj && ts > 0.0) ? -c_art * ts * irhoj * irhoj : 0.0;
            R[3 * m + n] = Ri[3 * m + n] + Rj;
          }
        fartx = pre * (delx * R[0] + dely * R[3] + delz * R[6]);
        farty = pre * (delx * R[1] + dely * R[4] + delz * R[7]);
        fartz = pre * (delx * R[2] + dely * R[5] + delz * R[8]);
      }
      // ---- Jaumann rate for solid i (:435-451)
      if (si) {
        double G0j = co.G0[tj];
        double geff;
        if (VARIANT == SPHBVF_FSI && SPECIES) {
          G0j = co.G0[tj] * (1.0 - 0.99 * d.pCs[(size_t)j * co.nspecies]);
          geff = (2.0 * G0i * G0j) / (G0i + G0j + 1e-12);
        } else geff = tb.geff[ti][tj];
        const double hw = -0.5 * Vj * wfd;   // 0.5 * Vj * wfd * (v_j - v_i) = hw * vel
        const double vel[3] = {velx, vely, velz}, del[3] = {delx, dely, delz};
        double eps[9], om[9];
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
            ddev[3 * m + n] += 2.0 * geff * (m == n ? e - (1. / 3.) * e : e) + dDotR + rDotD;
          }
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
      fx = fma(pj_, Bj.x, fx); fy = fma(pj_, Bj.y, fy); fz = fma(pj_, Bj.z, fz);
      if (SOLIDS) { fx += fartx; fy += farty; fz += fartz; }
      if (RANDOM) {
        // f_rand = sqrt(-4 kB e m_i m_j wfd / (rho_i rho_j dt)) / (r + 0.01 h) * (Wn . del)   (:403-431), with
        // Wn the symmetric traceless part of a d x d Gaussian matrix: off-diagonals N(0, 1/2), diagonal
        // g_ll - mean(g).  The matrix depends on (seed, step, min tag, max tag) only, so the partner
        // computes the same one with del -> -del: equal and opposite forces.
        const double eij = 0.5 * (ei + d.pD[j].w);
        const int tagj = d.ptag[j];
        const double pref = sqrt(fmax(-rand_pref * eij * (Vi * Vj) * wfd, 0.0)) * fast_rcp(r + 0.01 * h);
        const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
        const unsigned tlo = (unsigned)min(tagi, tagj), thi = (unsigned)max(tagi, tagj);
        const uint4 r0 = philox4x32(make_uint4(tlo, thi, (unsigned)ntimestep, (unsigned)(ntimestep >> 32) << 1), key);
        double g0, g1, g2, g3;
        gauss2(r0.x, r0.y, g0, g1);
        gauss2(r0.z, r0.w, g2, g3);
        double wxx, wyy, wzz, wxy, wxz = 0.0, wyz = 0.0;
        if (co.dim == 2) {
          // trace / dimension with the zz entry zero: Wxx = (gxx - gyy)/2 = -Wyy, Wzz irrelevant (delz = 0)
          wxx = 0.5 * (g0 - g1); wyy = -wxx; wzz = 0.0;
          wxy = 0.5 * (g2 + g3);
        } else {
          const uint4 r1 = philox4x32(make_uint4(tlo, thi, (unsigned)ntimestep, ((unsigned)(ntimestep >> 32) << 1) | 1u), key);
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
        fviscs = mmw * (-(c0i + co.c0[tj]) * mu + 2.0 * mu * mu) * fast_rcp(rhoi + rhoj);
      }
      const double cc = -(fpair + fviscs);
      fx += cc * delx + fartx;
      fy += cc * dely + farty;
      fz += cc * delz + fartz;
      if (SOLIDS == 2) {
        const double irhoj = Vj * tb.imass[tj];
        const double ii = irhoi * irhoi, jj = irhoj * irhoj;
        fx += mmw * (delx * (devi[0] * ii + devj[0] * jj) + dely * (devi[3] * ii + devj[3] * jj) + delz * (devi[6] * ii + devj[6] * jj));
        fy += mmw * (delx * (devi[1] * ii + devj[1] * jj) + dely * (devi[4] * ii + devj[4] * jj) + delz * (devi[7] * ii + devj[7] * jj));
        fz += mmw * (delx * (devi[2] * ii + devj[2] * jj) + dely * (devi[5] * ii + devj[5] * jj) + delz * (devi[8] * ii + devj[8] * jj));
      }
    }

    // ---- density rate (:548-555); (vt_i - vt_j).del = dvr - ai + aj
    {
      // rho_i (dvr - a_i + a_j) - (rho_i a_i + rho_j a_j) = rho_i (dvr + a_j) - 2 q_i - q_j
      double inner = fma(-2.0, qi, fma(rhoi, dvr + aj, -qj));
      if (VARIANT == SPHBVF_FSI)
        inner -= damp * 2.0 * h * c0i * (rhoj - rhoi) * (rsq * fast_rcp(rsq + h2eps));
      drho = fma(wfd * Vj, inner, drho);
    }

    // ---- BVF (:563-576)
    if (SOLIDS && !si && sj) {
      phi = fma(Vj2, wf, phi);
      const double cc = wfd * Vj2;
      nwx = fma(cc, delx, nwx); nwy = fma(cc, dely, nwy); nwz = fma(cc, delz, nwz);
    }

    // ---- species (:678-720)
    if (SPECIES) {
      const double hc = co.cutc[ti][tj];
      if (r < hc) {
        const double tc = hc - r;
        const double wfdc = tb.cwfdc[ti][tj] * tc * tc;
        const double irhoj = Vj * tb.imass[tj];
        const double q0 = tb.mred2[ti][tj] * (irhoi + irhoj) * rsq * wfdc / (rsq + tb.hc2eps[ti][tj]);
        for (int k = 0; k < co.nspecies; k++) {
          const double Cjk = d.pCs[(size_t)j * co.nspecies + k];
          double dq = co.kappa[ti][tj][k] * (Cspec_i[k] - Cjk) * q0;
          if (VARIANT == SPHBVF_TV) dq -= Vj * (Cspec_i[k] * ai + Cjk * aj) * wfdc;
          Qs[k] += dq;
        }
      }
    }
  };

  double vir[6] = {0, 0, 0, 0, 0, 0};
  auto visit = [&](const int ent, const Rec4 &Aj, const Rec4 &Bj, const Rec4 &Cj, const double rhoIj) {
    if (!VIRIAL) { body(ent, Aj, Bj, Cj, rhoIj); return; }
    const int g = (ent & NEIGH_JMASK) - d.nlocal;
    if (g < 0) return;
    const double sx = d.gshift[3 * (size_t)g], sy = d.gshift[3 * (size_t)g + 1], sz = d.gshift[3 * (size_t)g + 2];
    if (sx == 0.0 && sy == 0.0 && sz == 0.0) return;
    const double f0x = fx + spi * Bi.x, f0y = fy + spi * Bi.y, f0z = fz + spi * Bi.z;
    body(ent, Aj, Bj, Cj, rhoIj);
    const double Fx = (fx + spi * Bi.x) - f0x, Fy = (fy + spi * Bi.y) - f0y, Fz = (fz + spi * Bi.z) - f0z;
    vir[0] -= 0.5 * sx * Fx; vir[1] -= 0.5 * sy * Fy; vir[2] -= 0.5 * sz * Fz;
    vir[3] -= 0.5 * sx * Fy; vir[4] -= 0.5 * sx * Fz; vir[5] -= 0.5 * sy * Fz;
  };

  // ------------------------------------------------------------------ pipelined traversal
  // List entries stream from DRAM (4 B per neighbour, read once): each thread copies its own
  // entries RING rows ahead into a shared-memory ring with cp.async (no registers, no barrier:
  // a thread only ever reads the slots it wrote).  The records of neighbour k+1 are requested
  // before neighbour k is evaluated.
#if PAIR_TMA
  const int nn = live ? d.numneigh[i] : 0;
#else
  const int nn = d.numneigh[i];
#endif
  const int *np = d.neigh + i;
  const size_t stride = d.stride;
  int *myring = &ring[0][threadIdx.x];
  unsigned long long pol = 0;
#if PAIR_L2HINT >= 1
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#endif
#if PAIR_TMA
  // ------------------------------------------------------------------ TMA traversal
  // Per lane, the 96-byte record of entry k + D is copied global -> shared memory by the bulk-copy engine while entry k
  // is evaluated: D records in flight per thread without a single register, and the L1/LSU pipe only sees conflict-free
  // LDS.128 reads (24 wavefronts per visit instead of 66 for three gathers).  One mbarrier per (warp, slot): lane 0
  // posts the expected byte count of the warp's active lanes, each lane's copy completes its share, all lanes wait.
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int nnmax = nn;
#pragma unroll
    for (int o = 16; o; o >>= 1) nnmax = max(nnmax, __shfl_xor_sync(0xffffffffu, nnmax, o));
    unsigned char *myrec = pair_dyn + (size_t)threadIdx.x * TMA_RSTRIDE;
    unsigned long long *mbar = mbar_all + warp * PAIR_TMA;
    auto fetch1 = [&](int k) {
      if (k < nn) cp_async4(myring + (k % RING) * PAIR_T, np + (size_t)k * stride, pol);
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto issue = [&](int k) {   // warp-uniform call, k < nnmax, entry k has landed in the ring
      const int slot = k % PAIR_TMA;
      const bool act = k < nn;
      const unsigned m = __ballot_sync(0xffffffffu, act);
      if (lane == 0) mbar_expect_tx(mbar + slot, 96u * (unsigned)__popc(m));
      __syncwarp();
      if (act) {
        const int e = myring[(k % RING) * PAIR_T];
        bulk_g2s(myrec + (size_t)slot * PAIR_T * TMA_RSTRIDE, d.prec + (e & NEIGH_JMASK), 96u, mbar + slot);
      }
    };
#pragma unroll
    for (int k = 0; k < RING; k++) fetch1(k);
    asm volatile("cp.async.wait_group %0;" ::"n"(RING - PAIR_TMA) : "memory");   // entries < D landed
    __syncwarp();
    for (int k = 0; k < PAIR_TMA && k < nnmax; k++) issue(k);
    for (int k = 0; k < nnmax; k++) {
      const bool act = k < nn;
      const int e = act ? myring[(k % RING) * PAIR_T] : 0;
      fetch1(k + RING);
      asm volatile("cp.async.wait_group %0;" ::"n"(RING - PAIR_TMA) : "memory");   // entries <= k + D landed
      mbar_wait(mbar + (k % PAIR_TMA), (unsigned)((k / PAIR_TMA) & 1));
      if (act) {
        const double2 *r = reinterpret_cast<const double2 *>(myrec + (size_t)(k % PAIR_TMA) * PAIR_T * TMA_RSTRIDE);
        const double2 a0 = r[0], a1 = r[1], b0 = r[2], b1 = r[3], c0 = r[4], c1 = r[5];
        const Rec4 Aj = make_rec4(a0.x, a0.y, a1.x, a1.y), Bj = make_rec4(b0.x, b0.y, b1.x, b1.y),
                   Cj = make_rec4(c0.x, c0.y, c1.x, c1.y);
        visit(e, Aj, Bj, Cj, FILTER ? d.pD[e & NEIGH_JMASK].x : 0.0);
      }
      __syncwarp();
      if (k + PAIR_TMA < nnmax) issue(k + PAIR_TMA);
    }
  }
#else
  auto fetch2 = [&](int k) {   // entries k, k+1 -> ring slots k % RING, (k+1) % RING; one group
    if (k < nn) cp_async4(myring + (k % RING) * PAIR_T, np + (size_t)k * stride, pol);
    if (k + 1 < nn) cp_async4(myring + ((k + 1) % RING) * PAIR_T, np + (size_t)(k + 1) * stride, pol);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // groups allowed in flight after a wait: entries <= kk + 17 - 2 * PEND have landed
  constexpr int PEND = RING / 2 - 2 - PAIR_PFL2;
  static_assert(PEND >= 2, "prefetch distance too large for the ring");
#pragma unroll
  for (int k = 0; k < RING; k += 2) fetch2(k);
  asm volatile("cp.async.wait_group %0;" ::"n"(PEND) : "memory");   // entries 0..3 (+ 2 PFL2) landed
  int e0 = nn > 0 ? myring[0] : 0;
  int e1 = nn > 1 ? myring[PAIR_T] : 0;
#if PAIR_PFL2 > 0
#pragma unroll
  for (int q = 2; q < 4 + 2 * PAIR_PFL2; q++)
    if (q < nn) prefetch_rec_l2(d.prec + (myring[(q % RING) * PAIR_T] & NEIGH_JMASK));
#endif
  // Register pipeline: while the record of one neighbour is evaluated, the record of the next one is in flight.
  // ptxas puts all six record loads of the loop on ONE scoreboard, and a scoreboard is a counter: the first use of a
  // record waits until EVERY load issued before it has returned.  Issued in source order (next record's loads, then
  // the visit), the wait for record k therefore also waited for record k+1, requested five instructions earlier -- a
  // full memory latency exposed on every second visit (ncu source page: one DADD held 33 % of all stall samples).
  // The loads of record k+1 are therefore made DATA dependent on the first use of record k (`gate`: a NaN test the
  // compiler cannot fold), so they are issued right after record k has arrived and have a whole visit to complete.
#if PAIR_PIPE == 0
  // No register pipeline at all: one record buffer, the latency of every record is covered by other warps only
  // (<= 128 registers -> four CTAs, 16 warps per SM).
  for (int kk = 0; kk < nn; kk += 2) {
    fetch2(kk + RING);
    asm volatile("cp.async.wait_group %0;" ::"n"(PEND) : "memory");
    const int e2 = kk + 2 < nn ? myring[((kk + 2) % RING) * PAIR_T] : 0;
    const int e3 = kk + 3 < nn ? myring[((kk + 3) % RING) * PAIR_T] : 0;
    {
      const Prec *p = d.prec + (e0 & NEIGH_JMASK);
      const Rec4 A = ld_rec(&p->A), B = ld_rec(&p->B), C = ld_rec(&p->C);
      visit(e0, A, B, C, FILTER ? d.pD[e0 & NEIGH_JMASK].x : 0.0);
    }
    if (kk + 1 < nn) {
      const Prec *p = d.prec + (e1 & NEIGH_JMASK);
      const Rec4 A = ld_rec(&p->A), B = ld_rec(&p->B), C = ld_rec(&p->C);
      visit(e1, A, B, C, FILTER ? d.pD[e1 & NEIGH_JMASK].x : 0.0);
    }
    e0 = e2;
    e1 = e3;
  }

#elif PAIR_PIPE == 2
  // Two records in flight per warp: the records of the NEXT TWO neighbours are requested (one batch of six loads, gated
  // on the first use of the current pair so that the single scoreboard wait of the loop covers exactly one batch) while
  // the current two are evaluated.  Four record buffers: ~210 registers, two CTAs per SM -- 8 warps x 2 records = 16
  // records in flight per SM instead of 12 x 1.
  Rec4 A0, B0, C0, A1, B1, C1, A2, B2, C2, A3, B3, C3;
  double D0 = 0.0, D1 = 0.0, D2 = 0.0, D3 = 0.0;
  auto gate = [&](const Rec4 &A) { const double g = Ai.x - A.x; return g != g ? 1 : 0; };
  {
    const Prec *p = d.prec + (e0 & NEIGH_JMASK);
    A0 = ld_rec(&p->A); B0 = ld_rec(&p->B); C0 = ld_rec(&p->C);
    if (FILTER) D0 = d.pD[e0 & NEIGH_JMASK].x;
    const Prec *q = d.prec + (e1 & NEIGH_JMASK);
    A1 = ld_rec(&q->A); B1 = ld_rec(&q->B); C1 = ld_rec(&q->C);
    if (FILTER) D1 = d.pD[e1 & NEIGH_JMASK].x;
  }
  for (int kk = 0; kk < nn; kk += 4) {
    // ---- half 1: evaluate e0, e1 (records 0, 1); request records 2, 3 for e2, e3
    fetch2(kk + RING);
    asm volatile("cp.async.wait_group %0;" ::"n"(PEND) : "memory");   // entries <= kk+5 landed
    const int e2 = kk + 2 < nn ? myring[((kk + 2) % RING) * PAIR_T] : 0;
    const int e3 = kk + 3 < nn ? myring[((kk + 3) % RING) * PAIR_T] : 0;
    {
      const int g = gate(A0);
      const int j2 = (e2 & NEIGH_JMASK) + g, j3 = (e3 & NEIGH_JMASK) + g;
      const Prec *p = d.prec + j2, *q = d.prec + j3;
      A2 = ld_rec(&p->A); B2 = ld_rec(&p->B); C2 = ld_rec(&p->C);
      A3 = ld_rec(&q->A); B3 = ld_rec(&q->B); C3 = ld_rec(&q->C);
      if (FILTER) { D2 = d.pD[j2].x; D3 = d.pD[j3].x; }
    }
    visit(e0, A0, B0, C0, D0);
    if (kk + 1 < nn) visit(e1, A1, B1, C1, D1);
    if (kk + 2 >= nn) break;
    // ---- half 2: evaluate e2, e3 (records 2, 3); request records 0, 1 for e4, e5
    fetch2(kk + 2 + RING);
    asm volatile("cp.async.wait_group %0;" ::"n"(PEND) : "memory");   // entries <= kk+7 landed
    e0 = kk + 4 < nn ? myring[((kk + 4) % RING) * PAIR_T] : 0;
    e1 = kk + 5 < nn ? myring[((kk + 5) % RING) * PAIR_T] : 0;
    {
      const int g = gate(A2);
      const int j0 = (e0 & NEIGH_JMASK) + g, j1 = (e1 & NEIGH_JMASK) + g;
      const Prec *p = d.prec + j0, *q = d.prec + j1;
      A0 = ld_rec(&p->A); B0 = ld_rec(&p->B); C0 = ld_rec(&p->C);
      A1 = ld_rec(&q->A); B1 = ld_rec(&q->B); C1 = ld_rec(&q->C);
      if (FILTER) { D0 = d.pD[j0].x; D1 = d.pD[j1].x; }
    }
    visit(e2, A2, B2, C2, D2);
    if (kk + 3 < nn) visit(e3, A3, B3, C3, D3);
  }

#else
  Rec4 A0, B0, C0, A1, B1, C1;
  double D0 = 0.0, D1 = 0.0;   // rhoI_j of the Shepard numerator, part of the pipeline on filter steps
  auto gate = [&](const Rec4 &A) { const double g = Ai.x - A.x; return g != g ? 1 : 0; };
  {
    const Prec *p = d.prec + (e0 & NEIGH_JMASK);
    A0 = ld_rec(&p->A); B0 = ld_rec(&p->B); C0 = ld_rec(&p->C);
    if (FILTER) D0 = d.pD[e0 & NEIGH_JMASK].x;
  }
  for (int kk = 0; kk < nn; kk += 2) {
    {
      const int j1 = (e1 & NEIGH_JMASK) + gate(A0);
      const Prec *p = d.prec + j1;
      A1 = ld_rec(&p->A); B1 = ld_rec(&p->B); C1 = ld_rec(&p->C);
      if (FILTER) D1 = d.pD[j1].x;
    }
    fetch2(kk + RING);   // slots of entries kk, kk+1: already in e0, e1
    asm volatile("cp.async.wait_group %0;" ::"n"(PEND) : "memory");   // entries <= kk+5 (+ 2 PFL2) landed
    const int e2 = kk + 2 < nn ? myring[((kk + 2) % RING) * PAIR_T] : 0;
    const int e3 = kk + 3 < nn ? myring[((kk + 3) % RING) * PAIR_T] : 0;
#if PAIR_PFL2 > 0
    if (kk + 4 + 2 * PAIR_PFL2 < nn) prefetch_rec_l2(d.prec + (myring[((kk + 4 + 2 * PAIR_PFL2) % RING) * PAIR_T] & NEIGH_JMASK));
    if (kk + 5 + 2 * PAIR_PFL2 < nn) prefetch_rec_l2(d.prec + (myring[((kk + 5 + 2 * PAIR_PFL2) % RING) * PAIR_T] & NEIGH_JMASK));
#endif
    visit(e0, A0, B0, C0, D0);
    {
      const int j2 = (e2 & NEIGH_JMASK) + gate(A1);
      const Prec *p = d.prec + j2;
      A0 = ld_rec(&p->A); B0 = ld_rec(&p->B); C0 = ld_rec(&p->C);
      if (FILTER) D0 = d.pD[j2].x;
    }
    if (kk + 1 < nn) visit(e1, A1, B1, C1, D1);
    e0 = e2;
    e1 = e3;
  }

#endif
#endif   // PAIR_TMA

  if (VIRIAL) {
    // only atoms next to a periodic face have anything to add: plain atomics
#pragma unroll
    for (int q = 0; q < 6; q++)
      if (vir[q] != 0.0) atomicAdd(virial_out + q, vir[q]);
    return;
  }
#if PAIR_TMA
  if (!live) return;
#endif
  const double ddvc = 10.0 * 7.0 * co.B[ti];
  const size_t i3 = 3 * (size_t)i;
  st_out(d.f + i3, fma(spi, Bi.x, fx)); st_out(d.f + i3 + 1, fma(spi, Bi.y, fy)); st_out(d.f + i3 + 2, fma(spi, Bi.z, fz));
  st_out(d.drho + i, drho);
  st_out(d.nd + i, nd);
  st_out(d.rhoAux1 + i, rA1);
  st_out(d.rhoAux2 + i, rA2);
  st_out(d.phi + i, phi);
  st_out(d.nw + i3, nwx); st_out(d.nw + i3 + 1, nwy); st_out(d.nw + i3 + 2, nwz);
  st_out(d.ddv + i3, ddvc * ddvx); st_out(d.ddv + i3 + 1, ddvc * ddvy); st_out(d.ddv + i3 + 2, ddvc * ddvz);
  if (VARIANT != SPHBVF_TV) {
    d.ddx[i3] = ddxx; d.ddx[i3 + 1] = ddxy; d.ddx[i3 + 2] = ddxz;
    d.Pnew[i] = Pi;   // pair_ssa_tsdpd_bvf_mechanics.cpp:188
  }
  if (SOLIDS == 2)
#pragma unroll
    for (int k = 0; k < 9; k++) d.ddev[9 * (size_t)i + k] = si ? ddev[k] : 0.0;
  if (SPECIES)
    for (int k = 0; k < co.nspecies; k++) d.Q[(size_t)i * co.nspecies + k] = Qs[k];
}

template <int VARIANT, bool SPECIES, int SOLIDS, bool UNIFORM>
static void launch_filter(const DevState &d, const Coeffs &co, const PairTables &tb, const PairFlags &pf,
                          cudaStream_t st) {
  const int threads = PAIR_T;
  const int blocks = (d.nlocal + threads - 1) / threads;
#if PAIR_TMA
  constexpr int dyn = PAIR_TMA * PAIR_T * TMA_RSTRIDE + (PAIR_T / 32) * PAIR_TMA * 8;
  {   // per device: cheap enough to repeat on every launch of this tuning path
    cudaFuncSetAttribute(pair_kernel<VARIANT, SPECIES, SOLIDS, UNIFORM, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    cudaFuncSetAttribute(pair_kernel<VARIANT, SPECIES, SOLIDS, UNIFORM, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    cudaFuncSetAttribute(pair_kernel<VARIANT, SPECIES, SOLIDS, UNIFORM, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  }
#define PK(F, R) pair_kernel<VARIANT, SPECIES, SOLIDS, UNIFORM, F, R><<<blocks, threads, dyn, st>>>(d, co, tb, pf.damp, pf.rand_pref, pf.seed, pf.ntimestep)
#elif defined(PAIR_DIAG_SMEM)   // tools/: occupancy probe -- extra dynamic shared memory limits the resident CTAs per SM
  cudaFuncSetAttribute(pair_kernel<VARIANT, SPECIES, SOLIDS, UNIFORM, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_DIAG_SMEM);
#define PK(F, R) pair_kernel<VARIANT, SPECIES, SOLIDS, UNIFORM, F, R><<<blocks, threads, ((F) || (R)) ? 0 : PAIR_DIAG_SMEM, st>>>(d, co, tb, pf.damp, pf.rand_pref, pf.seed, pf.ntimestep)
#else
#define PK(F, R) pair_kernel<VARIANT, SPECIES, SOLIDS, UNIFORM, F, R><<<blocks, threads, 0, st>>>(d, co, tb, pf.damp, pf.rand_pref, pf.seed, pf.ntimestep)
#endif
  // the stochastic variant always carries the Shepard numerator (one instantiation less per case)
  if (pf.random) PK(true, true);
  else if (pf.filter_step) PK(true, false);
  else PK(false, false);
#undef PK
}

template <int VARIANT, bool SPECIES>
static void launch_solids(const DevState &d, const Coeffs &co, const PairTables &tb, const PairFlags &pf,
                          bool uniform, cudaStream_t st) {
  if (d.nlocal == 0) return;
  const int solids = !pf.any_solid ? 0 : (pf.with_dev ? 2 : 1);
  // the constant-coefficient fast path is instantiated for the rigid-wall / no-solid cases only
  if (solids == 0) {
    if (uniform) launch_filter<VARIANT, SPECIES, 0, true>(d, co, tb, pf, st);
    else launch_filter<VARIANT, SPECIES, 0, false>(d, co, tb, pf, st);
  } else if (solids == 1) {
    if (uniform) launch_filter<VARIANT, SPECIES, 1, true>(d, co, tb, pf, st);
    else launch_filter<VARIANT, SPECIES, 1, false>(d, co, tb, pf, st);
  } else launch_filter<VARIANT, SPECIES, 2, false>(d, co, tb, pf, st);
}

template <int VARIANT>
static void launch_species(const DevState &d, const Coeffs &co, const PairTables &tb, const PairFlags &pf,
                           bool uniform, cudaStream_t st) {
  if (co.nspecies > 0) launch_solids<VARIANT, true>(d, co, tb, pf, uniform, st);
  else launch_solids<VARIANT, false>(d, co, tb, pf, uniform, st);
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
                                 double *out, cudaStream_t st) {
  const int blocks = (d.nlocal + PAIR_T - 1) / PAIR_T;
  const int solids = !pf.any_solid ? 0 : (pf.with_dev ? 2 : 1);
#if PAIR_TMA
  constexpr int vdyn = PAIR_TMA * PAIR_T * TMA_RSTRIDE + (PAIR_T / 32) * PAIR_TMA * 8;
  {
    cudaFuncSetAttribute(pair_kernel<VARIANT, SPECIES, 0, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, vdyn);
    cudaFuncSetAttribute(pair_kernel<VARIANT, SPECIES, 1, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, vdyn);
    cudaFuncSetAttribute(pair_kernel<VARIANT, SPECIES, 2, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, vdyn);
  }
#else
  constexpr int vdyn = 0;
#endif
#define VK(S) pair_kernel<VARIANT, SPECIES, S, false, false, false, true><<<blocks, PAIR_T, vdyn, st>>>(d, co, tb, pf.damp, 0.0, 0ULL, pf.ntimestep, out)
  if (solids == 0) VK(0);
  else if (solids == 1) VK(1);
  else VK(2);
#undef VK
}

// Pair::virial_fdotr_compute for the gather formulation; out[6] must be zeroed by the caller
void launch_virial(const DevState &d, const Coeffs &co, const PairFlags &pf, double *out, cudaStream_t st) {
  if (!d.nlocal) return;
  virial_fdotr_kernel<<<(d.nlocal + 255) / 256, 256, 0, st>>>(d, out);
  if (!d.nghost) return;
  PairTables tb;
  make_tables(co, tb);
  const bool sp = co.nspecies > 0;
  switch (co.variant) {
    case SPHBVF_TV: sp ? launch_virial_solids<SPHBVF_TV, true>(d, co, tb, pf, out, st) : launch_virial_solids<SPHBVF_TV, false>(d, co, tb, pf, out, st); break;
    case SPHBVF_MECHANICS: sp ? launch_virial_solids<SPHBVF_MECHANICS, true>(d, co, tb, pf, out, st) : launch_virial_solids<SPHBVF_MECHANICS, false>(d, co, tb, pf, out, st); break;
    default: sp ? launch_virial_solids<SPHBVF_FSI, true>(d, co, tb, pf, out, st) : launch_virial_solids<SPHBVF_FSI, false>(d, co, tb, pf, out, st); break;
  }
}

void launch_pair(const DevState &d, const Coeffs &co, const PairFlags &pf, cudaStream_t st) {
  PairTables tb;
  const bool uniform = make_tables(co, tb);
  switch (co.variant) {
    case SPHBVF_TV: launch_species<SPHBVF_TV>(d, co, tb, pf, uniform, st); break;
    case SPHBVF_MECHANICS: launch_species<SPHBVF_MECHANICS>(d, co, tb, pf, uniform, st); break;
    default: launch_species<SPHBVF_FSI>(d, co, tb, pf, uniform, st); break;
  }
}

}  // namespace sphbvf

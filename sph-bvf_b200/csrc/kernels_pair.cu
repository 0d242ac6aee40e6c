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
// Mapping: one thread per owned atom, 128-thread CTAs.  Neighbour entries are read transposed
// (coalesced), neighbour state through three 32-byte records; all per-particle divisions were
// moved to the pack kernel (V = m/rho, P/rho^2), the loop body is divide-free except for species.
#include "sphbvf_internal.cuh"

namespace sphbvf {

struct PairTables {
  double cwfd[MAXT][MAXT];   // (1/r) dW/dr = cwfd * (h-r)^2
  double cwf[MAXT][MAXT];    // W = cwf * (h-r)^3 * (h+3r)
  double iwdelta[MAXT][MAXT];  // 1 / W(delta)
  double cwfdc[MAXT][MAXT];  // same as cwfd with h = cutc
  double h2eps[MAXT][MAXT];  // 0.01 h^2
  double hc2eps[MAXT][MAXT]; // 0.01 cutc^2
  double mimj[MAXT][MAXT];
  double mred2[MAXT][MAXT];  // 2 mi mj / (mi + mj)
  double geff[MAXT][MAXT];   // 2 Gi Gj / (Gi + Gj + 1e-12)
  double imass[MAXT];
};

static void make_tables(const Coeffs &co, PairTables &t) {
  const double delta_fac = co.variant == SPHBVF_TV ? (1.0 / 2.6) : (1.0 / 3.0);
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
      double h = co.cut[i][j], hc = co.cutc[i][j], dummy;
      coef(h, t.cwfd[i][j], t.cwf[i][j]);
      coef(hc > 0 ? hc : h, t.cwfdc[i][j], dummy);
      double delta = delta_fac * h, td = h - delta;
      double wdelta = t.cwf[i][j] * td * td * td * (h + 3. * delta);
      t.iwdelta[i][j] = 1.0 / wdelta;
      t.h2eps[i][j] = 0.01 * h * h;
      t.hc2eps[i][j] = 0.01 * hc * hc;
      t.mimj[i][j] = co.mass[i] * co.mass[j];
      t.mred2[i][j] = 2.0 * ((co.mass[i] * co.mass[j]) / (co.mass[i] + co.mass[j]));
      t.geff[i][j] = (2.0 * co.G0[i] * co.G0[j]) / (co.G0[i] + co.G0[j] + 1e-12);
    }
  }
}

// SOLIDS: 0 = no atom has solid_tag, 1 = solids whose deviatoric stress is identically zero
// (rigid walls: G0 == 0, dev == 0), 2 = elastic solids (deviatoric tensors gathered)
template <int VARIANT, bool SPECIES, int SOLIDS>
__global__ void __launch_bounds__(128)
pair_kernel(const DevState d, const __grid_constant__ Coeffs co, const __grid_constant__ PairTables tb,
            const int filter, const double damp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;

  const int fl = d.pflags[i];
  const int ti = fl & 7;
  const bool si = SOLIDS && ((fl >> 4) & 1);
  const double4 Ai = d.pA[i], Bi = d.pB[i], Ci = d.pC[i];
  const double rhoi = Ai.w, Vi = Bi.w, Vi2 = Vi * Vi, Prri = Ci.w;
  const double ddvc = 10.0 * 7.0 * co.B[ti];
  const double c0i = co.c0[ti];
  double Pi = Prri * rhoi * rhoi;
  double irhoi = Vi * tb.imass[ti];

  double devi[9];
  double arti = 0.0;        // scalar artificial stress when dev == 0
  double Ri[9];             // artificial stress tensor when dev != 0
  const double c_art = VARIANT == SPHBVF_FSI ? 0.1 : 0.35;
  if (SOLIDS == 1) arti = d.pD[i].y;
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
  double G0i = co.G0[ti];
  if (VARIANT == SPHBVF_FSI && SPECIES) G0i = co.G0[ti] * (1.0 - 0.99 * Cspec_i[0]);

  double fx = 0, fy = 0, fz = 0, drho = 0, nd = 0, rA1 = 0, rA2 = 0, phi = 0;
  double ddvx = 0, ddvy = 0, ddvz = 0, nwx = 0, nwy = 0, nwz = 0;
  double ddxx = 0, ddxy = 0, ddxz = 0;
  double ddev[9];
  double Qs[MAXS];
  if (SOLIDS == 2)
#pragma unroll
    for (int k = 0; k < 9; k++) ddev[k] = 0.0;
  if (SPECIES)
#pragma unroll
    for (int k = 0; k < MAXS; k++) Qs[k] = 0.0;

  const int nn = d.numneigh[i];
  const int *np = d.neigh + i;
  for (int kk = 0; kk < nn; kk++) {
    const int ent = __ldg(np + (size_t)kk * d.stride);
    const int j = ent & NEIGH_JMASK;
    const int tj = (ent >> NEIGH_JBITS) & 7;
    const bool sj = SOLIDS && ((ent >> 30) & 1);
    const double4 Aj = d.pA[j];
    const double delx = Ai.x - Aj.x, dely = Ai.y - Aj.y, delz = Ai.z - Aj.z;
    const double rsq = delx * delx + dely * dely + delz * delz;
    if (!(rsq < co.cutsq[ti][tj])) continue;

    const double4 Bj = d.pB[j], Cj = d.pC[j];
    const double h = co.cut[ti][tj];
    const double r = sqrt(rsq);
    const double t = h - r, t2 = t * t;
    const double wfd = tb.cwfd[ti][tj] * t2;
    const double wf = tb.cwf[ti][tj] * t2 * t * (h + 3. * r);
    const double rhoj = Aj.w, Vj = Bj.w, Vj2 = Vj * Vj, Prrj = Cj.w;
    const double velx = Bi.x - Bj.x, vely = Bi.y - Bj.y, velz = Bi.z - Bj.z;
    const double dvr = delx * velx + dely * vely + delz * velz;
    const double ai = Ci.x * delx + Ci.y * dely + Ci.z * delz;   // (v_i - vt_i) . del
    const double aj = Cj.x * delx + Cj.y * dely + Cj.z * delz;
    const double S2 = Vi2 + Vj2;
    const double S2w = S2 * wfd;

    // ---- sweep A (pair_...transport_velocity.cpp:243-254)
    nd += Vj2 * wf;
    rA2 += wf;
    if (filter) rA1 += d.pD[j].x * wf;
    {
      const double cc = ddvc * S2w;
      ddvx += cc * delx; ddvy += cc * dely; ddvz += cc * delz;
    }
    if (VARIANT != SPHBVF_TV) {   // ..._mechanics.cpp:250-252
      const double cc = -Vj2 * wf;
      ddxx += cc * velx; ddxy += cc * vely; ddxz += cc * velz;
    }

    // ---- pressure force (:396-399 / mechanics :408)
    const double mm = tb.mimj[ti][tj];
    const double mmw = mm * wfd;
    const double pij = Prrj + Prri;
    double fpair;
    if (VARIANT == SPHBVF_TV) fpair = (pij >= 0. || (si && sj)) ? mmw * pij : mmw * (Prrj - Prri);
    else fpair = mmw * pij;

    // ---- artificial stress (:454-494)
    double fartx = 0, farty = 0, fartz = 0;
    if (SOLIDS) {
      if (si || sj) {
        const double q = wf * tb.iwdelta[ti][tj], q2 = q * q;
        const double pre = mmw * q2 * q2;
        if (SOLIDS == 1) {
          const double artj = sj ? d.pD[j].y : 0.0;
          const double cc = pre * (arti + artj);
          fartx = cc * delx; farty = cc * dely; fartz = cc * delz;
        }
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
        const double q = wf * tb.iwdelta[ti][tj], q2 = q * q;
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
      const double fvisc = S2w * co.eta[ti][tj];
      const double s = -0.5 * S2w;
      const double pi_ = rhoi * ai, pj_ = rhoj * aj;
      fx += -delx * fpair + fvisc * velx + s * (pi_ * Bi.x + pj_ * Bj.x) + fartx;
      fy += -dely * fpair + fvisc * vely + s * (pi_ * Bi.y + pj_ * Bj.y) + farty;
      fz += -delz * fpair + fvisc * velz + s * (pi_ * Bi.z + pj_ * Bj.z) + fartz;
    } else {
      double fviscs = 0.;
      if (dvr < 0.) {
        const double mu = h * dvr / (rsq + tb.h2eps[ti][tj]);
        fviscs = mmw * (-(c0i + co.c0[tj]) * mu + 2.0 * mu * mu) / (rhoi + rhoj);
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
      double inner = rhoi * (dvr - ai + aj) - (rhoi * ai + rhoj * aj);
      if (VARIANT == SPHBVF_FSI)
        inner -= damp * 2.0 * h * c0i * (rhoj - rhoi) * (rsq / (rsq + tb.h2eps[ti][tj]));
      drho += wfd * Vj * inner;
    }

    // ---- BVF (:563-576)
    if (SOLIDS && !si && sj) {
      phi += Vj2 * wf;
      const double cc = wfd * Vj2;
      nwx += cc * delx; nwy += cc * dely; nwz += cc * delz;
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
  }

  const size_t i3 = 3 * (size_t)i;
  d.f[i3] = fx; d.f[i3 + 1] = fy; d.f[i3 + 2] = fz;
  d.drho[i] = drho;
  d.nd[i] = nd;
  d.rhoAux1[i] = rA1;
  d.rhoAux2[i] = rA2;
  d.phi[i] = phi;
  d.nw[i3] = nwx; d.nw[i3 + 1] = nwy; d.nw[i3 + 2] = nwz;
  d.ddv[i3] = ddvx; d.ddv[i3 + 1] = ddvy; d.ddv[i3 + 2] = ddvz;
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

template <int VARIANT, bool SPECIES>
static void launch_solids(const DevState &d, const Coeffs &co, const PairTables &tb, const PairFlags &pf,
                          cudaStream_t st) {
  const int threads = 128;
  const int blocks = (d.nlocal + threads - 1) / threads;
  if (blocks == 0) return;
  const int solids = !pf.any_solid ? 0 : (pf.with_dev ? 2 : 1);
  if (solids == 0) pair_kernel<VARIANT, SPECIES, 0><<<blocks, threads, 0, st>>>(d, co, tb, pf.filter_step, pf.damp);
  else if (solids == 1) pair_kernel<VARIANT, SPECIES, 1><<<blocks, threads, 0, st>>>(d, co, tb, pf.filter_step, pf.damp);
  else pair_kernel<VARIANT, SPECIES, 2><<<blocks, threads, 0, st>>>(d, co, tb, pf.filter_step, pf.damp);
}

template <int VARIANT>
static void launch_species(const DevState &d, const Coeffs &co, const PairTables &tb, const PairFlags &pf,
                           cudaStream_t st) {
  if (co.nspecies > 0) launch_solids<VARIANT, true>(d, co, tb, pf, st);
  else launch_solids<VARIANT, false>(d, co, tb, pf, st);
}

void launch_pair(const DevState &d, const Coeffs &co, const PairFlags &pf, cudaStream_t st) {
  PairTables tb;
  make_tables(co, tb);
  switch (co.variant) {
    case SPHBVF_TV: launch_species<SPHBVF_TV>(d, co, tb, pf, st); break;
    case SPHBVF_MECHANICS: launch_species<SPHBVF_MECHANICS>(d, co, tb, pf, st); break;
    default: launch_species<SPHBVF_FSI>(d, co, tb, pf, st); break;
  }
}

}  // namespace sphbvf

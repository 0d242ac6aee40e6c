// kernels_integrate.cu -- per-atom streaming kernels: integrator fixes, BC/source fixes, and the
// pack kernel that turns primary state into the pair kernel's 32-byte input records.
//
// Replaces FixSsaTsdpdBvf{TransportVelocity,Mechanics,Fsi}::{setup_pre_force,initial_integrate,
// final_integrate} (fix_ssa_tsdpd_bvf_transport_velocity.cpp:76-461, ..._mechanics.cpp:77-500,
// ..._fsi.cpp:77-470), FixSsaTsdpdBuoyancy::post_force (fix_ssa_tsdpd_buoyancy.cpp:113-140),
// FixSsaTsdpdForcing::post_integrate (fix_ssa_tsdpd_forcing.cpp:133-176),
// FixSsaTsdpdBuffer::{post_integrate,end_of_step} (fix_ssa_tsdpd_buffer.cpp:124-240) and
// FixSetForce::post_force with constants (fix_setforce.cpp:222-290).
// All are HBM-bound: one thread per owned atom, every array read and written once.
#include "sphbvf_internal.cuh"

namespace sphbvf {

static inline int nblocks(int n, int t) { return (n + t - 1) / t; }

__global__ void setup_pre_force_kernel(const DevState d, int groupbit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal || !(d.mask[i] & groupbit)) return;
  for (int k = 0; k < 3; k++) d.vest[3 * (size_t)i + k] = d.v[3 * (size_t)i + k];
  d.rhoI[i] = d.rho[i];
}

void launch_setup_pre_force(const DevState &d, int groupbit, cudaStream_t st) {
  if (!d.nlocal) return;
  setup_pre_force_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, groupbit);
  SPHBVF_LAUNCHED(1);
}

__device__ __forceinline__ void damp_factors(int variant, long ntimestep, double &damp, double &dampSolid) {
  const double tnow = (double)ntimestep;
  damp = tnow <= 1.0 ? tnow / 1.0 : 1.0;
  if (variant == SPHBVF_MECHANICS) dampSolid = tnow < 1e6 ? 0.0 : 1.0;   // ..._mechanics.cpp:151-153
  else dampSolid = tnow <= 1.0 ? 0.0 : 1.0;                              // ..._fsi.cpp:150-152
}

__device__ __forceinline__ double ldv(const double *p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ldi(const int *p) {
  int v;
  asm volatile("ld.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// pack: primary state -> pair input records.  Holds every per-particle division of the pair pass:
// V = m/rho, P/rho^2 with P = 7 B (rho/rho0 - 1) (pair_...transport_velocity.cpp:298-299: evaluated exactly like the
// reference, rho/rho0 - 1 first -- in a nearly quiescent fluid P is a difference of almost equal numbers and the flow
// is driven by it, so any reformulation through V = m/rho costs ten digits: measured, 2.5e-10 on the heated cavity),
// and the scalar artificial stress of a stress-free solid (:454-461 with dev = 0).  Shared by pack_kernel and the
// fused integrator so both produce the same bits.
__device__ __forceinline__ void pack_atom(const DevState &d, const Coeffs &co, const int i, const int t, const int solid,
                                          const int fixed, const double *x, const double *v, const double *vest,
                                          const double rho, const double rhoI, const double e, const int with_dev) {
  const double irho = 1.0 / rho;
  const double P = 7.0 * co.B[t] * (rho / co.rho0[t] - 1.0);
  const double Prr = P * irho * irho;
  Prec r;
  r.A = make_rec4(x[0], x[1], x[2], rho);
  r.B = make_rec4(vest[0], vest[1], vest[2], co.mass[t] * irho);
  r.C = make_rec4(vest[0] - v[0], vest[1] - v[1], vest[2] - v[2], Prr);
  d.prec[i] = r;
  double art = 0.0;
  if (solid) {
    const double c_art = co.variant == SPHBVF_FSI ? 0.1 : 0.35;
    const double Ps = co.variant == SPHBVF_MECHANICS ? fabs(P) : P;
    const double ts = -Ps;
    art = ts > 0.0 ? -c_art * ts * irho * irho : 0.0;
  }
  const double C0 = co.nspecies ? d.C[(size_t)i * co.nspecies] : 0.0;
  d.pD[i] = make_rec4(rhoI, art, C0, e);
  d.pflags[i] = t | (solid << 4) | (fixed << 5);
  for (int k = 0; k < co.nspecies; k++) d.pCs[(size_t)i * co.nspecies + k] = d.C[(size_t)i * co.nspecies + k];
  if (with_dev)
    for (int k = 0; k < 9; k++) d.pdev[9 * (size_t)i + k] = d.dev[9 * (size_t)i + k];
}

// The integrators are pure streaming (HBM bound): ONE kernel template holds both halves so that the
// three ways of running them execute the same arithmetic on the same values, bit for bit:
//   MODE 0  initial_integrate(step)                                  (fix_...:99-240)
//   MODE 1  final_integrate(step)                                    (fix_...:244-461)
//   MODE 2  final_integrate(step - 1) followed by initial_integrate(step) and, if asked, the pack of the
//           pair-input records: what Verlet::run executes back to back when nothing is scheduled
//           between two steps.  v, x, rho, rhoI stay in registers between the halves, f / drho / masks
//           are read once and the intermediate state is never written: 444 B per atom instead of 664.
// Every array an atom may need is loaded UNCONDITIONALLY at the top, before any branch on solid_tag /
// fixed_tag: with the loads inside the branches the kernels ran at 23-30 % of HBM peak, all warps
// waiting on one dependent round trip after another (ncu long_scoreboard 90 %, profiles/); the
// deviatoric tensors (solids only) stay behind their branch.
struct IntegArgs {
  double dt_final, dt_init;
  long step_final, step_init;
  int groupbit, do_pack, with_dev;
  // the atoms of this launch: positions [a0, a1) of `order` (nullptr: the atoms a0 .. a1 themselves).  The early halo
  // integrates the atoms along the brick faces first, so that their records travel while the interior is integrated.
  const int *order;
  int a0, a1;
};

template <int VARIANT, int MODE>
__global__ void __launch_bounds__(256)
integrate_kernel(const DevState d, const __grid_constant__ Coeffs co, const IntegArgs a) {
  const int p = a.a0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.a1) return;
  const int i = a.order ? a.order[p] : p;
  const size_t i3 = 3 * (size_t)i;
  constexpr bool FIN = MODE != 0, INI = MODE != 1;
  const bool filter = FIN && shepard_filter_step(VARIANT, a.step_final);
  // one batch of loads, none of them behind a branch on loaded data (ldv/ldi are volatile: the compiler may
  // neither sink them into the branches that use them nor split the batch)
  const int mask = ldi(d.mask + i), type = ldi(d.type + i), solid = ldi(d.solid + i), fixed = ldi(d.fixed + i);
  const double drho = ldv(d.drho + i);
  double rho = 0.0, rhoI = 0.0, nd = 1.0, phi0 = 0.0, shep = 0.0, e_i = 0.0;
  double v[3], f[3], x[3] = {0, 0, 0}, vest[3] = {0, 0, 0}, ddv[3] = {0, 0, 0}, ddx[3] = {0, 0, 0}, nw[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 3; k++) { v[k] = ldv(d.v + i3 + k); f[k] = ldv(d.f + i3 + k); }
  if (FIN) {
    nd = ldv(d.nd + i); phi0 = ldv(d.phi + i); rhoI = ldv(d.rhoI + i);
#pragma unroll
    for (int k = 0; k < 3; k++) { nw[k] = ldv(d.nw + i3 + k); vest[k] = ldv(d.vest + i3 + k); }
  }
  if (INI) {
    if (!FIN) rho = ldv(d.rho + i);
#pragma unroll
    for (int k = 0; k < 3; k++) { ddv[k] = ldv(d.ddv + i3 + k); x[k] = ldv(d.x + i3 + k); }
    if (VARIANT != SPHBVF_TV && !FIN) nd = ldv(d.nd + i);
  }
  if (VARIANT != SPHBVF_TV) {
#pragma unroll
    for (int k = 0; k < 3; k++) ddx[k] = ldv(d.ddx + i3 + k);
  }
  if (MODE == 2 && a.do_pack) e_i = ldv(d.e + i);
  if (filter) shep = ldv(d.rhoAux1 + i) / ldv(d.rhoAux2 + i);
  const bool ingroup = (mask & a.groupbit) != 0;
  if (MODE == 2 && !ingroup) { rho = d.rho[i]; rhoI = d.rhoI[i]; }
  // deviatoric tensors exist only when some solid can carry stress (with_dev); otherwise dev == ddev == 0 and
  // the reference's dev += c * ddev is the identity
  const bool devs = a.with_dev != 0;

  // ---------------------------------------------------------------- final_integrate(step_final)
  if (FIN && ingroup) {
    const double dtv = a.dt_final, dtf = 0.5 * dtv;
    const double dtfm = dtf / co.mass[type];
    double damp, dampSolid;
    damp_factors(VARIANT, a.step_final, damp, dampSolid);
    const double phi = phi0 / nd;
    d.phi[i] = phi;
#pragma unroll
    for (int k = 0; k < 3; k++) { nw[k] = nw[k] / nd; d.nw[i3 + k] = nw[k]; }
    if (fixed == 0) {
      if (solid == 0) {
        if (phi > 0.5) {   // BVF wall reflection (fix_...transport_velocity.cpp:310-342)
          double xr[3], vr[3];
#pragma unroll
          for (int k = 0; k < 3; k++) { vr[k] = v[k]; xr[k] = (INI ? x[k] : d.x[i3 + k]) - dtv * vr[k]; }
          const double norm = sqrt(nw[0] * nw[0] + nw[1] * nw[1] + nw[2] * nw[2]);
          const double en[3] = {-nw[0] / norm, -nw[1] / norm, -nw[2] / norm};
          const double vdot = vr[0] * en[0] + vr[1] * en[1] + vr[2] * en[2];
          const double mx = vdot > 0.0 ? vdot : 0.0;   // std::max(0.0, v_dot_en)
#pragma unroll
          for (int k = 0; k < 3; k++) {
            vr[k] = -vr[k] + 2.0 * mx * en[k];
            x[k] = xr[k] + dtv * vr[k];
            if (!INI) d.x[i3 + k] = x[k];
          }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          if (VARIANT == SPHBVF_TV) v[k] = vest[k] + dtfm * f[k];
          else v[k] = vest[k] + dtfm * f[k] * damp + 0.001 * ddx[k] / nd;
        }
        if (VARIANT == SPHBVF_TV) rho = filter ? shep + dtf * drho : rhoI + dtf * drho;
        else rho = filter ? shep + dtf * drho : rhoI + dtv * drho;
      } else {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double vn = v[k];
          if (VARIANT == SPHBVF_TV) vn += dtfm * f[k];
          else { vn += dtfm * f[k] + 0.001 * ddx[k] / nd; vn *= dampSolid; }
          v[k] = vn;
        }
        const double cdev = VARIANT == SPHBVF_TV ? 0.5 * dtv : dtf;
        if (devs) for (int k = 0; k < 9; k++) d.dev[9 * (size_t)i + k] += cdev * d.ddev[9 * (size_t)i + k];
        if (VARIANT == SPHBVF_TV) rho = filter ? shep + dtf * drho : rhoI + dtf * drho;
        else rho = rhoI + dtv * drho;
      }
    } else {
      if (solid == 0) {
        rho = filter ? shep + dtv * drho : rhoI + dtv * drho;
      } else {
        if (devs) for (int k = 0; k < 9; k++) d.dev[9 * (size_t)i + k] += dtf * d.ddev[9 * (size_t)i + k];
        rho = filter ? shep : rhoI;
      }
    }
    for (int k = 0; k < co.nspecies; k++) {
      const size_t q = (size_t)i * co.nspecies + k;
      const double c = d.C[q] + d.Q[q] * dtf;
      d.C[q] = c > 0 ? c : 0.0;
    }
    if (!INI) {
      if (fixed == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) d.v[i3 + k] = v[k];
      }
      d.rho[i] = rho;
    }
  }

  // ---------------------------------------------------------------- initial_integrate(step_init)
  if (INI && ingroup) {
    const double dtv = a.dt_init, dtf = 0.5 * dtv;
    const double dtfm = dtf / co.mass[type];
    double damp, dampSolid;
    damp_factors(VARIANT, a.step_init, damp, dampSolid);
    const double ndi = VARIANT != SPHBVF_TV ? nd : 1.0;
    if (fixed == 0) {
      if (solid == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double ve;
          if (VARIANT == SPHBVF_TV) ve = v[k] + dtfm * f[k];
          else ve = v[k] + dtfm * f[k] * damp + 0.001 * ddx[k] / ndi;
          const double vn = ve - dtfm * ddv[k];
          vest[k] = ve;
          v[k] = vn;
          x[k] = x[k] + dtv * vn;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double ve, vn = v[k];
          if (VARIANT == SPHBVF_TV) ve = vn + 2.0 * dtfm * f[k];
          else ve = vn + 2.0 * dtfm * f[k] + 0.001 * ddx[k] / ndi;
          vn += dtfm * f[k];
          if (VARIANT != SPHBVF_TV) { ve *= dampSolid; vn *= dampSolid; }
          vest[k] = ve;
          v[k] = vn;
          x[k] = x[k] + dtf * vn;     // sic: dtf (fix_...transport_velocity.cpp:183-185)
        }
        const double cdev = VARIANT == SPHBVF_TV ? 0.5 * dtv : dtf;
        if (devs) for (int k = 0; k < 9; k++) d.dev[9 * (size_t)i + k] += cdev * d.ddev[9 * (size_t)i + k];
      }
#pragma unroll
      for (int k = 0; k < 3; k++) { d.vest[i3 + k] = vest[k]; d.v[i3 + k] = v[k]; d.x[i3 + k] = x[k]; }
      rhoI = rho;
      rho = rho + dtf * drho;
      d.rhoI[i] = rhoI;
      d.rho[i] = rho;
    } else {
      // a fixed atom keeps x, v and vest; final_integrate only changed its rho
      if (solid == 0) {
        rhoI = rho;
        rho = rho + dtf * drho;
        d.rhoI[i] = rhoI;
        d.rho[i] = rho;
      } else {
        if (devs) for (int k = 0; k < 9; k++) d.dev[9 * (size_t)i + k] += dtf * d.ddev[9 * (size_t)i + k];
        rhoI = rho;
        d.rhoI[i] = rhoI;
        if (FIN) d.rho[i] = rho;
      }
    }
    for (int k = 0; k < co.nspecies; k++) {
      const size_t q = (size_t)i * co.nspecies + k;
      const double c = d.C[q] + d.Q[q] * dtf;
      d.C[q] = c > 0 ? c : 0.0;
    }
  }
  if (MODE == 2 && a.do_pack) pack_atom(d, co, i, type, solid, fixed, x, v, vest, rho, rhoI, e_i, a.with_dev);
}

template <int MODE>
static void launch_integrate(const DevState &d, const Coeffs &co, const IntegArgs &a, cudaStream_t st) {
  if (a.a1 <= a.a0) return;
  const int b = nblocks(a.a1 - a.a0, 256);
  if (co.variant == SPHBVF_TV) integrate_kernel<SPHBVF_TV, MODE><<<b, 256, 0, st>>>(d, co, a);
  else if (co.variant == SPHBVF_MECHANICS) integrate_kernel<SPHBVF_MECHANICS, MODE><<<b, 256, 0, st>>>(d, co, a);
  else integrate_kernel<SPHBVF_FSI, MODE><<<b, 256, 0, st>>>(d, co, a);
  SPHBVF_LAUNCHED(1);
}

void launch_initial_integrate(const DevState &d, const Coeffs &co, double dt, long ntimestep, int groupbit,
                              int with_dev, cudaStream_t st) {
  IntegArgs a = {0.0, dt, 0, ntimestep, groupbit, 0, with_dev, nullptr, 0, d.nlocal};
  launch_integrate<0>(d, co, a, st);
}

void launch_final_integrate(const DevState &d, const Coeffs &co, double dt, long ntimestep, int groupbit,
                            int with_dev, cudaStream_t st) {
  IntegArgs a = {dt, 0.0, ntimestep, 0, groupbit, 0, with_dev, nullptr, 0, d.nlocal};
  launch_integrate<1>(d, co, a, st);
}

void launch_final_initial(const DevState &d, const Coeffs &co, double dt_final, long step_final, double dt_init,
                          long step_init, int groupbit, int do_pack, int with_dev, cudaStream_t st, const int *order,
                          int a0, int a1) {
  IntegArgs a = {dt_final, dt_init, step_final, step_init, groupbit, do_pack, with_dev, order, a0, a1 < 0 ? d.nlocal : a1};
  launch_integrate<2>(d, co, a, st);
}

// ------------------------------------------------------------------------------------------
// BC / source fixes
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double buffer_phi(const FixDesc &fx, const double *x) {
  double phi;
  if (fx.ia[2] == 0) {
    const double xo = fx.a[0] - fx.a[2], xL = fx.a[0] + fx.a[2];
    phi = (x[0] - xo) / (xL - xo);
    phi = phi * phi * phi;
  } else {
    const double yo = fx.a[1] - fx.a[3], yL = fx.a[1] + fx.a[3];
    phi = (x[1] - yo) / (yL - yo);
    phi = 0.5 * (1.0 - tanh(8.0 - 16.0 * phi));
  }
  return phi;
}

__global__ void fix_kernel(const DevState d, const __grid_constant__ Coeffs co, const FixDesc fx, const int hook) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal || !(d.mask[i] & fx.groupbit)) return;
  const size_t i3 = 3 * (size_t)i;
  const int S = co.nspecies;
  if (fx.kind == FIX_BUOYANCY) {
    const double m = co.mass[d.type[i]];
    if (fx.ia[0]) d.f[i3 + fx.ia[1]] += m * fx.a[0];
    else d.f[i3 + fx.ia[1]] += m * fx.a[0] * (d.C[(size_t)i * S + fx.ia[2]] - fx.a[1]);
  } else if (fx.kind == FIX_SETFORCE) {
    d.f[i3] = fx.a[0]; d.f[i3 + 1] = fx.a[1]; d.f[i3 + 2] = fx.a[2];
  } else if (fx.kind == FIX_CHEMRXN) {
    // ia[0] = nreact | nprod << 8; ia[1] = reactants, ia[2] = products, one byte each
    const int nr = fx.ia[0] & 255, np = fx.ia[0] >> 8;
    double *C = d.C + (size_t)i * S, *Q = d.Q + (size_t)i * S;
    double flux = fx.a[0];
    if (nr == 2) flux = fx.a[0] * C[fx.ia[1] & 255] * C[(fx.ia[1] >> 8) & 255];
    else if (nr == 1) flux = fx.a[0] * C[fx.ia[1] & 255];
    for (int j = 0; j < nr; j++) Q[(fx.ia[1] >> (8 * j)) & 255] -= flux;
    for (int j = 0; j < np; j++) Q[(fx.ia[2] >> (8 * j)) & 255] += flux;
  } else if (fx.kind == FIX_FORCING) {
    const double drx = d.x[i3] - fx.a[0], dry = d.x[i3 + 1] - fx.a[1];
    bool inside;
    if (fx.ia[2] == 0) inside = (drx * drx + dry * dry) < fx.a[2] * fx.a[2];
    else inside = fabs(drx) < fx.a[2] && fabs(dry) < fx.a[3];
    if (!inside) return;
    if (fx.ia[0] == 0) d.C[(size_t)i * S + fx.ia[1]] = fx.a[4];
    else d.vest[i3 + fx.ia[1]] = fx.a[4];
  } else if (fx.kind == FIX_BUFFER) {
    const double drx = d.x[i3] - fx.a[0], dry = d.x[i3 + 1] - fx.a[1];
    if (!(fabs(drx) < fx.a[2] && fabs(dry) < fx.a[3])) return;
    const double phi = buffer_phi(fx, &d.x[i3]);
    double *t;
    if (hook == 2) t = &d.rho[i];
    else t = fx.ia[0] == 0 ? &d.C[(size_t)i * S + fx.ia[1]] : &d.vest[i3 + fx.ia[1]];
    *t = *t - phi * (*t - fx.a[4]);
  }
}

// |v|^2 >= 0, so the IEEE bit patterns order like unsigned integers: atomicMax on the raw bits
__global__ void max_vsq_kernel(const DevState d, const int groupbit, unsigned long long *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double vsq = 0.0;
  if (i < d.nlocal && (d.mask[i] & groupbit)) {
    const size_t i3 = 3 * (size_t)i;
    vsq = d.v[i3] * d.v[i3] + d.v[i3 + 1] * d.v[i3 + 1] + d.v[i3 + 2] * d.v[i3 + 2];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) vsq = fmax(vsq, __shfl_xor_sync(0xffffffffu, vsq, o));
  if ((threadIdx.x & 31) == 0 && vsq > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(vsq));
}

void launch_max_vsq(const DevState &d, int groupbit, unsigned long long *out, cudaStream_t st) {
  if (!d.nlocal) return;
  max_vsq_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, groupbit, out);
  SPHBVF_LAUNCHED(1);
}

__global__ void derive_flags_kernel(const DevState d, const __grid_constant__ Coeffs co, int *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
  const bool solid = d.solid[i] != 0;
  bool dev = solid && co.G0[d.type[i]] != 0.0;
  for (int k = 0; k < 9 && !dev; k++) dev = d.dev[9 * (size_t)i + k] != 0.0;
  if (solid && !out[0]) out[0] = 1;
  if (dev && !((out[1] >> d.type[i]) & 1)) atomicOr(&out[1], 1 << d.type[i]);   // mask of the types that can carry stress
  if (d.e[i] != 0.0 && !out[2]) out[2] = 1;
}

void launch_derive_flags(const DevState &d, const Coeffs &co, int *out, cudaStream_t st) {
  if (!d.nlocal) return;
  derive_flags_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, co, out);
  SPHBVF_LAUNCHED(1);
}

// sum over the atoms of `groupbit` of m v_a v_b (v = atom->v), LAMMPS order xx yy zz xy xz yz: what
// ComputeTemp::compute_scalar / compute_vector accumulate on the host (compute_temp.cpp:78-135).  Two stages with
// a fixed summation order (per-block partials, then one block) so that thermo output is reproducible.
__global__ void __launch_bounds__(256)
ke_partial_kernel(const DevState d, const __grid_constant__ Coeffs co, const int groupbit, double *partial) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double t[6] = {0, 0, 0, 0, 0, 0};
  if (i < d.nlocal && (d.mask[i] & groupbit)) {
    const size_t i3 = 3 * (size_t)i;
    const double m = co.mass[d.type[i]], vx = d.v[i3], vy = d.v[i3 + 1], vz = d.v[i3 + 2];
    t[0] = m * vx * vx; t[1] = m * vy * vy; t[2] = m * vz * vz;
    t[3] = m * vx * vy; t[4] = m * vx * vz; t[5] = m * vy * vz;
  }
  __shared__ double sh[6][8];
#pragma unroll
  for (int q = 0; q < 6; q++) {
    double x = t[q];
#pragma unroll
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double x = 0.0;
    for (int w = 0; w < 8; w++) x += sh[threadIdx.x][w];
    partial[6 * (size_t)blockIdx.x + threadIdx.x] = x;
  }
}

__global__ void __launch_bounds__(256)
ke_final_kernel(const double *partial, const int nblk, double *out6) {
  __shared__ double sh[6][256];
  double t[6] = {0, 0, 0, 0, 0, 0};
  for (int b = threadIdx.x; b < nblk; b += 256)
#pragma unroll
    for (int q = 0; q < 6; q++) t[q] += partial[6 * (size_t)b + q];
#pragma unroll
  for (int q = 0; q < 6; q++) sh[q][threadIdx.x] = t[q];
  __syncthreads();
  for (int s = 128; s; s >>= 1) {
    if (threadIdx.x < s)
#pragma unroll
      for (int q = 0; q < 6; q++) sh[q][threadIdx.x] += sh[q][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 6) out6[threadIdx.x] = sh[threadIdx.x][0];
}

// scratch: >= 6 * ceil(nlocal / 256) doubles
void launch_ke_tensor(const DevState &d, const Coeffs &co, int groupbit, double *scratch, double *out6, cudaStream_t st) {
  const int nb = nblocks(d.nlocal > 0 ? d.nlocal : 1, 256);
  ke_partial_kernel<<<nb, 256, 0, st>>>(d, co, groupbit, scratch);
  ke_final_kernel<<<1, 256, 0, st>>>(scratch, nb, out6);
  SPHBVF_LAUNCHED(2);
}

// hook: 0 post_integrate, 1 post_force, 2 end_of_step
bool fix_runs(const FixDesc &fx, int hook, long ntimestep) {
  switch (fx.kind) {
    case FIX_BUOYANCY:
    case FIX_CHEMRXN:
    case FIX_SETFORCE: return hook == 1;
    case FIX_FORCING: return hook == 0 && ntimestep > fx.step;
    case FIX_BUFFER:
      if (fx.ia[0] == 2) return hook == 2 && ntimestep > fx.step;
      return hook == 0 && ntimestep > fx.step;
  }
  return false;
}

void launch_fix(const DevState &d, const Coeffs &co, const FixDesc &fx, int hook, long ntimestep, cudaStream_t st) {
  if (!d.nlocal) return;
  if (!fix_runs(fx, hook, ntimestep)) return;
  fix_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, co, fx, hook);
  SPHBVF_LAUNCHED(1);
}

// ------------------------------------------------------------------------------------------
// pack kernel (steps where the fused integrator could not write the records: rebuilds, fixes in
// post_integrate, host uploads)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_kernel(const DevState d, const __grid_constant__ Coeffs co, const int with_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
  const size_t i3 = 3 * (size_t)i;
  const double x[3] = {d.x[i3], d.x[i3 + 1], d.x[i3 + 2]};
  const double v[3] = {d.v[i3], d.v[i3 + 1], d.v[i3 + 2]};
  const double vest[3] = {d.vest[i3], d.vest[i3 + 1], d.vest[i3 + 2]};
  pack_atom(d, co, i, d.type[i], d.solid[i], d.fixed[i], x, v, vest, d.rho[i], d.rhoI[i], d.e[i], with_dev);
}

void launch_pack(const DevState &d, const Coeffs &co, int with_dev, cudaStream_t st) {
  if (!d.nlocal) return;
  pack_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, co, with_dev);
  SPHBVF_LAUNCHED(1);
}

// self-image ghosts (periodic boundaries inside one rank): copy the owner's packed record and
// shift the position, as pack_comm with pbc does (atom_vec_ssa_tsdpd_atomic.cpp:487-549).
__global__ void ghost_refresh_kernel(const DevState d, const int S, const int with_dev) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= d.nghost) return;
  const int o = d.gowner[g];
  if (o < 0) return;   // ghost owned by another rank: filled by the halo exchange
  const int q = d.nlocal + g;
  Prec r = d.prec[o];
  // x + shift evaluated in the reference's order: one rounded add per shifted dimension
  r.A.x += d.gshift[3 * (size_t)g];
  r.A.y += d.gshift[3 * (size_t)g + 1];
  r.A.z += d.gshift[3 * (size_t)g + 2];
  d.prec[q] = r;
  d.pD[q] = d.pD[o];
  d.pflags[q] = d.pflags[o];
  for (int k = 0; k < S; k++) d.pCs[(size_t)q * S + k] = d.pCs[(size_t)o * S + k];
  if (with_dev)
    for (int k = 0; k < 9; k++) d.pdev[9 * (size_t)q + k] = d.pdev[9 * (size_t)o + k];
}

void launch_ghost_refresh(const DevState &d, const Coeffs &co, int with_dev, cudaStream_t st) {
  if (!d.nghost) return;
  ghost_refresh_kernel<<<nblocks(d.nghost, 256), 256, 0, st>>>(d, co.nspecies, with_dev);
  SPHBVF_LAUNCHED(1);
}

}  // namespace sphbvf

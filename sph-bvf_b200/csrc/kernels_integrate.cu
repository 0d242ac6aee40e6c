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
  if (d.nlocal) setup_pre_force_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, groupbit);
}

__device__ __forceinline__ void damp_factors(int variant, long ntimestep, double &damp, double &dampSolid) {
  const double tnow = (double)ntimestep;
  damp = tnow <= 1.0 ? tnow / 1.0 : 1.0;
  if (variant == SPHBVF_MECHANICS) dampSolid = tnow < 1e6 ? 0.0 : 1.0;   // ..._mechanics.cpp:151-153
  else dampSolid = tnow <= 1.0 ? 0.0 : 1.0;                              // ..._fsi.cpp:150-152
}

// Both integrators are pure streaming (HBM bound).  Every array an atom may need is loaded
// UNCONDITIONALLY at the top of the kernel, before any branch on solid_tag / fixed_tag: with the
// loads inside the branches the kernels ran at 23-30 % of HBM peak, all warps waiting on one
// dependent round trip after another (ncu long_scoreboard 90 %, profiles/); the deviatoric tensors
// (solids only) stay behind their branch.
template <int VARIANT>
__global__ void __launch_bounds__(256)
initial_integrate_kernel(const DevState d, const __grid_constant__ Coeffs co, const double dtv, const long ntimestep,
                         const int groupbit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
  const size_t i3 = 3 * (size_t)i;
  const int mask = d.mask[i], type = d.type[i], solid = d.solid[i], fixed = d.fixed[i];
  const double rho = d.rho[i], drho = d.drho[i];
  double v[3], f[3], ddv[3], x[3], ddx[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 3; k++) { v[k] = d.v[i3 + k]; f[k] = d.f[i3 + k]; ddv[k] = d.ddv[i3 + k]; x[k] = d.x[i3 + k]; }
  double ndi = 1.0;
  if (VARIANT != SPHBVF_TV) {
    ndi = d.nd[i];
#pragma unroll
    for (int k = 0; k < 3; k++) ddx[k] = d.ddx[i3 + k];
  }
  if (!(mask & groupbit)) return;
  const double dtf = 0.5 * dtv;
  const double dtfm = dtf / co.mass[type];
  double damp, dampSolid;
  damp_factors(VARIANT, ntimestep, damp, dampSolid);
  if (fixed == 0) {
    if (solid == 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double vest;
        if (VARIANT == SPHBVF_TV) vest = v[k] + dtfm * f[k];
        else vest = v[k] + dtfm * f[k] * damp + 0.001 * ddx[k] / ndi;
        const double vn = vest - dtfm * ddv[k];
        d.vest[i3 + k] = vest;
        d.v[i3 + k] = vn;
        d.x[i3 + k] = x[k] + dtv * vn;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double vest, vn = v[k];
        if (VARIANT == SPHBVF_TV) vest = vn + 2.0 * dtfm * f[k];
        else vest = vn + 2.0 * dtfm * f[k] + 0.001 * ddx[k] / ndi;
        vn += dtfm * f[k];
        if (VARIANT != SPHBVF_TV) { vest *= dampSolid; vn *= dampSolid; }
        d.vest[i3 + k] = vest;
        d.v[i3 + k] = vn;
        d.x[i3 + k] = x[k] + dtf * vn;     // sic: dtf (fix_...transport_velocity.cpp:183-185)
      }
      const double cdev = VARIANT == SPHBVF_TV ? 0.5 * dtv : dtf;
      for (int k = 0; k < 9; k++) d.dev[9 * (size_t)i + k] += cdev * d.ddev[9 * (size_t)i + k];
    }
    d.rhoI[i] = rho;
    d.rho[i] = rho + dtf * drho;
  } else {
    if (solid == 0) {
      d.rhoI[i] = rho;
      d.rho[i] = rho + dtf * drho;
    } else {
      for (int k = 0; k < 9; k++) d.dev[9 * (size_t)i + k] += dtf * d.ddev[9 * (size_t)i + k];
      d.rhoI[i] = rho;
    }
  }
  for (int k = 0; k < co.nspecies; k++) {
    const size_t q = (size_t)i * co.nspecies + k;
    const double c = d.C[q] + d.Q[q] * dtf;
    d.C[q] = c > 0 ? c : 0.0;
  }
}

void launch_initial_integrate(const DevState &d, const Coeffs &co, double dt, long ntimestep, int groupbit,
                              cudaStream_t st) {
  if (!d.nlocal) return;
  const int b = nblocks(d.nlocal, 256);
  if (co.variant == SPHBVF_TV) initial_integrate_kernel<SPHBVF_TV><<<b, 256, 0, st>>>(d, co, dt, ntimestep, groupbit);
  else if (co.variant == SPHBVF_MECHANICS) initial_integrate_kernel<SPHBVF_MECHANICS><<<b, 256, 0, st>>>(d, co, dt, ntimestep, groupbit);
  else initial_integrate_kernel<SPHBVF_FSI><<<b, 256, 0, st>>>(d, co, dt, ntimestep, groupbit);
}

template <int VARIANT>
__global__ void __launch_bounds__(256)
final_integrate_kernel(const DevState d, const __grid_constant__ Coeffs co, const double dtv, const long ntimestep,
                       const int groupbit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
  const size_t i3 = 3 * (size_t)i;
  // freqFilter 20 (TV :287, mechanics :311); fsi: 1e16 -> INT_MAX, never fires (..._fsi.cpp:304)
  const bool filter = VARIANT == SPHBVF_FSI ? (ntimestep % 2147483647L) == 0 : (ntimestep % 20) == 0;
  const int mask = d.mask[i], type = d.type[i], solid = d.solid[i], fixed = d.fixed[i];
  const double nd = d.nd[i], phi0 = d.phi[i], drho = d.drho[i], rhoI = d.rhoI[i];
  double nw[3], v[3], vest[3], f[3], ddx[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 3; k++) { nw[k] = d.nw[i3 + k]; v[k] = d.v[i3 + k]; vest[k] = d.vest[i3 + k]; f[k] = d.f[i3 + k]; }
  if (VARIANT != SPHBVF_TV) {
#pragma unroll
    for (int k = 0; k < 3; k++) ddx[k] = d.ddx[i3 + k];
  }
  double shep = 0.0;
  if (filter) shep = d.rhoAux1[i] / d.rhoAux2[i];
  if (!(mask & groupbit)) return;
  const double dtf = 0.5 * dtv;
  const double dtfm = dtf / co.mass[type];
  double damp, dampSolid;
  damp_factors(VARIANT, ntimestep, damp, dampSolid);
  const double phi = phi0 / nd;
  d.phi[i] = phi;
#pragma unroll
  for (int k = 0; k < 3; k++) { nw[k] = nw[k] / nd; d.nw[i3 + k] = nw[k]; }
  double rho;
  if (fixed == 0) {
    if (solid == 0) {
      if (phi > 0.5) {   // BVF wall reflection (fix_...transport_velocity.cpp:310-342)
        double x[3], vr[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { vr[k] = v[k]; x[k] = d.x[i3 + k] - dtv * vr[k]; }
        const double norm = sqrt(nw[0] * nw[0] + nw[1] * nw[1] + nw[2] * nw[2]);
        const double en[3] = {-nw[0] / norm, -nw[1] / norm, -nw[2] / norm};
        const double vdot = vr[0] * en[0] + vr[1] * en[1] + vr[2] * en[2];
        const double mx = vdot > 0.0 ? vdot : 0.0;   // std::max(0.0, v_dot_en)
#pragma unroll
        for (int k = 0; k < 3; k++) {
          vr[k] = -vr[k] + 2.0 * mx * en[k];
          d.x[i3 + k] = x[k] + dtv * vr[k];
        }
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (VARIANT == SPHBVF_TV) d.v[i3 + k] = vest[k] + dtfm * f[k];
        else d.v[i3 + k] = vest[k] + dtfm * f[k] * damp + 0.001 * ddx[k] / nd;
      }
      if (VARIANT == SPHBVF_TV) rho = filter ? shep + dtf * drho : rhoI + dtf * drho;
      else rho = filter ? shep + dtf * drho : rhoI + dtv * drho;
    } else {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double vn = v[k];
        if (VARIANT == SPHBVF_TV) vn += dtfm * f[k];
        else { vn += dtfm * f[k] + 0.001 * ddx[k] / nd; vn *= dampSolid; }
        d.v[i3 + k] = vn;
      }
      const double cdev = VARIANT == SPHBVF_TV ? 0.5 * dtv : dtf;
      for (int k = 0; k < 9; k++) d.dev[9 * (size_t)i + k] += cdev * d.ddev[9 * (size_t)i + k];
      if (VARIANT == SPHBVF_TV) rho = filter ? shep + dtf * drho : rhoI + dtf * drho;
      else rho = rhoI + dtv * drho;
    }
  } else {
    if (solid == 0) {
      rho = filter ? shep + dtv * drho : rhoI + dtv * drho;
    } else {
      for (int k = 0; k < 9; k++) d.dev[9 * (size_t)i + k] += dtf * d.ddev[9 * (size_t)i + k];
      rho = filter ? shep : rhoI;
    }
  }
  d.rho[i] = rho;
  for (int k = 0; k < co.nspecies; k++) {
    const size_t q = (size_t)i * co.nspecies + k;
    const double c = d.C[q] + d.Q[q] * dtf;
    d.C[q] = c > 0 ? c : 0.0;
  }
}

void launch_final_integrate(const DevState &d, const Coeffs &co, double dt, long ntimestep, int groupbit,
                            cudaStream_t st) {
  if (!d.nlocal) return;
  const int b = nblocks(d.nlocal, 256);
  if (co.variant == SPHBVF_TV) final_integrate_kernel<SPHBVF_TV><<<b, 256, 0, st>>>(d, co, dt, ntimestep, groupbit);
  else if (co.variant == SPHBVF_MECHANICS) final_integrate_kernel<SPHBVF_MECHANICS><<<b, 256, 0, st>>>(d, co, dt, ntimestep, groupbit);
  else final_integrate_kernel<SPHBVF_FSI><<<b, 256, 0, st>>>(d, co, dt, ntimestep, groupbit);
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
  if (d.nlocal) max_vsq_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, groupbit, out);
}

// hook: 0 post_integrate, 1 post_force, 2 end_of_step
void launch_fix(const DevState &d, const Coeffs &co, const FixDesc &fx, int hook, long ntimestep, cudaStream_t st) {
  if (!d.nlocal) return;
  bool run = false;
  switch (fx.kind) {
    case FIX_BUOYANCY:
    case FIX_CHEMRXN:
    case FIX_SETFORCE: run = hook == 1; break;
    case FIX_FORCING: run = hook == 0 && ntimestep > fx.step; break;
    case FIX_BUFFER:
      if (fx.ia[0] == 2) run = hook == 2 && ntimestep > fx.step;
      else run = hook == 0 && ntimestep > fx.step;
      break;
  }
  if (run) fix_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, co, fx, hook);
}

// ------------------------------------------------------------------------------------------
// pack: primary state -> pair input records.  Holds every per-particle division of the pair
// pass: V = m/rho, P/rho^2 with P = 7 B (rho/rho0 - 1) (pair_...transport_velocity.cpp:298-299),
// and the scalar artificial stress of a stress-free solid (:454-461 with dev = 0).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_kernel(const DevState d, const __grid_constant__ Coeffs co, const int with_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.nlocal) return;
  const size_t i3 = 3 * (size_t)i;
  const int t = d.type[i];
  const double rho = d.rho[i];
  const double vx = d.vest[i3], vy = d.vest[i3 + 1], vz = d.vest[i3 + 2];
  const double irho = 1.0 / rho;
  const double P = 7.0 * co.B[t] * (rho / co.rho0[t] - 1.0);
  const double Prr = P * irho * irho;
  Prec r;
  r.A = make_rec4(d.x[i3], d.x[i3 + 1], d.x[i3 + 2], rho);
  r.B = make_rec4(vx, vy, vz, co.mass[t] * irho);
  r.C = make_rec4(vx - d.v[i3], vy - d.v[i3 + 1], vz - d.v[i3 + 2], Prr);
  d.prec[i] = r;
  const int solid = d.solid[i];
  double art = 0.0;
  if (solid) {
    const double c_art = co.variant == SPHBVF_FSI ? 0.1 : 0.35;
    const double Ps = co.variant == SPHBVF_MECHANICS ? fabs(P) : P;
    const double ts = -Ps;
    art = ts > 0.0 ? -c_art * ts * irho * irho : 0.0;
  }
  const double C0 = co.nspecies ? d.C[(size_t)i * co.nspecies] : 0.0;
  d.pD[i] = make_rec4(d.rhoI[i], art, C0, d.e[i]);
  d.pflags[i] = t | (solid << 4) | (d.fixed[i] << 5);
  for (int k = 0; k < co.nspecies; k++) d.pCs[(size_t)i * co.nspecies + k] = d.C[(size_t)i * co.nspecies + k];
  if (with_dev)
    for (int k = 0; k < 9; k++) d.pdev[9 * (size_t)i + k] = d.dev[9 * (size_t)i + k];
}

void launch_pack(const DevState &d, const Coeffs &co, int with_dev, cudaStream_t st) {
  if (d.nlocal) pack_kernel<<<nblocks(d.nlocal, 256), 256, 0, st>>>(d, co, with_dev);
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
  if (d.nghost) ghost_refresh_kernel<<<nblocks(d.nghost, 256), 256, 0, st>>>(d, co.nspecies, with_dev);
}

}  // namespace sphbvf

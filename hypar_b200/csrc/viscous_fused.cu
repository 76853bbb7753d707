// viscous_fused.cu -- production path of the Navier-Stokes viscous terms (NavierStokes3DParabolicFunction.c:50-325).
//
// The reference evaluates the parabolic term in ~20 passes over memory (Q, three QDeriv arrays, per
// direction FViscous and FDeriv, all ghost-padded 5-component temporaries calloc'ed per call). Here:
//   (1) k_qderiv3 (this file): ONE kernel writes the (un-scaled) first differences of (u, v, w, T) in all three
//       directions (12 scalars per point; the density derivative is never used), each multiplied by the
//       local mu/Re, mu = T^0.76 -- every term of the viscous flux is (mu/Re) x a derivative, so the sweeps
//       need neither the viscosity nor exp/log. Q is evaluated in registers from the conserved variables.
//   (2) [halo exchange of the x- and y-derivative arrays; the z-derivatives are NOT exchanged: quirk Q1]
//   (3) the viscous flux of direction d and its derivative are evaluated INSIDE the hyperbolic sweep
//       kernel of direction d (sweep_fused.cuh, VISC = true), whose tile already holds the stencil cells.
// Semantics kept from the reference: a derivative along d exists only on lines
// whose transverse indices are interior; along the line it covers the ghost cells too, with the biased
// stencil of FirstDerivativeFourthOrder.c:92/:114 next to the line ends; every other location of the
// derivative arrays stays zero unless the halo exchange fills it. The outermost ghost layer (one-sided
// stencil, :82/:124) is never consumed and is not computed.
#include "hpb_internal.h"

namespace {

struct QD3Args {
  Geom G;
  double gamma, inv_Re;
  const double* u;
  const double* dxinv;
  double* qd;          // [dir][comp 0..3 = u, v, w, T][npg]
};

__device__ __forceinline__ void prim4(const double* __restrict__ u, long long npg, long long p, double gamma, double (&q)[4])
{
  const double rho = __ldg(u + p);
  const double rinv = 1.0 / rho;
  const double vx = (rho == 0) ? 0.0 : __ldg(u + npg + p) * rinv;
  const double vy = (rho == 0) ? 0.0 : __ldg(u + 2 * npg + p) * rinv;
  const double vz = (rho == 0) ? 0.0 : __ldg(u + 3 * npg + p) * rinv;
  const double e = __ldg(u + 4 * npg + p);
  const double P = (e - 0.5 * rho * (vx * vx + vy * vy + vz * vz)) * (gamma - 1.0);
  q[0] = vx; q[1] = vy; q[2] = vz; q[3] = gamma * P * rinv;
}

__global__ void __launch_bounds__(128) k_qderiv3(const QD3Args a)
{
  const Geom& G = a.G;
  const int g = G.g;
  const int i = blockIdx.x * blockDim.x + threadIdx.x - g, j = blockIdx.y - g, k = blockIdx.z - g;
  if (i >= G.N[0] + g) return;
  const bool in0 = (i >= 0 && i < G.N[0]), in1 = (j >= 0 && j < G.N[1]), in2 = (k >= 0 && k < G.N[2]);
  if ((int)in0 + (int)in1 + (int)in2 < 2) return;                     // edges / corners: never touched
  const long long p = (i + g) + (long long)G.P[0] * ((j + g) + (long long)G.P[1] * (k + g));
  const int idx[3] = { i, j, k };
  const bool inn[3] = { in0, in1, in2 };
  const double s12 = 1.0 / 12.0;
  // mu/Re at this point: mu = exp(0.76 log T) (raiseto, math_ops.h:37; NavierStokes3DParabolicFunction.c:174)
  double qc[4];
  prim4(a.u, G.npg, p, a.gamma, qc);
  const double muRe = exp(0.76 * log(qc[3])) * a.inv_Re;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    // transverse indices must be interior
    if (!(inn[(d + 1) % 3] && inn[(d + 2) % 3])) continue;
    const int x = idx[d], N = G.N[d];
    if (x == -g || x == N + g - 1) continue;                          // one-sided layer: never consumed
    const long long st = G.st[d];
    double D[4];
    if (x == -g + 1) {
      double f0[4], f1[4], f2[4], f3[4], f4[4];
      prim4(a.u, G.npg, p - st, a.gamma, f0); prim4(a.u, G.npg, p, a.gamma, f1); prim4(a.u, G.npg, p + st, a.gamma, f2);
      prim4(a.u, G.npg, p + 2 * st, a.gamma, f3); prim4(a.u, G.npg, p + 3 * st, a.gamma, f4);
#pragma unroll
      for (int c = 0; c < 4; c++) D[c] = (-3 * f0[c] - 10 * f1[c] + 18 * f2[c] - 6 * f3[c] + f4[c]) * s12;
    } else if (x == N + g - 2) {
      double f0[4], f1[4], f2[4], f3[4], f4[4];
      prim4(a.u, G.npg, p - 3 * st, a.gamma, f0); prim4(a.u, G.npg, p - 2 * st, a.gamma, f1); prim4(a.u, G.npg, p - st, a.gamma, f2);
      prim4(a.u, G.npg, p, a.gamma, f3); prim4(a.u, G.npg, p + st, a.gamma, f4);
#pragma unroll
      for (int c = 0; c < 4; c++) D[c] = (-f0[c] + 6 * f1[c] - 18 * f2[c] + 10 * f3[c] + 3 * f4[c]) * s12;
    } else {
      double fm2[4], fm1[4], fp1[4], fp2[4];
      prim4(a.u, G.npg, p - 2 * st, a.gamma, fm2); prim4(a.u, G.npg, p - st, a.gamma, fm1);
      prim4(a.u, G.npg, p + st, a.gamma, fp1); prim4(a.u, G.npg, p + 2 * st, a.gamma, fp2);
#pragma unroll
      for (int c = 0; c < 4; c++) D[c] = (fm2[c] - 8 * fm1[c] + 8 * fp1[c] - fp2[c]) * s12;
    }
    // NOT scaled by dxinv here: the reference scales AFTER the halo exchange with the receiver's local dxinv
    // (NavierStokes3DParabolicFunction.c:125-146); the sweeps apply it per cell
    const double dxi = muRe;
#pragma unroll
    for (int c = 0; c < 4; c++) a.qd[(long long)(d * 4 + c) * G.npg + p] = D[c] * dxi;
  }
}

// pack / unpack of one 4-component derivative array face (same face boxes as the solution exchange)
__global__ void k_face4(Geom G, double* __restrict__ a, int d, int off_d, double* __restrict__ buf, int to_buf)
{
  int b[3] = { G.N[0], G.N[1], G.N[2] };
  b[d] = G.g;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x, t1 = blockIdx.y, t2 = blockIdx.z;
  if (t0 >= b[0]) return;
  int s[3] = { t0, t1, t2 };
  s[d] += off_d;
  const long long p1 = (s[0] + G.g) + (long long)G.P[0] * ((s[1] + G.g) + (long long)G.P[1] * (s[2] + G.g));
  const long long nface = (long long)b[0] * b[1] * b[2];
  const long long p2 = t0 + (long long)b[0] * (t1 + (long long)b[1] * t2);
  if (to_buf) for (int v = 0; v < 4; v++) buf[v * nface + p2] = a[v * G.npg + p1];
  else        for (int v = 0; v < 4; v++) a[v * G.npg + p1] = buf[v * nface + p2];
}

} // namespace

namespace hpbk {

void qderiv_fused(hpb_solver* h, const double* u)
{
  ProfScope ps(h, HPB_PROF_VISCOUS);
  const Geom& G = h->geo;
  QD3Args a; a.G = G; a.gamma = h->phys.gamma; a.inv_Re = 1.0 / h->phys.Re; a.u = u; a.dxinv = h->d_dxinv; a.qd = h->d_qd4;
  dim3 grid((G.P[0] + 127) / 128, G.P[1], G.P[2]);
  k_qderiv3<<<grid, 128, 0, h->stream>>>(a);
  h->launches++;
}

static void face4(hpb_solver* h, double* a, int d, int off_d, double* buf, int to_buf)
{
  const Geom& G = h->geo;
  int b[3] = { G.N[0], G.N[1], G.N[2] };
  b[d] = G.g;
  k_face4<<<dim3((b[0] + 127) / 128, b[1], b[2]), 128, 0, h->stream>>>(G, a, d, off_d, buf, to_buf);
  h->launches++;
}

// field = HPB_FIELD_QDERIVX / HPB_FIELD_QDERIVY of the fused path: 4-component arrays inside d_qd4
void pack_qd4(hpb_solver* h, int field)
{
  ProfScope ps(h, HPB_PROF_HALO);
  const Geom& G = h->geo;
  double* a = h->d_qd4 + (long long)(field - 1) * 4 * G.npg;
  for (int d = 0; d < G.ndims; d++) {
    if (h->neighbor[2*d] >= 0)   face4(h, a, d, 0, h->d_send[field][2*d], 1);
    if (h->neighbor[2*d+1] >= 0) face4(h, a, d, G.N[d] - G.g, h->d_send[field][2*d+1], 1);
  }
}
void unpack_qd4(hpb_solver* h, int field)
{
  ProfScope ps(h, HPB_PROF_HALO);
  const Geom& G = h->geo;
  double* a = h->d_qd4 + (long long)(field - 1) * 4 * G.npg;
  for (int d = 0; d < G.ndims; d++) {
    if (h->neighbor[2*d] >= 0)   face4(h, a, d, -G.g, h->d_recv[field][2*d], 0);
    if (h->neighbor[2*d+1] >= 0) face4(h, a, d, G.N[d], h->d_recv[field][2*d+1], 0);
  }
}

} // namespace hpbk

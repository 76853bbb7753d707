// kernels.cu -- CUDA kernels (sm_100a) of the explicit-RHS path: layout transposes, boundary
// conditions, halo pack/unpack, the generic per-interface hyperbolic kernel (all models, component
// and characteristic WENO5, Rusanov/Roe), flux divergence + gravity source, Navier-Stokes viscous
// terms, LinearADR diffusion, RK stage updates and reductions.
// The fused per-cell sweep kernels for the component-wise hot configurations live in sweep_fused.cu.
#include <array>
#include <vector>
#include "hpb_internal.h"
#include "physics.cuh"
#include "weno.cuh"

namespace {

constexpr int TPB = 128;

#define MODEL_SWITCH(model, CALL)                                            \
  switch (model) {                                                           \
    case HPB_MODEL_LINEAR_ADR: { CALL(HPB_MODEL_LINEAR_ADR); } break;        \
    case HPB_MODEL_EULER1D:    { CALL(HPB_MODEL_EULER1D); } break;           \
    case HPB_MODEL_NS2D:       { CALL(HPB_MODEL_NS2D); } break;              \
    case HPB_MODEL_BURGERS:    { CALL(HPB_MODEL_BURGERS); } break;           \
    default:                   { CALL(HPB_MODEL_NS3D); } break;              \
  }

__device__ __forceinline__ long long cell_index(const Geom& G, int i0, int i1, int i2)
{
  // local interior indices (may be negative / >= N for ghosts)
  long long p = i0 + G.g;
  if (G.ndims > 1) p += (long long)G.P[0] * (i1 + G.g);
  if (G.ndims > 2) p += (long long)G.P[0] * G.P[1] * (i2 + G.g);
  return p;
}

// ------------------------------------------------------------------------------------------
// layout transposes: HyPar AoS (nvars innermost) <-> device SoA (component-major)
__global__ void k_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, long long n, int nv)
{
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  for (int v = 0; v < nv; v++) soa[v * n + p] = aos[p * nv + v];
}
__global__ void k_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, long long n, int nv)
{
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  for (int v = 0; v < nv; v++) aos[p * nv + v] = soa[v * n + p];
}

// ------------------------------------------------------------------------------------------
// boundary conditions: one thread per ghost point of the zone box [is, ie)
// BCPeriodic.c:19-63 (only iproc[dim]==1), BCExtrapolate.c, BCSlipWall.c
__global__ void k_bc_zone(Geom G, ZoneDev z, double gamma, double* __restrict__ phi)
{
  // linear thread index over the zone box, dim 0 fastest (x-faces are only g = 3 points wide)
  const int b0 = z.ie[0] - z.is[0], b1 = (G.ndims > 1 ? z.ie[1] - z.is[1] : 1), b2 = (G.ndims > 2 ? z.ie[2] - z.is[2] : 1);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)b0 * b1 * b2) return;
  const int t0 = (int)(t % b0), t1 = (int)((t / b0) % b1), t2 = (int)(t / ((long long)b0 * b1));
  int ib[3] = { t0, t1, t2 };
  int i1[3] = { t0 + z.is[0], t1 + z.is[1], t2 + z.is[2] };
  int i2[3] = { i1[0], i1[1], i1[2] };
  const int dim = z.dim, nv = G.nvars;
  if (z.type == HPB_BC_PERIODIC) {
    // source index uses the box-local index (no zone offset), as the reference does
    i2[0] = ib[0]; i2[1] = ib[1]; i2[2] = ib[2];
    if (z.face == 1) i2[dim] = ib[dim] + G.N[dim] - G.g;
  } else {
    if (z.face == 1) i2[dim] = G.g - 1 - ib[dim];
    else             i2[dim] = G.N[dim] - ib[dim] - 1;
  }
  const long long p1 = cell_index(G, i1[0], i1[1], i1[2]);
  const long long p2 = cell_index(G, i2[0], i2[1], i2[2]);
  if (z.type == HPB_BC_PERIODIC || z.type == HPB_BC_EXTRAPOLATE) {
    for (int v = 0; v < nv; v++) phi[v * G.npg + p1] = phi[v * G.npg + p2];
    return;
  }
  if (z.type == HPB_BC_DIRICHLET) {                 // BCDirichlet.c:20-38
    for (int v = 0; v < nv; v++) phi[v * G.npg + p1] = z.val[v];
    return;
  }
  if (z.type != HPB_BC_SLIP_WALL) {
    // BCNoslipWall.c, BCSubsonicInflow.c, BCSubsonicOutflow.c, BCSubsonicAmbivalent.c, BCSupersonicInflow.c,
    // BCSupersonicOutflow.c (2-D and 3-D branches): density, velocity and pressure of the ghost point come from the
    // mirrored interior point or from the zone's data; the energy is recomputed. 2-D: _Euler2DGetFlowVar_ (no
    // rho == 0 guard); 3-D: _NavierStokes3DGetFlowVar_.
    const int nd = G.ndims;
    const double inv_gamma_m1 = 1.0 / (gamma - 1.0);
    const double rho = phi[p2];
    double vel[3] = { 0.0, 0.0, 0.0 };
    for (int k = 0; k < nd; k++) vel[k] = (nd == 3 && rho == 0) ? 0.0 : phi[(1 + k) * G.npg + p2] / rho;
    const double energy = phi[(nv - 1) * G.npg + p2];
    const double vsq = (nd == 2) ? (vel[0] * vel[0]) + (vel[1] * vel[1]) : (vel[0] * vel[0]) + (vel[1] * vel[1]) + (vel[2] * vel[2]);
    const double pressure = (energy - 0.5 * rho * vsq) * (gamma - 1.0);
    bool inflow = (z.type == HPB_BC_SUBSONIC_INFLOW), outflow = (z.type == HPB_BC_SUBSONIC_OUTFLOW);
    if (z.type == HPB_BC_SUBSONIC_AMBIVALENT) {
      // face velocity by 2nd-order extrapolation from the two interior points next to the boundary, dotted with the
      // inward normal (BCSubsonicAmbivalent.c:66-92)
      int j1[3] = { i1[0], i1[1], i1[2] }, j2[3] = { i1[0], i1[1], i1[2] };
      if (z.face == 1) { j1[dim] = 0; j2[dim] = 1; } else { j1[dim] = G.N[dim] - 1; j2[dim] = G.N[dim] - 2; }
      const long long q1 = cell_index(G, j1[0], j1[1], j1[2]), q2 = cell_index(G, j2[0], j2[1], j2[2]);
      const double r1 = phi[q1], r2 = phi[q2];
      double vb[3] = { 0.0, 0.0, 0.0 };
      for (int k = 0; k < nd; k++) {
        const double a1 = (nd == 3 && r1 == 0) ? 0.0 : phi[(1 + k) * G.npg + q1] / r1;
        const double a2 = (nd == 3 && r2 == 0) ? 0.0 : phi[(1 + k) * G.npg + q2] / r2;
        vb[k] = 1.5 * a1 - 0.5 * a2;
      }
      double nrm[3] = { dim == 0 ? 1.0 : 0.0, dim == 1 ? 1.0 : 0.0, dim == 2 ? 1.0 : 0.0 };
      for (int k = 0; k < 3; k++) nrm[k] *= (double) z.face;
      const double vn = (nd == 2) ? vb[0] * nrm[0] + vb[1] * nrm[1] : vb[0] * nrm[0] + vb[1] * nrm[1] + vb[2] * nrm[2];
      if (vn > 0) inflow = true; else outflow = true;
    }
    double rho_gpt = rho, pressure_gpt = pressure, vg[3] = { vel[0], vel[1], vel[2] };
    if (z.type == HPB_BC_NOSLIP_WALL) {
      for (int k = 0; k < nd; k++) vg[k] = 2.0 * z.wall[k] - vel[k];
    } else if (inflow) {
      rho_gpt = z.rho;
      for (int k = 0; k < nd; k++) vg[k] = z.wall[k];
    } else if (outflow) {
      pressure_gpt = z.pressure;
    } else if (z.type == HPB_BC_SUPERSONIC_INFLOW) {
      rho_gpt = z.rho; pressure_gpt = z.pressure;
      for (int k = 0; k < nd; k++) vg[k] = z.wall[k];
    }
    const double vgsq = (nd == 2) ? vg[0] * vg[0] + vg[1] * vg[1] : vg[0] * vg[0] + vg[1] * vg[1] + vg[2] * vg[2];
    const double energy_gpt = inv_gamma_m1 * pressure_gpt + 0.5 * rho_gpt * vgsq;
    phi[p1] = rho_gpt;
    for (int k = 0; k < nd; k++) phi[(1 + k) * G.npg + p1] = rho_gpt * vg[k];
    phi[(nv - 1) * G.npg + p1] = energy_gpt;
    return;
  }
  // slip wall: rho, p copied; normal velocity 2*v_wall - v; energy recomputed
  const int ndv = nv - 2;
  const double rho = phi[p2];
  double vel[3] = { 0.0, 0.0, 0.0 }, vsq = 0.0;
  for (int k = 0; k < ndv; k++) {
    // 3-D uses _NavierStokes3DGetFlowVar_ (rho==0 guard); 1-D/2-D use the unguarded Euler macros
    vel[k] = (ndv == 3 && rho == 0) ? 0.0 : phi[(1 + k) * G.npg + p2] / rho;
  }
  for (int k = 0; k < ndv; k++) vsq += vel[k] * vel[k];
  const double energy = phi[(nv - 1) * G.npg + p2];
  // 1-D: _Euler1DGetFlowVar_ multiplies 0.5*rho*v*v left to right (BCSlipWall.c 1-D branch)
  const double pressure = (ndv == 1) ? (energy - 0.5 * rho * vel[0] * vel[0]) * (gamma - 1.0)
                                     : (energy - 0.5 * rho * vsq) * (gamma - 1.0);
  const double inv_gamma_m1 = 1.0 / (gamma - 1.0);
  double vg[3] = { vel[0], vel[1], vel[2] };
  vg[dim] = 2.0 * z.wall[dim] - vel[dim];
  double vgsq = 0.0;
  for (int k = 0; k < ndv; k++) vgsq += vg[k] * vg[k];
  const double energy_gpt = (ndv == 1) ? inv_gamma_m1 * pressure + 0.5 * rho * vg[0] * vg[0]
                                       : inv_gamma_m1 * pressure + 0.5 * rho * vgsq;
  phi[p1] = rho;
  for (int k = 0; k < ndv; k++) phi[(1 + k) * G.npg + p1] = rho * vg[k];
  phi[(nv - 1) * G.npg + p1] = energy_gpt;
}

// BCSponge.c:25-75 (BCSpongeSource): out -= sigma (u - u_ref) inside the zone box, sigma rising linearly from 0 at the
// zone's start to 1 at its end (x = the cell's coordinate along the zone's dimension)
__global__ void k_sponge(Geom G, ZoneDev z, const double* __restrict__ x, const double* __restrict__ u, double* __restrict__ out)
{
  const int b0 = z.ie[0] - z.is[0], b1 = (G.ndims > 1 ? z.ie[1] - z.is[1] : 1), b2 = (G.ndims > 2 ? z.ie[2] - z.is[2] : 1);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)b0 * b1 * b2) return;
  const int t0 = (int)(t % b0), t1 = (int)((t / b0) % b1), t2 = (int)(t / ((long long)b0 * b1));
  const int i1[3] = { t0 + z.is[0], t1 + (G.ndims > 1 ? z.is[1] : 0), t2 + (G.ndims > 2 ? z.is[2] : 0) };
  const double xc = x[G.xoff[z.dim] + G.g + i1[z.dim]];
  const double sigma = (z.face > 0) ? (xc - z.xs) / (z.xe - z.xs) : (xc - z.xe) / (z.xs - z.xe);
  const long long p = cell_index(G, i1[0], i1[1], i1[2]);
  for (int v = 0; v < G.nvars; v++) out[v * G.npg + p] -= (sigma * (u[v * G.npg + p] - z.val[v]));
}

// ------------------------------------------------------------------------------------------
// GENERIC hyperbolic kernel: one thread per interface. Restates ReconstructHyperbolic
// (HyperbolicFunction.c:167-222) without materialising fluxC, uC, the 12 weight arrays or
// uL/uR/fL/fR: flux + modified solution of the six stencil cells, the four weight sets
// (L/R x {flux, raw u}: WENOFifthOrderCalculateWeights.c:111-745; characteristic :760-1440),
// the four reconstructions (Interp1PrimFifthOrderWENO.c:74 / ...Char.c:86), the upwind flux,
// and -- in gravity directions -- the interface value of the well-balanced source function
// reconstructed with the SAME flux weights (NavierStokes3DSource.c:77-78, quirk Q5).
template <int MODEL>
__global__ void __launch_bounds__(TPB)
k_iface(Geom G, Phys ph, const double* __restrict__ u, const double* __restrict__ gf,
        const double* __restrict__ gg, int dir, double* __restrict__ fI, double* __restrict__ sI, int face)
{
  // face < 0: every interface of the sweep. face = 0 / 1: only the interfaces on the block's low / high face
  // along dir (interface index 0 / N_dir), written compactly (the face box has extent 1 along dir): the
  // boundary-flux bookkeeping of HyperbolicFunction.c:103-106 evaluates exactly these.
  constexpr int NV = ModelTraits<MODEL>::NV;
  int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (face >= 0) { if (dir == 0) M0 = 1; else if (dir == 1) M1 = 1; else M2 = 1; }
  if (i0 >= M0) return;
  const long long q = i0 + (long long)M0 * (i1 + (long long)M1 * i2);
  const long long ni = (long long)M0 * M1 * M2;
  if (face >= 0) { const int at = face ? G.N[dir] : 0; if (dir == 0) i0 = at; else if (dir == 1) i1 = at; else i2 = at; }
  const long long st = G.st[dir];
  const long long pm1 = cell_index(G, i0, i1, i2) - st;      // cell left of the interface

  double U[6][NV], F[6][NV], V[6][NV], GG[6];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const long long p = pm1 + (k - 2) * st;
#pragma unroll
    for (int v = 0; v < NV; v++) U[k][v] = u[v * G.npg + p];
    const double gfk = gf ? gf[p] : 1.0;
    GG[k] = gg ? gg[p] : 1.0;
    flux_fn<MODEL>(ph, U[k], dir, F[k], p);
    modified_fn<MODEL>(ph, U[k], gfk, GG[k], V[k]);
  }
  // kappa sources of the upwinding (grav fields); LinearADR with a varying advection field: the speeds of the two cells
  const bool lin_var = (MODEL == HPB_MODEL_LINEAR_ADR) && ph.advf != nullptr;
  const double kL = lin_var ? ph.advf[dir * ph.advf_npg + pm1] : (MODEL == HPB_MODEL_EULER1D) ? (gf ? gf[pm1] : 1.0) : GG[2];
  const double kR = lin_var ? ph.advf[dir * ph.advf_npg + pm1 + st] : (MODEL == HPB_MODEL_EULER1D) ? (gf ? gf[pm1 + st] : 1.0) : GG[3];

  double fL[NV], fR[NV], uL[NV], uR[NV];
  double wLF[NV][3], wRF[NV][3];
  double srcL[2] = { 0.0, 0.0 }, srcR[2] = { 0.0, 0.0 };      // characteristic reconstruction of the source function
  const bool use_char = ph.interp_char && (MODEL == HPB_MODEL_EULER1D || MODEL == HPB_MODEL_NS2D || MODEL == HPB_MODEL_NS3D);

  if (!use_char) {
#pragma unroll
    for (int v = 0; v < NV; v++) {
      double w1, w2, w3;
      if (ph.no_limiting) { w1 = 0.1; w2 = 0.6; w3 = 0.3; }
      else weno_weights_ref(ph.weno, ph.eps, F[0][v], F[1][v], F[2][v], F[3][v], F[4][v], w1, w2, w3);
      wLF[v][0] = w1; wLF[v][1] = w2; wLF[v][2] = w3;
      fL[v] = weno_combine(w1, w2, w3, F[0][v], F[1][v], F[2][v], F[3][v], F[4][v]);
      if (!ph.no_limiting) weno_weights_ref(ph.weno, ph.eps, F[5][v], F[4][v], F[3][v], F[2][v], F[1][v], w1, w2, w3);
      wRF[v][0] = w1; wRF[v][1] = w2; wRF[v][2] = w3;
      fR[v] = weno_combine(w1, w2, w3, F[5][v], F[4][v], F[3][v], F[2][v], F[1][v]);
      // weights from RAW u, applied to the modified solution (quirk Q4)
      if (!ph.no_limiting) weno_weights_ref(ph.weno, ph.eps, U[0][v], U[1][v], U[2][v], U[3][v], U[4][v], w1, w2, w3);
      uL[v] = weno_combine(w1, w2, w3, V[0][v], V[1][v], V[2][v], V[3][v], V[4][v]);
      if (!ph.no_limiting) weno_weights_ref(ph.weno, ph.eps, U[5][v], U[4][v], U[3][v], U[2][v], U[1][v], w1, w2, w3);
      uR[v] = weno_combine(w1, w2, w3, V[5][v], V[4][v], V[3][v], V[2][v], V[1][v]);
    }
  } else {
    // characteristic: project the stencils with L(Roe average of raw u at cells i-1, i)
    double uavg[NV], lam[NV], L[NV * NV], R[NV * NV];
    roe_average<MODEL>(ph, U[2], U[3], uavg);
    eigen<MODEL>(ph, uavg, dir, lam, L, R);
    double fLc[NV], fRc[NV], uLc[NV], uRc[NV];
    double sLc[NV], sRc[NV];          // gravity-source function in characteristic space (only when sI != nullptr)
#pragma unroll
    for (int v = 0; v < NV; v++) {
      double cF[6], cU[6], cV[6], cG[6];
#pragma unroll
      for (int k = 0; k < 6; k++) {
        double sF = 0.0, sU = 0.0, sV = 0.0, sG = 0.0;
#pragma unroll
        for (int j = 0; j < NV; j++) {
          sF += L[v * NV + j] * F[k][j];
          sU += L[v * NV + j] * U[k][j];
          sV += L[v * NV + j] * V[k][j];
          // source function G_k = g_k (0, d_x, d_y, d_z, 1) (Euler1D: S (0, 1, 1)), projected like the flux
          // (Interp1PrimFifthOrderWENOChar.c:160-176 applied to SourceC, NavierStokes3DSource.c:77-78)
          const double Gj = (j == 0) ? 0.0 : ((j == NV - 1) ? GG[k] : GG[k] * (dir == j - 1));
          sG += L[v * NV + j] * Gj;
        }
        cF[k] = sF; cU[k] = sU; cV[k] = sV; cG[k] = sG;
      }
      double w1, w2, w3;
      if (ph.no_limiting) { w1 = 0.1; w2 = 0.6; w3 = 0.3; }
      else weno_weights_ref(ph.weno, ph.eps, cF[0], cF[1], cF[2], cF[3], cF[4], w1, w2, w3);
      wLF[v][0] = w1; wLF[v][1] = w2; wLF[v][2] = w3;
      fLc[v] = weno_combine(w1, w2, w3, cF[0], cF[1], cF[2], cF[3], cF[4]);
      if (!ph.no_limiting) weno_weights_ref(ph.weno, ph.eps, cF[5], cF[4], cF[3], cF[2], cF[1], w1, w2, w3);
      wRF[v][0] = w1; wRF[v][1] = w2; wRF[v][2] = w3;
      fRc[v] = weno_combine(w1, w2, w3, cF[5], cF[4], cF[3], cF[2], cF[1]);
      sLc[v] = weno_combine(wLF[v][0], wLF[v][1], wLF[v][2], cG[0], cG[1], cG[2], cG[3], cG[4]);
      sRc[v] = weno_combine(wRF[v][0], wRF[v][1], wRF[v][2], cG[5], cG[4], cG[3], cG[2], cG[1]);
      if (!ph.no_limiting) weno_weights_ref(ph.weno, ph.eps, cU[0], cU[1], cU[2], cU[3], cU[4], w1, w2, w3);
      uLc[v] = weno_combine(w1, w2, w3, cV[0], cV[1], cV[2], cV[3], cV[4]);
      if (!ph.no_limiting) weno_weights_ref(ph.weno, ph.eps, cU[5], cU[4], cU[3], cU[2], cU[1], w1, w2, w3);
      uRc[v] = weno_combine(w1, w2, w3, cV[5], cV[4], cV[3], cV[2], cV[1]);
    }
#pragma unroll
    for (int i = 0; i < NV; i++) {
      double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
#pragma unroll
      for (int j = 0; j < NV; j++) {
        a += R[i * NV + j] * fLc[j]; b += R[i * NV + j] * fRc[j];
        c += R[i * NV + j] * uLc[j]; d += R[i * NV + j] * uRc[j];
      }
      fL[i] = a; fR[i] = b; uL[i] = c; uR[i] = d;
    }
    if (sI != nullptr) {
      // back-projection of the source function; only components dir+1 and NV-1 are consumed
      const int rows[2] = { dir + 1, NV - 1 };
#pragma unroll
      for (int r = 0; r < 2; r++) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int j = 0; j < NV; j++) { a += R[rows[r] * NV + j] * sLc[j]; b += R[rows[r] * NV + j] * sRc[j]; }
        srcL[r] = a; srcR[r] = b;
      }
    }
  }

  double fhat[NV];
  upwind_fn<MODEL>(ph, dir, fL, fR, uL, uR, U[2], U[3], kL, kR, fhat);
#pragma unroll
  for (int v = 0; v < NV; v++) fI[v * ni + q] = fhat[v];

  if ((MODEL == HPB_MODEL_NS3D || MODEL == HPB_MODEL_NS2D || MODEL == HPB_MODEL_EULER1D) && sI != nullptr) {
    // Euler1D (Euler1DSource.c:90-123): G = S (0, 1, 1) with the one field S = grav_field -- the same two components
    // source function G_j = g_grav_j * (0, d_x, d_y, d_z, 1): only components dir+1 and 4 are non-zero
    const int vm = dir + 1;
    double sLm, sRm, sLe, sRe;
    if (use_char) { sLm = srcL[0]; sRm = srcR[0]; sLe = srcL[1]; sRe = srcR[1]; }
    else {
      sLm = weno_combine(wLF[vm][0], wLF[vm][1], wLF[vm][2], GG[0], GG[1], GG[2], GG[3], GG[4]);
      sRm = weno_combine(wRF[vm][0], wRF[vm][1], wRF[vm][2], GG[5], GG[4], GG[3], GG[2], GG[1]);
      sLe = weno_combine(wLF[NV-1][0], wLF[NV-1][1], wLF[NV-1][2], GG[0], GG[1], GG[2], GG[3], GG[4]);
      sRe = weno_combine(wRF[NV-1][0], wRF[NV-1][1], wRF[NV-1][2], GG[5], GG[4], GG[3], GG[2], GG[1]);
    }
    sI[q]      = 0.5 * (sLm + sRm);
    sI[ni + q] = 0.5 * (sLe + sRe);
  }
}

// flux divergence (HyperbolicFunction.c:94-109): mode 0: out = -dxinv*(fI[p2]-fI[p1]); 1: out -= ...;
// 2: out = +...; 3: out += ... (the positive forms serve hpb_HyperbolicFunction)
__global__ void k_divergence(Geom G, const double* __restrict__ dxinv, const double* __restrict__ fI, int dir,
                             double* __restrict__ out, int mode)
{
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (i0 >= G.N[0]) return;
  const int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  const long long ni = (long long)M0 * M1 * M2;
  const long long q1 = i0 + (long long)M0 * (i1 + (long long)M1 * i2);
  const long long qs = (dir == 0 ? 1 : dir == 1 ? (long long)M0 : (long long)M0 * M1);
  const long long p = cell_index(G, i0, i1, i2);
  const int idx = (dir == 0 ? i0 : dir == 1 ? i1 : i2);
  const double dxi = dxinv[G.xoff[dir] + G.g + idx];
  for (int v = 0; v < G.nvars; v++) {
    const double t = dxi * (fI[v * ni + q1 + qs] - fI[v * ni + q1]);
    const long long a = v * G.npg + p;
    if (mode == 0) out[a] = -t;
    else if (mode == 1) out[a] -= t;
    else if (mode == 2) out[a] = t;
    else out[a] += t;
  }
}

// gravity source (NavierStokes3DSource.c:80-100): src_v += (term_v*f_grav) * (S_I[p2]-S_I[p1]) * dxinv
__global__ void k_ns3d_source(Geom G, Phys ph, const double* __restrict__ dxinv, const double* __restrict__ u,
                              const double* __restrict__ gf, const double* __restrict__ sI, int dir,
                              double* __restrict__ out)
{
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (i0 >= G.N[0]) return;
  const int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  const long long ni = (long long)M0 * M1 * M2;
  const long long q1 = i0 + (long long)M0 * (i1 + (long long)M1 * i2);
  const long long qs = (dir == 0 ? 1 : dir == 1 ? (long long)M0 : (long long)M0 * M1);
  const long long p = cell_index(G, i0, i1, i2);
  const int idx = (dir == 0 ? i0 : dir == 1 ? i1 : i2);
  const double dxi = dxinv[G.xoff[dir] + G.g + idx];
  const double rho = u[p];
  const double vd = (rho == 0 && ph.model != HPB_MODEL_EULER1D) ? 0.0 : u[(1 + dir) * G.npg + p] / rho;
  // Navier-Stokes: term = rho RT (1, v_dir), factor f_grav; Euler1D (Euler1DSource.c:62-66): term = rho (1, v), factor 1 / S
  const bool e1d = (ph.model == HPB_MODEL_EULER1D);
  const double f = e1d ? (1.0 / gf[p]) : gf[p];
  const double tm = e1d ? rho : rho * ph.RT, te = e1d ? rho * vd : rho * ph.RT * vd;
  out[(1 + dir) * G.npg + p] += ((tm * f) * (sI[q1 + qs] - sI[q1]) * dxi);
  out[(long long)(G.nvars - 1) * G.npg + p] += ((te * f) * (sI[ni + q1 + qs] - sI[ni + q1]) * dxi);   // energy: last component
}

// ------------------------------------------------------------------------------------------
// Navier-Stokes viscous terms (NavierStokes3DParabolicFunction.c:50-325, 2-D :38-230).
// phase 1: QD[dir] = dxinv * D_dir(Q), Q = (rho, u, v, w, T), on every line along dir whose transverse
// indices are interior; ghosts along the line included, one-sided at the line ends
// (FirstDerivativeFourthOrder.c:36-129). Locations never written stay 0 (the reference callocs them).
template <int MODEL>
__device__ __forceinline__ void prim_fn(double gamma, const double* __restrict__ u, long long npg, long long p, double* Q)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  double uu[NV], rho, vel[3], e, P;
#pragma unroll
  for (int v = 0; v < NV; v++) uu[v] = u[v * npg + p];
  flowvar<MODEL>(uu, gamma, rho, vel, e, P);
  Q[0] = rho;
#pragma unroll
  for (int k = 0; k < NV - 2; k++) Q[1 + k] = vel[k];
  Q[NV - 1] = gamma * P / rho;
}

__device__ __forceinline__ double d1_coeffs(int i, int N, int g, const double* f /* f[-4..4] centred */)
{
  const double one_twelve = 1.0 / 12.0;
  if (i == -g)            return (-25*f[0]+48*f[1]-36*f[2]+16*f[3]-3*f[4])*one_twelve;
  else if (i == -g + 1)   return (-3*f[-1]-10*f[0]+18*f[1]-6*f[2]+f[3])*one_twelve;
  else if (i < N + g - 2) return (f[-2]-8*f[-1]+8*f[1]-f[2])*one_twelve;
  else if (i == N + g - 2)return (-f[-3]+6*f[-2]-18*f[-1]+10*f[0]+3*f[1])*one_twelve;
  else                    return (3*f[-4]-16*f[-3]+36*f[-2]-48*f[-1]+25*f[0])*one_twelve;
}

template <int MODEL>
__global__ void k_ns_qderiv(Geom G, double gamma, const double* __restrict__ dxinv, const double* __restrict__ u,
                            int dir, double* __restrict__ QD)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  // thread box: full (with ghosts) along dir, interior along the others
  int B[3] = { G.N[0], G.N[1], G.N[2] };
  B[dir] = G.P[dir];
  int t[3] = { (int)(blockIdx.x * blockDim.x + threadIdx.x), (int)blockIdx.y, (int)blockIdx.z };
  if (t[0] >= B[0]) return;
  int ii[3] = { t[0], t[1], t[2] };
  ii[dir] -= G.g;
  const int i = ii[dir], N = G.N[dir], g = G.g;
  const long long p = cell_index(G, ii[0], ii[1], ii[2]);
  const long long st = G.st[dir];
  // stencil range needed
  int lo, hi;
  if (i == -g) { lo = 0; hi = 4; } else if (i == -g + 1) { lo = -1; hi = 3; }
  else if (i < N + g - 2) { lo = -2; hi = 2; } else if (i == N + g - 2) { lo = -3; hi = 1; } else { lo = -4; hi = 0; }
  double f[NV][9];
  for (int k = lo; k <= hi; k++) {
    double Q[NV];
    prim_fn<MODEL>(gamma, u, G.npg, p + k * st, Q);
#pragma unroll
    for (int v = 0; v < NV; v++) f[v][k + 4] = Q[v];
  }
  // un-scaled, as FirstDerivativePar leaves it: the reference exchanges the halos first and THEN scales every
  // point, ghosts included, by the LOCAL dxinv (NavierStokes3DParabolicFunction.c:125-146); on a periodic wrap
  // the receiver's ghost dxinv is a copy of its first interior value, not the sender's -- the scaling is applied
  // where the derivatives are consumed (fviscous_fn)
  (void)dxinv;
#pragma unroll
  for (int v = 0; v < NV; v++) QD[v * G.npg + p] = d1_coeffs(i, N, g, &f[v][4]);
}

// viscous flux of direction dir at one point (NavierStokes3DParabolicFunction.c:152-195 etc.)
template <int MODEL>
__device__ __forceinline__ void fviscous_fn(const Phys& ph, const Geom& G, const double* __restrict__ u,
                                            const double* __restrict__ QDx, const double* __restrict__ QDy,
                                            const double* __restrict__ QDz, long long p, int dir,
                                            double dxi, double dyi, double dzi, double* FV)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  const double two_third = 2.0 / 3.0;
  const double inv_gamma_m1 = 1.0 / (ph.gamma - 1.0), inv_Re = 1.0 / ph.Re, inv_Pr = 1.0 / ph.Pr;
  double Q[NV];
  prim_fn<MODEL>(ph.gamma, u, G.npg, p, Q);
  const double T = Q[NV - 1];
  const double mu = exp(0.76 * log(T));            // raiseto(T, 0.76), math_ops.h:37
  const long long n = G.npg;
  if (MODEL == HPB_MODEL_NS3D) {
    const double uvel = Q[1], vvel = Q[2], wvel = Q[3];
    const double ux = QDx[1*n+p] * dxi, vx = QDx[2*n+p] * dxi, wx = QDx[3*n+p] * dxi;
    const double uy = QDy[1*n+p] * dyi, vy = QDy[2*n+p] * dyi, wy = QDy[3*n+p] * dyi;
    const double uz = QDz[1*n+p] * dzi, vz = QDz[2*n+p] * dzi, wz = QDz[3*n+p] * dzi;
    double t1, t2, t3, q;
    if (dir == 0) {
      t1 = two_third * (mu*inv_Re) * (2*ux - vy - wz);
      t2 = (mu*inv_Re) * (uy + vx);
      t3 = (mu*inv_Re) * (uz + wx);
      q  = ( mu*inv_Re * inv_gamma_m1 * inv_Pr ) * (QDx[4*n+p] * dxi);
    } else if (dir == 1) {
      t1 = (mu*inv_Re) * (uy + vx);
      t2 = two_third * (mu*inv_Re) * (-ux + 2*vy - wz);
      t3 = (mu*inv_Re) * (vz + wy);
      q  = ( mu*inv_Re * inv_gamma_m1 * inv_Pr ) * (QDy[4*n+p] * dyi);
    } else {
      t1 = (mu*inv_Re) * (uz + wx);
      t2 = (mu*inv_Re) * (vz + wy);
      t3 = two_third * (mu*inv_Re) * (-ux - vy + 2*wz);
      q  = ( mu*inv_Re * inv_gamma_m1 * inv_Pr ) * (QDz[4*n+p] * dzi);
    }
    FV[0] = t1; FV[1] = t2; FV[2] = t3; FV[3] = uvel*t1 + vvel*t2 + wvel*t3 + q;
  } else {
    const double uvel = Q[1], vvel = Q[2];
    const double ux = QDx[1*n+p] * dxi, vx = QDx[2*n+p] * dxi;
    const double uy = QDy[1*n+p] * dyi, vy = QDy[2*n+p] * dyi;
    double t1, t2, q;
    if (dir == 0) {
      t1 = two_third * (mu*inv_Re) * (2*ux - vy);
      t2 = (mu*inv_Re) * (uy + vx);
      q  = ( (mu*inv_Re) * inv_gamma_m1 * inv_Pr ) * (QDx[3*n+p] * dxi);
    } else {
      t1 = (mu*inv_Re) * (uy + vx);
      t2 = two_third * (mu*inv_Re) * (-ux + 2*vy);
      q  = ( (mu*inv_Re) * inv_gamma_m1 * inv_Pr ) * (QDy[3*n+p] * dyi);
    }
    FV[0] = t1; FV[1] = t2; FV[2] = uvel*t1 + vvel*t2 + q;
  }
}

// phase 2a: FV (components 1..NV-1; component 0 is identically zero) for direction dir at the points
// the interior derivative needs: interior transverse, [-2, N+2) along dir
template <int MODEL>
__global__ void k_ns_fviscous(Geom G, Phys ph, const double* __restrict__ dxinv, const double* __restrict__ u,
                              const double* __restrict__ QDx, const double* __restrict__ QDy,
                              const double* __restrict__ QDz, int dir, double* __restrict__ FV)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  int B[3] = { G.N[0], G.N[1], G.N[2] };
  B[dir] += 4;
  int t[3] = { (int)(blockIdx.x * blockDim.x + threadIdx.x), (int)blockIdx.y, (int)blockIdx.z };
  if (t[0] >= B[0]) return;
  int ii[3] = { t[0], t[1], t[2] };
  ii[dir] -= 2;
  const long long p = cell_index(G, ii[0], ii[1], ii[2]);
  double F[NV - 1];
  const double dxi = dxinv[G.xoff[0] + G.g + ii[0]];
  const double dyi = dxinv[G.xoff[1] + G.g + ii[1]];
  const double dzi = (G.ndims > 2) ? dxinv[G.xoff[2] + G.g + ii[2]] : 1.0;
  fviscous_fn<MODEL>(ph, G, u, QDx, QDy, QDz, p, dir, dxi, dyi, dzi, F);
#pragma unroll
  for (int v = 0; v < NV - 1; v++) FV[v * G.npg + p] = F[v];
}

// phase 2b: out += dxinv * D_dir(FV) at interior points (central stencil only is reached there)
__global__ void k_ns_par_accum(Geom G, const double* __restrict__ dxinv, const double* __restrict__ FV, int dir,
                               double* __restrict__ out)
{
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (i0 >= G.N[0]) return;
  const long long p = cell_index(G, i0, i1, i2);
  const long long st = G.st[dir];
  const int idx = (dir == 0 ? i0 : dir == 1 ? i1 : i2);
  const double dxi = dxinv[G.xoff[dir] + G.g + idx];
  const double one_twelve = 1.0 / 12.0;
  for (int v = 1; v < G.nvars; v++) {
    const double* f = FV + (v - 1) * G.npg + p;
    const double d = (f[-2*st] - 8*f[-st] + 8*f[st] - f[2*st]) * one_twelve;
    out[v * G.npg + p] += (dxi * d);
  }
}

// LinearADR diffusion through ParabolicFunctionNC1Stage.c: out += dxinv^2 * D2_d(nu_d u)
__global__ void k_linadr_par(Geom G, Phys ph, const double* __restrict__ dxinv, const double* __restrict__ u,
                             double* __restrict__ out)
{
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (i0 >= G.N[0]) return;
  const long long p = cell_index(G, i0, i1, i2);
  const int idx[3] = { i0, i1, i2 };
  const double one_twelve = 1.0 / 12.0;
  for (int v = 0; v < G.nvars; v++) {
    double par = 0.0;
    for (int d = 0; d < G.ndims; d++) {
      const double nu = ph.diff[G.nvars * d + v];
      const double* f = u + v * G.npg + p;
      const long long st = G.st[d];
      double d2;
      if (ph.par_scheme == 2) d2 = nu*f[-st] - 2*(nu*f[0]) + nu*f[st];
      else d2 = (-(nu*f[-2*st]) + 16*(nu*f[-st]) - 30*(nu*f[0]) + 16*(nu*f[st]) - (nu*f[2*st])) * one_twelve;
      const double dxi = dxinv[G.xoff[d] + G.g + idx[d]];
      par += (dxi * dxi * d2);
    }
    out[v * G.npg + p] += par;
  }
}

// ------------------------------------------------------------------------------------------
// RK stage vector and step completion (TimeRK.c:126-195), over the ghost-padded length like the
// reference. Products are rounded before the add (no FMA contraction) so the update carries the
// reference's own rounding.
struct RKArgs { const double* k[4]; double a[4]; int n; };

__global__ void k_rk_combine(const double* __restrict__ u, RKArgs args, double* __restrict__ out, long long n)
{
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double t = u[i];
    for (int s = 0; s < args.n; s++) t = __dadd_rn(t, __dmul_rn(args.a[s], args.k[s][i]));
    out[i] = t;
  }
}

// GLM-GEE (TimeGLMGEE.c:66-80, 116-129): out = c0 u ; out += c1 aux ; out += a_i k_i in order -- _ArrayScaleCopy1D_
// followed by _ArrayAXPY_s, every product and sum rounded on its own like the reference's loops
struct GLMArgs { const double* k[HPB_MAX_STAGES]; double a[HPB_MAX_STAGES]; int n; double c0, c1; };
__global__ void k_glm_combine(const double* __restrict__ u, const double* __restrict__ aux, GLMArgs args,
                              double* __restrict__ out, long long n)
{
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double t = __dmul_rn(args.c0, u[i]);
    t = __dadd_rn(t, __dmul_rn(args.c1, aux[i]));
    for (int s = 0; s < args.n; s++) t = __dadd_rn(t, __dmul_rn(args.a[s], args.k[s][i]));
    out[i] = t;
  }
}
__global__ void k_swap(double* __restrict__ a, double* __restrict__ b, long long n)
{
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { const double t = a[i]; a[i] = b[i]; b[i] = t; }
}
// TimeError.c:47-52, 87-89: est = aux (yeps) or (u - aux) * (1/(1-gamma)) (yyt); dif = (1.0 u + (-1.0) uex) + (-1.0) est
__global__ void k_glm_error_fields(const double* __restrict__ u, const double* __restrict__ aux, const double* __restrict__ uex,
                                   int mode, double f, double* __restrict__ est, double* __restrict__ dif, long long n)
{
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const double e = (mode == HPB_GLM_YEPS) ? aux[i] : __dmul_rn(__dadd_rn(u[i], -aux[i]), f);
    est[i] = e;
    if (uex) dif[i] = __dadd_rn(__dadd_rn(__dmul_rn(1.0, u[i]), __dmul_rn(-1.0, uex[i])), __dmul_rn(-1.0, e));
    else dif[i] = 0.0;
  }
}

// exact path: rhs = (rhs + par) + src, the summation order of TimeRHSFunctionExplicit.c:89-92 (rhs holds -hyp)
__global__ void k_combine_rhs(double* __restrict__ rhs, const double* __restrict__ par, const double* __restrict__ src, long long n)
{
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double t = rhs[i];
    if (par) t = __dadd_rn(t, par[i]);
    if (src) t = __dadd_rn(t, src[i]);
    rhs[i] = t;
  }
}

__global__ void k_copy(double* __restrict__ dst, const double* __restrict__ src, long long n)
{
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------
// reductions: max CFL (NavierStokes3DComputeCFL.c:16-49 and the 2-D/1-D/LinearADR twins) and the
// interior sum of squares of (a - b) (TimePostStep.c:44-63)
__device__ __forceinline__ double block_reduce(double v, bool is_max)
{
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_down_sync(0xffffffffu, v, o);
    v = is_max ? fmax(v, other) : v + other;
  }
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = (lane < nw) ? sh[lane] : (is_max ? 0.0 : 0.0);
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_down_sync(0xffffffffu, v, o);
      v = is_max ? fmax(v, other) : v + other;
    }
  }
  return v;
}

__device__ __forceinline__ void atomic_max_double(double* addr, double val)
{
  // values are non-negative: the bit patterns order like the doubles
  atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(val));
}

template <int MODEL>
__global__ void k_cfl(Geom G, Phys ph, const double* __restrict__ dxinv, const double* __restrict__ u, double dt,
                      double* __restrict__ out)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  double m = 0.0;
  if (i0 < G.N[0]) {
    const int idx[3] = { i0, i1, i2 };
    if (MODEL == HPB_MODEL_LINEAR_ADR && ph.advf != nullptr) {
      // LinearADRComputeCFL.c:41-58 (sic: the spacing of dimension 0 for every direction)
      const long long p = cell_index(G, i0, i1, i2);
      const double dxi = dxinv[G.xoff[0] + G.g + idx[0]];
      for (int d = 0; d < G.ndims; d++) {
        const double c = ph.advf[d * ph.advf_npg + p] * dt * dxi;
        if (c > m) m = c;
      }
    } else if (MODEL == HPB_MODEL_LINEAR_ADR) {
      for (int d = 0; d < G.ndims; d++) {
        const double c = ph.adv[G.nvars * d] * dt * dxinv[G.xoff[d] + G.g + idx[d]];
        if (c > m) m = c;
      }
    } else if (MODEL == HPB_MODEL_BURGERS) {      // BurgersComputeCFL.c:14-48: u dt / dx (sic: no absolute value)
      const double uu = u[cell_index(G, i0, i1, i2)];
      for (int d = 0; d < G.ndims; d++) {
        const double c = uu * dt * dxinv[G.xoff[d] + G.g + idx[d]];
        if (c > m) m = c;
      }
    } else {
      const long long p = cell_index(G, i0, i1, i2);
      double uu[NV], rho, vel[3], e, P;
#pragma unroll
      for (int v = 0; v < NV; v++) uu[v] = u[v * G.npg + p];
      flowvar<MODEL>(uu, ph.gamma, rho, vel, e, P);
      const double c = sqrt(ph.gamma * P / rho);
      for (int d = 0; d < NV - 2; d++) {
        const double l = (fabs(vel[d]) + c) * dt * dxinv[G.xoff[d] + G.g + idx[d]];
        if (l > m) m = l;
      }
    }
  }
  m = block_reduce(m, true);
  if (threadIdx.x == 0) atomic_max_double(out, m);
}

__global__ void k_sumsq_diff(Geom G, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out)
{
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  double s = 0.0;
  if (i0 < G.N[0]) {
    const long long p = cell_index(G, i0, i1, i2);
    for (int v = 0; v < G.nvars; v++) { const double d = a[v * G.npg + p] - b[v * G.npg + p]; s += d * d; }
  }
  s = block_reduce(s, false);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

// ------------------------------------------------------------------------------------------
// conservation / error diagnostics. All reductions are DETERMINISTIC: a fixed launch
// shape, every thread walks its elements in a fixed order, fixed shuffle trees, no atomics -- two runs give the
// same bits. (The reference sums serially, dim 0 fastest; a parallel sum differs from it by rounding only.)
constexpr int DIAG_TPB = 256;
constexpr int DIAG_BLOCKS = 148 * 4;

// StageBoundaryIntegral[(2d+f)*nvars+v] = -/+ sum of the interface flux over the block's low / high face
// (HyperbolicFunction.c:103-106). face_fI: compact face array written by k_iface in face mode. One block per v.
__global__ void __launch_bounds__(1024)
k_face_sum(const double* __restrict__ face_fI, long long nface, double sign, double* __restrict__ out)
{
  const int v = blockIdx.x;
  double s = 0.0;
  for (long long i = threadIdx.x; i < nface; i += blockDim.x) s += face_fI[v * nface + i];
  s = block_reduce(s, false);
  if (threadIdx.x == 0) out[v] = sign * s;
}

// StepBoundaryIntegral = sum_s (dt b_s) BoundaryFlux[s]   (TimeRK.c:190-193, zeroed by TimePreStep.c:117)
__global__ void k_step_boundary_integral(const double* __restrict__ bf, int ns, int nbf, RKTableau rk, double dt,
                                         double* __restrict__ step_bi)
{
  const int k = threadIdx.x;
  if (k >= nbf) return;
  double acc = 0.0;
  for (int s = 0; s < ns; s++) acc += (dt * rk.b[s]) * bf[s * nbf + k];
  step_bi[k] = acc;
}

// per-block partial results over the interior points, part[ch*nblk + block].
// MODE 0: channel v = sum a[v,p] * dV(p), dV = prod_d (1/dxinv_d)            (VolumeIntegral.c:33-41)
// MODE 1: channels (sum |a-b|, sum (a-b)^2, max |a-b|) over all components; b == nullptr: norms of a itself
//         (ArraySumAbsnD / ArraySumSquarenD / ArrayMaxnD, arrayfunctions.h:486-545; CalculateError.c:68-104)
template <int MODE>
__global__ void __launch_bounds__(DIAG_TPB)
k_diag_partial(Geom G, const double* __restrict__ dxinv, const double* __restrict__ a, const double* __restrict__ b,
               double* __restrict__ part)
{
  const long long nint = (long long)G.N[0] * G.N[1] * G.N[2];
  double acc[HPB_MAX_NVARS] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nint; q += (long long)gridDim.x * blockDim.x) {
    const int i0 = (int)(q % G.N[0]), i1 = (int)((q / G.N[0]) % G.N[1]), i2 = (int)(q / ((long long)G.N[0] * G.N[1]));
    const long long p = cell_index(G, i0, i1, i2);
    if (MODE == 0) {
      const int idx[3] = { i0, i1, i2 };
      double dV = 1.0;
      for (int d = 0; d < G.ndims; d++) dV *= (1.0 / dxinv[G.xoff[d] + G.g + idx[d]]);
      for (int v = 0; v < G.nvars; v++) acc[v] += a[v * G.npg + p] * dV;
    } else {
      for (int v = 0; v < G.nvars; v++) {
        double e = a[v * G.npg + p];
        if (b) e -= b[v * G.npg + p];
        const double m = fabs(e);
        acc[0] += m; acc[1] += e * e; if (m > acc[2]) acc[2] = m;
      }
    }
  }
  const int nch = (MODE == 0) ? G.nvars : 3;
  for (int c = 0; c < nch; c++) {
    const double r = block_reduce(acc[c], MODE == 1 && c == 2);
    if (threadIdx.x == 0) part[c * gridDim.x + blockIdx.x] = r;
    __syncthreads();
  }
}

// TimePostStep.c:44-63: sum over the interior of (u^{n+1} - u^n)^2, evaluated from the stage right-hand sides the step
// has just combined, u^{n+1} - u^n = sum_s (dt b_s) k_s -- no copy of u^n is kept (2 passes over memory per step
// saved). It differs from the reference's difference of the two stored solutions by the rounding of the update
// (relative 1e-16 |u| / |u^{n+1} - u^n| per point).
__global__ void __launch_bounds__(DIAG_TPB)
k_step_norm_partial(Geom G, RKArgs args, double* __restrict__ part)
{
  const long long nint = (long long)G.N[0] * G.N[1] * G.N[2];
  double acc = 0.0;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nint; q += (long long)gridDim.x * blockDim.x) {
    const int i0 = (int)(q % G.N[0]), i1 = (int)((q / G.N[0]) % G.N[1]), i2 = (int)(q / ((long long)G.N[0] * G.N[1]));
    const long long p = cell_index(G, i0, i1, i2);
    for (int v = 0; v < G.nvars; v++) {
      double d = 0.0;
      for (int s = 0; s < args.n; s++) d += args.a[s] * args.k[s][v * G.npg + p];
      acc += d * d;
    }
  }
  const double r = block_reduce(acc, false);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}

// final pass: one block per channel over the nblk partials; channel `max_ch` is a maximum, the others sums
__global__ void __launch_bounds__(DIAG_TPB)
k_diag_final(const double* __restrict__ part, int nblk, int max_ch, double* __restrict__ out)
{
  const int c = blockIdx.x;
  const bool is_max = (c == max_ch);
  double s = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) { const double x = part[c * nblk + i]; s = is_max ? fmax(s, x) : s + x; }
  s = block_reduce(s, is_max);
  if (threadIdx.x == 0) out[c] = s;
}

// ------------------------------------------------------------------------------------------
// fine-grained API kernels (the reference's individual function pointers)
template <int MODEL>
__global__ void k_flux(Geom G, Phys ph, const double* __restrict__ u, int dir, double* __restrict__ f)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= G.npg) return;
  double uu[NV], ff[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) uu[v] = u[v * G.npg + p];
  flux_fn<MODEL>(ph, uu, dir, ff, p);
#pragma unroll
  for (int v = 0; v < NV; v++) f[v * G.npg + p] = ff[v];
}

template <int MODEL>
__global__ void k_modified(Geom G, Phys ph, const double* __restrict__ u, const double* __restrict__ gf,
                           const double* __restrict__ gg, double* __restrict__ uC)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= G.npg) return;
  double uu[NV], cc[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) uu[v] = u[v * G.npg + p];
  modified_fn<MODEL>(ph, uu, gf ? gf[p] : 1.0, gg ? gg[p] : 1.0, cc);
#pragma unroll
  for (int v = 0; v < NV; v++) uC[v * G.npg + p] = cc[v];
}

// SetInterpLimiterVar: the 12 weight arrays of one direction, w[(3*blk+k)*ni*NV + v*ni + q],
// blk = LF, LU, RF, RU (WENOFifthOrderCalculateWeights.c:147-158)
template <int MODEL>
__global__ void k_weights(Geom G, Phys ph, const double* __restrict__ fC, const double* __restrict__ u, int dir,
                          double* __restrict__ w)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  const int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (i0 >= M0) return;
  // optimal weights: WENO5 (0.1,0.6,0.3); CRWENO5 (0.2,0.5,0.3) except on the two physical-boundary interfaces
  // (WENOFifthOrderCalculateWeights.c:205-225: interface 0 of the first block and interface N of the last one along dir)
  double c1 = 0.1, c2 = 0.6, c3 = 0.3;
  {
    const int iI = (dir == 0 ? i0 : dir == 1 ? i1 : i2);
    const bool bnd = (iI == 0 && G.lo_phys[dir]) || (iI == G.N[dir] && G.hi_phys[dir]);
    if (ph.scheme == HPB_SCHEME_CRWENO5 && !bnd) { c1 = 0.2; c2 = 0.5; c3 = 0.3; }
  }
  const long long q = i0 + (long long)M0 * (i1 + (long long)M1 * i2);
  const long long ni = (long long)M0 * M1 * M2;
  const long long st = G.st[dir];
  const long long pm1 = cell_index(G, i0, i1, i2) - st;
  double U[6][NV], F[6][NV];
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const long long p = pm1 + (k - 2) * st;
#pragma unroll
    for (int v = 0; v < NV; v++) { U[k][v] = u[v * G.npg + p]; F[k][v] = fC[v * G.npg + p]; }
  }
  const bool use_char = ph.interp_char && (MODEL == HPB_MODEL_EULER1D || MODEL == HPB_MODEL_NS2D || MODEL == HPB_MODEL_NS3D);
  double L[NV * NV];
  if (use_char) {
    double uavg[NV], lam[NV], R[NV * NV];
    roe_average<MODEL>(ph, U[2], U[3], uavg);
    eigen<MODEL>(ph, uavg, dir, lam, L, R);
  }
#pragma unroll
  for (int v = 0; v < NV; v++) {
    double cF[6], cU[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
      if (use_char) {
        double sF = 0.0, sU = 0.0;
#pragma unroll
        for (int j = 0; j < NV; j++) { sF += L[v * NV + j] * F[k][j]; sU += L[v * NV + j] * U[k][j]; }
        cF[k] = sF; cU[k] = sU;
      } else { cF[k] = F[k][v]; cU[k] = U[k][v]; }
    }
    double ws[4][3];
    if (ph.no_limiting) {
      for (int b = 0; b < 4; b++) { ws[b][0] = c1; ws[b][1] = c2; ws[b][2] = c3; }
    } else {
      weno_weights_ref_c(ph.weno, ph.eps, c1, c2, c3, cF[0], cF[1], cF[2], cF[3], cF[4], ws[0][0], ws[0][1], ws[0][2]);
      weno_weights_ref_c(ph.weno, ph.eps, c1, c2, c3, cU[0], cU[1], cU[2], cU[3], cU[4], ws[1][0], ws[1][1], ws[1][2]);
      weno_weights_ref_c(ph.weno, ph.eps, c1, c2, c3, cF[5], cF[4], cF[3], cF[2], cF[1], ws[2][0], ws[2][1], ws[2][2]);
      weno_weights_ref_c(ph.weno, ph.eps, c1, c2, c3, cU[5], cU[4], cU[3], cU[2], cU[1], ws[3][0], ws[3][1], ws[3][2]);
    }
    for (int b = 0; b < 4; b++)
      for (int k = 0; k < 3; k++) w[((3 * b + k) * NV + v) * ni + q] = ws[b][k];
  }
}

// InterpolateInterfacesHyp with stored weights (Interp1PrimFifthOrderWENO.c:74, ...Char.c:86)
template <int MODEL>
__global__ void k_interp(Geom G, Phys ph, const double* __restrict__ fC, const double* __restrict__ u,
                         const double* __restrict__ w, int upw, int dir, int uflag, double* __restrict__ fI)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  const int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (i0 >= M0) return;
  const long long q = i0 + (long long)M0 * (i1 + (long long)M1 * i2);
  const long long ni = (long long)M0 * M1 * M2;
  const long long st = G.st[dir];
  const long long pm1 = cell_index(G, i0, i1, i2) - st;
  const int blk = (upw < 0 ? 2 : 0) + (uflag ? 1 : 0);
  // stencil (m3,m2,m1,p1,p2): left-biased = cells i-3..i+1, right-biased = i+2..i-2
  long long ps[5];
  for (int k = 0; k < 5; k++) ps[k] = (upw > 0) ? pm1 + (k - 2) * st : pm1 + (3 - k) * st;
  double S[5][NV];
#pragma unroll
  for (int k = 0; k < 5; k++)
#pragma unroll
    for (int v = 0; v < NV; v++) S[k][v] = fC[v * G.npg + ps[k]];
  const bool use_char = ph.interp_char && (MODEL == HPB_MODEL_EULER1D || MODEL == HPB_MODEL_NS2D || MODEL == HPB_MODEL_NS3D);
  double out[NV];
  if (!use_char) {
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const double w1 = w[((3 * blk + 0) * NV + v) * ni + q], w2 = w[((3 * blk + 1) * NV + v) * ni + q],
                   w3 = w[((3 * blk + 2) * NV + v) * ni + q];
      out[v] = weno_combine(w1, w2, w3, S[0][v], S[1][v], S[2][v], S[3][v], S[4][v]);
    }
  } else {
    double UL[NV], UR[NV], uavg[NV], lam[NV], L[NV * NV], R[NV * NV], fc[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) { UL[v] = u[v * G.npg + pm1]; UR[v] = u[v * G.npg + pm1 + st]; }
    roe_average<MODEL>(ph, UL, UR, uavg);
    eigen<MODEL>(ph, uavg, dir, lam, L, R);
#pragma unroll
    for (int v = 0; v < NV; v++) {
      double c[5];
#pragma unroll
      for (int k = 0; k < 5; k++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NV; j++) s += L[v * NV + j] * S[k][j];
        c[k] = s;
      }
      const double w1 = w[((3 * blk + 0) * NV + v) * ni + q], w2 = w[((3 * blk + 1) * NV + v) * ni + q],
                   w3 = w[((3 * blk + 2) * NV + v) * ni + q];
      fc[v] = weno_combine(w1, w2, w3, c[0], c[1], c[2], c[3], c[4]);
    }
#pragma unroll
    for (int i = 0; i < NV; i++) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < NV; j++) s += R[i * NV + j] * fc[j];
      out[i] = s;
    }
  }
#pragma unroll
  for (int v = 0; v < NV; v++) fI[v * ni + q] = out[v];
}

// ------------------------------------------------------------------------------------------
// Compact schemes, component-wise, one rank along the line:
// Interp1PrimFifthOrderCRWENO.c:80-237, Interp1PrimFifthOrderCompactUpwind.c:73-228. One tridiagonal system per
// grid line and component; rows of the two physical-boundary interfaces are the explicit WENO5 / fifth-order
// upwind value (a = c = 0, b = 1). Row arrays share the interface layout [v*ni + q]; r is solved in place.
// hcweno5 (Interp1PrimFifthOrderHCWENO.c:65-230): rows (sigma/2, 1, sigma/6) and the right-hand side
// sigma fCompact + (1 - sigma) fWENO with the hybridisation parameter sigma of :153-168 (0 on the physical boundary).
__device__ __forceinline__ double hcweno_sigma(const Phys& ph, bool bnd, double fm2, double fm1, double fp1, double fp2)
{
  if (bnd) return 0.0;
#define HC_ABS(a) ((a) < 0 ? -(a) : (a))
  const double cuckoo = (0.9 * ph.hc_rc / (1.0 - 0.9 * ph.hc_rc)) * ph.hc_xi * ph.hc_xi;
  const double df_jm12 = fm1 - fm2, df_jp12 = fp1 - fm1, df_jp32 = fp2 - fp1;
  const double r_j   = (HC_ABS(2 * df_jp12 * df_jm12) + cuckoo) / (df_jp12 * df_jp12 + df_jm12 * df_jm12 + cuckoo);
  const double r_jp1 = (HC_ABS(2 * df_jp32 * df_jp12) + cuckoo) / (df_jp32 * df_jp32 + df_jp12 * df_jp12 + cuckoo);
#undef HC_ABS
  const double r_int = (r_j < r_jp1 ? r_j : r_jp1);
  return ((r_int / ph.hc_rc) < 1.0 ? (r_int / ph.hc_rc) : 1.0);
}

__global__ void k_compact_rows(Geom G, Phys ph, const double* __restrict__ fC, const double* __restrict__ w, int upw,
                               int dir, int uflag, double* __restrict__ A, double* __restrict__ B,
                               double* __restrict__ Cc, double* __restrict__ R)
{
  const double one_third = 1.0 / 3.0, one_sixth = 1.0 / 6.0;
  const double thirteen_by_sixty = 13.0 / 60.0, fortyseven_by_sixty = 47.0 / 60.0, twentyseven_by_sixty = 27.0 / 60.0,
               one_by_twenty = 1.0 / 20.0, one_by_thirty = 1.0 / 30.0, nineteen_by_thirty = 19.0 / 30.0,
               three_by_ten = 3.0 / 10.0, six_by_ten = 6.0 / 10.0, one_by_ten = 1.0 / 10.0;
  const int NV = G.nvars;
  const int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (i0 >= M0) return;
  const long long q = i0 + (long long)M0 * (i1 + (long long)M1 * i2);
  const long long ni = (long long)M0 * M1 * M2;
  const long long st = G.st[dir];
  const long long pm1 = cell_index(G, i0, i1, i2) - st;
  const int blk = (upw < 0 ? 2 : 0) + (uflag ? 1 : 0);
  const int iI = (dir == 0 ? i0 : dir == 1 ? i1 : i2);
  const bool bnd = (iI == 0 && G.lo_phys[dir]) || (iI == G.N[dir] && G.hi_phys[dir]);
  long long ps[5];
  for (int k = 0; k < 5; k++) ps[k] = (upw > 0) ? pm1 + (k - 2) * st : pm1 + (3 - k) * st;
  for (int v = 0; v < NV; v++) {
    const double fm3 = fC[v * G.npg + ps[0]], fm2 = fC[v * G.npg + ps[1]], fm1 = fC[v * G.npg + ps[2]],
                 fp1 = fC[v * G.npg + ps[3]], fp2 = fC[v * G.npg + ps[4]];
    double a, b, c, r;
    if (ph.scheme == HPB_SCHEME_HCWENO5) {      // the candidates are always the WENO5 ones (HCWENO.c:140-143)
      const double one_half = 1.0 / 2.0;
      const double w1 = w[((3 * blk + 0) * NV + v) * ni + q], w2 = w[((3 * blk + 1) * NV + v) * ni + q],
                   w3 = w[((3 * blk + 2) * NV + v) * ni + q];
      const double f1 = (2 * one_sixth) * fm3 - (7.0 * one_sixth) * fm2 + (11.0 * one_sixth) * fm1;
      const double f2 = (-one_sixth) * fm2 + (5.0 * one_sixth) * fm1 + (2 * one_sixth) * fp1;
      const double f3 = (2 * one_sixth) * fm1 + (5 * one_sixth) * fp1 - (one_sixth) * fp2;
      const double sigma = hcweno_sigma(ph, bnd, fm2, fm1, fp1, fp2);
      if (upw > 0) { a = one_half * sigma; b = 1.0; c = one_sixth * sigma; }
      else         { c = one_half * sigma; b = 1.0; a = one_sixth * sigma; }
      const double fWENO = w1 * f1 + w2 * f2 + w3 * f3;
      const double fCompact = one_sixth * (one_third * fm2 + 19.0 * one_third * fm1 + 10.0 * one_third * fp1);
      r = sigma * fCompact + (1.0 - sigma) * fWENO;
    } else if (ph.scheme == HPB_SCHEME_CRWENO5) {
      const double w1 = w[((3 * blk + 0) * NV + v) * ni + q], w2 = w[((3 * blk + 1) * NV + v) * ni + q],
                   w3 = w[((3 * blk + 2) * NV + v) * ni + q];
      double f1, f2, f3;
      if (bnd) {
        f1 = (2 * one_sixth) * fm3 + (-7 * one_sixth) * fm2 + (11 * one_sixth) * fm1;
        f2 = (-one_sixth) * fm2 + (5 * one_sixth) * fm1 + (2 * one_sixth) * fp1;
        f3 = (2 * one_sixth) * fm1 + (5 * one_sixth) * fp1 + (-one_sixth) * fp2;
        a = 0.0; b = 1.0; c = 0.0;
      } else {
        f1 = (one_sixth) * fm2 + (5 * one_sixth) * fm1;
        f2 = (5 * one_sixth) * fm1 + (one_sixth) * fp1;
        f3 = (one_sixth) * fm1 + (5 * one_sixth) * fp1;
        const double lo = (2 * one_third) * w1 + (one_third) * w2;
        const double di = (one_third) * w1 + (2 * one_third) * w2 + (2 * one_third) * w3;
        const double hi = (one_third) * w3;
        if (upw > 0) { a = lo; b = di; c = hi; } else { c = lo; b = di; a = hi; }
      }
      r = w1 * f1 + w2 * f2 + w3 * f3;
    } else {
      if (bnd) {
        a = 0.0; b = 1.0; c = 0.0;
        r = one_by_thirty * fm3 - thirteen_by_sixty * fm2 + fortyseven_by_sixty * fm1 + twentyseven_by_sixty * fp1
          - one_by_twenty * fp2;
      } else {
        if (upw > 0) { a = three_by_ten; b = six_by_ten; c = one_by_ten; }
        else         { c = three_by_ten; b = six_by_ten; a = one_by_ten; }
        r = one_by_thirty * fm2 + nineteen_by_thirty * fm1 + one_third * fp1;
      }
    }
    A[v * ni + q] = a; B[v * ni + q] = b; Cc[v * ni + q] = c; R[v * ni + q] = r;
  }
}

// TridiagLU/tridiagLU.c:84-274 on one rank (stage 1 forward elimination, stage 4 back substitution; stages 2-3 are
// empty): one thread per system = (grid line, component), the line walked sequentially in the reference's operation
// order, so the solution carries the reference's bits. Threads of a warp own neighbouring lines (unit stride for the
// y/z sweeps). The `a*x[0]` products of stage 4 -- the reduced-system coupling of the multi-rank algorithm, zero on
// one rank -- are kept so that signed zeros come out as in the reference. err: set when a pivot is zero.
__global__ void k_tridiag(Geom G, int dir, double* __restrict__ A, double* __restrict__ B, const double* __restrict__ Cc,
                          double* __restrict__ X, int* __restrict__ err)
{
  const int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  const long long ni = (long long)M0 * M1 * M2;
  // the two transverse extents (t0 fastest) and the strides of a line in the interface array
  int T0, T1; long long s0, s1, qs;
  if (dir == 0)      { T0 = M1; T1 = M2; s0 = M0;  s1 = (long long)M0 * M1; qs = 1; }
  else if (dir == 1) { T0 = M0; T1 = M2; s0 = 1;   s1 = (long long)M0 * M1; qs = M0; }
  else               { T0 = M0; T1 = M1; s0 = 1;   s1 = M0;                 qs = (long long)M0 * M1; }
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x, t1 = blockIdx.y, v = blockIdx.z;
  if (t0 >= T0 || t1 >= T1) return;
  const int n = G.N[dir] + 1;
  double* a = A + v * ni + t0 * s0 + t1 * s1;
  double* b = B + v * ni + t0 * s0 + t1 * s1;
  const double* c = Cc + v * ni + t0 * s0 + t1 * s1;
  double* x = X + v * ni + t0 * s0 + t1 * s1;
  for (int i = 1; i < n; i++) {
    const double bm = b[(i - 1) * qs];
    if (bm == 0) { *err = 1; return; }
    const double factor = a[i * qs] / bm;
    b[i * qs] -= factor * c[(i - 1) * qs];
    a[i * qs] = -factor * a[(i - 1) * qs];
    x[i * qs] -= factor * x[(i - 1) * qs];
  }
  const int il = n - 1;
  if (b[il * qs] == 0) { *err = 1; return; }
  x[il * qs] = (x[il * qs] - a[il * qs] * x[0] - c[il * qs] * 0.0) / b[il * qs];
  for (int i = il - 1; i > -1; i--) {
    if (b[i * qs] == 0) { *err = 1; return; }
    x[i * qs] = (x[i * qs] - c[i * qs] * x[(i + 1) * qs] - a[i * qs] * x[0]) / b[i * qs];
  }
}

// ------------------------------------------------------------------------------------------
// The same systems when the grid line is split among ranks: TridiagLU/tridiagLU.c:84-274 with all four stages, the
// reduced system (one row per rank) solved by TridiagLU/tridiagIterJacobi.c:64-238 as the reference does by default
// (tridiagLUInit.c:70-76: jacobi, maxiter 10, atol 1e-12, rtol 1e-10, norm evaluated), and the hand-over of the shared
// interface of Interp1PrimFifthOrderCRWENO.c:206-223. One thread per system; `n` = rows on this rank (N, or N + 1 on
// the last rank of the line); `first` = this is the first rank of the line (tridiagLU's rank == 0). Exchange buffers:
// [k * nsys + sys]. Operation order follows the reference line by line (compiled without FMA contraction).
struct TriLine {
  int T0, T1; long long s0, s1, qs, ni;
};
__device__ __forceinline__ TriLine tri_line(const Geom& G, int dir)
{
  const int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  TriLine L;
  L.ni = (long long)M0 * M1 * M2;
  if (dir == 0)      { L.T0 = M1; L.T1 = M2; L.s0 = M0;  L.s1 = (long long)M0 * M1; L.qs = 1; }
  else if (dir == 1) { L.T0 = M0; L.T1 = M2; L.s0 = 1;   L.s1 = (long long)M0 * M1; L.qs = M0; }
  else               { L.T0 = M0; L.T1 = M1; L.s0 = 1;   L.s1 = M0;                 L.qs = (long long)M0 * M1; }
  return L;
}
// system index and base offset of this thread's system; false when out of range
__device__ __forceinline__ bool tri_sys(const Geom& G, int dir, const TriLine& L, long long& sys, long long& base)
{
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x, t1 = blockIdx.y, v = blockIdx.z;
  if (t0 >= L.T0 || t1 >= L.T1) return false;
  sys = t0 + (long long)L.T0 * (t1 + (long long)L.T1 * v);
  base = v * L.ni + t0 * L.s0 + t1 * L.s1;
  return true;
}

// stage 1 (tridiagLU.c:140-157) + the last row packed for the next rank (:166-171)
__global__ void k_mr_stage1(Geom G, int dir, int n, int first, double* __restrict__ A, double* __restrict__ B, double* __restrict__ Cc,
                            double* __restrict__ X, double* __restrict__ sendrow, long long nsys, int* __restrict__ err)
{
  const TriLine L = tri_line(G, dir);
  long long sys, base;
  if (!tri_sys(G, dir, L, sys, base)) return;
  double *a = A + base, *b = B + base, *c = Cc + base, *x = X + base;
  const long long qs = L.qs;
  for (int i = (first ? 1 : 2); i < n; i++) {
    const double bm = b[(i - 1) * qs];
    if (bm == 0) { *err = 1; return; }
    const double factor = a[i * qs] / bm;
    b[i * qs] -= factor * c[(i - 1) * qs];
    a[i * qs] = -factor * a[(i - 1) * qs];
    x[i * qs] -= factor * x[(i - 1) * qs];
    if (!first) {
      const double f2 = c[0] / b[(i - 1) * qs];
      c[0]  = -f2 * c[(i - 1) * qs];
      b[0] -=  f2 * a[(i - 1) * qs];
      x[0] -=  f2 * x[(i - 1) * qs];
    }
  }
  sendrow[0 * nsys + sys] = a[(n - 1) * qs];
  sendrow[1 * nsys + sys] = b[(n - 1) * qs];
  sendrow[2 * nsys + sys] = c[(n - 1) * qs];
  sendrow[3 * nsys + sys] = x[(n - 1) * qs];
}

// stage 2 (tridiagLU.c:182-202), ranks other than the first; then the start of the Jacobi iteration on the reduced row
// (tridiagIterJacobi.c:97-113: diagonal check, rhs saved, initial guess x = rhs / b). red = [rhs | x | sendL=sendR | recvL | recvR]
__global__ void k_mr_stage2(Geom G, int dir, int n, int first, double atol, double* __restrict__ A, double* __restrict__ B,
                            double* __restrict__ Cc, double* __restrict__ X, const double* __restrict__ recvrow,
                            double* __restrict__ red, long long nsys, int* __restrict__ err)
{
  const TriLine L = tri_line(G, dir);
  long long sys, base;
  if (!tri_sys(G, dir, L, sys, base)) return;
  if (first) {                       // tridiagLU.c:218: the first rank enters the reduced system with (0, 1, 0 | 0)
    red[0 * nsys + sys] = 0.0; red[1 * nsys + sys] = 0.0 / 1.0; red[2 * nsys + sys] = 0.0 / 1.0;
    return;
  }
  double *a = A + base, *b = B + base, *c = Cc + base, *x = X + base;
  const long long qs = L.qs;
  const double am1 = recvrow[0 * nsys + sys], bm1 = recvrow[1 * nsys + sys], cm1 = recvrow[2 * nsys + sys], xm1 = recvrow[3 * nsys + sys];
  if (bm1 == 0) { *err = 1; return; }
  double factor = a[0] / bm1;
  b[0] -= factor * cm1;
  a[0]  = -factor * am1;
  x[0] -= factor * xm1;
  if (b[(n - 1) * qs] == 0) { *err = 1; return; }
  factor = c[0] / b[(n - 1) * qs];
  b[0] -= factor * a[(n - 1) * qs];
  c[0]  = -factor * c[(n - 1) * qs];
  x[0] -= factor * x[(n - 1) * qs];
  if (b[0] * b[0] < atol * atol) { *err = 2; return; }
  red[0 * nsys + sys] = x[0];            // rhs
  x[0] /= b[0];                          // initial guess
  red[1 * nsys + sys] = x[0];
  red[2 * nsys + sys] = x[0];            // what the neighbours receive
}

// one Jacobi iteration on the reduced row (tridiagIterJacobi.c:150-160 norm, :186 update): pass 0 = this rank's part of
// the squared residual norm (summed over the systems by a fixed tree), pass 1 = the correction
__global__ void k_mr_jacobi(Geom G, int dir, int first, int pass, const double* __restrict__ A, const double* __restrict__ B,
                            const double* __restrict__ Cc, double* __restrict__ X, double* __restrict__ red, long long nsys,
                            double* __restrict__ part)
{
  const TriLine L = tri_line(G, dir);
  long long sys, base;
  double contrib = 0.0;
  if (tri_sys(G, dir, L, sys, base) && !first) {
    const double a = A[base], b = B[base], c = Cc[base];
    const double rhs = red[0 * nsys + sys], rl = red[3 * nsys + sys], rr = red[4 * nsys + sys];
    double* x = X + base;
    if (pass == 0) {
      const double r = a * rl + b * x[0] + c * rr - rhs;
      contrib = r * r;
    } else {
      x[0] = (rhs - a * rl - c * rr) / b;
      red[1 * nsys + sys] = x[0];
      red[2 * nsys + sys] = x[0];
    }
  }
  if (pass == 0) {
    const double v = block_reduce(contrib, false);
    if (threadIdx.x == 0) part[blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z)] = v;
  }
}
__global__ void k_mr_sum(const double* __restrict__ part, long long n, double* __restrict__ out)
{
  double v = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
  v = block_reduce(v, false);
  if (threadIdx.x == 0) out[0] = v;
}

// stage 4 (tridiagLU.c:244-257): xp1 = the next rank's reduced solution (0 on the last rank); then the first-interface
// solution packed for the previous rank (Interp1PrimFifthOrderCRWENO.c:214)
__global__ void k_mr_stage4(Geom G, int dir, int n, int first, const double* __restrict__ A, const double* __restrict__ B,
                            const double* __restrict__ Cc, double* __restrict__ X, const double* __restrict__ xp1,
                            double* __restrict__ sendfirst, long long nsys, int* __restrict__ err)
{
  const TriLine L = tri_line(G, dir);
  long long sys, base;
  if (!tri_sys(G, dir, L, sys, base)) return;
  const double *a = A + base, *b = B + base, *c = Cc + base;
  double* x = X + base;
  const long long qs = L.qs;
  const int il = n - 1;
  if (b[il * qs] == 0) { *err = 1; return; }
  x[il * qs] = (x[il * qs] - a[il * qs] * x[0] - c[il * qs] * xp1[sys]) / b[il * qs];
  for (int i = il - 1; i > (first ? 0 : 1) - 1; i--) {
    if (b[i * qs] == 0) { *err = 1; return; }
    x[i * qs] = (x[i * qs] - c[i * qs] * x[(i + 1) * qs] - a[i * qs] * x[0]) / b[i * qs];
  }
  sendfirst[sys] = x[0];
}
// the solution of the shared interface N arrives from the next rank (Interp1PrimFifthOrderCRWENO.c:222)
__global__ void k_mr_put_last(Geom G, int dir, double* __restrict__ X, const double* __restrict__ recvfirst)
{
  const TriLine L = tri_line(G, dir);
  long long sys, base;
  if (!tri_sys(G, dir, L, sys, base)) return;
  X[base + (long long)G.N[dir] * L.qs] = recvfirst[sys];
}

// ------------------------------------------------------------------------------------------
// Compact schemes on characteristic variables (Interp1PrimFifthOrderCRWENOChar.c:95-277,
// Interp1PrimFifthOrderCompactUpwindChar.c:85-264): one BLOCK tridiagonal system per grid line. Row blocks are
// (coefficient) x L(uavg), the right-hand side is the characteristic candidate combination, the unknown is the
// interface value itself. Block storage: element e of block i of system `sys` at [(i*NV*NV + e)*Nsys + sys], vector
// component at [(i*NV + v)*Nsys + sys] (neighbouring threads = neighbouring systems). sys = t0 + T0*t1 over the two
// transverse indices in increasing dimension order.
__device__ __forceinline__ void compact_sys_index(int dir, int i0, int i1, int i2, int M0, int M1, int M2, int& iI, long long& sys, long long& Nsys)
{
  if (dir == 0)      { iI = i0; sys = i1 + (long long)M1 * i2; Nsys = (long long)M1 * M2; }
  else if (dir == 1) { iI = i1; sys = i0 + (long long)M0 * i2; Nsys = (long long)M0 * M2; }
  else               { iI = i2; sys = i0 + (long long)M0 * i1; Nsys = (long long)M0 * M1; }
}

template <int MODEL>
__global__ void k_compact_rows_char(Geom G, Phys ph, const double* __restrict__ fC, const double* __restrict__ u,
                                    const double* __restrict__ w, int upw, int dir, int uflag, double* __restrict__ A,
                                    double* __restrict__ B, double* __restrict__ Cc, double* __restrict__ F)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  const double one_third = 1.0 / 3.0, one_sixth = 1.0 / 6.0;
  const double thirteen_by_sixty = 13.0 / 60.0, fortyseven_by_sixty = 47.0 / 60.0, twentyseven_by_sixty = 27.0 / 60.0,
               one_by_twenty = 1.0 / 20.0, one_by_thirty = 1.0 / 30.0, nineteen_by_thirty = 19.0 / 30.0,
               three_by_ten = 3.0 / 10.0, six_by_ten = 6.0 / 10.0, one_by_ten = 1.0 / 10.0;
  const int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (i0 >= M0) return;
  const long long q = i0 + (long long)M0 * (i1 + (long long)M1 * i2);
  const long long ni = (long long)M0 * M1 * M2;
  const long long st = G.st[dir];
  const long long pm1 = cell_index(G, i0, i1, i2) - st;
  const int blk = (upw < 0 ? 2 : 0) + (uflag ? 1 : 0);
  int iI; long long sys, Nsys;
  compact_sys_index(dir, i0, i1, i2, M0, M1, M2, iI, sys, Nsys);
  const bool bnd = (iI == 0 && G.lo_phys[dir]) || (iI == G.N[dir] && G.hi_phys[dir]);
  long long ps[5];
  for (int k = 0; k < 5; k++) ps[k] = (upw > 0) ? pm1 + (k - 2) * st : pm1 + (3 - k) * st;
  double UL[NV], UR[NV], uavg[NV], lam[NV], L[NV * NV], R[NV * NV];
#pragma unroll
  for (int v = 0; v < NV; v++) { UL[v] = u[v * G.npg + pm1]; UR[v] = u[v * G.npg + pm1 + st]; }
  roe_average<MODEL>(ph, UL, UR, uavg);
  eigen<MODEL>(ph, uavg, dir, lam, L, R);
  double S[5][NV];
#pragma unroll
  for (int k = 0; k < 5; k++)
#pragma unroll
    for (int v = 0; v < NV; v++) S[k][v] = fC[v * G.npg + ps[k]];
#pragma unroll
  for (int v = 0; v < NV; v++) {
    double c[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < NV; j++) s += L[v * NV + j] * S[k][j];
      c[k] = s;
    }
    const double fm3 = c[0], fm2 = c[1], fm1 = c[2], fp1 = c[3], fp2 = c[4];
    double lo, di, hi, f;
    if (ph.scheme == HPB_SCHEME_HCWENO5) {      // Interp1PrimFifthOrderHCWENOChar.c:171-230
      const double one_half = 1.0 / 2.0;
      const double w1 = w[((3 * blk + 0) * NV + v) * ni + q], w2 = w[((3 * blk + 1) * NV + v) * ni + q],
                   w3 = w[((3 * blk + 2) * NV + v) * ni + q];
      double f1, f2, f3;
      if (bnd) {
        f1 = (2 * one_sixth) * fm3 - (7.0 * one_sixth) * fm2 + (11.0 * one_sixth) * fm1;
        f2 = (-one_sixth) * fm2 + (5.0 * one_sixth) * fm1 + (2 * one_sixth) * fp1;
        f3 = (2 * one_sixth) * fm1 + (5 * one_sixth) * fp1 - (one_sixth) * fp2;
      } else {
        f1 = (one_sixth) * (fm2 + 5 * fm1);
        f2 = (one_sixth) * (5 * fm1 + fp1);
        f3 = (one_sixth) * (fm1 + 5 * fp1);
      }
      const double sigma = hcweno_sigma(ph, bnd, fm2, fm1, fp1, fp2);
      const double fWENO = w1 * f1 + w2 * f2 + w3 * f3;
      const double fCompact = one_sixth * (one_third * fm2 + 19.0 * one_third * fm1 + 10.0 * one_third * fp1);
      F[((long long)iI * NV + v) * Nsys + sys] = sigma * fCompact + (1.0 - sigma) * fWENO;
#pragma unroll
      for (int k = 0; k < NV; k++) {           // sigma = 0 rows keep the reference's signed zeros: no shortcut
        double a, b, cc;
        if (upw > 0) { a = (one_half * sigma) * L[v * NV + k]; b = (1.0) * L[v * NV + k]; cc = (one_sixth * sigma) * L[v * NV + k]; }
        else         { cc = (one_half * sigma) * L[v * NV + k]; b = (1.0) * L[v * NV + k]; a = (one_sixth * sigma) * L[v * NV + k]; }
        const long long e = ((long long)iI * NV * NV + v * NV + k) * Nsys + sys;
        A[e] = a; B[e] = b; Cc[e] = cc;
      }
      continue;
    }
    if (ph.scheme == HPB_SCHEME_CRWENO5) {
      const double w1 = w[((3 * blk + 0) * NV + v) * ni + q], w2 = w[((3 * blk + 1) * NV + v) * ni + q],
                   w3 = w[((3 * blk + 2) * NV + v) * ni + q];
      double f1, f2, f3;
      if (bnd) {
        f1 = (2 * one_sixth) * fm3 - (7.0 * one_sixth) * fm2 + (11.0 * one_sixth) * fm1;
        f2 = (-one_sixth) * fm2 + (5.0 * one_sixth) * fm1 + (2 * one_sixth) * fp1;
        f3 = (2 * one_sixth) * fm1 + (5 * one_sixth) * fp1 - (one_sixth) * fp2;
      } else {
        f1 = (one_sixth) * (fm2 + 5 * fm1);
        f2 = (one_sixth) * (5 * fm1 + fp1);
        f3 = (one_sixth) * (fm1 + 5 * fp1);
      }
      lo = ((2 * one_third) * w1 + (one_third) * w2);
      di = ((one_third) * w1 + (2 * one_third) * (w2 + w3));
      hi = ((one_third) * w3);
      f = w1 * f1 + w2 * f2 + w3 * f3;
    } else {
      lo = three_by_ten; di = six_by_ten; hi = one_by_ten;
      if (bnd) f = one_by_thirty * fm3 - thirteen_by_sixty * fm2 + fortyseven_by_sixty * fm1 + twentyseven_by_sixty * fp1
                 - one_by_twenty * fp2;
      else     f = one_by_thirty * fm2 + nineteen_by_thirty * fm1 + one_third * fp1;
    }
    F[((long long)iI * NV + v) * Nsys + sys] = f;
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double a, b, cc;
      if (bnd) { a = 0.0; cc = 0.0; b = L[v * NV + k]; }
      else if (upw > 0) { a = lo * L[v * NV + k]; b = di * L[v * NV + k]; cc = hi * L[v * NV + k]; }
      else              { cc = lo * L[v * NV + k]; b = di * L[v * NV + k]; a = hi * L[v * NV + k]; }
      const long long e = ((long long)iI * NV * NV + v * NV + k) * Nsys + sys;
      A[e] = a; B[e] = b; Cc[e] = cc;
    }
  }
}

// include/matops.h in its own operation order (row-major N x N)
template <int N> __device__ __forceinline__ void bt_invert(const double* A, double* B)      // _MatrixInvert_ :93-127
{
  double Ac[N * N];
  for (int i = 0; i < N * N; i++) { Ac[i] = A[i]; B[i] = 0.0; }
  for (int i = 0; i < N; i++) B[i * N + i] = 1.0;
  for (int i = 0; i < N - 1; i++)
    for (int j = i + 1; j < N; j++) {
      const double factor = Ac[j * N + i] / Ac[i * N + i];
      for (int k = i + 1; k < N; k++) Ac[j * N + k] -= (factor * Ac[i * N + k]);
      for (int k = 0; k < j; k++) B[j * N + k] -= (factor * B[i * N + k]);
    }
  for (int i = N - 1; i >= 0; i--)
    for (int k = 0; k < N; k++) {
      double sum = 0.0;
      for (int j = i + 1; j < N; j++) sum += (Ac[i * N + j] * B[j * N + k]);
      B[i * N + k] = (B[i * N + k] - sum) / Ac[i * N + i];
    }
}
template <int N> __device__ __forceinline__ void bt_mul(const double* A, const double* B, double* C)
{
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) {
    double s = 0;
    for (int k = 0; k < N; k++) s += (A[i * N + k] * B[k * N + j]);
    C[i * N + j] = s;
  }
}
template <int N> __device__ __forceinline__ void bt_mul_sub(double* C, const double* A, const double* B)
{
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) {
    double s = C[i * N + j];
    for (int k = 0; k < N; k++) s -= (A[i * N + k] * B[k * N + j]);
    C[i * N + j] = s;
  }
}
// y -= A x, element by element in place like _MatVecMultiplySubtract_ (:81-86): y and x may be the same array (row 0
// of the back substitution), and then the macro's in-place semantics are reproduced
template <int N> __device__ __forceinline__ void bt_matvec_sub(double* y, const double* A, const double* x)
{
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) y[i] -= (A[i * N + j] * x[j]);
}
template <int N> __device__ __forceinline__ void bt_load(const double* __restrict__ M, long long row, long long Nsys, long long sys, double* blk)
{
  for (int e = 0; e < N * N; e++) blk[e] = M[(row * N * N + e) * Nsys + sys];
}
template <int N> __device__ __forceinline__ void bt_store(double* __restrict__ M, long long row, long long Nsys, long long sys, const double* blk)
{
  for (int e = 0; e < N * N; e++) M[(row * N * N + e) * Nsys + sys] = blk[e];
}

// TridiagLU/blocktridiagLU.c:103-320 on one rank: block forward elimination from row 1 on, block back substitution, one
// thread per system in the reference's operation order; the solution is written to the interface array fI.
template <int MODEL>
__global__ void k_block_tridiag(Geom G, int dir, double* __restrict__ A, double* __restrict__ B, const double* __restrict__ Cc,
                                double* __restrict__ X, double* __restrict__ fI)
{
  constexpr int N = ModelTraits<MODEL>::NV;
  const int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  const long long ni = (long long)M0 * M1 * M2;
  int T0, T1; long long s0, s1, qs;
  if (dir == 0)      { T0 = M1; T1 = M2; s0 = M0;  s1 = (long long)M0 * M1; qs = 1; }
  else if (dir == 1) { T0 = M0; T1 = M2; s0 = 1;   s1 = (long long)M0 * M1; qs = M0; }
  else               { T0 = M0; T1 = M1; s0 = 1;   s1 = M0;                 qs = (long long)M0 * M1; }
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x, t1 = blockIdx.y;
  if (t0 >= T0 || t1 >= T1) return;
  const long long Nsys = (long long)T0 * T1, sys = t0 + (long long)T0 * t1;
  const int n = G.N[dir] + 1;
  double bprev[N * N], cprev[N * N], aprev[N * N], xprev[N], binv[N * N], factor[N * N], a[N * N], b[N * N], x[N];
  bt_load<N>(B, 0, Nsys, sys, bprev); bt_load<N>(Cc, 0, Nsys, sys, cprev); bt_load<N>(A, 0, Nsys, sys, aprev);
  for (int v = 0; v < N; v++) xprev[v] = X[((long long)0 * N + v) * Nsys + sys];
  for (int i = 1; i < n; i++) {
    bt_load<N>(A, i, Nsys, sys, a); bt_load<N>(B, i, Nsys, sys, b);
    for (int v = 0; v < N; v++) x[v] = X[((long long)i * N + v) * Nsys + sys];
    bt_invert<N>(bprev, binv);
    bt_mul<N>(a, binv, factor);
    bt_mul_sub<N>(b, factor, cprev);
    for (int e = 0; e < N * N; e++) a[e] = 0.0;
    bt_mul_sub<N>(a, factor, aprev);
    bt_matvec_sub<N>(x, factor, xprev);
    bt_store<N>(A, i, Nsys, sys, a); bt_store<N>(B, i, Nsys, sys, b);
    for (int v = 0; v < N; v++) X[((long long)i * N + v) * Nsys + sys] = x[v];
    for (int e = 0; e < N * N; e++) { bprev[e] = b[e]; aprev[e] = a[e]; }
    bt_load<N>(Cc, i, Nsys, sys, cprev);
    for (int v = 0; v < N; v++) xprev[v] = x[v];
  }
  // back substitution; x0 = row 0 of the system as it stands (the reduced-system coupling of the multi-rank algorithm)
  double x0[N], xnext[N], xt[N], zero[N];
  for (int v = 0; v < N; v++) { x0[v] = X[((long long)0 * N + v) * Nsys + sys]; zero[v] = 0.0; }
  {
    const int i = n - 1;
    bt_load<N>(B, i, Nsys, sys, b); bt_load<N>(A, i, Nsys, sys, a); bt_load<N>(Cc, i, Nsys, sys, cprev);
    for (int v = 0; v < N; v++) x[v] = X[((long long)i * N + v) * Nsys + sys];
    bt_invert<N>(b, binv);
    bt_matvec_sub<N>(x, a, x0);
    bt_matvec_sub<N>(x, cprev, zero);
    for (int r = 0; r < N; r++) { double s = 0; for (int j = 0; j < N; j++) s += (binv[r * N + j] * x[j]); xt[r] = s; }
    for (int v = 0; v < N; v++) { X[((long long)i * N + v) * Nsys + sys] = xt[v]; xnext[v] = xt[v]; }
  }
  for (int i = n - 2; i > -1; i--) {
    bt_load<N>(B, i, Nsys, sys, b); bt_load<N>(A, i, Nsys, sys, a); bt_load<N>(Cc, i, Nsys, sys, cprev);
    for (int v = 0; v < N; v++) x[v] = X[((long long)i * N + v) * Nsys + sys];
    bt_invert<N>(b, binv);
    bt_matvec_sub<N>(x, cprev, xnext);
    if (i > 0) bt_matvec_sub<N>(x, a, x0);
    else       bt_matvec_sub<N>(x, a, x);           // row 0: the reference's x + d*bs IS this row
    for (int r = 0; r < N; r++) { double s = 0; for (int j = 0; j < N; j++) s += (binv[r * N + j] * x[j]); xt[r] = s; }
    for (int v = 0; v < N; v++) { X[((long long)i * N + v) * Nsys + sys] = xt[v]; xnext[v] = xt[v]; }
  }
  // the solution in the interface layout
  const long long qbase = t0 * s0 + t1 * s1;
  for (int i = 0; i < n; i++)
    for (int v = 0; v < N; v++) fI[v * ni + qbase + i * qs] = X[((long long)i * N + v) * Nsys + sys];
}

// ------------------------------------------------------------------------------------------
// The block systems of the characteristic compact schemes when the grid line is split among ranks:
// TridiagLU/blocktridiagLU.c:103-320 with all four stages, the reduced system (one block row per rank) solved by
// TridiagLU/blocktridiagIterJacobi.c:92-297 as the reference does by default, and the hand-over of the shared interface
// (Interp1PrimFifthOrderCRWENOChar.c:246-263). One thread per grid line, block / vector storage of k_compact_rows_char,
// `n` = rows on this rank, `first` = first rank of the line. Exchange buffers: element e of block k of system sys at
// [(k*N*N + e)*Nsys + sys] (k = a, b, c of the last row), the vector after them at [(3*N*N + v)*Nsys + sys]; reduced-system
// vectors at [v*Nsys + sys]. Operation order follows the reference macro by macro (compiled without FMA contraction).
template <int N> __device__ __forceinline__ void bt_vload(const double* __restrict__ X, long long row, long long Nsys, long long sys, double* x)
{
  for (int v = 0; v < N; v++) x[v] = X[(row * N + v) * Nsys + sys];
}
template <int N> __device__ __forceinline__ void bt_vstore(double* __restrict__ X, long long row, long long Nsys, long long sys, const double* x)
{
  for (int v = 0; v < N; v++) X[(row * N + v) * Nsys + sys] = x[v];
}
template <int N> __device__ __forceinline__ void bt_matvec(const double* A, const double* x, double* y)     // _MatVecMultiply_
{
  for (int i = 0; i < N; i++) { double s = 0; for (int j = 0; j < N; j++) s += (A[i * N + j] * x[j]); y[i] = s; }
}
__device__ __forceinline__ bool bmr_sys(const Geom& G, int dir, long long& Nsys, long long& sys, long long& qbase, long long& qs, long long& ni)
{
  const int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  ni = (long long)M0 * M1 * M2;
  int T0, T1; long long s0, s1;
  if (dir == 0)      { T0 = M1; T1 = M2; s0 = M0;  s1 = (long long)M0 * M1; qs = 1; }
  else if (dir == 1) { T0 = M0; T1 = M2; s0 = 1;   s1 = (long long)M0 * M1; qs = M0; }
  else               { T0 = M0; T1 = M1; s0 = 1;   s1 = M0;                 qs = (long long)M0 * M1; }
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x, t1 = blockIdx.y;
  Nsys = (long long)T0 * T1;
  if (t0 >= T0 || t1 >= T1) return false;
  sys = t0 + (long long)T0 * t1;
  qbase = t0 * s0 + t1 * s1;
  return true;
}

// stage 1 (blocktridiagLU.c:150-171) + the last row packed for the next rank (:181-190)
template <int MODEL>
__global__ void k_bmr_stage1(Geom G, int dir, int n, int first, double* __restrict__ A, double* __restrict__ B, double* __restrict__ Cc,
                             double* __restrict__ X, double* __restrict__ sendrow)
{
  constexpr int N = ModelTraits<MODEL>::NV;
  long long Nsys, sys, qbase, qs, ni;
  if (!bmr_sys(G, dir, Nsys, sys, qbase, qs, ni)) return;
  double binv[N * N], factor[N * N], am[N * N], bm[N * N], cm[N * N], a[N * N], b[N * N], xm[N], x[N];
  for (int i = (first ? 1 : 2); i < n; i++) {
    bt_load<N>(B, i - 1, Nsys, sys, bm); bt_load<N>(Cc, i - 1, Nsys, sys, cm); bt_load<N>(A, i - 1, Nsys, sys, am);
    bt_vload<N>(X, i - 1, Nsys, sys, xm);
    bt_load<N>(A, i, Nsys, sys, a); bt_load<N>(B, i, Nsys, sys, b); bt_vload<N>(X, i, Nsys, sys, x);
    bt_invert<N>(bm, binv);
    bt_mul<N>(a, binv, factor);
    bt_mul_sub<N>(b, factor, cm);
    for (int e = 0; e < N * N; e++) a[e] = 0.0;
    bt_mul_sub<N>(a, factor, am);
    bt_matvec_sub<N>(x, factor, xm);
    bt_store<N>(A, i, Nsys, sys, a); bt_store<N>(B, i, Nsys, sys, b); bt_vstore<N>(X, i, Nsys, sys, x);
    if (!first) {                    // the same elimination into row 0 (:163-169)
      double c0[N * N], b0[N * N], x0[N];
      bt_load<N>(Cc, 0, Nsys, sys, c0); bt_load<N>(B, 0, Nsys, sys, b0); bt_vload<N>(X, 0, Nsys, sys, x0);
      bt_mul<N>(c0, binv, factor);
      for (int e = 0; e < N * N; e++) c0[e] = 0.0;
      bt_mul_sub<N>(c0, factor, cm);
      bt_mul_sub<N>(b0, factor, am);
      bt_matvec_sub<N>(x0, factor, xm);
      bt_store<N>(Cc, 0, Nsys, sys, c0); bt_store<N>(B, 0, Nsys, sys, b0); bt_vstore<N>(X, 0, Nsys, sys, x0);
    }
  }
  bt_load<N>(A, n - 1, Nsys, sys, a); bt_store<N>(sendrow, 0, Nsys, sys, a);
  bt_load<N>(B, n - 1, Nsys, sys, a); bt_store<N>(sendrow, 1, Nsys, sys, a);
  bt_load<N>(Cc, n - 1, Nsys, sys, a); bt_store<N>(sendrow, 2, Nsys, sys, a);
  bt_vload<N>(X, n - 1, Nsys, sys, x);
  for (int v = 0; v < N; v++) sendrow[(3LL * N * N + v) * Nsys + sys] = x[v];
}

// stage 2 (blocktridiagLU.c:199-222), ranks other than the first; then the start of the Jacobi iteration on the reduced
// block row (blocktridiagIterJacobi.c:116-126: rhs saved, initial guess x = b^-1 rhs). The first rank enters the reduced
// system with (0, I, 0 | 0) (blocktridiagLU.c:247). red = [rhs | x | sent x | recvL | recvR | xp1], N*Nsys each
template <int MODEL>
__global__ void k_bmr_stage2(Geom G, int dir, int n, int first, double* __restrict__ A, double* __restrict__ B, double* __restrict__ Cc,
                             double* __restrict__ X, const double* __restrict__ recvrow, double* __restrict__ red)
{
  constexpr int N = ModelTraits<MODEL>::NV;
  long long Nsys, sys, qbase, qs, ni;
  if (!bmr_sys(G, dir, Nsys, sys, qbase, qs, ni)) return;
  const long long NS = (long long)N * Nsys;
  if (first) {
    for (int v = 0; v < N; v++) { red[0 * NS + v * Nsys + sys] = 0.0; red[1 * NS + v * Nsys + sys] = 0.0; red[2 * NS + v * Nsys + sys] = 0.0; }
    return;
  }
  double am1[N * N], bm1[N * N], cm1[N * N], xm1[N], binv[N * N], factor[N * N], a0[N * N], b0[N * N], c0[N * N], x0[N];
  bt_load<N>(recvrow, 0, Nsys, sys, am1); bt_load<N>(recvrow, 1, Nsys, sys, bm1); bt_load<N>(recvrow, 2, Nsys, sys, cm1);
  for (int v = 0; v < N; v++) xm1[v] = recvrow[(3LL * N * N + v) * Nsys + sys];
  bt_load<N>(A, 0, Nsys, sys, a0); bt_load<N>(B, 0, Nsys, sys, b0); bt_load<N>(Cc, 0, Nsys, sys, c0); bt_vload<N>(X, 0, Nsys, sys, x0);
  bt_invert<N>(bm1, binv);
  bt_mul<N>(a0, binv, factor);
  bt_mul_sub<N>(b0, factor, cm1);
  for (int e = 0; e < N * N; e++) a0[e] = 0.0;
  bt_mul_sub<N>(a0, factor, am1);
  bt_matvec_sub<N>(x0, factor, xm1);
  double al[N * N], bl[N * N], cl[N * N], xl[N];
  bt_load<N>(A, n - 1, Nsys, sys, al); bt_load<N>(B, n - 1, Nsys, sys, bl); bt_load<N>(Cc, n - 1, Nsys, sys, cl); bt_vload<N>(X, n - 1, Nsys, sys, xl);
  bt_invert<N>(bl, binv);
  bt_mul<N>(c0, binv, factor);
  bt_mul_sub<N>(b0, factor, al);
  for (int e = 0; e < N * N; e++) c0[e] = 0.0;
  bt_mul_sub<N>(c0, factor, cl);
  bt_matvec_sub<N>(x0, factor, xl);
  bt_store<N>(A, 0, Nsys, sys, a0); bt_store<N>(B, 0, Nsys, sys, b0); bt_store<N>(Cc, 0, Nsys, sys, c0);
  // Jacobi: rhs = x0, x = b0^-1 rhs
  double xg[N];
  bt_invert<N>(b0, binv);
  bt_matvec<N>(binv, x0, xg);
  bt_vstore<N>(X, 0, Nsys, sys, xg);
  for (int v = 0; v < N; v++) { red[0 * NS + v * Nsys + sys] = x0[v]; red[1 * NS + v * Nsys + sys] = xg[v]; red[2 * NS + v * Nsys + sys] = xg[v]; }
}

// one Jacobi iteration on the reduced block row (blocktridiagIterJacobi.c:205-216 norm, :276-283 update), n = 1 per rank
template <int MODEL>
__global__ void k_bmr_jacobi(Geom G, int dir, int first, int pass, const double* __restrict__ A, const double* __restrict__ B,
                             const double* __restrict__ Cc, double* __restrict__ X, double* __restrict__ red, double* __restrict__ part)
{
  constexpr int N = ModelTraits<MODEL>::NV;
  long long Nsys, sys, qbase, qs, ni;
  double contrib = 0.0;
  if (bmr_sys(G, dir, Nsys, sys, qbase, qs, ni) && !first) {
    const long long NS = (long long)N * Nsys;
    double a[N * N], b[N * N], c[N * N], rhs[N], x[N], rl[N], rr[N];
    bt_load<N>(A, 0, Nsys, sys, a); bt_load<N>(B, 0, Nsys, sys, b); bt_load<N>(Cc, 0, Nsys, sys, c);
    for (int v = 0; v < N; v++) {
      rhs[v] = red[0 * NS + v * Nsys + sys]; x[v] = red[1 * NS + v * Nsys + sys];
      rl[v] = red[3 * NS + v * Nsys + sys]; rr[v] = red[4 * NS + v * Nsys + sys];
    }
    if (pass == 0) {
      double err[N];
      for (int v = 0; v < N; v++) err[v] = rhs[v];
      bt_matvec_sub<N>(err, a, rl);
      bt_matvec_sub<N>(err, b, x);
      bt_matvec_sub<N>(err, c, rr);
      for (int v = 0; v < N; v++) contrib += (err[v] * err[v]);
    } else {
      double xt[N], binv[N * N], xn[N];
      for (int v = 0; v < N; v++) xt[v] = rhs[v];
      bt_matvec_sub<N>(xt, a, rl);
      bt_matvec_sub<N>(xt, c, rr);
      bt_invert<N>(b, binv);
      bt_matvec<N>(binv, xt, xn);
      bt_vstore<N>(X, 0, Nsys, sys, xn);
      for (int v = 0; v < N; v++) { red[1 * NS + v * Nsys + sys] = xn[v]; red[2 * NS + v * Nsys + sys] = xn[v]; }
    }
  }
  if (pass == 0) {
    const double v = block_reduce(contrib, false);
    if (threadIdx.x == 0) part[blockIdx.x + (long long)gridDim.x * blockIdx.y] = v;
  }
}

// what each rank sends to the previous one after the reduced solve: its row-0 x (blocktridiagLU.c:262: the first rank's is
// its own row 0 as stage 1 left it -- its Jacobi ran on a private (0, I, 0 | 0))
template <int MODEL>
__global__ void k_bmr_row0(Geom G, int dir, const double* __restrict__ X, double* __restrict__ out)
{
  constexpr int N = ModelTraits<MODEL>::NV;
  long long Nsys, sys, qbase, qs, ni;
  if (!bmr_sys(G, dir, Nsys, sys, qbase, qs, ni)) return;
  for (int v = 0; v < N; v++) out[v * Nsys + sys] = X[((long long)0 * N + v) * Nsys + sys];
}

// stage 4 (blocktridiagLU.c:281-301): xp1 = the next rank's row-0 x (0 on the last rank); the solution goes to the interface
// array; row 0 packed for the previous rank (the shared interface, Interp1PrimFifthOrderCRWENOChar.c:254)
template <int MODEL>
__global__ void k_bmr_stage4(Geom G, int dir, int n, int first, const double* __restrict__ A, const double* __restrict__ B,
                             const double* __restrict__ Cc, double* __restrict__ X, const double* __restrict__ xp1,
                             double* __restrict__ sendfirst, double* __restrict__ fI)
{
  constexpr int N = ModelTraits<MODEL>::NV;
  long long Nsys, sys, qbase, qs, ni;
  if (!bmr_sys(G, dir, Nsys, sys, qbase, qs, ni)) return;
  double a[N * N], b[N * N], c[N * N], binv[N * N], x[N], x0[N], xn[N], xt[N];
  bt_vload<N>(X, 0, Nsys, sys, x0);
  for (int v = 0; v < N; v++) xn[v] = xp1[v * Nsys + sys];
  {
    const int i = n - 1;
    bt_load<N>(A, i, Nsys, sys, a); bt_load<N>(B, i, Nsys, sys, b); bt_load<N>(Cc, i, Nsys, sys, c); bt_vload<N>(X, i, Nsys, sys, x);
    bt_invert<N>(b, binv);
    bt_matvec_sub<N>(x, a, x0);
    bt_matvec_sub<N>(x, c, xn);
    bt_matvec<N>(binv, x, xt);
    bt_vstore<N>(X, i, Nsys, sys, xt);
    for (int v = 0; v < N; v++) xn[v] = xt[v];
  }
  for (int i = n - 2; i > (first ? 0 : 1) - 1; i--) {
    bt_load<N>(A, i, Nsys, sys, a); bt_load<N>(B, i, Nsys, sys, b); bt_load<N>(Cc, i, Nsys, sys, c); bt_vload<N>(X, i, Nsys, sys, x);
    bt_invert<N>(b, binv);
    bt_matvec_sub<N>(x, c, xn);
    if (i > 0) bt_matvec_sub<N>(x, a, x0);
    else       bt_matvec_sub<N>(x, a, x);           // row 0 of the first rank: the reference's x + d*bs IS this row
    bt_matvec<N>(binv, x, xt);
    bt_vstore<N>(X, i, Nsys, sys, xt);
    for (int v = 0; v < N; v++) xn[v] = xt[v];
  }
  for (int i = 0; i < n; i++)
    for (int v = 0; v < N; v++) fI[v * ni + qbase + i * qs] = X[((long long)i * N + v) * Nsys + sys];
  for (int v = 0; v < N; v++) sendfirst[v * Nsys + sys] = X[((long long)0 * N + v) * Nsys + sys];
}
// the solution of the shared interface N arrives from the next rank (Interp1PrimFifthOrderCRWENOChar.c:262)
template <int MODEL>
__global__ void k_bmr_put_last(Geom G, int dir, double* __restrict__ fI, const double* __restrict__ recvfirst)
{
  constexpr int N = ModelTraits<MODEL>::NV;
  long long Nsys, sys, qbase, qs, ni;
  if (!bmr_sys(G, dir, Nsys, sys, qbase, qs, ni)) return;
  for (int v = 0; v < N; v++) fI[v * ni + qbase + (long long)G.N[dir] * qs] = recvfirst[v * Nsys + sys];
}

// the linear schemes, component-wise: Interp1PrimFifthOrderUpwind.c:60-147, Interp1PrimFirstOrderUpwind.c:78-84,
// Interp1PrimSecondOrderCentral.c:80-88, Interp1PrimFourthOrderCentral.c:96-120
// ... and the MUSCL schemes: Interp1PrimSecondOrderMUSCL.c:118-156 (limiters of src/LimiterFunctions/),
// Interp1PrimThirdOrderMUSCL.c:118-160 (Koren's limiter with the epsilon of muscl.inp)
__device__ __forceinline__ double muscl_limiter_fn(int type, double r)
{
  if (type == HPB_LIMITER_MINMOD)   return fmax(0.0, fmin(1.0, r));
  if (type == HPB_LIMITER_VANLEER)  return (r + fabs(r)) / (1.0 + fabs(r));
  if (type == HPB_LIMITER_SUPERBEE) return fmax(fmax(0.0, fmin(2 * r, 1.0)), fmin(r, 2.0));
  return fmax(0.0, fmin(fmin(r, 0.5 * (1.0 + r)), 1.0));          // generalised minmod, theta = 1
}
// one interface value from the six cells around it, s[0..5] = cells i-3 .. i+2 (interface between cells i-1 and i)
__device__ __forceinline__ double linear_scheme_value(int scheme, int limiter, double meps, int upw, const double* s)
{
  const double c1 = 7.0 / 12.0, c2 = -1.0 / 12.0;
  const double one_by_thirty = 1.0 / 30.0, thirteen_by_sixty = 13.0 / 60.0, fortyseven_by_sixty = 47.0 / 60.0,
               twentyseven_by_sixty = 27.0 / 60.0, one_by_twenty = 1.0 / 20.0;
  // biased numbering (m3, m2, m1, p1, p2)
  const double fm3 = (upw > 0 ? s[0] : s[5]), fm2 = (upw > 0 ? s[1] : s[4]), fm1 = (upw > 0 ? s[2] : s[3]),
               fp1 = (upw > 0 ? s[3] : s[2]), fp2 = (upw > 0 ? s[4] : s[1]);
  if (scheme == HPB_SCHEME_MUSCL3) {
    const double one_third = 1.0 / 3.0, one_sixth = 1.0 / 6.0;
    if (upw > 0) {
      const double m2 = s[1], m1 = s[2], p1 = s[3];
      const double fdiff = p1 - m1, bdiff = m1 - m2;
      const double limit = (3 * fdiff * bdiff + meps) / (2 * (fdiff - bdiff) * (fdiff - bdiff) + 3 * fdiff * bdiff + meps);
      return m1 + limit * (one_third * fdiff + one_sixth * bdiff);
    } else {
      const double m1 = s[2], p1 = s[3], p2 = s[4];
      const double fdiff = p2 - p1, bdiff = p1 - m1;
      const double limit = (3 * fdiff * bdiff + meps) / (2 * (fdiff - bdiff) * (fdiff - bdiff) + 3 * fdiff * bdiff + meps);
      return p1 - limit * (one_third * fdiff + one_sixth * bdiff);
    }
  }
  if (scheme == HPB_SCHEME_MUSCL2) {
    if (upw > 0) {
      const double m2 = s[1], m1 = s[2], p1 = s[3];
      const double slope_ratio = (m1 - m2) / ((p1 - m1) + 1e-40);
      return m1 + 0.5 * muscl_limiter_fn(limiter, slope_ratio) * (p1 - m1);
    } else {
      const double m1 = s[2], p1 = s[3], p2 = s[4];
      const double slope_ratio = (p1 - m1) / ((p2 - p1) + 1e-40);
      return p1 + 0.5 * muscl_limiter_fn(limiter, slope_ratio) * (p1 - p2);
    }
  }
  if (scheme == HPB_SCHEME_FIRST)  return fm1;
  if (scheme == HPB_SCHEME_SECOND) return 0.5 * (s[2] + s[3]);
  if (scheme == HPB_SCHEME_FOURTH) return c2 * s[1] + c1 * s[2] + c1 * s[3] + c2 * s[4];
  return one_by_thirty * fm3 - thirteen_by_sixty * fm2 + fortyseven_by_sixty * fm1 + twentyseven_by_sixty * fp1 - one_by_twenty * fp2;
}

// component-wise, or characteristic-wise like the reference's Interp1Prim...Char.c twins: averaged state of the two cells
// next to the interface, its left eigenvectors applied to every stencil point, the scalar scheme per characteristic
// field, the right eigenvectors applied to the result
template <int MODEL>
__global__ void k_interp_upw5(Geom G, Phys ph, const double* __restrict__ fC, const double* __restrict__ u, int upw, int dir,
                              double* __restrict__ fI)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  const int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (i0 >= M0) return;
  const long long q = i0 + (long long)M0 * (i1 + (long long)M1 * i2);
  const long long ni = (long long)M0 * M1 * M2;
  const long long st = G.st[dir];
  const long long pm1 = cell_index(G, i0, i1, i2) - st;      // cell i-1
  const bool use_char = ph.interp_char && (MODEL == HPB_MODEL_EULER1D || MODEL == HPB_MODEL_NS2D || MODEL == HPB_MODEL_NS3D);
  double S[6][NV];
#pragma unroll
  for (int k = 0; k < 6; k++)
#pragma unroll
    for (int v = 0; v < NV; v++) S[k][v] = fC[v * G.npg + pm1 + (k - 2) * st];
  if (!use_char) {
#pragma unroll
    for (int v = 0; v < NV; v++) {
      double s[6];
#pragma unroll
      for (int k = 0; k < 6; k++) s[k] = S[k][v];
      fI[v * ni + q] = linear_scheme_value(ph.scheme, ph.muscl_limiter, ph.muscl_eps, upw, s);
    }
    return;
  }
  double UL[NV], UR[NV], uavg[NV], lam[NV], L[NV * NV], R[NV * NV], fc[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) { UL[v] = u[v * G.npg + pm1]; UR[v] = u[v * G.npg + pm1 + st]; }
  roe_average<MODEL>(ph, UL, UR, uavg);
  eigen<MODEL>(ph, uavg, dir, lam, L, R);
#pragma unroll
  for (int v = 0; v < NV; v++) {
    double s[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
      double a = 0.0;
#pragma unroll
      for (int j = 0; j < NV; j++) a += L[v * NV + j] * S[k][j];
      s[k] = a;
    }
    fc[v] = linear_scheme_value(ph.scheme, ph.muscl_limiter, ph.muscl_eps, upw, s);
  }
#pragma unroll
  for (int i = 0; i < NV; i++) {
    double a = 0.0;
#pragma unroll
    for (int j = 0; j < NV; j++) a += R[i * NV + j] * fc[j];
    fI[i * ni + q] = a;
  }
}

// NavierStokes3DSource.c:124-160: the source function G = g_grav * (0, d_x, d_y, d_z, 1) as a cell array
__global__ void k_ns3d_source_fn(Geom G, const double* __restrict__ gg, int dir, double* __restrict__ S)
{
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= G.npg) return;
  const double g = gg[p];
  S[p] = 0.0;
  for (int k = 0; k < G.ndims; k++) S[(long long)(1 + k) * G.npg + p] = g * (dir == k);
  S[(long long)(G.nvars - 1) * G.npg + p] = g;
}
// NavierStokes3DSource.c:174-205: interface value = average of the left- and right-biased reconstructions;
// only components dir+1 and 4 are consumed (k_ns3d_source)
__global__ void k_ns3d_source_avg(long long ni, int dir, int ve, const double* __restrict__ SL, const double* __restrict__ SR,
                                  double* __restrict__ sI)
{
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= ni) return;
  sI[q]      = 0.5 * (SL[(1 + dir) * ni + q] + SR[(1 + dir) * ni + q]);
  sI[ni + q] = 0.5 * (SL[ve * ni + q] + SR[ve * ni + q]);          // ve: the energy component
}
// boundary-face fluxes from a stored interface array (conservation bookkeeping of the piecewise path)
__global__ void k_face_from_iface(Geom G, int dir, int face, const double* __restrict__ fI, double* __restrict__ out)
{
  int M[3] = { G.N[0] + (dir == 0), G.N[1] + (dir == 1), G.N[2] + (dir == 2) };
  int F[3] = { M[0], M[1], M[2] }; F[dir] = 1;
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (i0 >= F[0]) return;
  int j[3] = { i0, i1, i2 };
  const long long qf = i0 + (long long)F[0] * (i1 + (long long)F[1] * i2);
  const long long nface = (long long)F[0] * F[1] * F[2];
  j[dir] = face ? G.N[dir] : 0;
  const long long q = j[0] + (long long)M[0] * (j[1] + (long long)M[1] * j[2]);
  const long long ni = (long long)M[0] * M[1] * M[2];
  for (int v = 0; v < G.nvars; v++) out[v * nface + qf] = fI[v * ni + q];
}

template <int MODEL>
__global__ void k_upwind(Geom G, Phys ph, const double* __restrict__ fL, const double* __restrict__ fR,
                         const double* __restrict__ uL, const double* __restrict__ uR, const double* __restrict__ u,
                         const double* __restrict__ gf, const double* __restrict__ gg, int dir, double* __restrict__ fI)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  const int M0 = G.N[0] + (dir == 0), M1 = G.N[1] + (dir == 1), M2 = G.N[2] + (dir == 2);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (i0 >= M0) return;
  const long long q = i0 + (long long)M0 * (i1 + (long long)M1 * i2);
  const long long ni = (long long)M0 * M1 * M2;
  const long long st = G.st[dir];
  const long long pL = cell_index(G, i0, i1, i2) - st, pR = pL + st;
  double a[NV], b[NV], c[NV], d[NV], cl[NV], cr[NV], out[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) {
    a[v] = fL[v * ni + q]; b[v] = fR[v * ni + q]; c[v] = uL[v * ni + q]; d[v] = uR[v * ni + q];
    cl[v] = u[v * G.npg + pL]; cr[v] = u[v * G.npg + pR];
  }
  const double* kk = (MODEL == HPB_MODEL_EULER1D) ? gf : gg;
  double kLv = kk ? kk[pL] : 1.0, kRv = kk ? kk[pR] : 1.0;
  if (MODEL == HPB_MODEL_LINEAR_ADR && ph.advf != nullptr) { kLv = ph.advf[dir * ph.advf_npg + pL]; kRv = ph.advf[dir * ph.advf_npg + pR]; }
  upwind_fn<MODEL>(ph, dir, a, b, c, d, cl, cr, kLv, kRv, out);
#pragma unroll
  for (int v = 0; v < NV; v++) fI[v * ni + q] = out[v];
}

// FirstDerivativeFourthOrderCentral over any nv-component SoA array (un-scaled)
__global__ void k_first_derivative(Geom G, const double* __restrict__ f, int dir, int nv, double* __restrict__ Df)
{
  int B[3] = { G.N[0], G.N[1], G.N[2] };
  B[dir] = G.P[dir];
  int t[3] = { (int)(blockIdx.x * blockDim.x + threadIdx.x), (int)blockIdx.y, (int)blockIdx.z };
  if (t[0] >= B[0]) return;
  int ii[3] = { t[0], t[1], t[2] };
  ii[dir] -= G.g;
  const int i = ii[dir];
  const long long p = cell_index(G, ii[0], ii[1], ii[2]);
  const long long st = G.st[dir];
  int lo, hi;
  const int N = G.N[dir], g = G.g;
  if (i == -g) { lo = 0; hi = 4; } else if (i == -g + 1) { lo = -1; hi = 3; }
  else if (i < N + g - 2) { lo = -2; hi = 2; } else if (i == N + g - 2) { lo = -3; hi = 1; } else { lo = -4; hi = 0; }
  for (int v = 0; v < nv; v++) {
    double s[9];
    for (int k = lo; k <= hi; k++) s[k + 4] = f[v * G.npg + p + k * st];
    Df[v * G.npg + p] = d1_coeffs(i, N, g, &s[4]);
  }
}

__global__ void k_second_derivative(Geom G, const double* __restrict__ f, int dir, int nv, int order, double* __restrict__ D2f)
{
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
  if (i0 >= G.N[0]) return;
  const long long p = cell_index(G, i0, i1, i2);
  const long long st = G.st[dir];
  const double one_twelve = 1.0 / 12.0;
  for (int v = 0; v < nv; v++) {
    const double* s = f + v * G.npg + p;
    if (order == 2) D2f[v * G.npg + p] = s[-st] - 2*s[0] + s[st];
    else D2f[v * G.npg + p] = (-s[-2*st] + 16*s[-st] - 30*s[0] + 16*s[st] - s[2*st]) * one_twelve;
  }
}

inline dim3 grid3(int n0, int n1, int n2) { return dim3((n0 + TPB - 1) / TPB, n1, n2); }
inline unsigned grid1(long long n) { long long b = (n + 255) / 256; return (unsigned)(b > 148LL * 64 ? 148 * 64 : (b < 1 ? 1 : b)); }

} // namespace

// =============================================================================================
// launchers
namespace hpbk {

#define LAUNCHED(h) ((h)->launches++)

void aos_to_soa(hpb_solver* h, const double* aos, double* soa, long long npts, int nv)
{
  k_aos_to_soa<<<(unsigned)((npts + 255) / 256), 256, 0, h->stream>>>(aos, soa, npts, nv); LAUNCHED(h);
}
void soa_to_aos(hpb_solver* h, const double* soa, double* aos, long long npts, int nv)
{
  k_soa_to_aos<<<(unsigned)((npts + 255) / 256), 256, 0, h->stream>>>(soa, aos, npts, nv); LAUNCHED(h);
}

void apply_bc(hpb_solver* h, double* u)
{
  ProfScope ps(h, HPB_PROF_BC);
  const Geom& G = h->geo;
  // Euler1DInitialize.c never sets boundary[n].gamma (NavierStokes2D/3DInitialize.c do): the reference's 1-D branches of the
  // boundary functions run with the calloc-ed 0, i.e. energy_gpt = -(-(e - K)) + K
  const double bc_gamma = (h->cfg.model == HPB_MODEL_EULER1D) ? 0.0 : h->phys.gamma;
  for (const ZoneDev& z : h->zones) {
    if (!z.on) continue;
    if (z.type == HPB_BC_PERIODIC && h->cfg.iproc[z.dim] != 1) continue;
    if (z.type == HPB_BC_SPONGE) continue;                     // BCSpongeUDummy: a source term, no ghost fill
    const int b0 = z.ie[0] - z.is[0], b1 = (G.ndims > 1 ? z.ie[1] - z.is[1] : 1), b2 = (G.ndims > 2 ? z.ie[2] - z.is[2] : 1);
    if (b0 <= 0 || b1 <= 0 || b2 <= 0) continue;
    k_bc_zone<<<(unsigned)(((long long)b0 * b1 * b2 + TPB - 1) / TPB), TPB, 0, h->stream>>>(G, z, bc_gamma, u); LAUNCHED(h);
  }
}

bool has_sponge(const hpb_solver* h)
{
  for (int n = 0; n < h->cfg.nzones; n++) if (h->cfg.zones[n].type == HPB_BC_SPONGE) return true;
  return false;
}

void sponge_source(hpb_solver* h, const double* u, double* out)
{
  const Geom& G = h->geo;
  for (const ZoneDev& z : h->zones) {
    if (!z.on || z.type != HPB_BC_SPONGE) continue;
    const int b0 = z.ie[0] - z.is[0], b1 = (G.ndims > 1 ? z.ie[1] - z.is[1] : 1), b2 = (G.ndims > 2 ? z.ie[2] - z.is[2] : 1);
    if (b0 <= 0 || b1 <= 0 || b2 <= 0) continue;
    k_sponge<<<(unsigned)(((long long)b0 * b1 * b2 + TPB - 1) / TPB), TPB, 0, h->stream>>>(G, z, h->d_x, u, out); LAUNCHED(h);
  }
}

void set_zero(hpb_solver* h, double* a, long long n) { cudaMemsetAsync(a, 0, n * sizeof(double), h->stream); }

void combine_rhs(hpb_solver* h, double* rhs, const double* par, const double* src)
{
  const long long n = h->geo.npg * h->geo.nvars;
  k_combine_rhs<<<grid1(n), 256, 0, h->stream>>>(rhs, par, src, n); LAUNCHED(h);
}

void copy(hpb_solver* h, double* dst, const double* src, long long n)
{
  k_copy<<<grid1(n), 256, 0, h->stream>>>(dst, src, n); LAUNCHED(h);
}

void hyperbolic_generic(hpb_solver* h, const double* u, double* out, bool negate, bool with_source, double* src)
{
  if (h->cfg.hyp_scheme != HPB_SCHEME_WENO5) { hyperbolic_pieces(h, u, out, negate, with_source, src); return; }
  const Geom& G = h->geo;
  const bool grav = h->phys.has_grav;
  const double* gf = (h->cfg.model == HPB_MODEL_LINEAR_ADR) ? nullptr : h->d_gravf;
  const double* gg = (h->cfg.model == HPB_MODEL_LINEAR_ADR) ? nullptr : h->d_gravg;
  for (int d = 0; d < G.ndims; d++) {
    const int M[3] = { G.N[0] + (d == 0), G.N[1] + (d == 1), G.N[2] + (d == 2) };
    double* sI = (with_source && grav && h->phys.grav[d] != 0.0) ? h->d_sI : nullptr;
    ProfScope ps(h, HPB_PROF_SWEEP_X + d);
#define CALL(M_) k_iface<M_><<<grid3(M[0], M[1], M[2]), TPB, 0, h->stream>>>(G, h->phys, u, gf, gg, d, h->d_fI, sI, -1)
    MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
    LAUNCHED(h);
    const int mode = negate ? (d == 0 ? 0 : 1) : (d == 0 ? 2 : 3);
    k_divergence<<<grid3(G.N[0], G.N[1], G.N[2]), TPB, 0, h->stream>>>(G, h->d_dxinv, h->d_fI, d, out, mode); LAUNCHED(h);
    if (sI) {
      k_ns3d_source<<<grid3(G.N[0], G.N[1], G.N[2]), TPB, 0, h->stream>>>(G, h->phys, h->d_dxinv, u, h->d_gravf, sI, d, src);
      LAUNCHED(h);
    }
  }
}

// HyperbolicFunction.c:81-109 + ReconstructHyperbolic :167-222 as the reference sequences them: flux, weights
// (SetInterpLimiterVar, WENO-type schemes only), modified solution, four InterpolateInterfacesHyp calls, Upwind,
// flux difference; then NavierStokes3DSource.c:54-101 for a gravity direction, with the same stored flux weights
// (quirk Q5) and fluxC / fL / fR as scratch like the reference (:54-57).
void hyperbolic_pieces(hpb_solver* h, const double* u, double* out, bool negate, bool with_source, double* src)
{
  const Geom& G = h->geo;
  const bool grav = h->phys.has_grav;
  const bool has_w = hpb_scheme_has_weights(h->cfg.hyp_scheme);
  double *fC = h->d_cell[0], *uC = h->d_cell[1];
  double *fL = h->d_iface[1], *fR = h->d_iface[2], *uL = h->d_iface[3], *uR = h->d_iface[4];
  long long wo = 0;              // the weights of direction d live where the fine-grained API keeps them (capi.cu: woff)
  for (int d = 0; d < G.ndims; d++) {
    const int M[3] = { G.N[0] + (d == 0), G.N[1] + (d == 1), G.N[2] + (d == 2) };
    const long long ni = (long long)M[0] * M[1] * M[2];
    double* w = h->d_w + wo;
    wo += 12 * ni * G.nvars;
    ProfScope ps(h, HPB_PROF_SWEEP_X + d);
    flux(h, u, fC, d);
    if (has_w) { weno_weights(h, fC, u, d, w); h->w_valid = true; }
    modified_solution(h, u, uC);
    weno_interp(h, uL, uC, u, w,  1, d, 1);
    weno_interp(h, uR, uC, u, w, -1, d, 1);
    weno_interp(h, fL, fC, u, w,  1, d, 0);
    weno_interp(h, fR, fC, u, w, -1, d, 0);
    upwind(h, h->d_fI, fL, fR, uL, uR, u, d);
    const int mode = negate ? (d == 0 ? 0 : 1) : (d == 0 ? 2 : 3);
    k_divergence<<<grid3(G.N[0], G.N[1], G.N[2]), TPB, 0, h->stream>>>(G, h->d_dxinv, h->d_fI, d, out, mode); LAUNCHED(h);
    if (with_source && grav && h->phys.grav[d] != 0.0) {
      k_ns3d_source_fn<<<(unsigned)((G.npg + 255) / 256), 256, 0, h->stream>>>(G, h->d_gravg, d, fC); LAUNCHED(h);
      weno_interp(h, fL, fC, u, w,  1, d, 0);
      weno_interp(h, fR, fC, u, w, -1, d, 0);
      k_ns3d_source_avg<<<(unsigned)((ni + 255) / 256), 256, 0, h->stream>>>(ni, d, G.nvars - 1, fL, fR, h->d_sI); LAUNCHED(h);
      k_ns3d_source<<<grid3(G.N[0], G.N[1], G.N[2]), TPB, 0, h->stream>>>(G, h->phys, h->d_dxinv, u, h->d_gravf, h->d_sI, d, src);
      LAUNCHED(h);
    }
  }
}

// ---- compact schemes with the grid line split among ranks. Group style (hpb_internal.h: hpbc): `hs` = one solver with
// the NCCL transport, every rank with the in-process one; each step loops over the group.
static long long mr_nsys(const hpb_solver* h, int dir)
{
  const Geom& G = h->geo;
  const long long M[3] = { G.N[0] + (dir == 0), G.N[1] + (dir == 1), G.N[2] + (dir == 2) };
  return M[0] * M[1] * M[2] / M[dir] * G.nvars;
}
static int mr_ensure(hpb_solver* h)
{
  if (h->d_mr) return HPB_OK;
  long long m = 0;
  for (int d = 0; d < h->geo.ndims; d++) { const long long k = mr_nsys(h, d); if (k > m) m = k; }
  const size_t bytes = (size_t)(17 * m + 4096) * sizeof(double);
  if (cudaMalloc((void**)&h->d_mr, bytes) != cudaSuccess) return hpb_fail(HPB_ERR_ALLOC, "compact schemes across ranks: scratch allocation failed");
  cudaMemset(h->d_mr, 0, bytes);
  cudaStreamSynchronize(cudaStreamLegacy);
  if (cudaMallocHost((void**)&h->h_mr, 128 * sizeof(double)) != cudaSuccess) return hpb_fail(HPB_ERR_ALLOC, "pinned alloc");
  return HPB_OK;
}

// the tridiagonal solve of the rows the group's k_compact_rows launches have left in d_tri[0..2] / X[r] (tridiagLU.c with
// the Jacobi reduced solve, then the shared interface from the next rank)
static int compact_solve_group(hpb_solver** hs, int n, int dir, double* const* X)
{
  const int tpb = 64;
  // lusolver.inp (tridiagLUInit.c:54-90; defaults jacobi, maxiter 10, atol 1e-12, rtol 1e-10, norm evaluated)
  const double atol = hs[0]->cfg.lu_atol, rtol = hs[0]->cfg.lu_rtol;
  const int maxiter = hs[0]->cfg.lu_maxiter;
  const bool evalnorm = hs[0]->cfg.lu_evaluate_norm != 0;
  std::vector<double*> sendrow(n), recvrow(n), red_x(n), red_sx(n), red_rl(n), red_rr(n), xp1(n), sfirst(n), rfirst(n), dnorm(n), dgath(n);
  std::vector<int> active(n, 1), first(n), last(n), rows(n);
  std::vector<dim3> grid(n);
  std::vector<long long> nsys(n), nsys4(n);
  for (int r = 0; r < n; r++) {
    hpb_solver* h = hs[r];
    int rc = mr_ensure(h); if (rc) return rc;
    const Geom& G = h->geo;
    const int M[3] = { G.N[0] + (dir == 0), G.N[1] + (dir == 1), G.N[2] + (dir == 2) };
    int T0, T1;
    if (dir == 0) { T0 = M[1]; T1 = M[2]; } else if (dir == 1) { T0 = M[0]; T1 = M[2]; } else { T0 = M[0]; T1 = M[1]; }
    grid[r] = dim3((T0 + tpb - 1) / tpb, T1, G.nvars);
    nsys[r] = mr_nsys(h, dir); nsys4[r] = 4 * nsys[r];
    first[r] = G.lo_phys[dir]; last[r] = G.hi_phys[dir];
    rows[r] = G.N[dir] + (last[r] ? 1 : 0);       // all ranks but the last leave the shared interface to the next one (:208-209)
    double* m = h->d_mr;
    sendrow[r] = m; recvrow[r] = m + 4 * nsys[r];
    double* red = m + 8 * nsys[r];                 // [rhs | x | sent x | recvL | recvR | xp1]
    red_x[r] = red + nsys[r]; red_sx[r] = red + 2 * nsys[r]; red_rl[r] = red + 3 * nsys[r]; red_rr[r] = red + 4 * nsys[r];
    xp1[r] = red + 5 * nsys[r];
    sfirst[r] = m + 14 * nsys[r]; rfirst[r] = m + 15 * nsys[r];
    dnorm[r] = m + 16 * nsys[r]; dgath[r] = dnorm[r] + 8;     // + partial sums from dnorm + 128 on
  }
#define EACH_R for (int r = 0; r < n; r++)
#define HCUR hpb_solver* h = hs[r]; cudaSetDevice(h->device); const Geom& G = h->geo; (void)G
  EACH_R { HCUR;
    cudaMemsetAsync(h->d_mr + 8 * nsys[r] + 3 * nsys[r], 0, 3 * nsys[r] * sizeof(double), h->stream);     // recvL, recvR, xp1
    k_mr_stage1<<<grid[r], tpb, 0, h->stream>>>(G, dir, rows[r], first[r], h->d_tri[0], h->d_tri[1], h->d_tri[2], X[r],
                                                sendrow[r], nsys[r], h->d_err); LAUNCHED(h); }
  { int rc = hpbc::line_shift(hs, n, dir, +1, sendrow.data(), recvrow.data(), nsys4.data(), nullptr); if (rc) return rc; }
  EACH_R { HCUR;
    k_mr_stage2<<<grid[r], tpb, 0, h->stream>>>(G, dir, rows[r], first[r], atol, h->d_tri[0], h->d_tri[1], h->d_tri[2], X[r],
                                                recvrow[r], h->d_mr + 8 * nsys[r], nsys[r], h->d_err); LAUNCHED(h); }
  // reduced system: Jacobi (tridiagIterJacobi.c:128-190); members of one line of ranks stop together
  std::vector<double> gnorm(n, 0.0), norm0(n, 0.0);
  std::vector<std::array<double, 64>> hn(n);
  for (int iter = 0; ; iter++) {
    bool any = false;
    EACH_R {
      if (!active[r]) continue;
      if (iter >= maxiter || (iter && evalnorm && gnorm[r] < atol) || (iter && evalnorm && gnorm[r] / norm0[r] < rtol)) active[r] = 0;
      any = any || active[r];
    }
    if (!any) break;
    { int rc = hpbc::line_swap(hs, n, dir, red_sx.data(), red_sx.data(), red_rl.data(), red_rr.data(), nsys.data(), active.data()); if (rc) return rc; }
    if (evalnorm) EACH_R { if (!active[r]) continue; HCUR;
      const long long nb = (long long)grid[r].x * grid[r].y * grid[r].z;
      k_mr_jacobi<<<grid[r], tpb, 0, h->stream>>>(G, dir, first[r], 0, h->d_tri[0], h->d_tri[1], h->d_tri[2], X[r], h->d_mr + 8 * nsys[r],
                                                  nsys[r], dnorm[r] + 128); LAUNCHED(h);
      if (nb > 3900) return hpb_fail(HPB_ERR_INVALID, "compact schemes across ranks: more than 3900 thread blocks per solve");
      k_mr_sum<<<1, 256, 0, h->stream>>>(dnorm[r] + 128, nb, dnorm[r]); LAUNCHED(h); }
    if (evalnorm) { int rc = hpbc::line_gather(hs, n, dir, dnorm.data(), dgath.data(), reinterpret_cast<double (*)[64]>(hn.data()), active.data()); if (rc) return rc; }
    EACH_R { if (!active[r]) continue;
      if (evalnorm) {
        const int np = hs[r]->cfg.iproc[dir];
        double sum = hn[r][0];                                   // the order of a rank-0-rooted reduction: ((r0 + r1) + r2) ...
        for (int k = 1; k < np; k++) sum += hn[r][k];
        gnorm[r] = sqrt(sum / np);                               // NT = sum of the local sizes = ranks on the line (:118)
        if (!iter) norm0[r] = gnorm[r];
      }
      HCUR;
      k_mr_jacobi<<<grid[r], tpb, 0, h->stream>>>(G, dir, first[r], 1, h->d_tri[0], h->d_tri[1], h->d_tri[2], X[r], h->d_mr + 8 * nsys[r],
                                                  nsys[r], nullptr); LAUNCHED(h); }
  }
  // each rank gets the first x of the next one (tridiagLU.c:228-233), back substitution, the shared interface
  { int rc = hpbc::line_shift(hs, n, dir, -1, red_x.data(), xp1.data(), nsys.data(), nullptr); if (rc) return rc; }
  EACH_R { HCUR;
    k_mr_stage4<<<grid[r], tpb, 0, h->stream>>>(G, dir, rows[r], first[r], h->d_tri[0], h->d_tri[1], h->d_tri[2], X[r], xp1[r],
                                                sfirst[r], nsys[r], h->d_err); LAUNCHED(h); }
  { int rc = hpbc::line_shift(hs, n, dir, -1, sfirst.data(), rfirst.data(), nsys.data(), nullptr); if (rc) return rc; }
  EACH_R { if (last[r]) continue; HCUR;
    k_mr_put_last<<<grid[r], tpb, 0, h->stream>>>(G, dir, X[r], rfirst[r]); LAUNCHED(h); }
#undef EACH_R
#undef HCUR
  return HPB_OK;
}

// the same for the block systems of the characteristic compact schemes (blocktridiagLU.c with the block Jacobi reduced
// solve): rows in d_tri[0..2], right-hand side / solution in d_bx, result into the interface arrays fI[r]
static int block_compact_solve_group(hpb_solver** hs, int n, int dir, double* const* fI)
{
  const int tpb = 64;
  // lusolver.inp (tridiagLUInit.c:54-90; defaults jacobi, maxiter 10, atol 1e-12, rtol 1e-10, norm evaluated)
  const double atol = hs[0]->cfg.lu_atol, rtol = hs[0]->cfg.lu_rtol;
  const int maxiter = hs[0]->cfg.lu_maxiter;
  const bool evalnorm = hs[0]->cfg.lu_evaluate_norm != 0;
  std::vector<double*> sendrow(n), recvrow(n), red(n), red_sx(n), red_rl(n), red_rr(n), xp1(n), xs1(n), sfirst(n), rfirst(n), dnorm(n), dgath(n);
  std::vector<int> active(n, 1), first(n), last(n), rows(n);
  std::vector<dim3> grid(n);
  std::vector<long long> nline(n), nrow(n), nvec(n);
  for (int r = 0; r < n; r++) {
    hpb_solver* h = hs[r];
    const Geom& G = h->geo;
    const int N = G.nvars;
    long long m = 0;
    for (int d = 0; d < G.ndims; d++) { const long long k = mr_nsys(h, d) / N; if (k > m) m = k; }
    const long long per = 2LL * (3 * N * N + N) + 9LL * N;       // send row, receive row, 6 reduced vectors, xs1, first-row send / receive
    if (!h->d_bmr) {
      cudaSetDevice(h->device);
      const size_t bytes = (size_t)(per * m + 8192) * sizeof(double);
      if (cudaMalloc((void**)&h->d_bmr, bytes) != cudaSuccess) return hpb_fail(HPB_ERR_ALLOC, "characteristic compact schemes across ranks: scratch allocation failed");
      cudaMemset(h->d_bmr, 0, bytes);
      cudaStreamSynchronize(cudaStreamLegacy);
      if (!h->h_mr && cudaMallocHost((void**)&h->h_mr, 128 * sizeof(double)) != cudaSuccess) return hpb_fail(HPB_ERR_ALLOC, "pinned alloc");
    }
    const int M[3] = { G.N[0] + (dir == 0), G.N[1] + (dir == 1), G.N[2] + (dir == 2) };
    int T0, T1;
    if (dir == 0) { T0 = M[1]; T1 = M[2]; } else if (dir == 1) { T0 = M[0]; T1 = M[2]; } else { T0 = M[0]; T1 = M[1]; }
    grid[r] = dim3((T0 + tpb - 1) / tpb, T1, 1);
    nline[r] = (long long)T0 * T1; nrow[r] = (3LL * N * N + N) * nline[r]; nvec[r] = (long long)N * nline[r];
    first[r] = G.lo_phys[dir]; last[r] = G.hi_phys[dir];
    rows[r] = G.N[dir] + (last[r] ? 1 : 0);
    double* p = h->d_bmr;
    sendrow[r] = p; p += nrow[r]; recvrow[r] = p; p += nrow[r];
    red[r] = p; red_sx[r] = p + 2 * nvec[r]; red_rl[r] = p + 3 * nvec[r]; red_rr[r] = p + 4 * nvec[r]; xp1[r] = p + 5 * nvec[r]; p += 6 * nvec[r];
    xs1[r] = p; p += nvec[r]; sfirst[r] = p; p += nvec[r]; rfirst[r] = p; p += nvec[r];
    dnorm[r] = h->d_bmr + per * m; dgath[r] = dnorm[r] + 8;        // + partial sums from dnorm + 128 on
  }
#define EACH_R for (int r = 0; r < n; r++)
#define HCUR hpb_solver* h = hs[r]; cudaSetDevice(h->device); const Geom& G = h->geo; (void)G
  EACH_R { HCUR;
    cudaMemsetAsync(red_rl[r], 0, 3 * nvec[r] * sizeof(double), h->stream);     // recvL, recvR, xp1
#define CALL(M_) k_bmr_stage1<M_><<<grid[r], tpb, 0, h->stream>>>(G, dir, rows[r], first[r], h->d_tri[0], h->d_tri[1], h->d_tri[2], h->d_bx, sendrow[r])
    MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
    LAUNCHED(h); }
  { int rc = hpbc::line_shift(hs, n, dir, +1, sendrow.data(), recvrow.data(), nrow.data(), nullptr); if (rc) return rc; }
  EACH_R { HCUR;
#define CALL(M_) k_bmr_stage2<M_><<<grid[r], tpb, 0, h->stream>>>(G, dir, rows[r], first[r], h->d_tri[0], h->d_tri[1], h->d_tri[2], h->d_bx, recvrow[r], red[r])
    MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
    LAUNCHED(h); }
  std::vector<double> gnorm(n, 0.0), norm0(n, 0.0);
  std::vector<std::array<double, 64>> hn(n);
  for (int iter = 0; ; iter++) {
    bool any = false;
    EACH_R {
      if (!active[r]) continue;
      if (iter >= maxiter || (iter && evalnorm && gnorm[r] < atol) || (iter && evalnorm && gnorm[r] / norm0[r] < rtol)) active[r] = 0;
      any = any || active[r];
    }
    if (!any) break;
    { int rc = hpbc::line_swap(hs, n, dir, red_sx.data(), red_sx.data(), red_rl.data(), red_rr.data(), nvec.data(), active.data()); if (rc) return rc; }
    if (evalnorm) EACH_R { if (!active[r]) continue; HCUR;
      const long long nb = (long long)grid[r].x * grid[r].y;
      if (nb > 3900) return hpb_fail(HPB_ERR_INVALID, "compact schemes across ranks: more than 3900 thread blocks per solve");
#define CALL(M_) k_bmr_jacobi<M_><<<grid[r], tpb, 0, h->stream>>>(G, dir, first[r], 0, h->d_tri[0], h->d_tri[1], h->d_tri[2], h->d_bx, red[r], dnorm[r] + 128)
      MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
      LAUNCHED(h);
      k_mr_sum<<<1, 256, 0, h->stream>>>(dnorm[r] + 128, nb, dnorm[r]); LAUNCHED(h); }
    if (evalnorm) { int rc = hpbc::line_gather(hs, n, dir, dnorm.data(), dgath.data(), reinterpret_cast<double (*)[64]>(hn.data()), active.data()); if (rc) return rc; }
    EACH_R { if (!active[r]) continue;
      if (evalnorm) {
        const int np = hs[r]->cfg.iproc[dir];
        double sum = hn[r][0];
        for (int k = 1; k < np; k++) sum += hn[r][k];
        gnorm[r] = sqrt(sum / np);                               // NT = ranks on the line (blocktridiagIterJacobi.c:139)
        if (!iter) norm0[r] = gnorm[r];
      }
      HCUR;
#define CALL(M_) k_bmr_jacobi<M_><<<grid[r], tpb, 0, h->stream>>>(G, dir, first[r], 1, h->d_tri[0], h->d_tri[1], h->d_tri[2], h->d_bx, red[r], nullptr)
      MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
      LAUNCHED(h); }
  }
  EACH_R { HCUR;
#define CALL(M_) k_bmr_row0<M_><<<grid[r], tpb, 0, h->stream>>>(G, dir, h->d_bx, xs1[r])
    MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
    LAUNCHED(h); }
  { int rc = hpbc::line_shift(hs, n, dir, -1, xs1.data(), xp1.data(), nvec.data(), nullptr); if (rc) return rc; }
  EACH_R { HCUR;
#define CALL(M_) k_bmr_stage4<M_><<<grid[r], tpb, 0, h->stream>>>(G, dir, rows[r], first[r], h->d_tri[0], h->d_tri[1], h->d_tri[2], h->d_bx, xp1[r], sfirst[r], fI[r])
    MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
    LAUNCHED(h); }
  { int rc = hpbc::line_shift(hs, n, dir, -1, sfirst.data(), rfirst.data(), nvec.data(), nullptr); if (rc) return rc; }
  EACH_R { if (last[r]) continue; HCUR;
#define CALL(M_) k_bmr_put_last<M_><<<grid[r], tpb, 0, h->stream>>>(G, dir, fI[r], rfirst[r])
    MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
    LAUNCHED(h); }
#undef EACH_R
#undef HCUR
  return HPB_OK;
}

static int weno_interp_group(hpb_solver** hs, int n, double* const* fI, double* const* fC, const double* const* u,
                             double* const* w, int upw, int dir, int uflag)
{
  for (int r = 0; r < n; r++) { cudaSetDevice(hs[r]->device); weno_interp(hs[r], fI[r], fC[r], u[r], w[r], upw, dir, uflag); }
  const hpb_solver* h0 = hs[0];
  const bool compact = hpb_scheme_is_compact(h0->cfg.hyp_scheme);
  if (compact && h0->cfg.iproc[dir] > 1)
    return h0->phys.interp_char ? block_compact_solve_group(hs, n, dir, fI) : compact_solve_group(hs, n, dir, fI);
  return HPB_OK;
}

// hyperbolic_pieces for a group of ranks (the reconstructions of a compact scheme couple them)
int hyperbolic_pieces_group(hpb_solver** hs, int n, const double* const* u, double* const* out, bool negate, bool with_source,
                            double* const* src)
{
  const int nd = hs[0]->geo.ndims;
  std::vector<double*> fC(n), uC(n), fL(n), fR(n), uL(n), uR(n), w(n);
  std::vector<long long> wo(n, 0);
  for (int d = 0; d < nd; d++) {
    for (int r = 0; r < n; r++) {
      hpb_solver* h = hs[r];
      cudaSetDevice(h->device);
      const Geom& G = h->geo;
      const long long ni = (long long)(G.N[0] + (d == 0)) * (G.N[1] + (d == 1)) * (G.N[2] + (d == 2));
      fC[r] = h->d_cell[0]; uC[r] = h->d_cell[1];
      fL[r] = h->d_iface[1]; fR[r] = h->d_iface[2]; uL[r] = h->d_iface[3]; uR[r] = h->d_iface[4];
      w[r] = h->d_w + wo[r];
      wo[r] += 12 * ni * G.nvars;
      flux(h, u[r], fC[r], d);
      if (hpb_scheme_has_weights(h->cfg.hyp_scheme)) { weno_weights(h, fC[r], u[r], d, w[r]); h->w_valid = true; }
      modified_solution(h, u[r], uC[r]);
    }
    int rc;
    if ((rc = weno_interp_group(hs, n, uL.data(), uC.data(), u, w.data(),  1, d, 1))) return rc;
    if ((rc = weno_interp_group(hs, n, uR.data(), uC.data(), u, w.data(), -1, d, 1))) return rc;
    if ((rc = weno_interp_group(hs, n, fL.data(), fC.data(), u, w.data(),  1, d, 0))) return rc;
    if ((rc = weno_interp_group(hs, n, fR.data(), fC.data(), u, w.data(), -1, d, 0))) return rc;
    for (int r = 0; r < n; r++) {
      hpb_solver* h = hs[r];
      cudaSetDevice(h->device);
      const Geom& G = h->geo;
      upwind(h, h->d_fI, fL[r], fR[r], uL[r], uR[r], u[r], d);
      const int mode = negate ? (d == 0 ? 0 : 1) : (d == 0 ? 2 : 3);
      k_divergence<<<grid3(G.N[0], G.N[1], G.N[2]), TPB, 0, h->stream>>>(G, h->d_dxinv, h->d_fI, d, out[r], mode); LAUNCHED(h);
    }
    const bool grav = hs[0]->phys.has_grav && with_source && hs[0]->phys.grav[d] != 0.0;
    if (grav) {
      for (int r = 0; r < n; r++) {
        hpb_solver* h = hs[r];
        cudaSetDevice(h->device);
        const Geom& G = h->geo;
        k_ns3d_source_fn<<<(unsigned)((G.npg + 255) / 256), 256, 0, h->stream>>>(G, h->d_gravg, d, fC[r]); LAUNCHED(h);
      }
      if ((rc = weno_interp_group(hs, n, fL.data(), fC.data(), u, w.data(),  1, d, 0))) return rc;
      if ((rc = weno_interp_group(hs, n, fR.data(), fC.data(), u, w.data(), -1, d, 0))) return rc;
      for (int r = 0; r < n; r++) {
        hpb_solver* h = hs[r];
        cudaSetDevice(h->device);
        const Geom& G = h->geo;
        const long long ni = (long long)(G.N[0] + (d == 0)) * (G.N[1] + (d == 1)) * (G.N[2] + (d == 2));
        k_ns3d_source_avg<<<(unsigned)((ni + 255) / 256), 256, 0, h->stream>>>(ni, d, G.nvars - 1, fL[r], fR[r], h->d_sI); LAUNCHED(h);
        k_ns3d_source<<<grid3(G.N[0], G.N[1], G.N[2]), TPB, 0, h->stream>>>(G, h->phys, h->d_dxinv, u[r], h->d_gravf, h->d_sI, d, src[r]);
        LAUNCHED(h);
      }
    }
  }
  return HPB_OK;
}

int tridiag_error(hpb_solver* h)
{
  if (!h->d_err) return 0;
  int e = 0;
  cudaMemcpyAsync(&e, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
  if (e) cudaMemsetAsync(h->d_err, 0, sizeof(int), h->stream);
  return e;
}

void parabolic_phase1(hpb_solver* h, const double* u)
{
  ProfScope ps(h, HPB_PROF_VISCOUS);
  const Geom& G = h->geo;
  for (int d = 0; d < G.ndims; d++) {
    int B[3] = { G.N[0], G.N[1], G.N[2] };
    B[d] = G.P[d];
    if (h->cfg.model == HPB_MODEL_NS3D)
      k_ns_qderiv<HPB_MODEL_NS3D><<<grid3(B[0], B[1], B[2]), TPB, 0, h->stream>>>(G, h->phys.gamma, h->d_dxinv, u, d, h->d_QD[d]);
    else
      k_ns_qderiv<HPB_MODEL_NS2D><<<grid3(B[0], B[1], B[2]), TPB, 0, h->stream>>>(G, h->phys.gamma, h->d_dxinv, u, d, h->d_QD[d]);
    LAUNCHED(h);
  }
}

void parabolic_phase2(hpb_solver* h, const double* u, double* out, bool accumulate)
{
  ProfScope ps(h, HPB_PROF_VISCOUS);
  const Geom& G = h->geo;
  if (!accumulate) set_zero(h, out, G.npg * G.nvars);
  for (int d = 0; d < G.ndims; d++) {
    int B[3] = { G.N[0], G.N[1], G.N[2] };
    B[d] += 4;
    if (h->cfg.model == HPB_MODEL_NS3D)
      k_ns_fviscous<HPB_MODEL_NS3D><<<grid3(B[0], B[1], B[2]), TPB, 0, h->stream>>>(G, h->phys, h->d_dxinv, u, h->d_QD[0], h->d_QD[1], h->d_QD[2], d, h->d_FV);
    else
      k_ns_fviscous<HPB_MODEL_NS2D><<<grid3(B[0], B[1], B[2]), TPB, 0, h->stream>>>(G, h->phys, h->d_dxinv, u, h->d_QD[0], h->d_QD[1], nullptr, d, h->d_FV);
    LAUNCHED(h);
    k_ns_par_accum<<<grid3(G.N[0], G.N[1], G.N[2]), TPB, 0, h->stream>>>(G, h->d_dxinv, h->d_FV, d, out); LAUNCHED(h);
  }
}

void parabolic_nc1(hpb_solver* h, const double* u, double* out, bool accumulate)
{
  const Geom& G = h->geo;
  if (!accumulate) set_zero(h, out, G.npg * G.nvars);
  bool any = false;
  for (int i = 0; i < G.ndims * G.nvars; i++) any = any || (h->phys.diff[i] != 0.0);
  if (h->cfg.model != HPB_MODEL_LINEAR_ADR || !any) return;   // identically zero term
  k_linadr_par<<<grid3(G.N[0], G.N[1], G.N[2]), TPB, 0, h->stream>>>(G, h->phys, h->d_dxinv, u, out); LAUNCHED(h);
}

void rk_stage(hpb_solver* h, int stage)
{
  ProfScope ps(h, HPB_PROF_RK);
  const long long n = h->geo.npg * h->geo.nvars;
  RKArgs a; a.n = 0;
  for (int i = 0; i < stage; i++) {
    // the reference adds every i < stage, including zero coefficients (adds 0*k: a no-op on finite data)
    const double c = h->cfg.dt * h->rk.A[stage * h->rk.ns + i];
    if (c == 0.0) continue;
    a.k[a.n] = h->d_Udot[i]; a.a[a.n] = c; a.n++;
  }
  k_rk_combine<<<grid1(n), 256, 0, h->stream>>>(h->d_u, a, h->d_U, n); LAUNCHED(h);
}

void rk_finish(hpb_solver* h)
{
  ProfScope ps(h, HPB_PROF_RK);
  const long long n = h->geo.npg * h->geo.nvars;
  RKArgs a; a.n = h->rk.ns;
  for (int s = 0; s < h->rk.ns; s++) { a.k[s] = h->d_Udot[s]; a.a[s] = h->cfg.dt * h->rk.b[s]; }
  k_rk_combine<<<grid1(n), 256, 0, h->stream>>>(h->d_u, a, h->d_u, n); LAUNCHED(h);
}

// ---- GLM-GEE (TimeGLMGEE.c:45-155; r = 2: the solution and one auxiliary solution)
void glm_aux_init(hpb_solver* h)
{
  // TimeInitialize.c:156-169: the auxiliary solution starts as a copy of u (yyt) or as zero (yeps)
  const long long n = h->geo.npg * h->geo.nvars;
  if (h->rk.mode == HPB_GLM_YYT) { k_copy<<<grid1(n), 256, 0, h->stream>>>(h->d_aux, h->d_u, n); LAUNCHED(h); }
  else set_zero(h, h->d_aux, n);
  h->aux_valid = true;
}

static GLMArgs glm_args(const hpb_solver* h, const double* row, int count, double c0, double c1)
{
  GLMArgs a; a.n = 0; a.c0 = c0; a.c1 = c1;
  for (int i = 0; i < count; i++) {
    const double c = h->cfg.dt * row[i];
    if (c == 0.0) continue;              // the reference adds 0 * k as well: a no-op on finite data
    a.k[a.n] = h->d_Udot[i]; a.a[a.n] = c; a.n++;
  }
  return a;
}

void glm_stage(hpb_solver* h, int j)
{
  ProfScope ps(h, HPB_PROF_RK);
  const long long n = h->geo.npg * h->geo.nvars;
  const int s = h->rk.ns;
  const GLMArgs a = glm_args(h, h->rk.A + j * s, j, h->rk.C[2 * j], h->rk.C[2 * j + 1]);     // TimeGLMGEE.c:70-79
  k_glm_combine<<<grid1(n), 256, 0, h->stream>>>(h->d_u, h->d_aux, a, h->d_U, n); LAUNCHED(h);
}

void glm_finish(hpb_solver* h)
{
  ProfScope ps(h, HPB_PROF_RK);
  const long long n = h->geo.npg * h->geo.nvars;
  const int s = h->rk.ns;
  // TimeGLMGEE.c:116-129: both new quantities from the OLD solution and auxiliary solution, then :143-151
  const GLMArgs a0 = glm_args(h, h->rk.b, s, h->rk.D[0], h->rk.D[1]);
  const GLMArgs a1 = glm_args(h, h->rk.b1, s, h->rk.D[2], h->rk.D[3]);
  k_glm_combine<<<grid1(n), 256, 0, h->stream>>>(h->d_u, h->d_aux, a0, h->d_U, n); LAUNCHED(h);
  k_glm_combine<<<grid1(n), 256, 0, h->stream>>>(h->d_u, h->d_aux, a1, h->d_aux2, n); LAUNCHED(h);
  k_swap<<<grid1(n), 256, 0, h->stream>>>(h->d_u, h->d_U, n); LAUNCHED(h);       // d_U keeps the previous solution (step norm)
  double* t = h->d_aux; h->d_aux = h->d_aux2; h->d_aux2 = t;
}

void glm_error_fields(hpb_solver* h, const double* uex, double* est, double* dif)
{
  const long long n = h->geo.npg * h->geo.nvars;
  k_glm_error_fields<<<grid1(n), 256, 0, h->stream>>>(h->d_u, h->d_aux, uex, h->rk.mode, (1.0 / (1.0 - h->rk.gamma)), est, dif, n);
  LAUNCHED(h);
}

void cfl(hpb_solver* h, const double* u, double dt, double* out_host)
{
  const Geom& G = h->geo;
  cudaMemsetAsync(h->d_red, 0, sizeof(double), h->stream);
#define CALL(M_) k_cfl<M_><<<grid3(G.N[0], G.N[1], G.N[2]), TPB, 0, h->stream>>>(G, h->phys, h->d_dxinv, u, dt, h->d_red)
  MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
  LAUNCHED(h);
  cudaMemcpyAsync(h->h_red, h->d_red, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
  *out_host = h->h_red[0];
}

void step_norm_sumsq(hpb_solver* h, double* out_host)
{
  const Geom& G = h->geo;
  RKArgs a; a.n = h->rk.ns;
  for (int s = 0; s < h->rk.ns; s++) { a.k[s] = h->d_Udot[s]; a.a[s] = h->cfg.dt * h->rk.b[s]; }
  k_step_norm_partial<<<DIAG_BLOCKS, DIAG_TPB, 0, h->stream>>>(G, a, h->d_part); LAUNCHED(h);
  k_diag_final<<<1, DIAG_TPB, 0, h->stream>>>(h->d_part, DIAG_BLOCKS, -1, h->d_red); LAUNCHED(h);
  cudaMemcpyAsync(h->h_red, h->d_red, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
  *out_host = h->h_red[0];
}

void sumsq_diff(hpb_solver* h, const double* a, const double* b, double* out_host)
{
  const Geom& G = h->geo;
  cudaMemsetAsync(h->d_red, 0, sizeof(double), h->stream);
  k_sumsq_diff<<<grid3(G.N[0], G.N[1], G.N[2]), TPB, 0, h->stream>>>(G, a, b, h->d_red); LAUNCHED(h);
  cudaMemcpyAsync(h->h_red, h->d_red, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
  *out_host = h->h_red[0];
}

void flux(hpb_solver* h, const double* u, double* f, int dir)
{
  const Geom& G = h->geo;
#define CALL(M_) k_flux<M_><<<(unsigned)((G.npg + 255) / 256), 256, 0, h->stream>>>(G, h->phys, u, dir, f)
  MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
  LAUNCHED(h);
}

void modified_solution(hpb_solver* h, const double* u, double* uC)
{
  const Geom& G = h->geo;
  const double* gf = (h->cfg.model == HPB_MODEL_LINEAR_ADR) ? nullptr : h->d_gravf;
  const double* gg = (h->cfg.model == HPB_MODEL_LINEAR_ADR) ? nullptr : h->d_gravg;
#define CALL(M_) k_modified<M_><<<(unsigned)((G.npg + 255) / 256), 256, 0, h->stream>>>(G, h->phys, u, gf, gg, uC)
  MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
  LAUNCHED(h);
}

void weno_weights(hpb_solver* h, const double* fC, const double* u, int dir, double* w)
{
  const Geom& G = h->geo;
  const int M[3] = { G.N[0] + (dir == 0), G.N[1] + (dir == 1), G.N[2] + (dir == 2) };
#define CALL(M_) k_weights<M_><<<grid3(M[0], M[1], M[2]), TPB, 0, h->stream>>>(G, h->phys, fC, u, dir, w)
  MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
  LAUNCHED(h);
}

void weno_interp(hpb_solver* h, double* fI, const double* fC, const double* u, const double* w, int upw, int dir, int uflag)
{
  const Geom& G = h->geo;
  const int M[3] = { G.N[0] + (dir == 0), G.N[1] + (dir == 1), G.N[2] + (dir == 2) };
  if (h->cfg.hyp_scheme >= HPB_SCHEME_UPW5 && !hpb_scheme_is_compact(h->cfg.hyp_scheme)) {
#define CALL(M_) k_interp_upw5<M_><<<grid3(M[0], M[1], M[2]), TPB, 0, h->stream>>>(G, h->phys, fC, u, upw, dir, fI)
    MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
    LAUNCHED(h);
    return;
  }
  if (hpb_scheme_is_compact(h->cfg.hyp_scheme) && h->phys.interp_char) {
    // characteristic variant: block rows, then one thread per grid line (block tridiagonal system)
    int T0, T1;
    if (dir == 0) { T0 = M[1]; T1 = M[2]; } else if (dir == 1) { T0 = M[0]; T1 = M[2]; } else { T0 = M[0]; T1 = M[1]; }
    const int tpb = 64;
    const bool split = h->cfg.iproc[dir] > 1;     // the line is split among ranks: block_compact_solve_group finishes the job
#define CALL(M_) { k_compact_rows_char<M_><<<grid3(M[0], M[1], M[2]), TPB, 0, h->stream>>>(G, h->phys, fC, u, w, upw, dir, uflag, \
                       h->d_tri[0], h->d_tri[1], h->d_tri[2], h->d_bx); \
                   if (!split) k_block_tridiag<M_><<<dim3((T0 + tpb - 1) / tpb, T1, 1), tpb, 0, h->stream>>>(G, dir, h->d_tri[0], h->d_tri[1], \
                       h->d_tri[2], h->d_bx, fI); }
    MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
    LAUNCHED(h); if (!split) LAUNCHED(h);
    return;
  }
  if (hpb_scheme_is_compact(h->cfg.hyp_scheme)) {
    // rows, then one thread per (line, component) system; the solution replaces the right-hand side in fI
    k_compact_rows<<<grid3(M[0], M[1], M[2]), TPB, 0, h->stream>>>(G, h->phys, fC, w, upw, dir, uflag,
                                                                    h->d_tri[0], h->d_tri[1], h->d_tri[2], fI); LAUNCHED(h);
    if (h->cfg.iproc[dir] > 1) return;          // the line is split among ranks: compact_solve_group finishes the job
    int T0, T1;
    if (dir == 0) { T0 = M[1]; T1 = M[2]; } else if (dir == 1) { T0 = M[0]; T1 = M[2]; } else { T0 = M[0]; T1 = M[1]; }
    const int tpb = 64;
    k_tridiag<<<dim3((T0 + tpb - 1) / tpb, T1, G.nvars), tpb, 0, h->stream>>>(G, dir, h->d_tri[0], h->d_tri[1], h->d_tri[2],
                                                                              fI, h->d_err); LAUNCHED(h);
    return;
  }
#define CALL(M_) k_interp<M_><<<grid3(M[0], M[1], M[2]), TPB, 0, h->stream>>>(G, h->phys, fC, u, w, upw, dir, uflag, fI)
  MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
  LAUNCHED(h);
}

void upwind(hpb_solver* h, double* fI, const double* fL, const double* fR, const double* uL, const double* uR,
            const double* u, int dir)
{
  const Geom& G = h->geo;
  const int M[3] = { G.N[0] + (dir == 0), G.N[1] + (dir == 1), G.N[2] + (dir == 2) };
  const double* gf = (h->cfg.model == HPB_MODEL_LINEAR_ADR) ? nullptr : h->d_gravf;
  const double* gg = (h->cfg.model == HPB_MODEL_LINEAR_ADR) ? nullptr : h->d_gravg;
#define CALL(M_) k_upwind<M_><<<grid3(M[0], M[1], M[2]), TPB, 0, h->stream>>>(G, h->phys, fL, fR, uL, uR, u, gf, gg, dir, fI)
  MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
  LAUNCHED(h);
}

void first_derivative(hpb_solver* h, double* Df, const double* f, int dir, int nv)
{
  const Geom& G = h->geo;
  int B[3] = { G.N[0], G.N[1], G.N[2] };
  B[dir] = G.P[dir];
  k_first_derivative<<<grid3(B[0], B[1], B[2]), TPB, 0, h->stream>>>(G, f, dir, nv, Df); LAUNCHED(h);
}

void second_derivative(hpb_solver* h, double* D2f, const double* f, int dir, int nv, int order)
{
  const Geom& G = h->geo;
  k_second_derivative<<<grid3(G.N[0], G.N[1], G.N[2]), TPB, 0, h->stream>>>(G, f, dir, nv, order, D2f); LAUNCHED(h);
}


// boundary-face fluxes of direction d of the state u (ghosts of dimension d valid) -> sbi[(2d+f)*nvars + v].
// The faces are evaluated by the per-interface kernel whichever path computes the sweep: 2 N^2 of the (N+1) N^2
// interfaces (0.4 % at 512^3). scratch: h->d_face (nvars * largest face).
void boundary_flux(hpb_solver* h, const double* u, int d, double* sbi)
{
  ProfScope ps(h, HPB_PROF_OTHER);
  const Geom& G = h->geo;
  const double* gf = (h->cfg.model == HPB_MODEL_LINEAR_ADR) ? nullptr : h->d_gravf;
  const double* gg = (h->cfg.model == HPB_MODEL_LINEAR_ADR) ? nullptr : h->d_gravg;
  int M[3] = { G.N[0], G.N[1], G.N[2] };
  M[d] = 1;
  const long long nface = (long long)M[0] * M[1] * M[2];
  if (h->cfg.hyp_scheme != HPB_SCHEME_WENO5) {
    // compact / linear schemes: the interface fluxes of direction d are re-evaluated by the piecewise sequence
    // (a boundary value of a compact scheme depends on the whole line) and the two faces are read from them
    double *fC = h->d_cell[0], *uC = h->d_cell[1];
    long long wo = 0;
    for (int k = 0; k < d; k++) wo += 12LL * (G.N[0] + (k == 0)) * (G.N[1] + (k == 1)) * (G.N[2] + (k == 2)) * G.nvars;
    double* w = h->d_w + wo;
    flux(h, u, fC, d);
    if (hpb_scheme_has_weights(h->cfg.hyp_scheme)) weno_weights(h, fC, u, d, w);
    modified_solution(h, u, uC);
    weno_interp(h, h->d_iface[3], uC, u, w,  1, d, 1);
    weno_interp(h, h->d_iface[4], uC, u, w, -1, d, 1);
    weno_interp(h, h->d_iface[1], fC, u, w,  1, d, 0);
    weno_interp(h, h->d_iface[2], fC, u, w, -1, d, 0);
    upwind(h, h->d_fI, h->d_iface[1], h->d_iface[2], h->d_iface[3], h->d_iface[4], u, d);
    for (int f = 0; f < 2; f++) {
      k_face_from_iface<<<grid3(M[0], M[1], M[2]), TPB, 0, h->stream>>>(G, d, f, h->d_fI, h->d_face); LAUNCHED(h);
      k_face_sum<<<G.nvars, 1024, 0, h->stream>>>(h->d_face, nface, f ? 1.0 : -1.0, sbi + (2 * d + f) * G.nvars); LAUNCHED(h);
    }
    return;
  }
  for (int f = 0; f < 2; f++) {
#define CALL(M_) k_iface<M_><<<grid3(M[0], M[1], M[2]), TPB, 0, h->stream>>>(G, h->phys, u, gf, gg, d, h->d_face, nullptr, f)
    MODEL_SWITCH(h->cfg.model, CALL)
#undef CALL
    LAUNCHED(h);
    k_face_sum<<<G.nvars, 1024, 0, h->stream>>>(h->d_face, nface, f ? 1.0 : -1.0, sbi + (2 * d + f) * G.nvars); LAUNCHED(h);
  }
}

void step_boundary_integral(hpb_solver* h, const double* bf, double* step_bi)
{
  const int nbf = 2 * h->geo.ndims * h->geo.nvars;
  k_step_boundary_integral<<<1, 32, 0, h->stream>>>(bf, h->rk.ns, nbf, h->rk, h->cfg.dt, step_bi); LAUNCHED(h);
}

// out_host[v] = this rank's part of VolumeIntegral (before the sum over ranks)
void volume_integral(hpb_solver* h, const double* u, double* out_host)
{
  const Geom& G = h->geo;
  k_diag_partial<0><<<DIAG_BLOCKS, DIAG_TPB, 0, h->stream>>>(G, h->d_dxinv, u, nullptr, h->d_part); LAUNCHED(h);
  k_diag_final<<<G.nvars, DIAG_TPB, 0, h->stream>>>(h->d_part, DIAG_BLOCKS, -1, h->d_red); LAUNCHED(h);
  cudaMemcpyAsync(h->h_red, h->d_red, G.nvars * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
  for (int v = 0; v < G.nvars; v++) out_host[v] = h->h_red[v];
}

// out_host = (sum |a-b|, sum (a-b)^2, max |a-b|) over this rank's interior points and all components
void diff_norm_sums(hpb_solver* h, const double* a, const double* b, double* out_host)
{
  const Geom& G = h->geo;
  k_diag_partial<1><<<DIAG_BLOCKS, DIAG_TPB, 0, h->stream>>>(G, h->d_dxinv, a, b, h->d_part); LAUNCHED(h);
  k_diag_final<<<3, DIAG_TPB, 0, h->stream>>>(h->d_part, DIAG_BLOCKS, 2, h->d_red); LAUNCHED(h);
  cudaMemcpyAsync(h->h_red, h->d_red, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
  for (int k = 0; k < 3; k++) out_host[k] = h->h_red[k];
}
int diag_partial_size() { return DIAG_BLOCKS * HPB_MAX_NVARS; }

} // namespace hpbk

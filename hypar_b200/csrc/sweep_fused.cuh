// sweep_fused.cuh -- the fused directional sweep: for one sweep direction, ONE kernel goes from the
// conserved variables to the flux divergence (and the well-balanced gravity source):
//
//   flux f_d(u), modified solution uC(u)                 NavierStokes3DFlux.c:24, ...ModifiedSolution.c:31
//   WENO5 weights, L/R-biased x {flux, raw u}            WENOFifthOrderCalculateWeights.c:111-745
//   fL, fR (flux, F-weights), uL, uR (uC, U-weights)     Interp1PrimFifthOrderWENO.c:74-168
//   Rusanov upwinding                                    NavierStokes3DUpwind.c:349-418
//   out -= dxinv (fhat_{j+1} - fhat_j)                   HyperbolicFunction.c:94-109
//   out += gravity source, same F-weights (quirk Q5)     NavierStokes3DSource.c:38-104
//
// without materialising fluxC, uC, the 12 weight arrays or the 5 interface arrays of the reference
// (about 9 kB per point-stage of memory traffic there; here: read u once, update out once).
//
// Work decomposition. A CTA owns TW = 8 grid lines along the sweep direction and a chunk of
// TL - 2 = 30 cells of each line. One thread per "reconstruction cell" j (TL = 32 per line: the 30
// output cells plus one on either side):
//   phase 1  every stencil cell (36 per line) gets its record in shared memory, computed ONCE:
//            u, f_d(u), uC_energy, sqrt(rho), velocity, total enthalpy, c + |v_d|   (17 doubles)
//   phase 2  thread j reads the 5-cell windows of f and u centred on j and reconstructs BOTH values
//            the centred stencil serves: the left-biased value at j+1/2 and the right-biased value
//            at j-1/2. The three smoothness indicators are the same for both (the stencil is only
//            mirrored), so they -- and (beta+eps)^2 and their pair products -- are computed once.
//   phase 3  thread j owns interface j+1/2: its own left-biased values + the right-biased values of
//            thread j+1 (through shared memory) -> Rusanov flux
//   phase 4  thread j differences the interface fluxes j+1/2 (own) and j-1/2 (thread j-1) and
//            updates `out` (and the source) with coalesced stores.
// The x-sweep maps the 32 lanes of a warp along the line (contiguous x); y- and z-sweeps map 8
// consecutive threads to 8 x-contiguous lines, so global accesses are 64-256 B contiguous runs.
//
// Arithmetic. FP64 throughout, FMA contraction on. The weights are evaluated in a division-free
// homogeneous form (one reciprocal per weight set instead of 4 (JS/Z/YC) or 9 (mapped) divisions):
// with q_k = (beta_k+eps)^2 and p_k the product of the other two q, alpha_k = c_k/q_k is proportional
// to n_k = c_k p_k, so  sum_k w_k f_k = (sum_k n_k f_k) / (sum_k n_k). This is algebraically the
// reference's formula; results agree to a few ulp (tests: <= 1e-12 relative per RHS evaluation).
#pragma once
#include "hpb_internal.h"

namespace hpbf {

constexpr int TL = 32;            // reconstruction cells per line chunk (threads along the line)
constexpr int TW = 8;             // lines per CTA
constexpr int NT = TL * TW;       // threads per CTA
constexpr int SC = TL + 4;        // stencil cells per line chunk (j0-3 .. j0+TL)
constexpr int OUTL = TL - 2;      // output cells per line chunk

// weight type as template parameter: HPB_WENO_JS/M/Z/YC, 4 = no_limiting (optimal weights)
constexpr int WT_NOLIM = 4;

struct SweepArgs {
  Geom G;
  Phys ph;
  const double* u;
  const double* gf;      // gravity field f (ghost-padded scalar) or nullptr
  const double* gg;      // gravity field g
  const double* dxinv;   // concatenated, with ghosts
  double* out;           // rhs / hyp accumulator (SoA, ghosts)
  double* src;           // gravity source accumulator (may alias out) or nullptr
  int dir;
  int mode;              // 0: out = -div ; 1: out -= div ; 2: out = +div ; 3: out += div
  int with_source;       // add the gravity source of this direction to src
  int nlines;            // number of grid lines along dir
  const double* qd;      // VISC: scaled derivatives of (u,v,w,T): qd[(dir*4 + comp) * npg + p] (viscous_fused.cu)
};

__device__ __forceinline__ double rcp_fast(double x)
{
  // MUFU.RCP64H seed (relative error 9.8e-7 measured) + one cubic Newton step: relative error <= 2.2e-16
  // over 2^20 random operands across 400 binades (tools/microbench.cu, profiles/r01_microbench.txt).
  // Operands here are sums of positive normal numbers: no zero / inf / denormal handling needed.
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  e = fma(e, e, e);
  return fma(r, e, r);
}

// ------------------------------------------------------------------------------------------
// Both reconstructions a centred 5-cell stencil serves. X = values the weights are computed from,
// Y = values that are reconstructed (Y == X for the flux; Y = uC, X = raw u for the solution: Q4).
// Z (optional) = a second field reconstructed with the same weights (gravity source function: Q5).
// Returns L = left-biased value at j+1/2, R = right-biased value at j-1/2.
template <int WT, bool HASZ>
__device__ __forceinline__ void recon_pair(const double (&X)[5], const double (&Y)[5], const double (&Z)[5],
                                           double eps, double& L, double& R, double& ZL, double& ZR)
{
  const double s6 = 1.0 / 6.0;
  // candidate values: left-biased (at j+1/2) and right-biased (at j-1/2); index k = sub-stencil
  // (j-2,j-1,j), (j-1,j,j+1), (j,j+1,j+2)
  const double fL1 = (2*s6) * Y[0] + (-7*s6) * Y[1] + (11*s6) * Y[2];
  const double fL2 = (-s6) * Y[1] + (5*s6) * Y[2] + (2*s6) * Y[3];
  const double fL3 = (2*s6) * Y[2] + (5*s6) * Y[3] + (-s6) * Y[4];
  const double fR3 = (2*s6) * Y[4] + (-7*s6) * Y[3] + (11*s6) * Y[2];
  const double fR2 = (-s6) * Y[3] + (5*s6) * Y[2] + (2*s6) * Y[1];
  const double fR1 = (2*s6) * Y[2] + (5*s6) * Y[1] + (-s6) * Y[0];
  double zL1 = 0, zL2 = 0, zL3 = 0, zR1 = 0, zR2 = 0, zR3 = 0;
  if (HASZ) {
    zL1 = (2*s6) * Z[0] + (-7*s6) * Z[1] + (11*s6) * Z[2];
    zL2 = (-s6) * Z[1] + (5*s6) * Z[2] + (2*s6) * Z[3];
    zL3 = (2*s6) * Z[2] + (5*s6) * Z[3] + (-s6) * Z[4];
    zR3 = (2*s6) * Z[4] + (-7*s6) * Z[3] + (11*s6) * Z[2];
    zR2 = (-s6) * Z[3] + (5*s6) * Z[2] + (2*s6) * Z[1];
    zR1 = (2*s6) * Z[2] + (5*s6) * Z[1] + (-s6) * Z[0];
  }
  const double c1 = 0.1, c2 = 0.6, c3 = 0.3;
  if (WT == WT_NOLIM) {
    L = c1 * fL1 + c2 * fL2 + c3 * fL3;
    R = c3 * fR1 + c2 * fR2 + c1 * fR3;     // right-biased: optimal weights mirrored
    if (HASZ) { ZL = c1 * zL1 + c2 * zL2 + c3 * zL3; ZR = c3 * zR1 + c2 * zR2 + c1 * zR3; }
    return;
  }
  // smoothness indicators of the three sub-stencils (shared by both biases)
  const double t12 = 13.0 / 12.0;
  const double d1 = X[0] - 2*X[1] + X[2], e1 = X[0] - 4*X[1] + 3*X[2];
  const double d2 = X[1] - 2*X[2] + X[3], e2 = X[1] - X[3];
  const double d3 = X[2] - 2*X[3] + X[4], e3 = 3*X[2] - 4*X[3] + X[4];
  const double b1 = t12 * d1 * d1 + 0.25 * e1 * e1;
  const double b2 = t12 * d2 * d2 + 0.25 * e2 * e2;
  const double b3 = t12 * d3 * d3 + 0.25 * e3 * e3;
  const double s1 = b1 + eps, s2 = b2 + eps, s3 = b3 + eps;
  const double q1 = s1 * s1, q2 = s2 * s2, q3 = s3 * s3;
  // h_k proportional to 1/q_k (JS, M) or (1 + tau^2/q_k) (Z, YC), common positive factor dropped
  double h1, h2, h3;
  if (WT == HPB_WENO_JS || WT == HPB_WENO_M) {
    h1 = q2 * q3; h2 = q1 * q3; h3 = q1 * q2;
  } else {
    double tau;
    if (WT == HPB_WENO_Z) tau = fabs(b3 - b1);
    else { const double t = X[0] - 4*X[1] + 6*X[2] - 4*X[3] + X[4]; tau = t * t; }
    const double tt = tau * tau;
    h1 = (q1 + tt) * (q2 * q3); h2 = (q2 + tt) * (q1 * q3); h3 = (q3 + tt) * (q1 * q2);
  }
  // un-normalised nonlinear weights; left-biased: optimal (c1,c2,c3); right-biased: (c3,c2,c1)
  double nL1 = c1 * h1, n2 = c2 * h2, nL3 = c3 * h3;
  double nR1 = c3 * h1, nR3 = c1 * h3;
  if (WT != HPB_WENO_M) {
    const double rL = rcp_fast(nL1 + n2 + nL3), rR = rcp_fast(nR1 + n2 + nR3);
    L = (nL1 * fL1 + n2 * fL2 + nL3 * fL3) * rL;
    R = (nR1 * fR1 + n2 * fR2 + nR3 * fR3) * rR;
    if (HASZ) {
      ZL = (nL1 * zL1 + n2 * zL2 + nL3 * zL3) * rL;
      ZR = (nR1 * zR1 + n2 * zR2 + nR3 * zR3) * rR;
    }
    return;
  }
  // mapped weights (Henrick et al.): w~ = JS weights; alpha_k = w~(c + c^2 - 3 c w~ + w~^2)/(c^2 + w~(1-2c)) = A_k/B_k;
  // sum_k w_k f_k = (sum_k A_k B_l B_m f_k) / (sum_k A_k B_l B_m)
  {
    const double rL = rcp_fast(nL1 + n2 + nL3);
    const double w1 = nL1 * rL, w2 = n2 * rL, w3 = nL3 * rL;
    const double A1 = w1 * (c1 + c1*c1 + w1 * (w1 - 3*c1)), B1 = c1*c1 + w1 * (1.0 - 2*c1);
    const double A2 = w2 * (c2 + c2*c2 + w2 * (w2 - 3*c2)), B2 = c2*c2 + w2 * (1.0 - 2*c2);
    const double A3 = w3 * (c3 + c3*c3 + w3 * (w3 - 3*c3)), B3 = c3*c3 + w3 * (1.0 - 2*c3);
    const double t1 = A1 * (B2 * B3), t2 = A2 * (B1 * B3), t3 = A3 * (B1 * B2);
    const double r = rcp_fast(t1 + t2 + t3);
    L = (t1 * fL1 + t2 * fL2 + t3 * fL3) * r;
    if (HASZ) ZL = (t1 * zL1 + t2 * zL2 + t3 * zL3) * r;
  }
  {
    const double rR = rcp_fast(nR1 + n2 + nR3);
    const double w1 = nR1 * rR, w2 = n2 * rR, w3 = nR3 * rR;   // optimal weights (c3, c2, c1)
    const double A1 = w1 * (c3 + c3*c3 + w1 * (w1 - 3*c3)), B1 = c3*c3 + w1 * (1.0 - 2*c3);
    const double A2 = w2 * (c2 + c2*c2 + w2 * (w2 - 3*c2)), B2 = c2*c2 + w2 * (1.0 - 2*c2);
    const double A3 = w3 * (c1 + c1*c1 + w3 * (w3 - 3*c1)), B3 = c1*c1 + w3 * (1.0 - 2*c1);
    const double t1 = A1 * (B2 * B3), t2 = A2 * (B1 * B3), t3 = A3 * (B1 * B2);
    const double r = rcp_fast(t1 + t2 + t3);
    R = (t1 * fR1 + t2 * fR2 + t3 * fR3) * r;
    if (HASZ) ZR = (t1 * zR1 + t2 * zR2 + t3 * zR3) * r;
  }
}

// ------------------------------------------------------------------------------------------
// record layout in shared memory (field-major: rec[field * NCELL + cell])
template <int MODEL> struct RecLayout;
template <> struct RecLayout<HPB_MODEL_LINEAR_ADR> { enum { NV = 1, U = 0, F = 1, NF = 2, V4 = 0, SR = 0, VEL = 0, H = 0, A = 0, GF = 0, GG = 0 }; };
template <> struct RecLayout<HPB_MODEL_NS2D> { enum { NV = 4, U = 0, F = 4, V4 = 8, SR = 9, VEL = 10, H = 12, A = 13, NF = 14, GF = 14, GG = 14 }; };
template <> struct RecLayout<HPB_MODEL_NS3D> { enum { NV = 5, U = 0, F = 5, V4 = 10, SR = 11, VEL = 12, H = 15, A = 16, GF = 17, GG = 18, NF = 17 }; };

template <int MODEL, bool GRAV, bool VISC>
constexpr int rec_fields()
{
  return RecLayout<MODEL>::NF + ((GRAV && MODEL == HPB_MODEL_NS3D) ? 2 : 0) + ((VISC && MODEL == HPB_MODEL_NS3D) ? 4 : 0);
}

// exchange buffers: per reconstruction cell: right-biased values fR, uR (+ sR x2) ; interface flux (+ S x2)
template <int MODEL, bool GRAV>
constexpr int exch_fields() { return 3 * RecLayout<MODEL>::NV + ((GRAV && MODEL == HPB_MODEL_NS3D) ? 4 : 0); }

template <int MODEL, bool GRAV, bool VISC>
constexpr size_t sweep_smem_bytes()
{
  return sizeof(double) * ((size_t)rec_fields<MODEL, GRAV, VISC>() * SC * TW + (size_t)exch_fields<MODEL, GRAV>() * NT);
}

// MAPX: true  -> lanes run along the sweep line (x-sweep);
//       false -> 8 consecutive threads = 8 x-contiguous lines (y-, z-sweeps)
// VISC (NavierStokes3D): the viscous flux of the sweep direction is evaluated per stencil cell from the
// derivative arrays of viscous_fused.cu and its fourth-order central derivative is added to `out`
// (NavierStokes3DParabolicFunction.c:152-208, :210-261, :263-314)
template <int MODEL, int WT, bool MAPX, bool GRAV, bool VISC>
__global__ void __launch_bounds__(NT, 2) k_sweep(const SweepArgs a)
{
  using RL = RecLayout<MODEL>;
  constexpr int NV = RL::NV;
  constexpr bool FLUID = (MODEL != HPB_MODEL_LINEAR_ADR);
  constexpr bool G3 = GRAV && MODEL == HPB_MODEL_NS3D;
  constexpr bool V3 = VISC && MODEL == HPB_MODEL_NS3D;
  constexpr int FVI = RL::NF + (G3 ? 2 : 0);          // record fields of the viscous flux (4)
  constexpr int NCELL = SC * TW;
  extern __shared__ double smem[];
  double* rec = smem;
  double* exR = smem + rec_fields<MODEL, GRAV, VISC>() * NCELL;      // [2*NV (+2)][NT]
  double* exF = exR + (2 * NV + (G3 ? 2 : 0)) * NT;            // [NV (+2)][NT]

  const Geom& G = a.G;
  const int dir = a.dir;
  const int N = G.N[dir];
  const long long st = G.st[dir];
  const int tid = threadIdx.x;
  const int j0 = blockIdx.y * OUTL;                     // first output cell of this chunk
  const int line0 = blockIdx.x * TW;

  // line -> offset of its cell j = 0 in the ghost-padded array
  auto line_base = [&](int line) -> long long {
    int ia, ib;       // the two transverse indices, in increasing dimension order
    long long p;
    if (dir == 0)      { ia = line % G.N[1]; ib = line / G.N[1]; p = G.g + (long long)G.P[0] * ((ia + (G.ndims > 1 ? G.g : 0)) + (long long)G.P[1] * (ib + (G.ndims > 2 ? G.g : 0))); }
    else if (dir == 1) { ia = line % G.N[0]; ib = line / G.N[0]; p = (ia + G.g) + (long long)G.P[0] * (G.g + (long long)G.P[1] * (ib + (G.ndims > 2 ? G.g : 0))); }
    else               { ia = line % G.N[0]; ib = line / G.N[0]; p = (ia + G.g) + (long long)G.P[0] * ((ib + G.g) + (long long)G.P[1] * G.g); }
    return p;
  };
  auto cidx = [&](int l, int w) -> int { return MAPX ? (w * SC + l) : (l * TW + w); };

  const double gamma = a.ph.gamma;

  // ---------------- phase 1: records of the stencil cells
  for (int c = tid; c < NCELL; c += NT) {
    int l, w;
    if (MAPX) { l = c % SC; w = c / SC; } else { w = c % TW; l = c / TW; }
    const int j = j0 - 3 + l;
    const int line = line0 + w;
    if (line >= a.nlines || j > N + 2) continue;
    const long long p = line_base(line) + (long long)j * st;
    const int ci = cidx(l, w);
    double U[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) U[v] = a.u[v * G.npg + p];
#pragma unroll
    for (int v = 0; v < NV; v++) rec[(RL::U + v) * NCELL + ci] = U[v];
    if (!FLUID) {
      rec[RL::F * NCELL + ci] = a.ph.adv[dir] * U[0];
    } else {
      constexpr int NDV = NV - 2;
      const double rho = U[0];
      double vel[3] = { 0.0, 0.0, 0.0 };
      const double rinv = 1.0 / rho;
#pragma unroll
      for (int k = 0; k < NDV; k++) vel[k] = (rho == 0) ? 0.0 : U[1 + k] * rinv;
      double vsq = 0.0;
#pragma unroll
      for (int k = 0; k < NDV; k++) vsq += vel[k] * vel[k];
      const double e = U[NV - 1];
      const double ke = 0.5 * rho * vsq;
      const double P = (e - ke) * (gamma - 1.0);
      const double vn = (dir == 0) ? vel[0] : (dir == 1 ? vel[1] : vel[2]);
      rec[(RL::F + 0) * NCELL + ci] = rho * vn;
#pragma unroll
      for (int k = 0; k < NDV; k++) rec[(RL::F + 1 + k) * NCELL + ci] = rho * vn * vel[k] + (k == dir ? P : 0.0);
      rec[(RL::F + NV - 1) * NCELL + ci] = (e + P) * vn;
      double gfv = 1.0, ggv = 1.0;
      if (G3) { gfv = a.gf[p]; ggv = a.gg[p]; rec[RL::GF * NCELL + ci] = gfv; rec[RL::GG * NCELL + ci] = ggv; }
      const double igm1 = 1.0 / (gamma - 1.0);
      rec[RL::V4 * NCELL + ci] = G3 ? ((P * igm1) * (1.0 / ggv) + ke * gfv) : (P * igm1 + ke);
      const double c2 = gamma * P * rinv;
      rec[RL::SR * NCELL + ci] = sqrt(rho);
#pragma unroll
      for (int k = 0; k < NDV; k++) rec[(RL::VEL + k) * NCELL + ci] = vel[k];
      rec[RL::H * NCELL + ci] = 0.5 * vsq + c2 * igm1;
      rec[RL::A * NCELL + ci] = sqrt(c2) + fabs(vn);
      if (V3 && l >= 1 && l <= TL + 2) {
        // viscous flux of direction dir at this cell: mu = T^0.76 (raiseto, math_ops.h:37), T = gamma P / rho
        const double T = gamma * P / rho;
        const double mu = exp(0.76 * log(T));
        const double two_third = 2.0 / 3.0;
        const double muRe = mu * (1.0 / a.ph.Re);
        const double kq = muRe * igm1 * (1.0 / a.ph.Pr);
        const long long n = G.npg;
        const double* qx = a.qd + p;              // + (0*4 + c) * n
        const double* qy = a.qd + 4 * n + p;
        const double* qz = a.qd + 8 * n + p;
        double t1, t2, t3, q;
        if (dir == 0) {
          const double ux = qx[0], vx = qx[n], wx = qx[2*n], Tx = qx[3*n];
          const double uy = qy[0], vy = qy[n], uz = qz[0], wz = qz[2*n];
          t1 = two_third * muRe * (2 * ux - vy - wz); t2 = muRe * (uy + vx); t3 = muRe * (uz + wx); q = kq * Tx;
        } else if (dir == 1) {
          const double uy = qy[0], vy = qy[n], wy = qy[2*n], Ty = qy[3*n];
          const double ux = qx[0], vx = qx[n], vz = qz[n], wz = qz[2*n];
          t1 = muRe * (uy + vx); t2 = two_third * muRe * (-ux + 2 * vy - wz); t3 = muRe * (vz + wy); q = kq * Ty;
        } else {
          const double uz = qz[0], vz = qz[n], wz = qz[2*n], Tz = qz[3*n];
          const double ux = qx[0], wx = qx[2*n], vy = qy[n], wy = qy[2*n];
          t1 = muRe * (uz + wx); t2 = muRe * (vz + wy); t3 = two_third * muRe * (-ux - vy + 2 * wz); q = kq * Tz;
        }
        rec[(FVI + 0) * NCELL + ci] = t1;
        rec[(FVI + 1) * NCELL + ci] = t2;
        rec[(FVI + 2) * NCELL + ci] = t3;
        rec[(FVI + 3) * NCELL + ci] = vel[0] * t1 + vel[1] * t2 + vel[2] * t3 + q;
      }
    }
  }
  __syncthreads();

  // ---------------- phase 2: both reconstructions of the centred stencil of cell j
  int l, w;
  if (MAPX) { l = tid % TL; w = tid / TL; } else { w = tid % TW; l = tid / TW; }
  const int j = j0 - 1 + l;                 // reconstruction cell
  const int line = line0 + w;
  const bool line_ok = line < a.nlines;
  const bool rc_ok = line_ok && j <= N;     // j >= -1 always
  const int cc = cidx(l + 2, w);            // record index of cell j
  const int cs = MAPX ? 1 : TW;             // record stride along the line
  double fLv[NV], uLv[NV], sL[2] = { 0.0, 0.0 };
  if (rc_ok) {
    double Zg[5] = { 0, 0, 0, 0, 0 };
    if (G3) {
#pragma unroll
      for (int k = 0; k < 5; k++) Zg[k] = rec[RL::GG * NCELL + cc + (k - 2) * cs];
    }
#pragma unroll
    for (int v = 0; v < NV; v++) {
      double X[5], Y[5], R, zl, zr;
#pragma unroll
      for (int k = 0; k < 5; k++) X[k] = rec[(RL::F + v) * NCELL + cc + (k - 2) * cs];
      const bool zsrc = G3 && a.with_source && (v == dir + 1 || v == NV - 1);
      if (G3 && zsrc) {
        recon_pair<WT, true>(X, X, Zg, a.ph.eps, fLv[v], R, zl, zr);
        const int s = (v == NV - 1) ? 1 : 0;
        sL[s] = zl;
        exR[(2 * NV + s) * NT + tid] = zr;
      } else {
        recon_pair<WT, false>(X, X, X, a.ph.eps, fLv[v], R, zl, zr);
      }
      exR[v * NT + tid] = R;
      // solution: weights from raw u, applied to the modified solution (Q4)
#pragma unroll
      for (int k = 0; k < 5; k++) X[k] = rec[(RL::U + v) * NCELL + cc + (k - 2) * cs];
      if (FLUID && (G3 || v == NV - 1)) {
#pragma unroll
        for (int k = 0; k < 5; k++) {
          if (v == NV - 1) Y[k] = rec[RL::V4 * NCELL + cc + (k - 2) * cs];
          else Y[k] = X[k] * rec[RL::GF * NCELL + cc + (k - 2) * cs];
        }
        recon_pair<WT, false>(X, Y, Y, a.ph.eps, uLv[v], R, zl, zr);
      } else {
        recon_pair<WT, false>(X, X, X, a.ph.eps, uLv[v], R, zl, zr);
      }
      exR[(NV + v) * NT + tid] = R;
    }
  }
  __syncthreads();

  // ---------------- phase 3: interface j+1/2 (between cells j and j+1): upwind flux
  const int tnb = MAPX ? tid + 1 : tid + TW;        // thread of cell j+1
  const bool if_ok = rc_ok && (l < TL - 1) && (j + 1 <= N);
  if (if_ok) {
    double fh[NV];
    if (!FLUID) {
      fh[0] = (a.ph.adv[dir] > 0) ? fLv[0] : exR[0 * NT + tnb];
    } else {
      constexpr int NDV = NV - 2;
      const int cL = cc, cR = cc + cs;
      const double tL = rec[RL::SR * NCELL + cL], tR = rec[RL::SR * NCELL + cR];
      const double rs = 1.0 / (tL + tR);
      double vsq = 0.0, vn = 0.0;
#pragma unroll
      for (int k = 0; k < NDV; k++) {
        const double v = (tL * rec[(RL::VEL + k) * NCELL + cL] + tR * rec[(RL::VEL + k) * NCELL + cR]) * rs;
        vsq += v * v;
        if (k == dir) vn = v;
      }
      const double H = (tL * rec[RL::H * NCELL + cL] + tR * rec[RL::H * NCELL + cR]) * rs;
      const double cavg = sqrt((gamma - 1.0) * (H - 0.5 * vsq));
      const double aavg = cavg + fabs(vn);
      double alpha = fmax(fmax(rec[RL::A * NCELL + cL], rec[RL::A * NCELL + cR]), aavg);
      if (G3) alpha *= fmax(rec[RL::GG * NCELL + cL], rec[RL::GG * NCELL + cR]);
#pragma unroll
      for (int v = 0; v < NV; v++) {
        const double fR = exR[v * NT + tnb], uR = exR[(NV + v) * NT + tnb];
        fh[v] = 0.5 * (fLv[v] + fR) - alpha * (0.5 * (uR - uLv[v]));
      }
    }
#pragma unroll
    for (int v = 0; v < NV; v++) exF[v * NT + tid] = fh[v];
    if (G3 && a.with_source) {
      exF[(NV + 0) * NT + tid] = 0.5 * (sL[0] + exR[(2 * NV + 0) * NT + tnb]);
      exF[(NV + 1) * NT + tid] = 0.5 * (sL[1] + exR[(2 * NV + 1) * NT + tnb]);
    }
  }
  __syncthreads();

  // ---------------- phase 4: flux difference of cell j (interfaces j+1/2: own, j-1/2: thread of cell j-1)
  if (line_ok && l >= 1 && l <= OUTL && j < N) {
    const int tpv = MAPX ? tid - 1 : tid - TW;
    const long long p = line_base(line) + (long long)j * st;
    const double dxi = a.dxinv[G.xoff[dir] + G.g + j];
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const double t = dxi * (exF[v * NT + tid] - exF[v * NT + tpv]);
      const long long q = v * G.npg + p;
      if (a.mode == 0) a.out[q] = -t;
      else if (a.mode == 1) a.out[q] -= t;
      else if (a.mode == 2) a.out[q] = t;
      else a.out[q] += t;
    }
    if (V3) {
      // par_v += dxinv * (FV[j-2] - 8 FV[j-1] + 8 FV[j+1] - FV[j+2]) / 12   (components 1..4)
      const double s12 = 1.0 / 12.0;
#pragma unroll
      for (int v = 0; v < 4; v++) {
        const double* f = rec + (FVI + v) * NCELL + cc;
        const double dfv = (f[-2 * cs] - 8 * f[-cs] + 8 * f[cs] - f[2 * cs]) * s12;
        a.out[(1 + v) * G.npg + p] += dxi * dfv;
      }
    }
    if (G3 && a.with_source) {
      // NavierStokes3DSource.c:80-100
      const double rho = rec[RL::U * NCELL + cc];
      const double vd = rec[(RL::VEL + dir) * NCELL + cc];
      const double f = rec[RL::GF * NCELL + cc];
      const double tm = rho * a.ph.RT, te = rho * a.ph.RT * vd;
      a.src[(1 + dir) * G.npg + p] += (tm * f) * (exF[(NV + 0) * NT + tid] - exF[(NV + 0) * NT + tpv]) * dxi;
      a.src[(NV - 1) * G.npg + p]  += (te * f) * (exF[(NV + 1) * NT + tid] - exF[(NV + 1) * NT + tpv]) * dxi;
    }
  }
}

// host-side launcher of one (MODEL, WT) family; defined in sweep_fused_wt.inc per weight type
template <int WT>
bool launch_sweep(hpb_solver* h, const SweepArgs& a);

} // namespace hpbf

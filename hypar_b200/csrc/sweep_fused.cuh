// sweep_fused.cuh -- the fused directional sweep: for one sweep direction, ONE kernel goes from the
// conserved variables to the flux divergence, the viscous flux derivative and the gravity source:
//
//   flux f_d(u), modified solution uC(u)                 NavierStokes3DFlux.c:24, ...ModifiedSolution.c:31
//   WENO5 weights, L/R-biased x {flux, raw u}            WENOFifthOrderCalculateWeights.c:111-745
//   fL, fR (flux, F-weights), uL, uR (uC, U-weights)     Interp1PrimFifthOrderWENO.c:74-168
//   Rusanov upwinding                                    NavierStokes3DUpwind.c:349-418
//   out -= dxinv (fhat_{j+1} - fhat_j)                   HyperbolicFunction.c:94-109
//   out += dxinv D_d(FViscous_d)                         NavierStokes3DParabolicFunction.c:152-314
//   out += gravity source, same F-weights (quirk Q5)     NavierStokes3DSource.c:38-104
//
// without materialising fluxC, uC, the 12 weight arrays, the 5 interface arrays or FViscous/FDeriv of the
// reference (about 9 kB per point-stage of memory traffic there; here: read u (+8 derivative scalars) once,
// update out once).
//
// Work decomposition: a persistent line march, one WARP per grid line. A CTA (8 warps) owns TW = 8 grid lines
// along the sweep direction and marches along them in steps of TL = 32 cells; lane l of warp w works on cell l
// of the step on line w. Everything between loading the inputs and storing the result is private to the warp
// (its own region of shared memory, __syncwarp only); the CTA only cooperates to move data between global
// and shared memory with contiguous accesses (the lines of a CTA are x-adjacent for the y-/z-sweeps).
// Per step m (cells are numbered along the line, -3..N+2 with ghosts):
//   B1  cp.async of step m landed; results of step m-1 staged                                  (CTA barrier)
//   C   cooperative: results of step m-1 -> global (read-modify-write when accumulating; the old values were
//       requested before B1)
//   P1  record of cell 32m+3+l in shared memory, computed ONCE per cell: u, f_d(u), uC_energy, sqrt(rho),
//       velocity, total enthalpy, c+|v_d| (, viscous flux, gravity fields)
//   B2  staging consumed                                                                         (CTA barrier)
//   P0  cp.async prefetch of the raw inputs of step m+1 (u, and for the viscous terms 8 derivative scalars):
//       overlaps the arithmetic of P2..P4
//   P2  lane reconstructs, from the 5-cell windows centred on cell j = 32m+1+l, BOTH values that centred
//       stencil serves: the left-biased value at j+1/2 and the right-biased value at j-1/2. The three
//       smoothness indicators are the same for both (the stencil is only mirrored), so they -- and
//       (beta+eps)^2 and their pair products -- are computed once.
//   P3  lane owns interface j-1/2: its own right-biased values + the left-biased values of cell j-1
//       (neighbour lane through shared memory; across steps through a carry slot) -> Rusanov flux
//   P4  lane forms the update of cell j-1 = 32m+l: difference of the interface fluxes j-1/2 (own) and j-3/2
//       (neighbour / carry), central derivative of the viscous flux, gravity source -> staged for C
//   then the last 5 records and the carry slots are shifted to the front for the next step (warp-private).
// Every record, reconstruction, interface and output is computed exactly once per cell (one mostly idle
// start-up step per line produces the records of the low ghost cells and the first interface).
//
// Arithmetic. FP64 throughout, FMA contraction on. The weights are evaluated in a division-free
// homogeneous form (one reciprocal per weight set instead of 4 (JS/Z/YC) or 9 (mapped) divisions):
// with q_k = (beta_k+eps)^2 and p_k the product of the other two q, alpha_k = c_k/q_k is proportional
// to n_k = c_k p_k, so  sum_k w_k f_k = (sum_k n_k f_k) / (sum_k n_k). This is algebraically the
// reference's formula; results agree to a few ulp (tests: <= 1e-12 relative per RHS evaluation).
#pragma once
#include "hpb_internal.h"

namespace hpbf {

constexpr int TL = 32;            // cells per line per march step
constexpr int TW = 8;             // lines per CTA
constexpr int NT = TL * TW;       // threads per CTA
constexpr int RP = TL + 5;        // record positions per line: cells [32m-2, 32m+35) of step m
constexpr int NREC = RP * TW;     // record slots per field  (index: line * RP + position)
constexpr int SS = TW + 1;        // y-/z-sweeps: staging stride per cell (padded; index: cell * SS + line)
constexpr int NSTG = TL * SS;     // staging slots per field (x-sweep: index line * TL + cell)
constexpr int XP = TL + 1;        // exchange slots per line (slot 0 = carry from the previous step)
constexpr int NEX = XP * TW;      // exchange slots per field (index: line * XP + slot)

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
  double* src;           // gravity source accumulator: must be `out` itself (or nullptr)
  int dir;
  int mode;              // 0: out = -div ; 1: out -= div ; 2: out = +div ; 3: out += div
  int with_source;       // add the gravity source of this direction to src
  int nlines;            // number of grid lines along dir
  const double* qd;      // VISC: (mu/Re) x first differences of (u,v,w,T): qd[(dir*4 + comp) * npg + p] (viscous_fused.cu)
  int upw;               // fluid models: 0 = Rusanov, 1 = Roe with Harten's entropy fix
  int qidx[8];           // VISC: the 8 derivative scalars the viscous flux of `dir` needs (indices dir*4+comp into qd):
                         //   dir 0: ux vx wx Tx | uy vy | uz wz   dir 1: uy vy wy Ty | ux vx | vz wz   dir 2: uz vz wz Tz | ux wx | vy wy
  double* unext;         // last direction only (sweep_tma.cuh, RKF): where the next stage solution ubase + adt * out goes, or nullptr
  const double* ubase;   //   the solution at the start of the step (u^n; the sweep's input u is the stage solution)
  double adt;            //   a_{s+1,s} dt of the explicit RK tableau (TimeRK.c:131-141)
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

__device__ __forceinline__ double sqrt_fast(double x)
{
  // MUFU.RSQ64H seed (relative error ~1e-6) + two coupled Newton steps on (sqrt x, 1 / (2 sqrt x)): the error
  // is squared twice (below 1e-20 before rounding). x > 0 and normal (densities, squared sound speeds).
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-h, g, 0.5);
  return fma(g, r, g);
}

// ------------------------------------------------------------------------------------------
// Roe dissipation R |Lambda| L (uR - uL) of NavierStokes3DUpwind.c:40-125 / NavierStokes2DUpwind.c (Roe), in closed form.
// The reference multiplies the three 5x5 matrices out (MatMult5, MatMult5, MatVecMult5: 275 multiply-adds per interface);
// with the eigenvectors of navierstokes3d.h:288-470 the product collapses to the wave decomposition
//   beta = (gamma-1) (ek drho - v.dm + dE),   dn = vn drho - dm_n
//   w0 = drho - beta / a^2                      entropy wave, speed vn,       r0 = (1, v, ek)
//   w-+ = (beta +- a dn) / (2 a^2)              acoustic waves, speeds vn -+ a, r-+ = (1, v -+ a n, h0 -+ a vn)
//   s_t = dm_t - v_t drho                       shear waves (t != n), speed vn, r_t = (0, e_t, v_t)
// (the signs of the reference's shear rows and columns cancel in the product), ~60 operations. |lambda| carries Harten's
// fix with delta = 1e-6 (:95-101). The Roe-averaged state comes from the sqrt(rho)-weighted cell records the Rusanov
// flux uses too; a^2 = (gamma-1)(H - ek) is what _NavierStokes3DRoeAverage_ + GetFlowVar return to rounding.
// Input du = uR - uL (TWICE the reference's udiff: the caller's flux accumulates 2 x the interface flux).
__device__ __forceinline__ double harten_abs(double lam)
{
  const double delta = 0.000001;
  const double al = fabs(lam);
  return (al < delta) ? (lam * lam + delta * delta) * (0.5 / delta) : al;
}
template <int NV>
__device__ __forceinline__ void roe_dissipation(const double (&du)[NV], const double (&vh)[3], double vn, double vsq, double H,
                                                double gamma, int dir, double (&diss)[NV])
{
  constexpr int NDV = NV - 2;
  const double gm1 = gamma - 1.0;
  const double ek = 0.5 * vsq;
  const double a2 = gm1 * (H - ek);
  double a = sqrt_fast(a2);
  const double ia2 = rcp_fast(a2);
  double vdm = 0.0;
#pragma unroll
  for (int k = 0; k < NDV; k++) vdm = fma(vh[k], du[1 + k], vdm);
  const double beta = gm1 * (fma(ek, du[0], du[NV - 1]) - vdm);
  double dmn = 0.0;
#pragma unroll
  for (int k = 0; k < NDV; k++) if (k == dir) dmn = du[1 + k];
  const double dn = fma(vn, du[0], -dmn);
  const double l0 = harten_abs(vn), lm = harten_abs(vn - a), lp = harten_abs(vn + a);
  const double c0 = l0 * fma(-beta, ia2, du[0]);
  const double hi = 0.5 * ia2;
  const double cm = lm * (hi * fma(a, dn, beta));
  const double cp = lp * (hi * fma(-a, dn, beta));
  const double csum = c0 + cm + cp, cdif = a * (cp - cm);
  diss[0] = csum;
  double shear_e = 0.0;
#pragma unroll
  for (int k = 0; k < NDV; k++) {
    if (k == dir) diss[1 + k] = fma(vh[k], csum, cdif);
    else {
      const double st = l0 * fma(-vh[k], du[0], du[1 + k]);
      diss[1 + k] = fma(vh[k], csum, st);
      shear_e = fma(vh[k], st, shear_e);
    }
  }
  diss[NV - 1] = fma(ek, c0, fma(H, cm + cp, fma(vn, cdif, shear_e)));
}

// ------------------------------------------------------------------------------------------
// Both reconstructions a centred 5-cell stencil serves. X = values the weights are computed from,
// Y = values that are reconstructed (SAME: Y is X -- the flux, and every component of the solution the modified
// solution leaves unchanged; otherwise Y = uC, X = raw u: Q4).
// Z (optional) = a second field reconstructed with the same weights (gravity source function: Q5).
// Returns L = left-biased value at j+1/2, R = right-biased value at j-1/2.
//
// The weights of a set sum to one, so a reconstruction is its CENTRAL candidate plus a weighted correction by the
// two third differences of the stencil, D3a = Y0 - 3 Y1 + 3 Y2 - Y3 and D3b = Y1 - 3 Y2 + 3 Y3 - Y4:
//   left-biased  candidates (Interp1PrimFifthOrderWENO.c:148-158): fL1 - fL2 =  D3a/3 ,  fL3 - fL2 =  D3b/6
//   right-biased candidates (mirrored stencil)                   : fR1 - fR2 = -D3a/6 ,  fR3 - fR2 = -D3b/3
//   L = fL2 + wL1 D3a/3 + wL3 D3b/6 ,   R = fR2 - wR1 D3a/6 - wR3 D3b/3
// (index k = sub-stencil (j-2,j-1,j), (j-1,j,j+1), (j,j+1,j+2) for both biases). With Y = X the third differences
// are differences of the second differences the smoothness indicators need anyway. 6 instead of 18 operations for
// the candidates of a pair; algebraically the reference's formula.
template <int WT, bool HASZ, bool SAME>
__device__ __forceinline__ void recon_pair(const double (&X)[5], const double (&Y)[5], const double (&Z)[5],
                                           double eps, double& L, double& R, double& ZL, double& ZR)
{
  const double s6 = 1.0 / 6.0;
  // 6 x the central candidates
  const double gL = fma(2.0, Y[3], fma(5.0, Y[2], -Y[1]));
  const double gR = fma(2.0, Y[1], fma(5.0, Y[2], -Y[3]));
  double zgL = 0, zgR = 0, zDa = 0, zDb = 0;
  if (HASZ) {
    zgL = fma(2.0, Z[3], fma(5.0, Z[2], -Z[1]));
    zgR = fma(2.0, Z[1], fma(5.0, Z[2], -Z[3]));
    zDa = fma(3.0, Z[2] - Z[1], Z[0] - Z[3]);
    zDb = fma(3.0, Z[3] - Z[2], Z[1] - Z[4]);
  }
  if (WT == WT_NOLIM) {
    // optimal weights (0.1, 0.6, 0.3), mirrored for the right-biased value
    const double Da = fma(3.0, Y[2] - Y[1], Y[0] - Y[3]), Db = fma(3.0, Y[3] - Y[2], Y[1] - Y[4]);
    L = s6 * fma(0.3, Db, fma(0.2, Da, gL));
    R = s6 * (gR - fma(0.2, Db, 0.3 * Da));
    if (HASZ) { ZL = s6 * fma(0.3, zDb, fma(0.2, zDa, zgL)); ZR = s6 * (zgR - fma(0.2, zDb, 0.3 * zDa)); }
    return;
  }
  // smoothness indicators of the three sub-stencils (shared by both biases), scaled by 4: b_k = 4 beta_k =
  // 13/3 d^2 + e^2 (one operation less each). The weights are homogeneous of degree 0 in (beta + eps), so
  // eps and tau are scaled alike and nothing else changes.
  const double t133 = 13.0 / 3.0;
  const double d1 = X[0] - 2*X[1] + X[2], e1 = X[0] - 4*X[1] + 3*X[2];
  const double d2 = X[1] - 2*X[2] + X[3], e2 = X[1] - X[3];
  const double d3 = X[2] - 2*X[3] + X[4], e3 = 3*X[2] - 4*X[3] + X[4];
  const double Da = SAME ? (d1 - d2) : fma(3.0, Y[2] - Y[1], Y[0] - Y[3]);
  const double Db = SAME ? (d2 - d3) : fma(3.0, Y[3] - Y[2], Y[1] - Y[4]);
  const double b1 = fma(e1, e1, (t133 * d1) * d1);
  const double b2 = fma(e2, e2, (t133 * d2) * d2);
  const double b3 = fma(e3, e3, (t133 * d3) * d3);
  const double eps4 = 4.0 * eps;
  const double s1 = b1 + eps4, s2 = b2 + eps4, s3 = b3 + eps4;
  const double q1 = s1 * s1, q2 = s2 * s2, q3 = s3 * s3;
  // h_k proportional to 1/q_k (JS, M) or (1 + tau^2/q_k) (Z, YC), common positive factor dropped
  double h1, h2, h3;
  if (WT == HPB_WENO_JS || WT == HPB_WENO_M) {
    h1 = q2 * q3; h2 = q1 * q3; h3 = q1 * q2;
  } else {
    double tau;                                           // 4 tau
    if (WT == HPB_WENO_Z) tau = fabs(b3 - b1);
    else { const double t = 2*X[0] - 8*X[1] + 12*X[2] - 8*X[3] + 2*X[4]; tau = t * t; }
    const double tt = tau * tau;
    h1 = (q1 + tt) * (q2 * q3); h2 = (q2 + tt) * (q1 * q3); h3 = (q3 + tt) * (q1 * q2);
  }
  // un-normalised nonlinear weights, the common factor 0.1 of the optimal weights dropped:
  // left-biased (h1, 6 h2, 3 h3), right-biased (3 h1, 6 h2, h3)
  const double h26 = 6.0 * h2;
  const double rL = rcp_fast(fma(3.0, h3, h26 + h1)), rR = rcp_fast(fma(3.0, h1, h26 + h3));
  if (WT != HPB_WENO_M) {
    // 6 L = gL + rL (2 h1 Da + 3 h3 Db) ; 6 R = gR - rR (3 h1 Da + 2 h3 Db)
    const double P = h1 * Da, Q = h3 * Db;
    L = s6 * fma(rL, fma(3.0, Q, 2.0 * P), gL);
    R = s6 * fma(-rR, fma(3.0, P, 2.0 * Q), gR);
    if (HASZ) {
      const double zP = h1 * zDa, zQ = h3 * zDb;
      ZL = s6 * fma(rL, fma(3.0, zQ, 2.0 * zP), zgL);
      ZR = s6 * fma(-rR, fma(3.0, zP, 2.0 * zQ), zgR);
    }
    return;
  }
  // mapped weights (Henrick et al.): w~ = JS weights; alpha_k = w~(c + c^2 - 3 c w~ + w~^2)/(c^2 + w~(1-2c)) = A_k/B_k;
  // w_k = A_k B_l B_m / sum_k A_k B_l B_m =: t_k / sum t. The factor 2 of the correction term is folded into the
  // constants of A (T1 = 2 t1 on the left-biased side, T3 = 2 t3 on the right-biased side).
  const double c1 = 0.1, c2 = 0.6, c3 = 0.3;
  {
    const double w1 = h1 * rL, w2 = h26 * rL, w3 = h3 * (3.0 * rL);
    const double A1 = w1 * (2*(c1 + c1*c1) + w1 * fma(2.0, w1, -6*c1)), B1 = c1*c1 + w1 * (1.0 - 2*c1);   // 2 A1
    const double A2 = w2 * (c2 + c2*c2 + w2 * (w2 - 3*c2)), B2 = c2*c2 + w2 * (1.0 - 2*c2);
    const double A3 = w3 * (c3 + c3*c3 + w3 * (w3 - 3*c3)), B3 = c3*c3 + w3 * (1.0 - 2*c3);
    const double T1 = A1 * (B2 * B3), t2 = A2 * (B1 * B3), t3 = A3 * (B1 * B2);
    const double r = rcp_fast(fma(0.5, T1, t2) + t3);
    L = s6 * fma(r, fma(t3, Db, T1 * Da), gL);
    if (HASZ) ZL = s6 * fma(r, fma(t3, zDb, T1 * zDa), zgL);
  }
  {
    const double w1 = h1 * (3.0 * rR), w2 = h26 * rR, w3 = h3 * rR;   // optimal weights (c3, c2, c1)
    const double A1 = w1 * (c3 + c3*c3 + w1 * (w1 - 3*c3)), B1 = c3*c3 + w1 * (1.0 - 2*c3);
    const double A2 = w2 * (c2 + c2*c2 + w2 * (w2 - 3*c2)), B2 = c2*c2 + w2 * (1.0 - 2*c2);
    const double A3 = w3 * (2*(c1 + c1*c1) + w3 * fma(2.0, w3, -6*c1)), B3 = c1*c1 + w3 * (1.0 - 2*c1);   // 2 A3
    const double t1 = A1 * (B2 * B3), t2 = A2 * (B1 * B3), T3 = A3 * (B1 * B2);
    const double r = rcp_fast(fma(0.5, T3, t2) + t1);
    R = s6 * fma(-r, fma(t1, Da, T3 * Db), gR);
    if (HASZ) ZR = s6 * fma(-r, fma(t1, zDa, T3 * zDb), zgR);
  }
}

// ------------------------------------------------------------------------------------------
// record layout in shared memory (field-major: rec[field * NREC + pos * TW + line])
template <int MODEL> struct RecLayout;
template <> struct RecLayout<HPB_MODEL_LINEAR_ADR> { enum { NV = 1, U = 0, F = 1, NF = 2, V4 = 0, SR = 0, VEL = 0, H = 0, A = 0 }; };
template <> struct RecLayout<HPB_MODEL_NS2D> { enum { NV = 4, U = 0, F = 4, V4 = 8, SR = 9, VEL = 10, H = 12, A = 13, NF = 14 }; };
template <> struct RecLayout<HPB_MODEL_NS3D> { enum { NV = 5, U = 0, F = 5, V4 = 10, SR = 11, VEL = 12, H = 15, A = 16, NF = 17 }; };

template <int MODEL, bool GRAV, bool VISC>
struct SweepLayout {
  using RL = RecLayout<MODEL>;
  static constexpr bool G3 = GRAV && MODEL == HPB_MODEL_NS3D;
  static constexpr bool V3 = VISC && MODEL == HPB_MODEL_NS3D;
  static constexpr int NV = RL::NV;
  static constexpr int GFI = RL::NF;                       // record fields: gravity f, g
  static constexpr int FVI = RL::NF + (G3 ? 2 : 0);        // record fields: viscous flux (4)
  static constexpr int NFREC = FVI + (V3 ? 4 : 0);
  static constexpr int SGI = NV;                           // staging fields: gravity f, g
  static constexpr int SQI = NV + (G3 ? 2 : 0);            // staging fields: 8 derivative scalars
  static constexpr int NFSTG = SQI + (V3 ? 8 : 0);
  static constexpr int NFL = 2 * NV + (G3 ? 2 : 0);        // exchange: left-biased fL, uL (, sL x2)
  static constexpr int NFF = NV + (G3 ? 2 : 0);            // exchange: interface flux (, S x2)
  static constexpr size_t smem_bytes = sizeof(double) * ((size_t)NFREC * NREC + (size_t)NFSTG * NSTG + (size_t)(NFL + NFF) * NEX);
};

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// LOADX: the cooperative global<->shared mapping runs lanes along the sweep line (x-sweep: the line itself is
//        contiguous); otherwise 8 consecutive threads = the 8 x-adjacent lines of the CTA (y-, z-sweeps).
template <int MODEL, int WT, bool LOADX, bool GRAV, bool VISC>
__global__ void __launch_bounds__(NT, 2) k_sweep(const SweepArgs a)
{
  using SL = SweepLayout<MODEL, GRAV, VISC>;
  using RL = RecLayout<MODEL>;
  constexpr int NV = RL::NV;
  constexpr bool FLUID = (MODEL != HPB_MODEL_LINEAR_ADR);
  constexpr bool G3 = SL::G3, V3 = SL::V3;
  extern __shared__ double smem[];
  double* rec = smem;
  double* stg = rec + SL::NFREC * NREC;
  double* exL = stg + SL::NFSTG * NSTG;
  double* exF = exL + SL::NFL * NEX;

  const Geom& G = a.G;
  const int dir = a.dir;
  const int N = G.N[dir];
  const long long st = G.st[dir];
  const long long npg = G.npg;
  const int tid = threadIdx.x;
  const int w = tid >> 5, l = tid & 31;                      // compute mapping: warp = line, lane = cell of the step
  const int lw = LOADX ? w : (tid & (TW - 1));               // cooperative mapping
  const int ll = LOADX ? l : (tid >> 3);
  const int line0 = blockIdx.x * TW;
  const bool accumulate = (a.mode & 1) != 0;

  // line -> offset of its cell j = 0 in the ghost-padded array
  auto line_base = [&](int line) -> long long {
    int ia, ib;       // the two transverse indices, in increasing dimension order
    long long p;
    if (dir == 0)      { ia = line % G.N[1]; ib = line / G.N[1]; p = G.g + (long long)G.P[0] * ((ia + (G.ndims > 1 ? G.g : 0)) + (long long)G.P[1] * (ib + (G.ndims > 2 ? G.g : 0))); }
    else if (dir == 1) { ia = line % G.N[0]; ib = line / G.N[0]; p = (ia + G.g) + (long long)G.P[0] * (G.g + (long long)G.P[1] * (ib + (G.ndims > 2 ? G.g : 0))); }
    else               { ia = line % G.N[0]; ib = line / G.N[0]; p = (ia + G.g) + (long long)G.P[0] * ((ib + G.g) + (long long)G.P[1] * G.g); }
    return p;
  };
  const bool line_ok = (line0 + w) < a.nlines;
  const bool lline_ok = (line0 + lw) < a.nlines;
  const long long lbase = lline_ok ? line_base(line0 + lw) : 0;
  const double gamma = a.ph.gamma;
  const int M = (N + TL - 1) / TL;
  const int sidx_c = LOADX ? (lw * TL + ll) : (ll * SS + lw);   // staging slot, cooperative mapping
  const int sidx_p = LOADX ? (w * TL + l) : (l * SS + w);       // staging slot, compute mapping
  // VISC: dxinv of the two transverse directions of this warp's line, in the order the staged scalars use them
  // (dir 0: y, z ; dir 1: x, z ; dir 2: x, y)
  double vsc0 = 1.0, vsc1 = 1.0;
  if (V3 && line_ok) {
    const int line = line0 + w;
    const int ta = (dir == 0) ? 1 : 0, tb = (dir == 2) ? 1 : 2;
    const int ia = line % G.N[ta], ib = line / G.N[ta];
    vsc0 = a.dxinv[G.xoff[ta] + G.g + ia];
    vsc1 = a.dxinv[G.xoff[tb] + G.g + ib];
  }
  const int rbase = w * RP;                                      // this warp's records
  const int xbase = w * XP;                                      // this warp's exchange slots

  auto prefetch = [&](int m) {
    const int c = TL * m + 3 + ll;
    if (lline_ok && c >= -3 && c <= N + 2) {
      const long long p = lbase + (long long)c * st;
      double* d = stg + sidx_c;
#pragma unroll
      for (int v = 0; v < NV; v++) cp_async8(d + v * NSTG, a.u + v * npg + p);
      if (G3) { cp_async8(d + SL::SGI * NSTG, a.gf + p); cp_async8(d + (SL::SGI + 1) * NSTG, a.gg + p); }
      if (V3) {
#pragma unroll
        for (int k = 0; k < 8; k++) cp_async8(d + (SL::SQI + k) * NSTG, a.qd + (long long)a.qidx[k] * npg + p);
      }
    }
    cp_async_commit();
  };
  // results of step m are staged in the exchange slots 1..32 of fields 0..NV-1 (free after P3) and written to
  // global by the cooperative mapping at the top of step m+1: cell 32m+ll of line lw
  double old[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) old[v] = 0.0;
  auto request_old = [&](int m) {
    const int jo = TL * m + ll;
    if (accumulate && lline_ok && jo >= 0 && jo < N) {
      const long long po = lbase + (long long)jo * st;
#pragma unroll
      for (int v = 0; v < NV; v++) old[v] = a.out[v * npg + po];
    }
  };
  auto flush = [&](int m) {
    const int jo = TL * m + ll;
    if (lline_ok && jo >= 0 && jo < N) {
      const long long po = lbase + (long long)jo * st;
      const int xs = lw * XP + 1 + ll;
#pragma unroll
      for (int v = 0; v < NV; v++) a.out[v * npg + po] = old[v] + exL[v * NEX + xs];
    }
  };

  prefetch(-1);
  for (int m = -1; m < M; m++) {
    cp_async_wait_all();
    __syncthreads();                 // B1: staging of step m visible; results of step m-1 staged

    if (m >= 0) flush(m - 1);        // (step -1 stages nothing)

    // ---------------- P1: record of cell c1 = 32m+3+l (record position 5+l)
    const int c1 = TL * m + 3 + l;
    if (line_ok && c1 >= -3 && c1 <= N + 2) {
      const double* s = stg + sidx_p;
      const int ci = rbase + 5 + l;
      double U[NV];
#pragma unroll
      for (int v = 0; v < NV; v++) U[v] = s[v * NSTG];
#pragma unroll
      for (int v = 0; v < NV; v++) rec[(RL::U + v) * NREC + ci] = U[v];
      if (!FLUID) {
        rec[RL::F * NREC + ci] = a.ph.adv[dir] * U[0];
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
        rec[(RL::F + 0) * NREC + ci] = rho * vn;
#pragma unroll
        for (int k = 0; k < NDV; k++) rec[(RL::F + 1 + k) * NREC + ci] = rho * vn * vel[k] + (k == dir ? P : 0.0);
        rec[(RL::F + NV - 1) * NREC + ci] = (e + P) * vn;
        double gfv = 1.0, ggv = 1.0;
        if (G3) {
          gfv = s[SL::SGI * NSTG]; ggv = s[(SL::SGI + 1) * NSTG];
          rec[SL::GFI * NREC + ci] = gfv; rec[(SL::GFI + 1) * NREC + ci] = ggv;
        }
        const double igm1 = 1.0 / (gamma - 1.0);
        rec[RL::V4 * NREC + ci] = G3 ? ((P * igm1) * (1.0 / ggv) + ke * gfv) : (P * igm1 + ke);
        const double c2 = gamma * P * rinv;
        rec[RL::SR * NREC + ci] = sqrt(rho);
#pragma unroll
        for (int k = 0; k < NDV; k++) rec[(RL::VEL + k) * NREC + ci] = vel[k];
        rec[RL::H * NREC + ci] = 0.5 * vsq + c2 * igm1;
        rec[RL::A * NREC + ci] = sqrt(c2) + fabs(vn);
        if (V3) {
          // viscous flux of direction dir from the (mu/Re)-weighted derivatives (viscous_fused.cu)
          // each difference is scaled by the LOCAL dxinv of its own direction (after the halo exchange, as the
          // reference does): along the sweep line for the first four, the line's transverse indices for the rest
          double q[8];
          const double dd = a.dxinv[G.xoff[dir] + G.g + c1];
#pragma unroll
          for (int k = 0; k < 8; k++) q[k] = s[(SL::SQI + k) * NSTG] * (k < 4 ? dd : (k < 6 ? vsc0 : vsc1));
          const double two_third = 2.0 / 3.0;
          const double kq = igm1 * (1.0 / a.ph.Pr);
          double t1, t2, t3;
          if (dir == 0)      { t1 = two_third * (2 * q[0] - q[5] - q[7]); t2 = q[4] + q[1]; t3 = q[6] + q[2]; }
          else if (dir == 1) { t1 = q[0] + q[5]; t2 = two_third * (-q[4] + 2 * q[1] - q[7]); t3 = q[6] + q[2]; }
          else               { t1 = q[0] + q[5]; t2 = q[1] + q[7]; t3 = two_third * (-q[4] - q[6] + 2 * q[2]); }
          rec[(SL::FVI + 0) * NREC + ci] = t1;
          rec[(SL::FVI + 1) * NREC + ci] = t2;
          rec[(SL::FVI + 2) * NREC + ci] = t3;
          rec[(SL::FVI + 3) * NREC + ci] = vel[0] * t1 + vel[1] * t2 + vel[2] * t3 + kq * q[3];
        }
      }
    }
    __syncthreads();                         // B2: staging and staged results consumed
    if (m + 1 < M) prefetch(m + 1);          // overlaps P2..P4
    request_old(m);                          // old values of the cells this step updates: in flight until the next B1

    // ---------------- P2: both reconstructions of the centred stencil of cell j = 32m+1+l (position l+3)
    const int j = TL * m + 1 + l;
    const bool rc_ok = line_ok && j >= -1 && j <= N;
    const int cc = rbase + l + 3;
    double fRv[NV], uRv[NV], sR[2] = { 0.0, 0.0 };
    if (rc_ok) {
      double Zg[5] = { 0, 0, 0, 0, 0 };
      if (G3) {
#pragma unroll
        for (int k = 0; k < 5; k++) Zg[k] = rec[(SL::GFI + 1) * NREC + cc + (k - 2)];
      }
      const int ex = xbase + l + 1;
#pragma unroll
      for (int v = 0; v < NV; v++) {
        double X[5], Y[5], L, zl, zr;
#pragma unroll
        for (int k = 0; k < 5; k++) X[k] = rec[(RL::F + v) * NREC + cc + (k - 2)];
        const bool zsrc = G3 && a.with_source && (v == dir + 1 || v == NV - 1);
        if (G3 && zsrc) {
          recon_pair<WT, true, true>(X, X, Zg, a.ph.eps, L, fRv[v], zl, zr);
          const int sidx = (v == NV - 1) ? 1 : 0;
          sR[sidx] = zr;
          exL[(2 * NV + sidx) * NEX + ex] = zl;
        } else {
          recon_pair<WT, false, true>(X, X, X, a.ph.eps, L, fRv[v], zl, zr);
        }
        exL[v * NEX + ex] = L;
        // solution: weights from raw u, applied to the modified solution (Q4)
#pragma unroll
        for (int k = 0; k < 5; k++) X[k] = rec[(RL::U + v) * NREC + cc + (k - 2)];
        if (FLUID && (G3 || v == NV - 1)) {
#pragma unroll
          for (int k = 0; k < 5; k++) {
            if (v == NV - 1) Y[k] = rec[RL::V4 * NREC + cc + (k - 2)];
            else Y[k] = X[k] * rec[SL::GFI * NREC + cc + (k - 2)];
          }
          recon_pair<WT, false, false>(X, Y, Y, a.ph.eps, L, uRv[v], zl, zr);
        } else {
          recon_pair<WT, false, true>(X, X, X, a.ph.eps, L, uRv[v], zl, zr);
        }
        exL[(NV + v) * NEX + ex] = L;
      }
    }
    __syncwarp();

    // ---------------- P3: interface j-1/2 (between cells j-1 and j): upwind flux
    const bool if_ok = line_ok && j >= 0 && j <= N;
    double fh[NV], Sh[2] = { 0.0, 0.0 };
#pragma unroll
    for (int v = 0; v < NV; v++) fh[v] = 0.0;
    if (if_ok) {
      const int exl = xbase + l;              // left-biased values of cell j-1 (slot 0: carry)
      if (!FLUID) {
        fh[0] = (a.ph.adv[dir] > 0) ? exL[0 * NEX + exl] : fRv[0];
      } else {
        constexpr int NDV = NV - 2;
        const int cL = cc - 1, cR = cc;
        const double tL = rec[RL::SR * NREC + cL], tR = rec[RL::SR * NREC + cR];
        const double rs = 1.0 / (tL + tR);
        double vsq = 0.0, vn = 0.0, vhat[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
        for (int k = 0; k < NDV; k++) {
          const double v = (tL * rec[(RL::VEL + k) * NREC + cL] + tR * rec[(RL::VEL + k) * NREC + cR]) * rs;
          vsq += v * v;
          vhat[k] = v;
          if (k == dir) vn = v;
        }
        const double H = (tL * rec[RL::H * NREC + cL] + tR * rec[RL::H * NREC + cR]) * rs;
        if (a.upw == 1) {
          double du[NV], diss[NV];
#pragma unroll
          for (int v = 0; v < NV; v++) du[v] = uRv[v] - exL[(NV + v) * NEX + exl];
          roe_dissipation<NV>(du, vhat, vn, vsq, H, gamma, dir, diss);
#pragma unroll
          for (int v = 0; v < NV; v++) fh[v] = 0.5 * ((exL[v * NEX + exl] + fRv[v]) - diss[v]);
        } else {
        const double cavg = sqrt((gamma - 1.0) * (H - 0.5 * vsq));
        const double aavg = cavg + fabs(vn);
        double alpha = fmax(fmax(rec[RL::A * NREC + cL], rec[RL::A * NREC + cR]), aavg);
        if (G3) alpha *= fmax(rec[(SL::GFI + 1) * NREC + cL], rec[(SL::GFI + 1) * NREC + cR]);
#pragma unroll
        for (int v = 0; v < NV; v++) {
          const double fL = exL[v * NEX + exl], uL = exL[(NV + v) * NEX + exl];
          fh[v] = 0.5 * (fL + fRv[v]) - alpha * (0.5 * (uRv[v] - uL));
        }
        }
      }
      const int exo = xbase + l + 1;
#pragma unroll
      for (int v = 0; v < NV; v++) exF[v * NEX + exo] = fh[v];
      if (G3 && a.with_source) {
        Sh[0] = 0.5 * (exL[(2 * NV + 0) * NEX + exl] + sR[0]);
        Sh[1] = 0.5 * (exL[(2 * NV + 1) * NEX + exl] + sR[1]);
        exF[(NV + 0) * NEX + exo] = Sh[0];
        exF[(NV + 1) * NEX + exo] = Sh[1];
      }
    }
    __syncwarp();
    // carry of the left-biased values (slot 32 -> slot 0) before the slots 1..32 are reused for the results
    if (l < SL::NFL) exL[l * NEX + xbase] = exL[l * NEX + xbase + TL];
    __syncwarp();

    // ---------------- P4: cell jo = j-1 = 32m+l (position l+2): interfaces j-1/2 (own) and j-3/2 (neighbour / carry)
    const int jo = j - 1;
    if (line_ok && jo >= 0 && jo < N) {
      const int exl = xbase + l;
      const int co = cc - 1;
      const double dxi = a.dxinv[G.xoff[dir] + G.g + jo];
      double res[NV];
#pragma unroll
      for (int v = 0; v < NV; v++) {
        const double t = dxi * (fh[v] - exF[v * NEX + exl]);
        res[v] = (a.mode < 2) ? -t : t;
      }
      if (V3) {
        // par_v += dxinv * (FV[j-2] - 8 FV[j-1] + 8 FV[j+1] - FV[j+2]) / 12   (components 1..4)
        const double s12 = 1.0 / 12.0;
#pragma unroll
        for (int v = 0; v < 4; v++) {
          const double* f = rec + (SL::FVI + v) * NREC + co;
          const double dfv = (f[-2] - 8 * f[-1] + 8 * f[1] - f[2]) * s12;
          res[1 + v] += dxi * dfv;
        }
      }
      if (G3 && a.with_source) {
        // NavierStokes3DSource.c:80-100 (the source accumulates into the same array as the flux divergence)
        const double rho = rec[RL::U * NREC + co];
        const double vd = rec[(RL::VEL + dir) * NREC + co];
        const double f = rec[SL::GFI * NREC + co];
        const double tm = rho * a.ph.RT, te = rho * a.ph.RT * vd;
        const double sm = (tm * f) * (Sh[0] - exF[(NV + 0) * NEX + exl]) * dxi;
        const double se = (te * f) * (Sh[1] - exF[(NV + 1) * NEX + exl]) * dxi;
#pragma unroll
        for (int v = 1; v < NV; v++) res[v] += ((v == dir + 1) ? sm : 0.0) + ((v == NV - 1) ? se : 0.0);
      }
#pragma unroll
      for (int v = 0; v < NV; v++) exL[v * NEX + xbase + 1 + l] = res[v];      // staged for the cooperative store
    }
    __syncwarp();

    // ---------------- shift (warp-private): the last 5 records and the interface-flux carry move to the front
    for (int i = l; i < SL::NFREC * 5; i += 32) {
      const int f = i / 5, r = i % 5;
      rec[f * NREC + rbase + r] = rec[f * NREC + rbase + TL + r];
    }
    if (l < SL::NFF) exF[l * NEX + xbase] = exF[l * NEX + xbase + TL];
    __syncwarp();
  }
  __syncthreads();
  flush(M - 1);
}

// host-side launcher of one (MODEL, WT) family; defined in sweep_fused_inst.cuh per weight type
template <int WT>
bool launch_sweep(hpb_solver* h, const SweepArgs& a);

} // namespace hpbf

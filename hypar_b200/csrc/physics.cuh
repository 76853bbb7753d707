// physics.cuh -- device functions for the physical models on the hot path.
// Algebra follows the reference's macros (cited per function); evaluated in registers.
#pragma once
#include "hpb_internal.h"

#define HPB_DEV __device__ __forceinline__

template <int MODEL> struct ModelTraits;
template <> struct ModelTraits<HPB_MODEL_LINEAR_ADR> { static constexpr int NV = 1; static constexpr int ND = 1; };
template <> struct ModelTraits<HPB_MODEL_EULER1D>    { static constexpr int NV = 3; static constexpr int ND = 1; };
template <> struct ModelTraits<HPB_MODEL_NS2D>       { static constexpr int NV = 4; static constexpr int ND = 2; };
template <> struct ModelTraits<HPB_MODEL_NS3D>       { static constexpr int NV = 5; static constexpr int ND = 3; };
template <> struct ModelTraits<HPB_MODEL_BURGERS>    { static constexpr int NV = 1; static constexpr int ND = 1; };

HPB_DEV double hpb_abs(double a) { return fabs(a); }
HPB_DEV double hpb_max3(double a, double b, double c) { return fmax(fmax(a, b), c); }

// ---- flow variables. NS: velocity guarded against rho == 0 (navierstokes3d.h:98-108,
// navierstokes2d.h:81-90); Euler1D unguarded (euler1d.h _Euler1DGetFlowVar_).
// prim = (rho, v[0..ND-1], e, P)
template <int MODEL>
HPB_DEV void flowvar(const double* u, double gamma, double& rho, double* vel, double& e, double& P)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  constexpr int NDV = NV - 2;     // number of velocity components
  rho = u[0];
  double vsq = 0.0;
  if (MODEL == HPB_MODEL_EULER1D) {
    vel[0] = u[1] / rho;
    vsq = vel[0] * vel[0];
  } else {
#pragma unroll
    for (int k = 0; k < NDV; k++) vel[k] = (rho == 0) ? 0.0 : u[1 + k] / rho;
#pragma unroll
    for (int k = 0; k < NDV; k++) vsq += vel[k] * vel[k];
  }
  e = u[NV - 1];
  if (MODEL == HPB_MODEL_EULER1D) P = (e - 0.5 * rho * vel[0] * vel[0]) * (gamma - 1.0);   // euler1d.h: 0.5*rho*v*v, left to right
  else P = (e - 0.5 * rho * vsq) * (gamma - 1.0);
}

// ---- FFunction: NavierStokes3DFlux.c:24 (_NavierStokes3DSetFlux_ navierstokes3d.h:114-139),
// NavierStokes2DFlux.c:21, Euler1DFlux.c:16, LinearADRAdvection.c:37
template <int MODEL>
HPB_DEV void flux_fn(const Phys& ph, const double* u, int dir, double* f, long long p = -1)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  if (MODEL == HPB_MODEL_LINEAR_ADR) {
    // p: the cell (with ghosts) -- only the spatially varying advection field needs it (LinearADRAdvection.c:72-80)
    const double a = (ph.advf != nullptr && p >= 0) ? ph.advf[dir * ph.advf_npg + p] : ph.adv[dir];
    f[0] = a * u[0];
  } else if (MODEL == HPB_MODEL_BURGERS) {        // BurgersAdvection.c:17-49: the same flux in every direction
    f[0] = 0.5 * u[0] * u[0];
  } else {
    constexpr int NDV = NV - 2;
    double rho, vel[3], e, P;
    flowvar<MODEL>(u, ph.gamma, rho, vel, e, P);
    const double vn = vel[dir];
    f[0] = rho * vn;
#pragma unroll
    for (int k = 0; k < NDV; k++) f[1 + k] = rho * vn * vel[k] + (k == dir ? P : 0.0);
    f[NV - 1] = (e + P) * vn;
  }
}

// generic-nvars LinearADR (nvars may exceed 1)
HPB_DEV void linadr_flux(const Phys& ph, int nv, const double* u, int dir, double* f)
{
  for (int v = 0; v < nv; v++) f[v] = ph.adv[nv * dir + v] * u[v];
}

// ---- UFunction: NavierStokes3DModifiedSolution.c:31, NavierStokes2DModifiedSolution.c:31,
// Euler1DModifiedSolution.c (uC = u / grav_field); LinearADR has none (uC = u)
template <int MODEL>
HPB_DEV void modified_fn(const Phys& ph, const double* u, double gf, double gg, double* uC)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  if (MODEL == HPB_MODEL_LINEAR_ADR || MODEL == HPB_MODEL_BURGERS) {
    uC[0] = u[0];
  } else if (MODEL == HPB_MODEL_EULER1D) {
    const double a = 1.0 / gf;
#pragma unroll
    for (int v = 0; v < NV; v++) uC[v] = a * u[v];
  } else {
    constexpr int NDV = NV - 2;
    double rho, vel[3], e, P;
    flowvar<MODEL>(u, ph.gamma, rho, vel, e, P);
    double vsq = 0.0;
#pragma unroll
    for (int k = 0; k < NDV; k++) vsq += vel[k] * vel[k];
    const double inv_gamma_m1 = 1.0 / (ph.gamma - 1.0);
#pragma unroll
    for (int v = 0; v < NV - 1; v++) uC[v] = u[v] * gf;
    uC[NV - 1] = (P * inv_gamma_m1) * (1.0 / gg) + (0.5 * rho * vsq) * gf;
  }
}

// ---- Roe average: navierstokes3d.h:144-169, navierstokes2d.h _NavierStokes2DRoeAverage_,
// euler1d.h _Euler1DRoeAverage_ (the 3-D macro keeps the rho==0 guard of GetFlowVar)
template <int MODEL>
HPB_DEV void roe_average(const Phys& ph, const double* uL, const double* uR, double* uavg)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  constexpr int NDV = NV - 2;
  const double gamma = ph.gamma;
  double rhoL, vL[3], eL, PL, rhoR, vR[3], eR, PR;
  if (MODEL == HPB_MODEL_NS3D) {
    flowvar<MODEL>(uL, gamma, rhoL, vL, eL, PL);
    flowvar<MODEL>(uR, gamma, rhoR, vR, eR, PR);
  } else {
    rhoL = uL[0]; rhoR = uR[0];
#pragma unroll
    for (int k = 0; k < NDV; k++) { vL[k] = uL[1 + k] / rhoL; vR[k] = uR[1 + k] / rhoR; }
    eL = uL[NV - 1]; eR = uR[NV - 1];
    double sL = 0.0, sR = 0.0;
#pragma unroll
    for (int k = 0; k < NDV; k++) { sL += vL[k] * vL[k]; sR += vR[k] * vR[k]; }
    if (MODEL == HPB_MODEL_EULER1D) {                       // _Euler1DRoeAverage_: 0.5*rho*v*v, left to right
      PL = (eL - 0.5 * rhoL * vL[0] * vL[0]) * (gamma - 1.0);
      PR = (eR - 0.5 * rhoR * vR[0] * vR[0]) * (gamma - 1.0);
    } else {
      PL = (eL - 0.5 * rhoL * sL) * (gamma - 1.0);
      PR = (eR - 0.5 * rhoR * sR) * (gamma - 1.0);
    }
  }
  double vsqL = 0.0, vsqR = 0.0;
#pragma unroll
  for (int k = 0; k < NDV; k++) { vsqL += vL[k] * vL[k]; vsqR += vR[k] * vR[k]; }
  const double cLsq = gamma * PL / rhoL, cRsq = gamma * PR / rhoR;
  const double HL = 0.5 * vsqL + cLsq / (gamma - 1.0);
  const double HR = 0.5 * vsqR + cRsq / (gamma - 1.0);
  const double tL = sqrt(rhoL), tR = sqrt(rhoR);
  const double rho = tL * tR;
  const double tsum = tL + tR;
  double v[3], vsq = 0.0;
#pragma unroll
  for (int k = 0; k < NDV; k++) { v[k] = (tL * vL[k] + tR * vR[k]) / tsum; vsq += v[k] * v[k]; }
  const double H = (tL * HL + tR * HR) / tsum;
  double P;
  if (MODEL == HPB_MODEL_NS3D) {
    P = (H - 0.5 * vsq) * (rho * (gamma - 1.0)) / gamma;
  } else {
    const double csq = (gamma - 1.0) * (H - 0.5 * vsq);
    P = csq * rho / gamma;
  }
  const double e = (MODEL == HPB_MODEL_EULER1D) ? (P / (gamma - 1.0) + 0.5 * rho * v[0] * v[0])
                                                : (P / (gamma - 1.0) + 0.5 * rho * vsq);
  uavg[0] = rho;
#pragma unroll
  for (int k = 0; k < NDV; k++) uavg[1 + k] = rho * v[k];
  uavg[NV - 1] = e;
}

// ---- eigen-structure. Euler1D: euler1d.h _Euler1DEigenvalues_/_LeftEigenvectors_/_RightEigenvectors_;
// NS3D: navierstokes3d.h:254-471 (ordering chosen by the reference per direction).
// lam = eigenvalues (diagonal of D), L rows = left eigenvectors, R columns = right eigenvectors.
HPB_DEV void e1d_eigen(double gamma, const double* u, double* lam, double* L, double* R)
{
  const double rho = u[0], v = u[1] / rho, e = u[2];
  const double P = (e - 0.5 * rho * v * v) * (gamma - 1.0);
  const double c = sqrt(gamma * P / rho);
  lam[0] = v; lam[1] = v - c; lam[2] = v + c;
  const double k = (gamma - 1) / (rho * c);
  L[3] = k * (-(v * v) / 2 - c * v / (gamma - 1));
  L[4] = k * (v + c / (gamma - 1));
  L[5] = k * (-1);
  L[0] = k * (rho * (-(v * v) / 2 + c * c / (gamma - 1)) / c);
  L[1] = k * (rho * v / c);
  L[2] = k * (-rho / c);
  L[6] = k * ((v * v) / 2 - c * v / (gamma - 1));
  L[7] = k * (-v + c / (gamma - 1));
  L[8] = k * (1);
  R[1] = -rho / (2 * c); R[4] = -rho * (v - c) / (2 * c); R[7] = -rho * ((v * v) / 2 + (c * c) / (gamma - 1) - c * v) / (2 * c);
  R[0] = 1;              R[3] = v;                        R[6] = v * v / 2;
  R[2] = rho / (2 * c);  R[5] = rho * (v + c) / (2 * c);  R[8] = rho * ((v * v) / 2 + (c * c) / (gamma - 1) + c * v) / (2 * c);
}

HPB_DEV void ns3d_eigen(double ga, const double* u, int dir, double* lam, double* L, double* R)
{
  double rho, vel[3], e, P;
  flowvar<HPB_MODEL_NS3D>(u, ga, rho, vel, e, P);
  const double vx = vel[0], vy = vel[1], vz = vel[2];
  const double gm1 = ga - 1.0;
  const double ek = 0.5 * (vx * vx + vy * vy + vz * vz);
  const double a = sqrt(ga * P / rho);
  const double h0 = a * a / gm1 + ek;
  const double vn = vel[dir];
  // acoustic rows/cols: index am (v-a) and ap (v+a); entropy 0; shear s1, s2
  const int am = dir + 1, ap = 4;
#pragma unroll
  for (int k = 0; k < 5; k++) lam[k] = vn;
  lam[am] = vn - a; lam[ap] = vn + a;
#pragma unroll
  for (int k = 0; k < 25; k++) { L[k] = 0.0; R[k] = 0.0; }
  const double i2aa = 2 * a * a, aa = a * a;
  // entropy wave (row 0)
  L[0] = (aa - gm1 * ek) / aa; L[1] = (gm1 * vx) / aa; L[2] = (gm1 * vy) / aa; L[3] = (gm1 * vz) / aa; L[4] = (-gm1) / aa;
  // acoustic v-a
  L[am*5+0] = (gm1 * ek + a * vn) / i2aa;
  L[am*5+1] = ((-gm1) * vx - (dir == 0 ? a : 0.0)) / i2aa;
  L[am*5+2] = ((-gm1) * vy - (dir == 1 ? a : 0.0)) / i2aa;
  L[am*5+3] = ((-gm1) * vz - (dir == 2 ? a : 0.0)) / i2aa;
  L[am*5+4] = gm1 / i2aa;
  // acoustic v+a
  L[ap*5+0] = (gm1 * ek - a * vn) / i2aa;
  L[ap*5+1] = ((-gm1) * vx + (dir == 0 ? a : 0.0)) / i2aa;
  L[ap*5+2] = ((-gm1) * vy + (dir == 1 ? a : 0.0)) / i2aa;
  L[ap*5+3] = ((-gm1) * vz + (dir == 2 ? a : 0.0)) / i2aa;
  L[ap*5+4] = gm1 / i2aa;
  // right eigenvectors: entropy, acoustic
  R[0*5+0] = 1.0; R[1*5+0] = vx; R[2*5+0] = vy; R[3*5+0] = vz; R[4*5+0] = ek;
  R[0*5+am] = 1.0; R[1*5+am] = vx - (dir == 0 ? a : 0.0); R[2*5+am] = vy - (dir == 1 ? a : 0.0);
  R[3*5+am] = vz - (dir == 2 ? a : 0.0); R[4*5+am] = h0 - a * vn;
  R[0*5+ap] = 1.0; R[1*5+ap] = vx + (dir == 0 ? a : 0.0); R[2*5+ap] = vy + (dir == 1 ? a : 0.0);
  R[3*5+ap] = vz + (dir == 2 ? a : 0.0); R[4*5+ap] = h0 + a * vn;
  // shear waves (signs and slots exactly as navierstokes3d.h)
  if (dir == 0) {
    L[2*5+0] = vy;  L[2*5+2] = -1.0;   R[2*5+2] = -1.0; R[4*5+2] = -vy;
    L[3*5+0] = -vz; L[3*5+3] = 1.0;    R[3*5+3] = 1.0;  R[4*5+3] = vz;
  } else if (dir == 1) {
    L[1*5+0] = -vx; L[1*5+1] = 1.0;    R[1*5+1] = 1.0;  R[4*5+1] = vx;
    L[3*5+0] = vz;  L[3*5+3] = -1.0;   R[3*5+3] = -1.0; R[4*5+3] = -vz;
  } else {
    L[1*5+0] = vx;  L[1*5+1] = -1.0;   R[1*5+1] = -1.0; R[4*5+1] = -vx;
    L[2*5+0] = -vy; L[2*5+2] = 1.0;    R[2*5+2] = 1.0;  R[4*5+2] = vy;
  }
}

// NS2D: navierstokes2d.h:213-400 (_NavierStokes2DEigenvalues_ / LeftEigenvectors_ / RightEigenvectors_), written out
// with the reference's nx, ny so that every entry rounds as there. Ordering: x: (v-a, v+a, shear, entropy);
// y: (v-a, shear, v+a, entropy).
HPB_DEV void ns2d_eigen(double ga, const double* u, int dir, double* lam, double* L, double* R)
{
  const double ga_minus_one = ga - 1.0;
  const double rho = u[0], vx = u[1] / rho, vy = u[2] / rho, e = u[3];
  const double vsq = (vx * vx) + (vy * vy);
  const double P = (e - 0.5 * rho * vsq) * (ga - 1.0);
  const double ek = 0.5 * (vx * vx + vy * vy);
  const double a = sqrt(ga * P / rho);
  const double h0 = a * a / ga_minus_one + ek;
  double nx = 0, ny = 0, un;
  if (dir == 0) {
    un = vx; nx = 1.0;
    lam[0] = un - a; lam[1] = un + a; lam[2] = un; lam[3] = un;
    L[0]  = (ga_minus_one * ek + a * un) / (2 * a * a);
    L[1]  = ((-ga_minus_one) * vx - a * nx) / (2 * a * a);
    L[2]  = ((-ga_minus_one) * vy - a * ny) / (2 * a * a);
    L[3]  = ga_minus_one / (2 * a * a);
    L[12] = (a * a - ga_minus_one * ek) / (a * a);
    L[13] = (ga_minus_one * vx) / (a * a);
    L[14] = (ga_minus_one * vy) / (a * a);
    L[15] = (-ga_minus_one) / (a * a);
    L[4]  = (ga_minus_one * ek - a * un) / (2 * a * a);
    L[5]  = ((-ga_minus_one) * vx + a * nx) / (2 * a * a);
    L[6]  = ((-ga_minus_one) * vy + a * ny) / (2 * a * a);
    L[7]  = ga_minus_one / (2 * a * a);
    L[8]  = (vy - un * ny) / nx;
    L[9]  = ny;
    L[10] = (ny * ny - 1.0) / nx;
    L[11] = 0.0;
    R[0] = 1.0; R[4] = vx - a * nx; R[8]  = vy - a * ny; R[12] = h0 - a * un;
    R[3] = 1.0; R[7] = vx;          R[11] = vy;          R[15] = ek;
    R[1] = 1.0; R[5] = vx + a * nx; R[9]  = vy + a * ny; R[13] = h0 + a * un;
    R[2] = 0.0; R[6] = ny;          R[10] = -nx;         R[14] = vx * ny - vy * nx;
  } else {
    un = vy; ny = 1.0;
    lam[0] = un - a; lam[1] = un; lam[2] = un + a; lam[3] = un;
    L[0]  = (ga_minus_one * ek + a * un) / (2 * a * a);
    L[1]  = ((1.0 - ga) * vx - a * nx) / (2 * a * a);
    L[2]  = ((1.0 - ga) * vy - a * ny) / (2 * a * a);
    L[3]  = ga_minus_one / (2 * a * a);
    L[12] = (a * a - ga_minus_one * ek) / (a * a);
    L[13] = ga_minus_one * vx / (a * a);
    L[14] = ga_minus_one * vy / (a * a);
    L[15] = (1.0 - ga) / (a * a);
    L[8]  = (ga_minus_one * ek - a * un) / (2 * a * a);
    L[9]  = ((1.0 - ga) * vx + a * nx) / (2 * a * a);
    L[10] = ((1.0 - ga) * vy + a * ny) / (2 * a * a);
    L[11] = ga_minus_one / (2 * a * a);
    L[4]  = (un * nx - vx) / ny;
    L[5]  = (1.0 - nx * nx) / ny;
    L[6]  = -nx;
    L[7]  = 0;
    R[0] = 1.0; R[4] = vx - a * nx; R[8]  = vy - a * ny; R[12] = h0 - a * un;
    R[3] = 1.0; R[7] = vx;          R[11] = vy;          R[15] = ek;
    R[2] = 1.0; R[6] = vx + a * nx; R[10] = vy + a * ny; R[14] = h0 + a * un;
    R[1] = 0;   R[5] = ny;          R[9]  = -nx;         R[13] = vx * ny - vy * nx;
  }
}

// eigenvalues only (_Euler1DEigenvalues_, _NavierStokes3DEigenvalues_ navierstokes3d.h:254-278): same ordering as eigen()
template <int MODEL>
HPB_DEV void eigenvalues(const Phys& ph, const double* u, int dir, double* lam)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  double rho, vel[3], e, P;
  flowvar<MODEL>(u, ph.gamma, rho, vel, e, P);
  const double c = sqrt(ph.gamma * P / rho);
  if (MODEL == HPB_MODEL_EULER1D) { lam[0] = vel[0]; lam[1] = vel[0] - c; lam[2] = vel[0] + c; }
  else if (MODEL == HPB_MODEL_NS2D) {          // navierstokes2d.h:213-240
    const double vn = (dir == 0) ? vel[0] : vel[1];
    lam[0] = vn - c; lam[1] = (dir == 0) ? vn + c : vn; lam[2] = (dir == 0) ? vn : vn + c; lam[3] = vn;
  } else {
    const double vn = (dir == 0) ? vel[0] : (dir == 1 ? vel[1] : vel[2]);
#pragma unroll
    for (int k = 0; k < NV; k++) lam[k] = (k == dir + 1) ? (vn - c) : ((k == NV - 1) ? (vn + c) : vn);
  }
}

template <int MODEL>
HPB_DEV void eigen(const Phys& ph, const double* u, int dir, double* lam, double* L, double* R)
{
  if (MODEL == HPB_MODEL_EULER1D) e1d_eigen(ph.gamma, u, lam, L, R);
  else if (MODEL == HPB_MODEL_NS3D) ns3d_eigen(ph.gamma, u, dir, lam, L, R);
  else if (MODEL == HPB_MODEL_NS2D) ns2d_eigen(ph.gamma, u, dir, lam, L, R);
}

// ---- Upwind. Inputs: reconstructed fL,fR,uL,uR at the interface, raw u of the two adjacent cells,
// kappa sources (grav field values at the two cells).
// NavierStokes3DUpwind.c:349-418 (Rusanov), :40-125 (Roe + Harten fix), NavierStokes2DUpwind.c
// (Rusanov), Euler1DUpwind.c:30-95 (Roe, kappa|lambda|, no entropy fix), :376-443 (Rusanov: the
// reference uses c = gamma*P/rho WITHOUT the square root there), LinearADRUpwind.c:16-100.
template <int MODEL>
HPB_DEV void upwind_fn(const Phys& ph, int dir, const double* fL, const double* fR, const double* uL,
                       const double* uR, const double* ucL, const double* ucR, double kL, double kR, double* fI)
{
  constexpr int NV = ModelTraits<MODEL>::NV;
  if (MODEL == HPB_MODEL_LINEAR_ADR) {
    if (ph.advf != nullptr) {       // LinearADRUpwind.c:56-82; kL, kR carry the advection speed of the two adjacent cells
      const double eigL = kL, eigR = kR;
      if ((eigL > 0) && (eigR > 0))      fI[0] = fL[0];
      else if ((eigL < 0) && (eigR < 0)) fI[0] = fR[0];
      else {
        const double alpha = fmax(hpb_abs(eigL), hpb_abs(eigR));
        fI[0] = 0.5 * (fL[0] + fR[0] - alpha * (uR[0] - uL[0]));
      }
      return;
    }
    fI[0] = (ph.adv[dir] > 0 ? fL[0] : fR[0]);
    return;
  }
  if (MODEL == HPB_MODEL_BURGERS) {               // BurgersUpwind.c:17-70: the wave speed is u itself
    const double eigL = ucL[0], eigR = ucR[0];
    if ((eigL > 0) && (eigR > 0))      fI[0] = fL[0];
    else if ((eigL < 0) && (eigR < 0)) fI[0] = fR[0];
    else {
      const double alpha = fmax(hpb_abs(eigL), hpb_abs(eigR));
      fI[0] = 0.5 * (fL[0] + fR[0] - alpha * (uR[0] - uL[0]));
    }
    return;
  }
  double udiff[NV], uavg[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) udiff[v] = 0.5 * (uR[v] - uL[v]);
  roe_average<MODEL>(ph, ucL, ucR, uavg);
  const double kappa = fmax(kL, kR);
  if (ph.upwind == HPB_UPWIND_RUSANOV) {
    double rho, vel[3], e, P, c;
    flowvar<MODEL>(ucL, ph.gamma, rho, vel, e, P);
    c = (MODEL == HPB_MODEL_EULER1D) ? ph.gamma * P / rho : sqrt(ph.gamma * P / rho);
    const double alphaL = c + hpb_abs(vel[dir]);
    flowvar<MODEL>(ucR, ph.gamma, rho, vel, e, P);
    c = (MODEL == HPB_MODEL_EULER1D) ? ph.gamma * P / rho : sqrt(ph.gamma * P / rho);
    const double alphaR = c + hpb_abs(vel[dir]);
    flowvar<MODEL>(uavg, ph.gamma, rho, vel, e, P);
    c = (MODEL == HPB_MODEL_EULER1D) ? ph.gamma * P / rho : sqrt(ph.gamma * P / rho);
    const double alphaavg = c + hpb_abs(vel[dir]);
    const double alpha = kappa * hpb_max3(alphaL, alphaR, alphaavg);
#pragma unroll
    for (int v = 0; v < NV; v++) fI[v] = 0.5 * (fL[v] + fR[v]) - alpha * udiff[v];
  } else if (ph.upwind == HPB_UPWIND_RF || ph.upwind == HPB_UPWIND_LLF) {
    // characteristic-based Roe-fixed / local Lax-Friedrichs: NavierStokes3DUpwind.c:140-237, :244-330;
    // Euler1DUpwind.c:115-205, :212-286. NavierStokes3D takes the Roe average and the one-sided eigenvalues
    // from the RECONSTRUCTED interface states uL, uR; Euler1D from the two adjacent cells, with kappa.
    // NavierStokes2D (NavierStokes2DUpwind.c:120-196, :223-300): RF like NavierStokes3D (interface states, no kappa);
    // LLF from the two adjacent cells, kappa on the FIRST characteristic field only (sic).
    if (MODEL == HPB_MODEL_EULER1D || MODEL == HPB_MODEL_NS3D || MODEL == HPB_MODEL_NS2D) {
      double eigL[NV], eigC[NV], eigR[NV], L[NV * NV], R[NV * NV], uroe[NV];
      if (MODEL == HPB_MODEL_NS3D || (MODEL == HPB_MODEL_NS2D && ph.upwind == HPB_UPWIND_RF)) {
        roe_average<MODEL>(ph, uL, uR, uroe);
        eigenvalues<MODEL>(ph, uL, dir, eigL);
        eigenvalues<MODEL>(ph, uR, dir, eigR);
      } else {
#pragma unroll
        for (int v = 0; v < NV; v++) uroe[v] = uavg[v];
        eigenvalues<MODEL>(ph, ucL, dir, eigL);
        eigenvalues<MODEL>(ph, ucR, dir, eigR);
      }
      eigen<MODEL>(ph, uroe, dir, eigC, L, R);
      double wL[NV], wR[NV], gL[NV], gR[NV], fc[NV];
#pragma unroll
      for (int i = 0; i < NV; i++) {           // MatVecMult3/5 (matops.h): left to right
        double a = L[i * NV] * uL[0], b = L[i * NV] * uR[0], c = L[i * NV] * fL[0], d = L[i * NV] * fR[0];
#pragma unroll
        for (int j = 1; j < NV; j++) {
          a += L[i * NV + j] * uL[j]; b += L[i * NV + j] * uR[j];
          c += L[i * NV + j] * fL[j]; d += L[i * NV + j] * fR[j];
        }
        wL[i] = a; wR[i] = b; gL[i] = c; gR[i] = d;
      }
#pragma unroll
      for (int k = 0; k < NV; k++) {
        if (ph.upwind == HPB_UPWIND_RF && eigL[k] > 0 && eigC[k] > 0 && eigR[k] > 0)      fc[k] = gL[k];
        else if (ph.upwind == HPB_UPWIND_RF && eigL[k] < 0 && eigC[k] < 0 && eigR[k] < 0) fc[k] = gR[k];
        else {
          double alpha = hpb_max3(hpb_abs(eigL[k]), hpb_abs(eigC[k]), hpb_abs(eigR[k]));
          if (MODEL == HPB_MODEL_EULER1D) alpha = kappa * alpha;
          if (MODEL == HPB_MODEL_NS2D && ph.upwind == HPB_UPWIND_LLF && k == 0) alpha = kappa * alpha;
          fc[k] = 0.5 * (gL[k] + gR[k] + alpha * (wL[k] - wR[k]));
        }
      }
#pragma unroll
      for (int i = 0; i < NV; i++) {
        double s2 = R[i * NV] * fc[0];
#pragma unroll
        for (int j = 1; j < NV; j++) s2 += R[i * NV + j] * fc[j];
        fI[i] = s2;
      }
    }
  } else {
    if (MODEL == HPB_MODEL_EULER1D || MODEL == HPB_MODEL_NS3D || MODEL == HPB_MODEL_NS2D) {
      double lam[NV], L[NV * NV], R[NV * NV];
      eigen<MODEL>(ph, uavg, dir, lam, L, R);
      if (MODEL == HPB_MODEL_EULER1D) {
#pragma unroll
        for (int k = 0; k < NV; k++) lam[k] = kappa * hpb_abs(lam[k]);
      } else {
        const double delta = 0.000001, delta2 = delta * delta;
#pragma unroll
        for (int k = 0; k < NV; k++) {
          lam[k] = (hpb_abs(lam[k]) < delta ? (lam[k] * lam[k] + delta2) / (2 * delta) : hpb_abs(lam[k]));
          if (MODEL == HPB_MODEL_NS2D) lam[k] = kappa * lam[k];        // NavierStokes2DUpwind.c:86-91
        }
      }
      // udiss = (R (|D| L)) udiff, multiplied out in the reference's own order (MatMult, MatMult, MatVecMult:
      // NavierStokes3DUpwind.c:104-106, Euler1DUpwind.c:84-86) so that the result carries its rounding
      double DL[NV * NV], modA[NV * NV];
#pragma unroll
      for (int i = 0; i < NV; i++)
#pragma unroll
        for (int j = 0; j < NV; j++) DL[i * NV + j] = lam[i] * L[i * NV + j];
#pragma unroll
      for (int i = 0; i < NV; i++)
#pragma unroll
        for (int j = 0; j < NV; j++) {
          double s = R[i * NV + 0] * DL[0 * NV + j];
#pragma unroll
          for (int k = 1; k < NV; k++) s += R[i * NV + k] * DL[k * NV + j];
          modA[i * NV + j] = s;
        }
#pragma unroll
      for (int i = 0; i < NV; i++) {
        double s = modA[i * NV + 0] * udiff[0];
#pragma unroll
        for (int j = 1; j < NV; j++) s += modA[i * NV + j] * udiff[j];
        fI[i] = 0.5 * (fL[i] + fR[i]) - s;
      }
    }
  }
}

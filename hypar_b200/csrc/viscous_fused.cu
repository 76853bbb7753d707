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
  int lo[3], ext[3];   // k_qderiv3: box of points (interior-relative origin, extent)
  int zchunk;          // k_qderiv_int: interior planes per CTA
};

__device__ __forceinline__ double rcp_nr(double x)
{
  // MUFU.RCP64H seed + one cubic Newton step (relative error <= 2.2e-16, profiles/r01_microbench.txt);
  // x is a density: positive and normal wherever the result is used
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  e = fma(e, e, e);
  return fma(r, e, r);
}

__device__ __forceinline__ void load5(const double* __restrict__ u, long long npg, long long p, double (&r)[5])
{
#pragma unroll
  for (int v = 0; v < 5; v++) r[v] = __ldg(u + v * npg + p);
}
// (u, v, w, T = gamma p / rho) from the conserved variables; shared by both kernels so that a point gets the same
// bits whichever kernel evaluates it (the overlapped multi-GPU schedule splits the interior differently)
__device__ __forceinline__ void prim_of(const double (&r)[5], double gamma, double (&q)[4])
{
  // every operation spelled out (explicit fma / _rn intrinsics): no compiler-chosen contraction, so the two kernels
  // cannot round differently
  const double rho = r[0];
  const double rinv = rcp_nr(rho);
  const double vx = (rho == 0) ? 0.0 : __dmul_rn(r[1], rinv);
  const double vy = (rho == 0) ? 0.0 : __dmul_rn(r[2], rinv);
  const double vz = (rho == 0) ? 0.0 : __dmul_rn(r[3], rinv);
  const double vsq = fma(vz, vz, fma(vy, vy, __dmul_rn(vx, vx)));
  const double P = __dmul_rn(fma(__dmul_rn(-0.5, rho), vsq, r[4]), gamma - 1.0);
  q[0] = vx; q[1] = vy; q[2] = vz; q[3] = __dmul_rn(__dmul_rn(gamma, P), rinv);
}
__device__ __forceinline__ double central4(double m2, double m1, double p1, double p2)
{
  return __dmul_rn(__dsub_rn(fma(8.0, p1, fma(-8.0, m1, m2)), p2), 1.0 / 12.0);
}
// mu = T^0.76 = exp(0.76 log T) (raiseto, math_ops.h:37; NavierStokes3DParabolicFunction.c:174). HPB_MU_FAST = 1 (a measured
// variant, NOT the default: profiles/r02s_bench_mufast.json against r02s_bench_muslow.json on the same box -- 18.75 against
// 18.58 ms per step in these kernels, which wait for memory, not for the FP64 pipe; the whole GPU suite passes with it,
// profiles/r02s_pytest_gpu.log): the 25th root of T^19 instead of the double-precision exp and log (~100 FP64 instructions per
// point) -- single-precision seed y0 = powf(T, 0.76) (relative error e0 <~ 2e-6), then ONE Halley step for y^25 = T^19,
// y1 = y0 ((n-1) y0^n + (n+1) a) / ((n+1) y0^n + (n-1) a), whose error is (n^2-1)/12 e0^3 = 52 e0^3 < 5e-16: ~25 FP64
// instructions. 19/25 differs from the double 0.76 by 2.7e-17 (times |log T|: nothing). Both Q-derivative kernels share it, so a
// point still gets the same bits whichever kernel evaluates it; the exact path (kernels.cu) keeps exp(0.76 log T).
#ifndef HPB_MU_FAST
#define HPB_MU_FAST 0
#endif
__device__ __forceinline__ double mu_over_Re(double T, double inv_Re)
{
#if HPB_MU_FAST
  const double y0 = (double)powf((float)T, 0.76f);
  const double y2 = __dmul_rn(y0, y0), y4 = __dmul_rn(y2, y2), y8 = __dmul_rn(y4, y4), y16 = __dmul_rn(y8, y8);
  const double yn = __dmul_rn(__dmul_rn(y16, y8), y0);                       // y0^25
  const double t2 = __dmul_rn(T, T), t4 = __dmul_rn(t2, t2), t8 = __dmul_rn(t4, t4), t16 = __dmul_rn(t8, t8);
  const double an = __dmul_rn(__dmul_rn(t16, t2), T);                        // T^19
  const double num = fma(26.0, an, __dmul_rn(24.0, yn)), den = fma(24.0, an, __dmul_rn(26.0, yn));
  return __dmul_rn(__dmul_rn(__dmul_rn(y0, num), rcp_nr(den)), inv_Re);
#else
  return __dmul_rn(exp(__dmul_rn(0.76, log(T))), inv_Re);
#endif
}

__device__ __forceinline__ void prim4(const double* __restrict__ u, long long npg, long long p, double gamma, double (&q)[4])
{
  double r[5];
  load5(u, npg, p, r);
  prim_of(r, gamma, q);
}

// One thread per point of a box, general in position (biased stencils next to the line ends, transverse
// indices must be interior). The production path uses it for the six ghost slabs only (3.5 % of the points at
// 512^3); the interior is k_qderiv_int below.
__global__ void __launch_bounds__(128) k_qderiv3(const QD3Args a)
{
  const Geom& G = a.G;
  const int g = G.g;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)a.ext[0] * a.ext[1] * a.ext[2]) return;
  const int i = a.lo[0] + (int)(t % a.ext[0]);
  const int j = a.lo[1] + (int)((t / a.ext[0]) % a.ext[1]);
  const int k = a.lo[2] + (int)(t / ((long long)a.ext[0] * a.ext[1]));
  const bool in0 = (i >= 0 && i < G.N[0]), in1 = (j >= 0 && j < G.N[1]), in2 = (k >= 0 && k < G.N[2]);
  if ((int)in0 + (int)in1 + (int)in2 < 2) return;                     // edges / corners: never touched
  const long long p = (i + g) + (long long)G.P[0] * ((j + g) + (long long)G.P[1] * (k + g));
  const int idx[3] = { i, j, k };
  const bool inn[3] = { in0, in1, in2 };
  const double s12 = 1.0 / 12.0;
  // mu/Re at this point: mu = exp(0.76 log T) (raiseto, math_ops.h:37; NavierStokes3DParabolicFunction.c:174)
  double qc[4];
  prim4(a.u, G.npg, p, a.gamma, qc);
  const double muRe = mu_over_Re(qc[3], a.inv_Re);
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
      for (int c = 0; c < 4; c++) D[c] = central4(fm2[c], fm1[c], fp1[c], fp2[c]);
    }
    // NOT scaled by dxinv here: the reference scales AFTER the halo exchange with the receiver's local dxinv
    // (NavierStokes3DParabolicFunction.c:125-146); the sweeps apply it per cell
    const double dxi = muRe;
#pragma unroll
    for (int c = 0; c < 4; c++) a.qd[(long long)(d * 4 + c) * G.npg + p] = __dmul_rn(D[c], dxi);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Interior points: all three differences are the central one, (f[-2] - 8 f[-1] + 8 f[+1] - f[+2]) / 12
// (FirstDerivativeFourthOrder.c:103). A CTA owns a 32(x) x 16(y) column (measured at 512^3: 4.5 ms against 5.1 ms with 32 x 8) and marches over `zchunk` planes:
//   * the primitive variables (u, v, w, T) of a point are evaluated ONCE per column and kept in a 5-plane
//     register window (z-difference from registers);
//   * the centre plane plus a 2-cell x/y halo (evaluated by the first 160 threads) is staged in shared memory,
//     double-buffered: one __syncthreads per plane (x- and y-differences from shared memory);
//   * 12 coalesced stores per point.
// Work per point: 1.6 primitive evaluations instead of 13, one reciprocal (Newton, no IEEE division),
// exp + log for mu = T^0.76. HBM: 40 B read + 96 B written per point.
#ifndef Q_TY
#define Q_TY 16
#endif
#ifndef Q_DEPTH
#define Q_DEPTH 4
#endif
#ifndef Q_ZCHUNK
#define Q_ZCHUNK 64
#endif
#ifndef Q_MINB
#define Q_MINB 1
#endif
#ifndef Q_STCS
#define Q_STCS 0       // streaming stores measured no faster (profiles/r02f_variants.txt: 18.7 vs 18.3 ms per step)
#endif
#if Q_STCS
#define QD_STORE(ptr, val) __stcs((ptr), (val))
#else
#define QD_STORE(ptr, val) (*(ptr) = (val))
#endif
#ifndef Q_TX
#define Q_TX 32       // 64 x 8 tiles (longer contiguous rows per output stream) measured in profiles/r02v_variants.txt
#endif
constexpr int QTX = Q_TX, QTY = Q_TY, QH = 2;
static_assert(4 * QTY + 4 * QTX <= QTX * QTY, "every halo point of a plane needs a thread");
constexpr int QSX = QTX + 2 * QH + 1;      // padded row (37): conflict-free column access is not needed, rows are read along x
constexpr int QSY = QTY + 2 * QH;

constexpr int QD = Q_DEPTH;                // depth of the cp.async input ring (planes in flight per thread)
constexpr int QNH = 4 * QTY + 4 * QTX;     // halo points per plane (160)
constexpr size_t QSMEM = sizeof(double) * (2 * 4 * QSY * QSX + QD * 5 * (QTX * QTY) + QD * 5 * QNH);

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gmem_src) : "memory");
}

__global__ void __launch_bounds__(QTX * QTY, Q_MINB) k_qderiv_int(const QD3Args a)
{
  extern __shared__ double qsm[];
  double (*Q)[4][QSY][QSX] = reinterpret_cast<double (*)[4][QSY][QSX]>(qsm);            // [2][4][QSY][QSX]
  double* rawA = qsm + 2 * 4 * QSY * QSX;                                               // [QD][5][256]: own column
  double* rawB = rawA + QD * 5 * (QTX * QTY);                                           // [QD][5][160]: halo point
  const Geom& G = a.G;
  const int g = G.g;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * QTX + tx;
  // the box of points this launch owns: [lo, lo + ext) in interior-relative indices (the whole interior, or the
  // part of it that does not read ghost cells when the halo exchange is still in flight)
  const int i0 = a.lo[0] + blockIdx.x * QTX, j0 = a.lo[1] + blockIdx.y * QTY;
  const int i = i0 + tx, j = j0 + ty;
  const int ie = a.lo[0] + a.ext[0], je = a.lo[1] + a.ext[1];
  const int kb = a.lo[2] + blockIdx.z * a.zchunk, ke = min(kb + a.zchunk, a.lo[2] + a.ext[2]);
  const bool ok = (i < ie) && (j < je);
  // partial tiles: the two columns / rows just outside the box lie INSIDE the thread tile; their threads
  // own no output but must stage the centre plane for their neighbours
  const bool edge = !ok && ((i < ie + QH && j < je) || (i < ie && j < je + QH));
  const long long npg = G.npg, sz = G.st[2];
  const long long pcol = (i + g) + (long long)G.P[0] * (j + g);             // + P0*P1*(k+g)
  // halo assignment of threads 0..159: 4 columns x 8 rows (x-halo), then 32 columns x 4 rows (y-halo)
  int hi = 0, hj = 0; bool hok = false;
  if (tid < 4 * QTY) {
    const int c = tid % 4, r = tid / 4;
    hi = (c < 2) ? (c - 2) : (QTX + c - 2); hj = r; hok = true;
  } else if (tid < QNH) {
    const int t = tid - 4 * QTY, c = t % QTX, r = t / QTX;
    hi = c; hj = (r < 2) ? (r - 2) : (QTY + r - 2); hok = true;
  }
  const int ghi = i0 + hi, ghj = j0 + hj;
  hok = hok && (ghi + g < G.P[0]) && (ghj + g < G.P[1]);
  const long long phalo = (ghi + g) + (long long)G.P[0] * (ghj + g);

  // Inputs arrive through a QD-deep cp.async ring, each thread copying (and later reading) only its own slots, so no
  // barrier is involved: stream A = the thread's own column (plane k+2 for owners: it enters the z-window; plane k for
  // the edge threads), stream B = its halo point (plane k). QD-1 planes are in flight per thread.
  const bool actA = ok || edge;
  const int offA = ok ? 2 : 0;
  auto issue = [&](int it) {                                                 // inputs of iteration it (plane kb + it)
    const int k = kb + it;
    if (k < ke) {
      const int slot = it % QD;
      if (actA) {
        const long long p = pcol + sz * (k + offA + g);
#pragma unroll
        for (int v = 0; v < 5; v++) cp_async8(rawA + (slot * 5 + v) * (QTX * QTY) + tid, a.u + v * npg + p);
      }
      if (hok) {
        const long long p = phalo + sz * (k + g);
#pragma unroll
        for (int v = 0; v < 5; v++) cp_async8(rawB + (slot * 5 + v) * QNH + tid, a.u + v * npg + p);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int it = 0; it < QD - 1; it++) issue(it);

  double w[5][4];                                                            // planes k-2 .. k+2
#pragma unroll
  for (int s = 0; s < 5; s++) { w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0.0; }
  if (ok) {
#pragma unroll
    for (int s = 1; s < 5; s++) {                                            // w[1..4] = planes kb-2 .. kb+1
      double r[5];
      load5(a.u, npg, pcol + sz * (kb + s - 3 + g), r);
      prim_of(r, a.gamma, w[s]);
    }
  }
  for (int k = kb; k < ke; k++) {
    const int it = k - kb;
    const int b = it & 1, slot = it % QD;
    const long long pk = sz * (k + g);
    asm volatile("cp.async.wait_group %0;" :: "n"(QD - 2) : "memory");      // this thread's copies of iteration `it` landed
    // shift the window; plane k+2 enters
#pragma unroll
    for (int s = 0; s < 4; s++) { w[s][0] = w[s + 1][0]; w[s][1] = w[s + 1][1]; w[s][2] = w[s + 1][2]; w[s][3] = w[s + 1][3]; }
    if (actA) {
      double r[5];
#pragma unroll
      for (int v = 0; v < 5; v++) r[v] = rawA[(slot * 5 + v) * (QTX * QTY) + tid];
      if (ok) prim_of(r, a.gamma, w[4]); else prim_of(r, a.gamma, w[2]);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) Q[b][c][ty + QH][tx + QH] = w[2][c];
    if (hok) {
      double r[5], h[4];
#pragma unroll
      for (int v = 0; v < 5; v++) r[v] = rawB[(slot * 5 + v) * QNH + tid];
      prim_of(r, a.gamma, h);
#pragma unroll
      for (int c = 0; c < 4; c++) Q[b][c][hj + QH][hi + QH] = h[c];
    }
    issue(it + QD - 1);                                                       // refills the slot consumed one iteration ago
    __syncthreads();
    if (ok) {
      const double muRe = mu_over_Re(w[2][3], a.inv_Re);
      const long long p = pcol + pk;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const double* r = &Q[b][c][ty + QH][tx + QH];
        const double dx = central4(r[-2], r[-1], r[1], r[2]);
        const double dy = central4(r[-2 * QSX], r[-QSX], r[QSX], r[2 * QSX]);
        const double dz = central4(w[0][c], w[1][c], w[3][c], w[4][c]);
        // streaming stores: 12 output streams per CTA that nothing re-reads before the whole array has passed through L2
        QD_STORE(a.qd + (long long)(0 * 4 + c) * npg + p, __dmul_rn(dx, muRe));
        QD_STORE(a.qd + (long long)(1 * 4 + c) * npg + p, __dmul_rn(dy, muRe));
        QD_STORE(a.qd + (long long)(2 * 4 + c) * npg + p, __dmul_rn(dz, muRe));
      }
    }
  }
}

} // namespace

namespace hpbk {

// All Q-derivatives of one stage: the interior in one tiled launch, the six ghost slabs (normal derivative only; the
// outermost layer is skipped inside the kernel) with the per-point kernel.
int qderiv_fused(hpb_solver* h, const double* u, int part)
{
  (void)part;
  ProfScope ps(h, HPB_PROF_VISCOUS);
  const Geom& G = h->geo;
  // the shared-memory opt-in is a per-device attribute: one bit per device ordinal
  static unsigned long long configured = 0ull;
  const unsigned long long dbit = 1ull << (h->device & 63);
  if (!(configured & dbit)) {
    if (cudaFuncSetAttribute(k_qderiv_int, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QSMEM) != cudaSuccess) {
      cudaGetLastError();
      return hpb_fail(HPB_ERR_CUDA, "k_qderiv_int: cannot reserve %zu bytes of shared memory", QSMEM);
    }
    configured |= dbit;
  }
  QD3Args a; a.G = G; a.gamma = h->phys.gamma; a.inv_Re = 1.0 / h->phys.Re; a.u = u; a.dxinv = h->d_dxinv; a.qd = h->d_qd4;
  a.zchunk = Q_ZCHUNK;
  {
    for (int d = 0; d < 3; d++) { a.lo[d] = 0; a.ext[d] = G.N[d]; }
    dim3 grid((G.N[0] + QTX - 1) / QTX, (G.N[1] + QTY - 1) / QTY, (G.N[2] + a.zchunk - 1) / a.zchunk);
    k_qderiv_int<<<grid, dim3(QTX, QTY, 1), QSMEM, h->stream>>>(a);
    h->launches++;
  }
  for (int d = 0; d < 3; d++) {
    for (int f = 0; f < 2; f++) {
      for (int k = 0; k < 3; k++) { a.lo[k] = 0; a.ext[k] = G.N[k]; }
      a.lo[d] = f ? G.N[d] : -G.g;
      a.ext[d] = G.g;
      const long long n = (long long)a.ext[0] * a.ext[1] * a.ext[2];
      if (n <= 0) continue;
      k_qderiv3<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(a);
      h->launches++;
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return hpb_fail(HPB_ERR_CUDA, "Q-derivative kernels: launch failed: %s", cudaGetErrorString(e));
  return HPB_OK;
}

} // namespace hpbk

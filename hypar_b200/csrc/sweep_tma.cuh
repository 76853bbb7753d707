// sweep_tma.cuh -- the fused directional sweep of sweep_fused.cuh with all global<->shared traffic moved
// to the TMA unit and every CTA-wide barrier of the march removed (NavierStokes2D / NavierStokes3D).
//
// Same arithmetic and work decomposition as k_sweep (one WARP per grid line, 8 lines per CTA, 32 cells
// per march step, records / exchanges warp-private in shared memory, reference citations in
// sweep_fused.cuh). What changes is how the data moves (profiles/r01b: in k_sweep 22 % of the warp-stall
// samples of the y/z sweeps sat in the cp.async / LDG / STG phases and their address arithmetic):
//
//   inputs   one elected thread issues cp.async.bulk.tensor (4-D tiled, ghost-padded SoA array seen as
//            [field][z][y][x]) for the whole CTA tile of the next step: u (all components in one box),
//            8 derivative scalars, gravity fields. x-sweep: box 32(x) x 8 lines(y); y-/z-sweep: box
//            8 lines(x) x 32 cells, written with the 64-byte swizzle so that a warp reading ITS line
//            (stride 64 B) is 2-way instead of 16-way bank conflicted. Completion: mbarrier expect_tx.
//            Ghost cells beyond the array and the start-up step use the TMA's zero fill (negative
//            coordinates are legal for loads).
//   results  y-/z-sweep (always accumulating): the CTA's 32x8 tile of each component is staged in shared
//            memory (same swizzle) and ONE cp.reduce.async.bulk.tensor ... .add per step performs the
//            read-modify-write in L2: no loads of the old right-hand side, no registers for them. Cells
//            outside the interior are staged as +0.0 (adding zero to a ghost entry changes nothing; the
//            box is clipped at the array bounds by the TMA).  x-sweep (always the first direction:
//            overwrites): lanes run along x, so each warp stores its own 256 contiguous bytes from
//            registers.
//   stage    RKF variant of the LAST direction's sweep (y in 2-D, z in 3-D): instead of adding its flux difference to the
//            right-hand side in L2, the CTA loads the tile of the right-hand side accumulated so far into the result tile
//            and the tile of u^n into the tile of the next stage solution (TMA, issued as soon as the previous step's stores
//            have read them), completes k = rhs + own part in place, forms the next stage solution u^n + (a dt) k in place
//            beside it -- the two roundings of k_rk_combine -- and stores BOTH
//            tiles with plain cp.async.bulk.tensor stores: the stage-vector pass over memory (TimeRK.c:131-141) disappears
//            under a kernel that is bound by the FP64 pipe. Cells of the tile outside the interior are written back
//            unchanged (k) or as zero (next stage solution: ghost lines / ghost cells, filled by the boundary conditions
//            and the halo exchange afterwards).
//   sync     no __syncthreads in the march. The last warp to finish reading the input tile of step m
//            (shared atomic counter) issues the loads of step m+1; the last warp to stage its results
//            issues the reduce and, when the TMA has read the tile, frees it through an mbarrier. Warps
//            of a CTA can drift by up to one step.
//
// TMA constraints (tools/tma_probe.cu, checked on the B200): the innermost box coordinate must be
// 16-byte aligned (even index) for loads and stores, stores must not start at negative coordinates,
// FP64 reduce-add is supported. Hence: x-sweep loads start at padded index 32m+6; the 8-line tiles of
// the y-/z-sweeps start at padded x index 8t+2 (grid line x = 8t-1: the first line of the first tile is
// a ghost line and idles); the padded row length P0 must be even (otherwise the caller falls back to
// k_sweep).
#pragma once
#include <cuda.h>
#include "sweep_fused.cuh"

// where the thread that issued the result stores takes the owed wait: after P1 of the next step (0, default) or at its top (1)
#ifndef HPB_PEND_TOP
#define HPB_PEND_TOP 0
#endif
// x-sweep result stores: plain (default) or streaming (-DHPB_XS_STCS=1; measured in profiles/r02f_variants.txt)
#ifndef HPB_XS_STCS
#define HPB_XS_STCS 0
#endif
#if HPB_XS_STCS
#define XS_STORE(ptr, val) __stcs((ptr), (val))
#else
#define XS_STORE(ptr, val) (*(ptr) = (val))
#endif

namespace hpbf {

struct TmaMaps {
  CUtensorMap u, qd, out, gf, gg;
  CUtensorMap un, u0;      // RKF: the next stage solution; the solution u^n at the start of the step
};
// host (sweep_fused.cu): descriptors for the arrays of one launch; false when the TMA path does not apply
// (driver entry point missing, odd padded row length, misaligned array)
bool tma_maps_for(hpb_solver* h, const SweepArgs& a, bool xs, bool grav, bool visc, TmaMaps* tm);
bool tma_available();      // the driver's cuTensorMapEncodeTiled entry point was found (and HPB_NO_TMA is not set)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  unsigned done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load4(void* dst, const CUtensorMap* m, unsigned long long* bar, int c0, int c1, int c2, int c3)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               :: "r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_reduce_add4(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3)
{
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               :: "l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store4(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3)
{
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               :: "l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory carve-up (doubles). TMA tiles first (the 64-byte swizzle is a function of address bits 4..8: the
// dynamic shared memory is declared 1024-byte aligned and every tile is 2048 B), then the warp-private
// records and exchanges of k_sweep. Without gravity the mass flux rho*v_d IS the conserved momentum component
// (to rounding), so its record field, reconstruction and exchange slot are dropped (SKIPF0).
template <int MODEL, bool GRAV, bool VISC, bool RKF = false>
struct TmaLayout {
  using SL = SweepLayout<MODEL, GRAV, VISC>;
  static constexpr int NV = RecLayout<MODEL>::NV;
  static constexpr int NDV = NV - 2;
  static constexpr bool G3 = SL::G3, V3 = SL::V3;
  static constexpr bool SKIPF0 = !G3;
  static constexpr int SK = SKIPF0 ? 1 : 0;
  // record fields
  static constexpr int rU = 0;
  static constexpr int rF = NV - SK;                        // flux component v at rF + v (v >= SK)
  static constexpr int rV4 = 2 * NV - SK;
  static constexpr int rSR = rV4 + 1;
  static constexpr int rVEL = rSR + 1;
  static constexpr int rH = rVEL + NDV;
  static constexpr int rA = rH + 1;
  static constexpr int rGF = rA + 1;                        // gravity f, g
  static constexpr int rFV = rGF + (G3 ? 2 : 0);            // viscous flux (4)
  static constexpr int NFREC = rFV + (V3 ? 4 : 0);
  // exchange fields (left-biased values): flux v at xF + v (v >= SK), solution v at xU + v, gravity source xZ + {0,1}
  static constexpr int xF = -SK;
  static constexpr int xU = NV - SK;
  static constexpr int xZ = 2 * NV - SK;
  static constexpr int NFL = xZ + (G3 ? 2 : 0);
  static constexpr int NFF = SL::NFF;                       // interface flux (, S x2)
  static constexpr int TILE = TL * TW;                      // doubles per field tile (2048 B)
  static constexpr int NFIN = SL::NFSTG;                    // input fields: u, (gf, gg), (8 derivative scalars)
  static constexpr int o_in = 0;
  static constexpr int o_out = o_in + NFIN * TILE;          // result tile (y-/z-sweeps)
  static constexpr int o_ust = o_out + NV * TILE;           // RKF: tile of the next stage solution
  static constexpr int o_rec = o_ust + (RKF ? NV * TILE : 0);
  static constexpr int o_exL = o_rec + NFREC * NREC;
  static constexpr int o_car = o_exL + NFL * NEX;           // interface-flux carry of every line (lane 31 -> lane 0 of the next step)
  static constexpr int o_sync = o_car + NFF * TW;           // 3 mbarriers (full, free, rin) + 2 counters
  static constexpr size_t smem_bytes = sizeof(double) * (size_t)(o_sync + 4);
};

// CHR: characteristic-wise reconstruction (Interp1PrimFifthOrderWENOChar.c:86-198 with the weights of
// WENOFifthOrderCalculateWeights.c:760-1440). The lane owns an INTERFACE instead of a centred stencil: the six cells
// around it are projected on the left eigenvectors of the Roe average of the two adjacent cells, every characteristic
// field is reconstructed from both sides (flux with flux weights, solution with solution weights), and the upwind flux is
// formed in characteristic space -- 1/2 (fL + fR) - 1/2 |lambda| (uR - uL) per field (Roe; alpha instead of |lambda| for
// Rusanov) -- so that ONE multiplication by the right eigenvectors returns the interface flux (the reference multiplies
// fL, fR, uL, uR back one by one and Roe's scheme projects uR - uL again: the same numbers, L R = I). The eigenvectors
// are applied in closed form (see roe_dissipation): with beta = (gamma-1)(ek q0 - v.qm + qE) and dn = vn q0 - q_n the
// five amplitudes of a vector q are q0 - beta/a^2, (beta +- a dn)/(2 a^2) and q_t - v_t q0 -- the reference's rows up to
// the sign of the shear rows, which cancels against the sign of the matching column (WENO is odd in its data and its
// weights are even). Only (q0, beta, dn) of the six cells are kept in registers; the shear fields re-read their
// component. No gravity in this variant.
template <int MODEL, int WT, bool XS, bool GRAV, bool VISC, bool CHR = false, bool RKF = false>
__global__ void __launch_bounds__(NT, 2) k_sweep_tma(const SweepArgs a, const __grid_constant__ TmaMaps tm)
{
  static_assert(!(RKF && XS), "the stage-vector variant belongs to the accumulating sweeps");
  using LY = TmaLayout<MODEL, GRAV, VISC, RKF>;
  using SL = SweepLayout<MODEL, GRAV, VISC>;
  constexpr int NV = LY::NV;
  constexpr int NDV = NV - 2;
  constexpr bool G3 = SL::G3, V3 = SL::V3;
  constexpr bool SKIPF0 = LY::SKIPF0;
  constexpr int TILE = LY::TILE;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  double* stg = smem + LY::o_in;
  double* ost = smem + LY::o_out;
  double* rec = smem + LY::o_rec;
  double* exL = smem + LY::o_exL;
  double* car = smem + LY::o_car;
  double* ust = smem + LY::o_ust;
  unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(smem + LY::o_sync);
  unsigned long long* free_bar = full_bar + 1;
  unsigned* cons_cnt = reinterpret_cast<unsigned*>(full_bar + 2);
  unsigned* res_cnt = cons_cnt + 1;
  unsigned long long* rin_bar = full_bar + 3;

  const Geom& G = a.G;
  const int dir = a.dir;
  const int N = G.N[dir];
  const long long npg = G.npg;
  const int tid = threadIdx.x;
  const int w = tid >> 5, l = tid & 31;
  const int g1 = (G.ndims > 1) ? G.g : 0, g2 = (G.ndims > 2) ? G.g : 0;

  // tile -> lines. x-sweep: 8 consecutive y at one z; y-sweep: 8 consecutive x (from x = 8t-1) at one z;
  // z-sweep: 8 consecutive x at one y.
  int tc0, tc1, tc2;               // TMA coordinates of the tile origin (the entry along `dir` is set per step)
  bool line_ok;
  long long pline = 0;             // x-sweep: offset of cell 0 of this warp's line
  int ia, ib;                      // transverse indices of this warp's line, increasing dimension order
  if (XS) {
    const int nty = (G.N[1] + TW - 1) / TW;
    const int ty = blockIdx.x % nty, z = blockIdx.x / nty;
    ia = ty * TW + w; ib = z;
    line_ok = ia < G.N[1];
    tc0 = 0; tc1 = ty * TW + g1; tc2 = z + g2;
    pline = G.g + (long long)G.P[0] * ((ia + g1) + (long long)G.P[1] * (ib + g2));
  } else {
    const int ntx = (G.N[0] + 1 + TW - 1) / TW;
    const int tx = blockIdx.x % ntx, o = blockIdx.x / ntx;
    ia = tx * TW - 1 + w; ib = o;
    line_ok = ia >= 0 && ia < G.N[0];
    tc0 = tx * TW + 2;
    if (dir == 1) { tc1 = 0; tc2 = o + g2; } else { tc1 = o + G.g; tc2 = 0; }
  }
  const double gamma = a.ph.gamma;
  const int M = (N + TL - 1) / TL;

  // staging slot of (cell l of the step, line w): x-sweep [w][l]; y-/z-sweep [l][w] with the 64-byte swizzle
  // (16-byte chunk index ^= bits 7..8 of the byte offset = (l >> 1) & 3)
  const int sidx = XS ? (w * TL + l) : (l * TW + (w ^ (((l >> 1) & 3) << 1)));

  double vsc0 = 1.0, vsc1 = 1.0;
  if (V3 && line_ok) {
    const int ta = (dir == 0) ? 1 : 0, tb = (dir == 2) ? 1 : 2;
    vsc0 = a.dxinv[G.xoff[ta] + G.g + ia];
    vsc1 = a.dxinv[G.xoff[tb] + G.g + ib];
  }
  const int rbase = w * RP;
  const int xbase = w * XP;
  const double* dxl = a.dxinv + G.xoff[dir] + G.g;

  auto issue_loads = [&](int m) {
    // executed by ONE thread: the input tile of step m (cells 32m+3 .. 32m+34 of the 8 lines)
    mbar_expect_tx(full_bar, (unsigned)(LY::NFIN * TILE * sizeof(double)));
    const int cd = TL * m + 3 + G.g;
    int c0 = tc0, c1 = tc1, c2 = tc2;
    if (dir == 0) c0 = cd; else if (dir == 1) c1 = cd; else c2 = cd;
    tma_load4(stg, &tm.u, full_bar, c0, c1, c2, 0);
    if (G3) {
      tma_load4(stg + SL::SGI * TILE, &tm.gf, full_bar, c0, c1, c2, 0);
      tma_load4(stg + (SL::SGI + 1) * TILE, &tm.gg, full_bar, c0, c1, c2, 0);
    }
    if (V3) {
#pragma unroll
      for (int k = 0; k < 8; k++) tma_load4(stg + (SL::SQI + k) * TILE, &tm.qd, full_bar, c0, c1, c2, a.qidx[k]);
    }
  };

  auto issue_rin = [&](int m) {
    // RKF, executed by ONE thread: the right-hand side accumulated by the earlier directions on the cells step m writes
    // and u^n there (the stage solution in the records is U_s, not u^n, from the second stage on)
    mbar_expect_tx(rin_bar, (unsigned)(2 * NV * TILE * sizeof(double)));
    const int cd = TL * m + G.g;
    int c0 = tc0, c1t = tc1, c2t = tc2;
    if (dir == 1) c1t = cd; else c2t = cd;
    tma_load4(ost, &tm.out, rin_bar, c0, c1t, c2t, 0);
    tma_load4(ust, &tm.u0, rin_bar, c0, c1t, c2t, 0);
  };

  if ((smem_u32(smem) & 1023u) != 0) __trap();      // the swizzled tiles assume the declared alignment
  if (tid == 0) {
    mbar_init(full_bar, 1);
    mbar_init(free_bar, 1);
    mbar_init(rin_bar, 1);
    *cons_cnt = 0; *res_cnt = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();
  if (tid == 0) { issue_loads(-1); if (RKF && M > 0) issue_rin(0); }
  // this thread has issued the result stores / the reduce of the previous step and still owes the wait for their shared-memory
  // reads (then: the tiles' next loads, or the release of the result tile). The wait is taken after P1 of the NEXT step, not
  // right after the issue: the warp that stages last is the slowest of its CTA, and everything it does alone delays them all
  // (profiles/r02l: the input-tile wait of the other warps doubled when it also waited for the stores and issued the loads)
  bool pend = false;

  for (int m = -1; m < M; m++) {
    const int c1 = TL * m + 3 + l;
    const int j = TL * m + 1 + l;
    const int jo = j - 1;
    // grid metrics of this step's cells (L1 hits: the same values serve all lines), requested before the wait
    double dd = 0.0, dxi = 0.0;
    if (V3 && c1 >= -3 && c1 <= N + 2) dd = dxl[c1];
    if (jo >= 0 && jo < N) dxi = dxl[jo];
    const double dxih = 0.5 * dxi;

#if HPB_PEND_TOP
    // (variant: the owed wait at the top of the step, before this warp waits for its own input tile)
    if (!XS && pend) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      if (RKF) issue_rin(m);
      else mbar_arrive(free_bar);
      pend = false;
    }
#endif
    mbar_wait(full_bar, (unsigned)(m + 1) & 1u);

    // ---------------- P1: record of cell c1 = 32m+3+l (record position 5+l)
    if (line_ok && c1 >= -3 && c1 <= N + 2) {
      const double* s = stg + sidx;
      const int ci = rbase + 5 + l;
      double U[NV];
#pragma unroll
      for (int v = 0; v < NV; v++) U[v] = s[v * TILE];
#pragma unroll
      for (int v = 0; v < NV; v++) rec[(LY::rU + v) * NREC + ci] = U[v];
      const double rho = U[0];
      double vel[3] = { 0.0, 0.0, 0.0 };
      const double rinv = rcp_fast(rho);          // face ghosts are filled: rho > 0 wherever a record is built
#pragma unroll
      for (int k = 0; k < NDV; k++) vel[k] = U[1 + k] * rinv;
      double vsq = 0.0;
#pragma unroll
      for (int k = 0; k < NDV; k++) vsq += vel[k] * vel[k];
      const double e = U[NV - 1];
      const double ke = 0.5 * rho * vsq;
      const double P = (e - ke) * (gamma - 1.0);
      const double vn = (dir == 0) ? vel[0] : (dir == 1 ? vel[1] : vel[2]);
      if (!SKIPF0) rec[(LY::rF + 0) * NREC + ci] = rho * vn;
#pragma unroll
      for (int k = 0; k < NDV; k++) rec[(LY::rF + 1 + k) * NREC + ci] = rho * vn * vel[k] + (k == dir ? P : 0.0);
      rec[(LY::rF + NV - 1) * NREC + ci] = (e + P) * vn;
      double gfv = 1.0, ggv = 1.0;
      if (G3) {
        gfv = s[SL::SGI * TILE]; ggv = s[(SL::SGI + 1) * TILE];
        rec[LY::rGF * NREC + ci] = gfv; rec[(LY::rGF + 1) * NREC + ci] = ggv;
      }
      const double igm1 = 1.0 / (gamma - 1.0);
      // modified solution (NavierStokes3DModifiedSolution.c:65): without gravity its energy P/(gamma-1) + rho|v|^2/2 is
      // the conserved energy itself (to rounding), like the other components: nothing to store
      if (G3) rec[LY::rV4 * NREC + ci] = (P * igm1) * rcp_fast(ggv) + ke * gfv;
      const double c2 = gamma * P * rinv;
      rec[LY::rSR * NREC + ci] = sqrt_fast(rho);
#pragma unroll
      for (int k = 0; k < NDV; k++) rec[(LY::rVEL + k) * NREC + ci] = vel[k];
      rec[LY::rH * NREC + ci] = 0.5 * vsq + c2 * igm1;
      rec[LY::rA * NREC + ci] = sqrt_fast(c2) + fabs(vn);
      if (V3) {
        double q[8];
#pragma unroll
        for (int k = 0; k < 8; k++) q[k] = s[(SL::SQI + k) * TILE] * (k < 4 ? dd : (k < 6 ? vsc0 : vsc1));
        const double two_third = 2.0 / 3.0;
        const double kq = igm1 * (1.0 / a.ph.Pr);
        double t1, t2, t3;
        if (dir == 0)      { t1 = two_third * (2 * q[0] - q[5] - q[7]); t2 = q[4] + q[1]; t3 = q[6] + q[2]; }
        else if (dir == 1) { t1 = q[0] + q[5]; t2 = two_third * (-q[4] + 2 * q[1] - q[7]); t3 = q[6] + q[2]; }
        else               { t1 = q[0] + q[5]; t2 = q[1] + q[7]; t3 = two_third * (-q[4] - q[6] + 2 * q[2]); }
        rec[(LY::rFV + 0) * NREC + ci] = t1;
        rec[(LY::rFV + 1) * NREC + ci] = t2;
        rec[(LY::rFV + 2) * NREC + ci] = t3;
        rec[(LY::rFV + 3) * NREC + ci] = vel[0] * t1 + vel[1] * t2 + vel[2] * t3 + kq * q[3];
      }
    }
    __syncwarp();
    // the input tile is consumed by this warp; the last warp of the CTA to get here requests the next one
    if (l == 0) {
      __threadfence_block();
      const unsigned old = atomicAdd(cons_cnt, 1u);
      if ((old & (TW - 1)) == TW - 1 && m + 1 < M) {
        __threadfence_block();
        issue_loads(m + 1);
      }
#if !HPB_PEND_TOP
      if (!XS && pend) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (RKF) issue_rin(m);               // the tiles of step m-1 have been read: this step's right-hand side and u^n
        else mbar_arrive(free_bar);
        pend = false;
      }
#endif
    }

    const int cc = rbase + l + 3;
    double fh[NV], Sh[2] = { 0.0, 0.0 };
#pragma unroll
    for (int v = 0; v < NV; v++) fh[v] = 0.0;
    if constexpr (CHR) {
    // ---------------- P2+P3 (characteristic-wise): interface j-1/2 between cells jo = j-1 (position l+2) and j (l+3);
    // stencil cells jo-2 .. jo+3 = positions l .. l+5
    const bool if_ok = line_ok && j >= 0 && j <= N;
    if (if_ok) {
      const int cL = cc - 1, cR = cc;
      const int p0 = rbase + l;
      const double gm1 = gamma - 1.0;
      const double tL = rec[LY::rSR * NREC + cL], tR = rec[LY::rSR * NREC + cR];
      const double rs = rcp_fast(tL + tR);
      double vsq = 0.0, vn = 0.0, vhat[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
      for (int k = 0; k < NDV; k++) {
        const double v = (tL * rec[(LY::rVEL + k) * NREC + cL] + tR * rec[(LY::rVEL + k) * NREC + cR]) * rs;
        vsq += v * v;
        vhat[k] = v;
        if (k == dir) vn = v;
      }
      const double H = (tL * rec[LY::rH * NREC + cL] + tR * rec[LY::rH * NREC + cR]) * rs;
      const double ek = 0.5 * vsq;
      const double a2 = gm1 * (H - ek);
      const double aa = sqrt_fast(a2);
      const double ia2 = rcp_fast(a2), hia2 = 0.5 * ia2;
      // dissipation speed per field: Roe |lambda| with Harten's fix, or the Rusanov alpha for all of them
      double lam0, lamm, lamp;
      if (a.upw == 1) { lam0 = harten_abs(vn); lamm = harten_abs(vn - aa); lamp = harten_abs(vn + aa); }
      else {
        const double alpha = fmax(fmax(rec[LY::rA * NREC + cL], rec[LY::rA * NREC + cR]), aa + fabs(vn));
        lam0 = lamm = lamp = alpha;
      }
      // Two passes over the six cells, flux first, then the solution (one pass keeps (q0, beta, dn) of the six cells in
      // registers: 18 doubles; both at once would spill under the 128-register cap of two CTAs per SM). Per field k the
      // pass returns fL + fR (flux) or uR - uL (solution); g[k] = (fL + fR) - lam_k (uR - uL) = 2 x the upwind amplitude.
      // Field order: entropy, vn - a, vn + a, then the shear wave of every tangential component.
      double g[5] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
      auto pass = [&](const bool isU) {
        const int rQ = isU ? LY::rU : LY::rF;
        double Q0[6], QB[6], QD[6];
#pragma unroll
        for (int c = 0; c < 6; c++) {
          const int pc = p0 + c;
          double qm[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
          for (int k = 0; k < NDV; k++) qm[k] = rec[(rQ + 1 + k) * NREC + pc];
          // without gravity the mass flux is the normal momentum component itself (no record field for it)
          const double q0 = (!isU && SKIPF0) ? rec[(LY::rU + 1 + dir) * NREC + pc] : rec[rQ * NREC + pc];
          const double qE = rec[(rQ + NV - 1) * NREC + pc];
          double vq = 0.0;
#pragma unroll
          for (int k = 0; k < NDV; k++) vq = fma(vhat[k], qm[k], vq);
          double qn = 0.0;
#pragma unroll
          for (int k = 0; k < NDV; k++) if (k == dir) qn = qm[k];
          Q0[c] = q0; QB[c] = gm1 * (fma(ek, q0, qE) - vq); QD[c] = fma(vn, q0, -qn);
        }
        auto both = [&](const double (&X)[6]) -> double {
          double vL, vR, d0, d1, d2;
          const double XL[5] = { X[0], X[1], X[2], X[3], X[4] }, XR[5] = { X[1], X[2], X[3], X[4], X[5] };
          recon_pair<WT, false, true>(XL, XL, XL, a.ph.eps, vL, d0, d1, d2);    // left-biased value of the stencil centred on jo
          recon_pair<WT, false, true>(XR, XR, XR, a.ph.eps, d0, vR, d1, d2);    // right-biased value of the stencil centred on j
          return isU ? (vR - vL) : (vL + vR);
        };
        auto put = [&](int k, double val, double lam) { g[k] = isU ? fma(-lam, val, g[k]) : val; };
        double X[6];
#pragma unroll
        for (int c = 0; c < 6; c++) X[c] = fma(-QB[c], ia2, Q0[c]);
        put(0, both(X), lam0);
#pragma unroll
        for (int c = 0; c < 6; c++) X[c] = hia2 * fma(aa, QD[c], QB[c]);
        put(1, both(X), lamm);
#pragma unroll
        for (int c = 0; c < 6; c++) X[c] = hia2 * fma(-aa, QD[c], QB[c]);
        put(2, both(X), lamp);
        int kf = 3;
#pragma unroll
        for (int k = 0; k < NDV; k++) {
          if (k == dir) continue;
#pragma unroll
          for (int c = 0; c < 6; c++) X[c] = fma(-vhat[k], Q0[c], rec[(rQ + 1 + k) * NREC + p0 + c]);
          put(kf, both(X), lam0);
          kf++;
        }
      };
      pass(false);
      pass(true);
      // back to conserved variables: one multiplication by the right eigenvectors
      const double g0 = g[0], gm = g[1], gp = g[2];
      const double gsum = g0 + gm + gp, gdif = aa * (gp - gm);
      fh[0] = gsum;
      double shear_e = 0.0;
      {
        int kf = 3;
#pragma unroll
        for (int k = 0; k < NDV; k++) {
          if (k == dir) fh[1 + k] = fma(vhat[k], gsum, gdif);
          else {
            fh[1 + k] = fma(vhat[k], gsum, g[kf]);
            shear_e = fma(vhat[k], g[kf], shear_e);
            kf++;
          }
        }
      }
      fh[NV - 1] = fma(ek, g0, fma(H, gm + gp, fma(vn, gdif, shear_e)));
    }
    __syncwarp();
    } else {
    // ---------------- P2: both reconstructions of the centred stencil of cell j = 32m+1+l (position l+3)
    const bool rc_ok = line_ok && j >= -1 && j <= N;
    double fRv[NV], uRv[NV], sR[2] = { 0.0, 0.0 };
#pragma unroll
    for (int v = 0; v < NV; v++) { fRv[v] = 0.0; uRv[v] = 0.0; }
    if (!G3 && m == -1) {
      // start-up step: only the stencils centred on cells -1 and 0 exist (lanes 30, 31 in the regular mapping, which
      // would cost the warp a full step for two cells). Spread their 2 x (2 NV - 1) reconstructions over the lanes
      // instead: lane t < 2 NP handles (cell t / NP, pair t % NP), pairs = flux components 1..NV-1, then solution
      // components 0..NV-1. Left-biased values go to the exchange slots the regular P3 reads (31, 32), the
      // right-biased values of cell 0 to slot 1 (free during this step), from where lane 31 picks them up.
      constexpr int NP = 2 * NV - 1;
      if (line_ok && l < 2 * NP) {
        const int ce = l / NP, pr = l - ce * NP;
        const bool isU = pr >= NV - 1;
        const int v = isU ? pr - (NV - 1) : pr + 1;
        const int fx = isU ? LY::rU + v : LY::rF + v;
        const int fy = (G3 && isU && v == NV - 1) ? LY::rV4 : fx;
        const int cp = rbase + 33 + ce;
        double X[5], Y[5], L, R, zl, zr;
#pragma unroll
        for (int k = 0; k < 5; k++) { X[k] = rec[fx * NREC + cp + (k - 2)]; Y[k] = rec[fy * NREC + cp + (k - 2)]; }
        recon_pair<WT, false, false>(X, Y, Y, a.ph.eps, L, R, zl, zr);
        const int xf = isU ? LY::xU + v : LY::xF + v;
        exL[xf * NEX + xbase + 31 + ce] = L;
        if (ce == 1) exL[xf * NEX + xbase + 1] = R;
      }
      __syncwarp();
      if (line_ok && l == 31) {
#pragma unroll
        for (int v = 0; v < NV; v++) uRv[v] = exL[(LY::xU + v) * NEX + xbase + 1];
#pragma unroll
        for (int v = 1; v < NV; v++) fRv[v] = exL[(LY::xF + v) * NEX + xbase + 1];
        fRv[0] = (dir == 0) ? uRv[1] : ((dir == 1 || NDV < 3) ? uRv[2] : uRv[NDV]);
      }
    } else if (rc_ok) {
      double Zg[5] = { 0, 0, 0, 0, 0 };
      if (G3) {
#pragma unroll
        for (int k = 0; k < 5; k++) Zg[k] = rec[(LY::rGF + 1) * NREC + cc + (k - 2)];
      }
      const int ex = xbase + l + 1;
#pragma unroll
      for (int v = 0; v < NV; v++) {
        double X[5], Y[5], L, zl, zr;
        if (!(SKIPF0 && v == 0)) {
#pragma unroll
          for (int k = 0; k < 5; k++) X[k] = rec[(LY::rF + v) * NREC + cc + (k - 2)];
          const bool zsrc = G3 && a.with_source && (v == dir + 1 || v == NV - 1);
          if (G3 && zsrc) {
            recon_pair<WT, true, true>(X, X, Zg, a.ph.eps, L, fRv[v], zl, zr);
            const int si = (v == NV - 1) ? 1 : 0;
            sR[si] = zr;
            exL[(LY::xZ + si) * NEX + ex] = zl;
          } else {
            recon_pair<WT, false, true>(X, X, X, a.ph.eps, L, fRv[v], zl, zr);
          }
          exL[(LY::xF + v) * NEX + ex] = L;
        }
        // solution: weights from raw u, applied to the modified solution (Q4)
#pragma unroll
        for (int k = 0; k < 5; k++) X[k] = rec[(LY::rU + v) * NREC + cc + (k - 2)];
        if (G3) {
#pragma unroll
          for (int k = 0; k < 5; k++) {
            if (v == NV - 1) Y[k] = rec[LY::rV4 * NREC + cc + (k - 2)];
            else Y[k] = X[k] * rec[LY::rGF * NREC + cc + (k - 2)];
          }
          recon_pair<WT, false, false>(X, Y, Y, a.ph.eps, L, uRv[v], zl, zr);
        } else {
          recon_pair<WT, false, true>(X, X, X, a.ph.eps, L, uRv[v], zl, zr);
        }
        exL[(LY::xU + v) * NEX + ex] = L;
      }
      if (SKIPF0) fRv[0] = (dir == 0) ? uRv[1] : ((dir == 1 || NDV < 3) ? uRv[2] : uRv[NDV]);       // f_0 = rho v_d is the conserved momentum component itself
    }
    __syncwarp();

    // ---------------- P3: interface j-1/2 (between cells j-1 and j): Rusanov flux
    const bool if_ok = line_ok && j >= 0 && j <= N;
    if (if_ok) {
      const int exl = xbase + l;              // left-biased values of cell j-1 (slot 0: carry)
      const int cL = cc - 1, cR = cc;
      const double tL = rec[LY::rSR * NREC + cL], tR = rec[LY::rSR * NREC + cR];
      const double rs = rcp_fast(tL + tR);
      double vsq = 0.0, vn = 0.0, vhat[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
      for (int k = 0; k < NDV; k++) {
        const double v = (tL * rec[(LY::rVEL + k) * NREC + cL] + tR * rec[(LY::rVEL + k) * NREC + cR]) * rs;
        vsq += v * v;
        vhat[k] = v;
        if (k == dir) vn = v;
      }
      const double H = (tL * rec[LY::rH * NREC + cL] + tR * rec[LY::rH * NREC + cR]) * rs;
      if (a.upw == 1) {
        // Roe (NavierStokes3DUpwind.c:40-125): 2 x the interface flux = (fL + fR) - R |Lambda| L (uR - uL)
        double du[NV], diss[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) du[v] = uRv[v] - exL[(LY::xU + v) * NEX + exl];
        roe_dissipation<NV>(du, vhat, vn, vsq, H, gamma, dir, diss);
#pragma unroll
        for (int v = 0; v < NV; v++) {
          const double fL = (SKIPF0 && v == 0) ? exL[(LY::xU + 1 + dir) * NEX + exl] : exL[(LY::xF + v) * NEX + exl];
          fh[v] = (fL + fRv[v]) - diss[v];
        }
      } else {
        const double cavg = sqrt_fast((gamma - 1.0) * (H - 0.5 * vsq));
        const double aavg = cavg + fabs(vn);
        double alpha = fmax(fmax(rec[LY::rA * NREC + cL], rec[LY::rA * NREC + cR]), aavg);
        if (G3) alpha *= fmax(rec[(LY::rGF + 1) * NREC + cL], rec[(LY::rGF + 1) * NREC + cR]);
#pragma unroll
        for (int v = 0; v < NV; v++) {
          const double uL = exL[(LY::xU + v) * NEX + exl];
          const double fL = (SKIPF0 && v == 0) ? exL[(LY::xU + 1 + dir) * NEX + exl] : exL[(LY::xF + v) * NEX + exl];
          fh[v] = (fL + fRv[v]) - alpha * (uRv[v] - uL);          // 2 x the interface flux; the factor 1/2 is in dxih
        }
      }
      if (G3 && a.with_source) {
        Sh[0] = 0.5 * (exL[(LY::xZ + 0) * NEX + exl] + sR[0]);
        Sh[1] = 0.5 * (exL[(LY::xZ + 1) * NEX + exl] + sR[1]);
      }
    }
    __syncwarp();
    // carry of the left-biased values (slot 32 -> slot 0)
    if (l < LY::NFL) exL[l * NEX + xbase] = exL[l * NEX + xbase + TL];
    __syncwarp();

    }

    // ---------------- P4: cell jo = j-1 = 32m+l (position l+2): interfaces j-1/2 (own) and j-3/2 (the lane below; lane 0:
    // the carry lane 31 left in the previous step). The interface fluxes travel by warp shuffle, not through shared memory.
    if (m >= 0) {
      const bool out_ok = line_ok && jo >= 0 && jo < N;
      double res[NV];
      double Sl[2] = { 0.0, 0.0 };
#pragma unroll
      for (int v = 0; v < NV; v++) {
        double fl = __shfl_up_sync(0xffffffffu, fh[v], 1);
        if (l == 0) fl = car[v * TW + w];
        const double t = dxih * (fh[v] - fl);
        res[v] = out_ok ? ((a.mode < 2) ? -t : t) : 0.0;
      }
      if (G3 && a.with_source) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
          Sl[k] = __shfl_up_sync(0xffffffffu, Sh[k], 1);
          if (l == 0) Sl[k] = car[(NV + k) * TW + w];
        }
      }
      if (out_ok) {
        const int co = cc - 1;
        if (V3) {
          // par_v += dxinv * (FV[j-2] - 8 FV[j-1] + 8 FV[j+1] - FV[j+2]) / 12   (components 1..4)
          const double dxi12 = dxi * (1.0 / 12.0);
#pragma unroll
          for (int v = 0; v < 4; v++) {
            const double* f = rec + (LY::rFV + v) * NREC + co;
            res[1 + v] = fma(dxi12, fma(8.0, f[1] - f[-1], f[-2] - f[2]), res[1 + v]);
          }
        }
        if (G3 && a.with_source) {
          // NavierStokes3DSource.c:80-100 (the source accumulates into the same array as the flux divergence)
          const double rho = rec[LY::rU * NREC + co];
          const double vd = rec[(LY::rVEL + dir) * NREC + co];
          const double f = rec[LY::rGF * NREC + co];
          const double tmm = rho * a.ph.RT, te = rho * a.ph.RT * vd;
          const double sm = (tmm * f) * (Sh[0] - Sl[0]) * dxi;
          const double se = (te * f) * (Sh[1] - Sl[1]) * dxi;
#pragma unroll
          for (int v = 1; v < NV; v++) res[v] += ((v == dir + 1) ? sm : 0.0) + ((v == NV - 1) ? se : 0.0);
        }
      }
      if (XS) {
        // first direction: overwrite; the warp's 32 cells are contiguous
        if (out_ok) {
#pragma unroll
          for (int v = 0; v < NV; v++) XS_STORE(a.out + v * npg + pline + jo, res[v]);
        }
      } else {
        if (RKF) {
          // the tiles hold the right-hand side of the earlier directions and u^n (and the stores of step m-1 have read
          // them): complete k in place, next stage solution in place beside it -- k_rk_combine's two roundings, no contraction
          mbar_wait(rin_bar, (unsigned)m & 1u);
#pragma unroll
          for (int v = 0; v < NV; v++) {
            const double kf = __dadd_rn(ost[v * TILE + sidx], res[v]);
            ost[v * TILE + sidx] = kf;
            const double un = ust[v * TILE + sidx];
            ust[v * TILE + sidx] = out_ok ? __dadd_rn(un, __dmul_rn(a.adt, kf)) : 0.0;
          }
        } else {
          if (m >= 1) mbar_wait(free_bar, (unsigned)(m - 1) & 1u);     // the reduce of step m-1 has read the tile
#pragma unroll
          for (int v = 0; v < NV; v++) ost[v * TILE + sidx] = res[v];
        }
        fence_proxy_async();
        __syncwarp();
        if (l == 0) {
          __threadfence_block();
          const unsigned old = atomicAdd(res_cnt, 1u);
          if ((old & (TW - 1)) == TW - 1) {
            __threadfence_block();
            fence_proxy_async();
            const int cd = TL * m + G.g;
            int c0 = tc0, c1t = tc1, c2t = tc2;
            if (dir == 1) c1t = cd; else c2t = cd;
            if (RKF) {
              tma_store4(&tm.out, ost, c0, c1t, c2t, 0);
              tma_store4(&tm.un, ust, c0, c1t, c2t, 0);
            } else tma_reduce_add4(&tm.out, ost, c0, c1t, c2t, 0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            pend = true;                     // the wait and what follows it: after P1 of the next step (above)
          }
        }
      }
    }
    __syncwarp();
    // carry of the interface fluxes for lane 0 of the next step (also from the start-up step, whose lane 31 owns interface -1/2)
    if (l == 31) {
#pragma unroll
      for (int v = 0; v < NV; v++) car[v * TW + w] = fh[v];
      if (G3 && a.with_source) { car[(NV + 0) * TW + w] = Sh[0]; car[(NV + 1) * TW + w] = Sh[1]; }
    }

    // ---------------- shift (warp-private): the last 5 records and the interface-flux carry move to the front
    // Lane l < 25 owns record r = l % 5 of the fields l / 5 + {0, 5, 10, ...}: one address per lane, the fields at
    // compile-time offsets, all loads before the stores (profiles/r02c: the i / 5, i % 5 loop over 5 NFREC elements was
    // 5.6 % of the warp-stall samples of every sweep)
    if (l < 25) {
      const int f0 = (l * 13) >> 6;                       // l / 5 for l < 32
      double* p = rec + f0 * NREC + rbase + (l - 5 * f0);
      constexpr int NG = (LY::NFREC + 4) / 5;
      double t[NG];
#pragma unroll
      for (int k = 0; k < NG; k++) if (5 * k + 5 <= LY::NFREC || f0 + 5 * k < LY::NFREC) t[k] = p[5 * k * NREC + TL];
#pragma unroll
      for (int k = 0; k < NG; k++) if (5 * k + 5 <= LY::NFREC || f0 + 5 * k < LY::NFREC) p[5 * k * NREC] = t[k];
    }
    __syncwarp();
  }
  if (!XS && pend) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // shared memory stays valid until the last tiles are read
}

} // namespace hpbf

// sweep_fused_inst.cuh -- launcher of the fused sweep for one weight type (one translation unit per
// weight type keeps the build parallel: sweep_fused_wt{0..4}.cu).
#pragma once
#include "sweep_fused.cuh"
#include "sweep_tma.cuh"

namespace hpbf {

template <int MODEL, int WT, bool MAPX, bool GRAV, bool VISC>
static bool launch_one(hpb_solver* h, const SweepArgs& a)
{
  constexpr size_t smem = SweepLayout<MODEL, GRAV, VISC>::smem_bytes;
  // the shared-memory opt-in is a per-device attribute: one bit per device ordinal (one process may hold solvers on
  // several GPUs)
  static unsigned long long configured = 0ull;
  auto kern = k_sweep<MODEL, WT, MAPX, GRAV, VISC>;
  const unsigned long long dbit = 1ull << (h->device & 63);
  if (!(configured & dbit)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured |= dbit;
  }
  dim3 grid((a.nlines + TW - 1) / TW, 1, 1);
  kern<<<grid, NT, smem, h->stream>>>(a);
  h->launches++;
  return true;
}

// TMA variant (sweep_tma.cuh): NavierStokes2D/3D, even padded row length, x-sweep overwriting / y-,z-sweeps
// accumulating (what hyperbolic_fused always asks for). Returns false when not applicable: the caller then
// launches k_sweep.
template <int MODEL, int WT, bool XS, bool GRAV, bool VISC, bool CHR = false, bool RKF = false>
static bool launch_one_tma(hpb_solver* h, const SweepArgs& a)
{
  using LY = TmaLayout<MODEL, GRAV, VISC, RKF>;
  constexpr size_t smem = LY::smem_bytes;
  if (smem > 227 * 1024) return false;
  const bool accumulate = (a.mode & 1) != 0;
  if (XS == accumulate) return false;
  TmaMaps tm;
  if (RKF != (a.unext != nullptr)) return false;
  if (!tma_maps_for(h, a, XS, GRAV, VISC, &tm)) return false;
  static unsigned long long configured = 0ull;        // per device ordinal, as above
  auto kern = k_sweep_tma<MODEL, WT, XS, GRAV, VISC, CHR, RKF>;
  const unsigned long long dbit = 1ull << (h->device & 63);
  if (!(configured & dbit)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured |= dbit;
  }
  const Geom& G = a.G;
  long long nblk;
  if (XS) nblk = (long long)((G.N[1] + TW - 1) / TW) * G.N[2];
  else    nblk = (long long)((G.N[0] + 1 + TW - 1) / TW) * (a.dir == 1 ? G.N[2] : G.N[1]);
  if (nblk <= 0 || nblk > 2147483647LL) return false;
  kern<<<dim3((unsigned)nblk, 1, 1), NT, smem, h->stream>>>(a, tm);
  h->launches++;
  h->tma_launches++;
  return true;
}

template <int MODEL, int WT, bool MAPX, bool GRAV, bool VISC>
static bool launch_pick(hpb_solver* h, const SweepArgs& a)
{
  if (MODEL != HPB_MODEL_LINEAR_ADR && h->phys.interp_char) {
    // characteristic-wise reconstruction: the TMA-fed kernel only (fused_available has checked that it applies), no gravity
    constexpr int M2 = (MODEL == HPB_MODEL_LINEAR_ADR) ? HPB_MODEL_NS3D : MODEL;
    if (GRAV || a.unext != nullptr) return false;
    return launch_one_tma<M2, WT, MAPX, false, VISC, true>(h, a);
  }
  if (MODEL != HPB_MODEL_LINEAR_ADR && h->cfg.use_fused != 2) {
    constexpr int M2 = (MODEL == HPB_MODEL_LINEAR_ADR) ? HPB_MODEL_NS3D : MODEL;   // never instantiated for LinearADR
    // the last direction's sweep that also forms the next RK stage solution (accumulating sweeps only; the caller has
    // checked that the TMA-fed kernel applies -- there is no other kernel that writes a.unext)
    if (!MAPX && a.unext != nullptr) return launch_one_tma<M2, WT, false, GRAV, VISC, false, true>(h, a);
    if (launch_one_tma<M2, WT, MAPX, GRAV, VISC>(h, a)) return true;
  }
  if (a.unext != nullptr) return false;
  return launch_one<MODEL, WT, MAPX, GRAV, VISC>(h, a);
}

template <int MODEL, int WT, bool MAPX>
static bool launch_map(hpb_solver* h, const SweepArgs& a, bool grav, bool visc)
{
  constexpr bool NS3 = (MODEL == HPB_MODEL_NS3D);
  if (NS3 && grav && visc) return launch_pick<MODEL, WT, MAPX, NS3, NS3>(h, a);
  if (NS3 && grav)         return launch_pick<MODEL, WT, MAPX, NS3, false>(h, a);
  if (NS3 && visc)         return launch_pick<MODEL, WT, MAPX, false, NS3>(h, a);
  return launch_pick<MODEL, WT, MAPX, false, false>(h, a);
}

template <int MODEL, int WT>
static bool launch_model(hpb_solver* h, const SweepArgs& a, bool grav, bool visc)
{
  return (a.dir == 0) ? launch_map<MODEL, WT, true>(h, a, grav, visc) : launch_map<MODEL, WT, false>(h, a, grav, visc);
}

template <int WT>
bool launch_sweep(hpb_solver* h, const SweepArgs& a)
{
  const bool grav = h->phys.has_grav != 0;
  switch (h->cfg.model) {
    case HPB_MODEL_LINEAR_ADR: return launch_model<HPB_MODEL_LINEAR_ADR, WT>(h, a, false, false);
    case HPB_MODEL_NS2D:       return launch_model<HPB_MODEL_NS2D, WT>(h, a, false, false);
    case HPB_MODEL_NS3D:       return launch_model<HPB_MODEL_NS3D, WT>(h, a, grav, a.qd != nullptr);
    default: return false;
  }
}

} // namespace hpbf

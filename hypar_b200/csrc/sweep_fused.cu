// sweep_fused.cu -- dispatch of the fused directional sweep kernels (sweep_fused.cuh).
// Covered: component-wise WENO5 (all weight types, no_limiting) for LinearADR, NavierStokes2D and
// NavierStokes3D with Rusanov or Roe upwinding (with or without gravity), and characteristic-wise WENO5 for
// NavierStokes2D / 3D without gravity. Everything else (rf-char / llf-char upwinding, Euler1D, characteristic
// reconstruction with gravity) is served by the generic per-interface kernels.
#include "sweep_fused.cuh"
#include "sweep_tma.cuh"
#include <cstdlib>
#include <cstring>

namespace hpbf {

// ---- TMA descriptors (cuTensorMapEncodeTiled through the runtime's driver entry point: no libcuda link)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn()
{
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* off = getenv("HPB_NO_TMA");
    if (!(off && off[0] == '1')) {
      void* p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
          q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
      else cudaGetLastError();
    }
  }
  return fn;
}

// ghost-padded SoA array of nf fields as a 4-D tensor (x, y, z, field); kind 0: x-sweep box 32 x 8 lines(y),
// kind 1 / 2: y- / z-sweep box 8 lines(x) x 32 cells with the 64-byte swizzle. boxf = fields per box.
static bool get_map(hpb_solver* h, const void* ptr, int nf, int boxf, int kind, CUtensorMap* out)
{
  for (const auto& e : h->tma_cache)
    if (e.ptr == ptr && e.nf == nf * 16 + boxf && e.kind == kind) { memcpy(out, e.map.b, sizeof(CUtensorMap)); return true; }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  const Geom& G = h->geo;
  if ((G.P[0] & 1) || (reinterpret_cast<uintptr_t>(ptr) & 15)) return false;
  cuuint64_t dims[4] = { (cuuint64_t)G.P[0], (cuuint64_t)G.P[1], (cuuint64_t)G.P[2], (cuuint64_t)nf };
  cuuint64_t strides[3] = { (cuuint64_t)G.P[0] * 8, (cuuint64_t)G.P[0] * G.P[1] * 8, (cuuint64_t)G.npg * 8 };
  cuuint32_t box[4] = { 1, 1, 1, (cuuint32_t)boxf };
  if (kind == 0) { box[0] = TL; box[1] = TW; }
  else { box[0] = TW; box[kind] = TL; }
  cuuint32_t es[4] = { 1, 1, 1, 1 };
  CUtensorMap m;
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void*>(ptr), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, kind == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  hpb_solver::TmaEntry e;
  e.ptr = ptr; e.nf = nf * 16 + boxf; e.kind = kind;
  memcpy(e.map.b, &m, sizeof(CUtensorMap));
  if (h->tma_cache.size() > 256) h->tma_cache.clear();
  h->tma_cache.push_back(e);
  *out = m;
  return true;
}

bool tma_available() { return encode_fn() != nullptr; }

bool tma_maps_for(hpb_solver* h, const SweepArgs& a, bool xs, bool grav, bool visc, TmaMaps* tm)
{
  static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
  const Geom& G = a.G;
  if (G.ndims < 2 || G.g != HPB_G) return false;
  if (xs != (a.dir == 0)) return false;
  const int kind = a.dir;
  const int nv = G.nvars;
  memset(tm, 0, sizeof(*tm));
  if (!get_map(h, a.u, nv, nv, kind, &tm->u)) return false;
  if (xs) tm->out = tm->u;    // unused by the x-sweep (direct stores)
  else if (!get_map(h, a.out, nv, nv, kind, &tm->out)) return false;
  tm->qd = tm->u; tm->gf = tm->u; tm->gg = tm->u;
  if (visc && !get_map(h, a.qd, 12, 1, kind, &tm->qd)) return false;
  if (grav && (!get_map(h, a.gf, 1, 1, kind, &tm->gf) || !get_map(h, a.gg, 1, 1, kind, &tm->gg))) return false;
  tm->un = tm->u; tm->u0 = tm->u;
  if (a.unext != nullptr && (xs || !get_map(h, a.unext, nv, nv, kind, &tm->un) || !get_map(h, a.ubase, nv, nv, kind, &tm->u0))) return false;
  return true;
}

} // namespace hpbf

namespace hpbk {

bool fused_available(const hpb_solver* h)
{
  const hpb_config& c = h->cfg;
  if (!c.use_fused) return false;
  if (c.hyp_scheme != HPB_SCHEME_WENO5) return false;       // compact / linear schemes: reference-exact kernels only
  if (c.model == HPB_MODEL_BURGERS) return false;
  if (h->phys.advf != nullptr || c.advection_field != nullptr) return false;     // spatially varying advection: exact kernels
  if (c.model == HPB_MODEL_EULER1D) return false;
  if (h->phys.interp_char) {
    // characteristic-wise WENO5: NavierStokes2D / 3D without gravity, Roe or Rusanov, on the TMA-fed kernel only (even padded
    // row length, 2-D or 3-D, driver entry point available)
    if (c.model != HPB_MODEL_NS2D && c.model != HPB_MODEL_NS3D) return false;
    if (h->phys.has_grav || c.use_fused == 2) return false;
    if (c.upwind != HPB_UPWIND_RUSANOV && c.upwind != HPB_UPWIND_ROE) return false;
    if ((h->geo.P[0] & 1) || h->geo.ndims < 2 || h->geo.g != HPB_G || !hpbf::tma_available()) return false;
  }
  if ((c.model == HPB_MODEL_NS2D || c.model == HPB_MODEL_NS3D) && c.upwind != HPB_UPWIND_RUSANOV && c.upwind != HPB_UPWIND_ROE) return false;
  if (c.model == HPB_MODEL_LINEAR_ADR && c.nvars != 1) return false;
  if (c.model == HPB_MODEL_NS2D && h->phys.has_grav) return false;   // 2-D gravity source: reference-exact kernels only
  return true;
}

// Can the last direction's sweep of this configuration also write the next RK stage solution (sweep_tma.cuh, RKF)? The
// TMA-fed kernel must be the one that runs (2-D / 3-D fluid model, component-wise, even padded row length, driver entry
// point) and the sweeps must be the last contribution to the right-hand side (the caller checks the rest: capi.cu).
bool stage_fusion_available(const hpb_solver* h)
{
  const hpb_config& c = h->cfg;
  if (!fused_available(h) || c.use_fused == 2 || h->phys.interp_char) return false;
  if (c.model != HPB_MODEL_NS2D && c.model != HPB_MODEL_NS3D) return false;
  if ((h->geo.P[0] & 1) || h->geo.ndims < 2 || h->geo.g != HPB_G || !hpbf::tma_available()) return false;
  return true;
}

bool hyperbolic_fused(hpb_solver* h, const double* u, double* out, bool negate, bool with_source, double* src,
                      const double* qd, int only_dir, double* unext, double adt, const double* ubase)
{
  if (with_source && src != nullptr && src != out) return false;     // the fused kernels accumulate the source into `out`
  const Geom& G = h->geo;
  if (unext != nullptr && (!negate || !ubase || unext == u || unext == out || unext == ubase || !stage_fusion_available(h))) return false;
  const int wt = h->phys.no_limiting ? hpbf::WT_NOLIM : h->phys.weno;
  for (int d = 0; d < G.ndims; d++) {
    if (only_dir >= 0 && d != only_dir) continue;
    hpbf::SweepArgs a;
    a.G = G; a.ph = h->phys; a.u = u;
    a.gf = h->d_gravf; a.gg = h->d_gravg; a.dxinv = h->d_dxinv;
    a.out = out; a.src = src; a.dir = d;
    a.mode = negate ? (d == 0 ? 0 : 1) : (d == 0 ? 2 : 3);
    a.with_source = (with_source && h->phys.has_grav && h->phys.grav[d] != 0.0 && src != nullptr) ? 1 : 0;
    a.qd = qd;
    a.upw = (h->cfg.upwind == HPB_UPWIND_ROE) ? 1 : 0;
    a.unext = (d == G.ndims - 1) ? unext : nullptr; a.adt = adt; a.ubase = ubase;
    {
      static const int tab[3][8] = { { 0, 1, 2, 3, 4, 5, 8, 10 }, { 4, 5, 6, 7, 0, 1, 9, 10 }, { 8, 9, 10, 11, 0, 2, 5, 6 } };
      for (int k = 0; k < 8; k++) a.qidx[k] = tab[d][k];
    }
    a.nlines = (d == 0 ? G.N[1] * G.N[2] : d == 1 ? G.N[0] * G.N[2] : G.N[0] * G.N[1]);
    ProfScope ps(h, a.unext ? HPB_PROF_SWEEP_FUSED : HPB_PROF_SWEEP_X + d);
    bool ok;
    switch (wt) {
      case HPB_WENO_JS: ok = hpbf::launch_sweep<HPB_WENO_JS>(h, a); break;
      case HPB_WENO_M:  ok = hpbf::launch_sweep<HPB_WENO_M>(h, a); break;
      case HPB_WENO_Z:  ok = hpbf::launch_sweep<HPB_WENO_Z>(h, a); break;
      case HPB_WENO_YC: ok = hpbf::launch_sweep<HPB_WENO_YC>(h, a); break;
      default:          ok = hpbf::launch_sweep<hpbf::WT_NOLIM>(h, a); break;
    }
    if (!ok) {
      // only possible before the first direction has written anything (attribute / grid limits are
      // direction-independent for a given configuration)
      if (d == 0 || only_dir >= 0) return false;
      hpb_fail(HPB_ERR_CUDA, "fused sweep launch failed in direction %d", d);
      return true;
    }
  }
  return true;
}

} // namespace hpbk

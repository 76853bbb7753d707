// sweep_fused.cu -- dispatch of the fused directional sweep kernels (sweep_fused.cuh).
// Covered: component-wise WENO5 (all weight types, no_limiting) for LinearADR, NavierStokes2D and
// NavierStokes3D with Rusanov upwinding (with or without gravity). Everything else (characteristic
// reconstruction, Roe upwinding, Euler1D) is served by the generic per-interface kernels.
#include "sweep_fused.cuh"

namespace hpbk {

bool fused_available(const hpb_solver* h)
{
  const hpb_config& c = h->cfg;
  if (!c.use_fused) return false;
  if (h->phys.interp_char) return false;
  if (c.model == HPB_MODEL_EULER1D) return false;
  if ((c.model == HPB_MODEL_NS2D || c.model == HPB_MODEL_NS3D) && c.upwind != HPB_UPWIND_RUSANOV) return false;
  if (c.model == HPB_MODEL_LINEAR_ADR && c.nvars != 1) return false;
  return true;
}

bool hyperbolic_fused(hpb_solver* h, const double* u, double* out, bool negate, bool with_source, double* src,
                      const double* qd)
{
  if (with_source && src != nullptr && src != out) return false;     // the fused kernels accumulate the source into `out`
  const Geom& G = h->geo;
  const int wt = h->phys.no_limiting ? hpbf::WT_NOLIM : h->phys.weno;
  for (int d = 0; d < G.ndims; d++) {
    hpbf::SweepArgs a;
    a.G = G; a.ph = h->phys; a.u = u;
    a.gf = h->d_gravf; a.gg = h->d_gravg; a.dxinv = h->d_dxinv;
    a.out = out; a.src = src; a.dir = d;
    a.mode = negate ? (d == 0 ? 0 : 1) : (d == 0 ? 2 : 3);
    a.with_source = (with_source && h->phys.has_grav && h->phys.grav[d] != 0.0 && src != nullptr) ? 1 : 0;
    a.qd = qd;
    {
      static const int tab[3][8] = { { 0, 1, 2, 3, 4, 5, 8, 10 }, { 4, 5, 6, 7, 0, 1, 9, 10 }, { 8, 9, 10, 11, 0, 2, 5, 6 } };
      for (int k = 0; k < 8; k++) a.qidx[k] = tab[d][k];
    }
    a.nlines = (d == 0 ? G.N[1] * G.N[2] : d == 1 ? G.N[0] * G.N[2] : G.N[0] * G.N[1]);
    ProfScope ps(h, HPB_PROF_SWEEP_X + d);
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
      if (d == 0) return false;
      hpb_fail(HPB_ERR_CUDA, "fused sweep launch failed in direction %d", d);
      return true;
    }
  }
  return true;
}

} // namespace hpbk

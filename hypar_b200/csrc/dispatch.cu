// dispatch.cu -- chooses the hyperbolic implementation: the fused per-cell sweep kernels where one
// exists for the configuration (sweep_fused.cu), otherwise the generic per-interface kernels.
#include "hpb_internal.h"

namespace hpbk {
void hyperbolic(hpb_solver* h, const double* u, double* out, bool negate, bool with_source, double* src)
{
  if (h->cfg.use_fused && fused_available(h) && hyperbolic_fused(h, u, out, negate, with_source, src, nullptr)) return;
  hyperbolic_generic(h, u, out, negate, with_source, src);
}
}

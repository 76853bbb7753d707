// sweep_fused.cu -- fused per-cell directional sweep kernels (placeholder until implemented)
#include "hpb_internal.h"
namespace hpbk {
bool hyperbolic_fused(hpb_solver*, const double*, double*, bool, bool, double*) { return false; }
}

// fused sweep kernels, weight type 4 (0 JS, 1 mapped, 2 Z, 3 YC, 4 no_limiting)
#include "sweep_fused_inst.cuh"
namespace hpbf { template bool launch_sweep<4>(hpb_solver*, const SweepArgs&); }

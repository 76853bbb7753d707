"""Per-kernel cost of the staged multi-GPU step on ONE GPU (no NCCL): rank 0 of a 2x2x2 decomposition of 1024^3
(512^3 local) runs its pack / unpack / BC / derivative / sweep kernels on whatever the halo buffers hold.
Run under `ncu --metrics gpu__time_duration.sum --csv` and summarise with tools/launch_summary.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from hypar_b200.solver import Solver, FIELD_U, FIELD_QDERIVX, FIELD_QDERIVY

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
size, iproc = bench.weak_grid(n, 8)
s, b, ph, w, x = bench.c4_inputs(size, iproc)
sv = Solver(s, b, ph, w, x, rank=0, device=0)
g, nloc = sv.ghosts, sv.dim_local
x_loc = [x[d][sv.is_global[d]:sv.is_global[d] + nloc[d]] for d in range(3)]
u_host_t = torch.zeros(sv.npoints_local_wghosts * 5, dtype=torch.float64).pin_memory()
fld = bench.synth_field_torch(x_loc, torch.device("cuda", 0))
u_host_t.view(nloc[2] + 2 * g, nloc[1] + 2 * g, nloc[0] + 2 * g, 5)[g:-g, g:-g, g:-g, :].copy_(fld)
# ghost cells: fill with the nearest interior value so that the kernels see physical data without an exchange
v = u_host_t.view(nloc[2] + 2 * g, nloc[1] + 2 * g, nloc[0] + 2 * g, 5)
v[:g] = v[g:g + 1]; v[-g:] = v[-g - 1:-g]
v[:, :g] = v[:, g:g + 1]; v[:, -g:] = v[:, -g - 1:-g]
v[:, :, :g] = v[:, :, g:g + 1]; v[:, :, -g:] = v[:, :, -g - 1:-g]
del fld
sv.set_solution(u_host_t.numpy())
L, h = sv.L, sv.h
for rep in range(2):
    sv._ck(L.hpb_stage_begin(h, 0))
    sv._ck(L.hpb_stage_halo_done(h, FIELD_U))
    sv._ck(L.hpb_stage_rhs_a(h, 0))
    sv._ck(L.hpb_stage_halo_done(h, FIELD_QDERIVX))
    sv._ck(L.hpb_stage_halo_done(h, FIELD_QDERIVY))
    sv._ck(L.hpb_stage_rhs_b(h, 0))
    sv.synchronize()
print("done", sv.kernel_launches)

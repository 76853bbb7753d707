"""Pipelined host-array stepping (hpb_pipe_upload / hpb_TimeIntegrateAsync / hpb_pipe_download / hpb_pipe_wait): a
sequence of independent fields through one solver, the copies of neighbouring fields overlapping the steps. Every
field's result must be the blocking hpb_TimeIntegrate result of the same input, bit for bit, whatever the overlap."""
import numpy as np
import pytest

from hypar_b200 import cases
from hypar_b200.solver import Solver
from oracle import hpo

pytestmark = pytest.mark.gpu

CASES = [cases.linear_advection_sine(256, "mapped"),
         cases.euler1d_sod(101, "js"),
         cases.ns2d_vortex((64, 48), "yc"),
         cases.ns3d_turbulence((32, 24, 20), "mapped"),
         cases.ns3d_rising_bubble((16, 20, 12), "yc"),
         cases.ns2d_vortex((40, 28), "z", scheme="crweno5")]


def _pinned(n):
    import torch
    t = torch.zeros(n, dtype=torch.float64).pin_memory()
    return t, t.numpy()


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
@pytest.mark.parametrize("nsteps", [1, 3])
def test_pipelined_fields_equal_blocking_calls(need_gpu, case, nsteps):
    S = hpo.Setup(case)
    u0 = S.local_u0()
    nfields = 5
    rng = np.random.RandomState(11)
    # independent fields: the case's initial solution, smoothly perturbed
    fields = [u0 * (1.0 + 1e-3 * k * np.cos(0.37 * k + np.arange(u0.size) * 1e-3)) for k in range(nfields)]
    sv = Solver.from_case(case)
    ref = []
    for f in fields:
        u = f.copy()
        sv.TimeIntegrate(u, nsteps, 0.0)
        ref.append(sv.interior(u).copy())
    keep_in, keep_out, ins, outs = [], [], [], []
    for f in fields:
        ti, a = _pinned(f.size)
        to, b = _pinned(f.size)
        a[:] = f
        keep_in.append(ti); keep_out.append(to); ins.append(a); outs.append(b)
    l0 = sv.kernel_launches
    for a, b in zip(ins, outs):
        sv.TimeIntegrateAsync(a, b, nsteps, 0.0)          # enqueue only
    sv.pipe_wait()
    assert sv.kernel_launches > l0
    for k in range(nfields):
        got = sv.interior(outs[k])
        assert np.isfinite(got).all()
        assert np.array_equal(got, ref[k]), f"field {k}: max abs diff {np.abs(got - ref[k]).max():.3e}"
    # the pieces, with the steps issued by the caller in between, and a blocking call afterwards on the same solver
    sv.pipe_upload(ins[1], 0.0)
    sv.TimeSteps(nsteps)
    sv.pipe_download(outs[0])
    sv.pipe_join()
    sv.pipe_wait()
    assert np.array_equal(sv.interior(outs[0]), ref[1])
    u = fields[2].copy()
    sv.TimeIntegrate(u, nsteps, 0.0)
    assert np.array_equal(sv.interior(u), ref[2])
    sv.close()

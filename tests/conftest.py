"""pytest configuration: markers, import paths, shared helpers.

  -m "not gpu" : oracle vs golden vectors / vs the compiled reference, host logic, C-ABI symbols
                 (runs on a CPU-only box in a few minutes)
  -m gpu       : parity of the CUDA path against the oracle through the C ABI (needs a B200)
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

# parity tolerance of BASELINE.json's north_star: relative L2 / Linf <= 1e-12 per RHS evaluation (FP64)
RHS_TOL = 1e-12


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def assert_exact(got, ref, what, libm_ulp=False):
    """The exact path is bit-identical to the reference wherever only +, -, *, /, sqrt are involved (IEEE, no FMA).
    The one exception is the viscosity law mu = exp(0.76 log T) (NavierStokes3DParabolicFunction.c:174): CUDA's exp / log
    and glibc's differ by one ulp on a few percent of the arguments (measured: 85 of 2016 temperatures in [1, 1.08]),
    which moves the parabolic term by ~1e-16 relative. `libm_ulp` = True allows exactly that much (1e-14 of max|ref|)
    for viscous cases whose temperatures hit such arguments; it is never set for inviscid cases."""
    got, ref = np.asarray(got), np.asarray(ref)
    if np.array_equal(got, ref):
        return
    d = float(np.abs(got - ref).max())
    assert libm_ulp and d <= 1e-14 * float(np.abs(ref).max()), f"{what}: not bit-identical, max abs diff {d:.3e}"


def rel_linf(a, b):
    s = np.abs(b).max()
    return float(np.abs(a - b).max() / (s if s > 0 else 1.0))


def rel_l2(a, b):
    s = np.sqrt((b * b).sum())
    d = a - b
    return float(np.sqrt((d * d).sum()) / (s if s > 0 else 1.0))


def assert_close(a, b, tol=RHS_TOL, what=""):
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert np.isfinite(a).all(), f"{what}: non-finite values in result"
    li, l2 = rel_linf(a, b), rel_l2(a, b)
    assert li <= tol and l2 <= tol, f"{what}: rel Linf {li:.3e}, rel L2 {l2:.3e} > {tol:.1e}"
    return li, l2


def gpu_available():
    try:
        from hypar_b200 import _lib
        return _lib.load().hpb_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def need_gpu():
    from hypar_b200 import _lib
    L = _lib.load()          # raises if the CUDA library has not been built: no silent fallback
    if L.hpb_device_count() <= 0:
        pytest.fail("test marked gpu but no CUDA device is visible")
    return L

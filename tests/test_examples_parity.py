"""The reference's OWN example directories as parity cases (tools/examples_parity.py; the full list of 72 bit-identical
directories is profiles/examples_parity.txt): initial solution from the directory's own aux/init.c, the unmodified reference
(oracle/_ref/hypar_ref_mpi1, one thread) reads the directory with its own readers; the package's readers feed the oracle.
u after the boundary conditions, hyp, par, source, rhs: bit-identical. Runs where /root/reference exists (this container)."""
import os

import pytest

EXAMPLES = "/root/reference/Examples"
DIRS = [
    "1D/Euler1D/SodShockTube",                              # characteristic WENO5 + Roe, extrapolate
    "1D/Euler1D/SodShockTubeWithGravity",                   # Euler1D gravity source, slip walls
    "1D/Euler1D/ShuOsherProblem",
    "1D/LinearAdvection/SineWave_NonConstantAdvection",     # advection.inp (ascii)
    "1D/LinearAdvection/TestSponge",                        # sponge zone
    "2D/Burgers/SineWave",
    "2D/LinearAdvection/GaussianPulse",
    "2D/NavierStokes2D/1DSodShockTubeWithGravity/X",        # 3 cells across; one wall velocity per zone in boundary.inp
    "2D/NavierStokes2D/HydrostaticBalance_3_2D",            # HB 3 with N_bv
    "2D/NavierStokes2D/LidDrivenCavity",                    # viscous, moving no-slip wall
    "2D/NavierStokes2D/RadialExpansionWave",                # mapped characteristic WENO5 + rf-char (the reference's OpenMP race)
    "2D/NavierStokes2D/RisingThermalBubble",
    "3D/NavierStokes3D/2D_RiemannCase4/XZ",
    "3D/NavierStokes3D/DensitySineWave",
    "LaSDI/1d_burgers_sinewave/test",
]


@pytest.mark.parametrize("rel", DIRS)
def test_reference_example_directory(rel):
    import examples_parity as ep
    d = os.path.join(EXAMPLES, rel)
    if not os.path.isdir(d) or not os.access(ep.EXE, os.X_OK):
        pytest.skip("needs /root/reference and oracle/_ref (authoring container)")
    assert ep.classify(d) is None, "the B200 path refuses this directory"
    r = ep.check(d)
    assert r.startswith("BIT-IDENTICAL"), r

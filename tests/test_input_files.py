"""HyPar's input files read the way the reference scans them (host side, CPU).

  * initial.inp in its text flavour (``ip_file_type ascii``, ReadArray.c:173-217: grid, then the field one VARIABLE after the
    other) -- most of the reference's Examples use it. Pinned live: the reference reads the file our writer produced and
    steps from it to the same bits as the oracle from the arrays; our reader gives the arrays back.
  * boundary.inp zone values are unchecked fscanf("%lf") reads (InitializeBoundaries.c:108-200): a missing value stays 0 and
    consumes nothing. The reference's own Examples rely on it (one wall velocity in a 2-D run, no outflow pressure, blank lines).
    Pinned live on such a file.
  * solver.inp: ``input_mode parallel N`` / ``output_mode parallel N`` take a second token (ReadInputs.c:358-366).
"""
import os

import numpy as np
import pytest

from hypar_b200 import cases, hypario
from hypar_b200.solver import Solver
from oracle import hpo


def _ascii(case):
    case.solver["ip_file_type"] = "ascii"
    return case


ASCII_CASES = [
    _ascii(cases.linear_advection_sine(64, "js")),
    _ascii(cases.euler1d_sod(101, "mapped")),
    _ascii(cases.ns2d_vortex((20, 16), "z")),
    _ascii(cases.ns3d_density_wave((8, 10, 12), "js")),
    _ascii(cases.linear_advection_varying((16, 12), "js")),          # advection.inp in the same flavour
]


@pytest.mark.parametrize("case", ASCII_CASES, ids=[c.name for c in ASCII_CASES])
def test_ascii_initial_round_trip(case, tmp_path):
    d = str(tmp_path / "run")
    case.write(d)
    with open(os.path.join(d, "initial.inp"), "rb") as f:
        assert b"\x00" not in f.read(4096)                     # text
    x, u = hypario.read_initial(os.path.join(d, "initial.inp"), case.solver["size"], case.nvars, "ascii")
    assert np.array_equal(u, case.u0) and all(np.array_equal(a, b) for a, b in zip(x, case.x))
    sv = Solver.from_directory(d)
    assert np.array_equal(sv.u0_global, case.u0)
    gx, _ = sv.grid()
    S = hpo.Setup(case)
    assert np.array_equal(gx, S.x)
    if case.advection_field is not None:
        assert np.array_equal(sv.advection_field(), S.adv_field)
    sv.close()


@pytest.mark.parametrize("case", ASCII_CASES, ids=[c.name for c in ASCII_CASES])
def test_reference_reads_our_ascii_files(case):
    from refrun import ref_available, run_reference
    if not ref_available("hypar_ref"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    o = run_reference(case, "steps", [2])
    assert "ASCII file initial.inp" in o["stdout"]
    S = hpo.Setup(case, mpi_semantics=False)
    O = hpo.Oracle(S)
    u = S.local_u0()
    for _ in range(2):
        O.time_step(u, float(case.solver["dt"]), hpo.rk_type_of(case))
    assert np.array_equal(S.interior(u).ravel(), S.interior(o["ufinal"]["data"]).ravel())


SLOPPY = """4

subsonic-inflow  0  1  0.0 0.0  0.0 1.0
1.0 0.5 0.025

noslip-wall  1  1  0.0 1.0  0.0 0.0
0.0
slip-wall  1  -1  0.0 1.0  0.0 0.0

subsonic-outflow  0  -1  0.0 0.0  0.0 1.0
"""


def test_boundary_inp_missing_values_stay_zero(tmp_path):
    p = str(tmp_path / "boundary.inp")
    with open(p, "w") as f:
        f.write(SLOPPY)
    z = hypario.read_boundary_inp(p, 2, 4)
    assert [t["type"] for t in z] == ["subsonic-inflow", "noslip-wall", "slip-wall", "subsonic-outflow"]
    assert z[0]["density"] == 1.0 and z[0]["velocity"] == [0.5, 0.025]
    assert z[1]["wall_velocity"] == [0.0, 0.0] and z[2]["wall_velocity"] == [0.0, 0.0]
    assert z[3]["pressure"] == 0.0 and z[3]["dim"] == 0 and z[3]["face"] == -1
    assert z[2]["xmin"] == [0.0, 0.0] and z[2]["xmax"] == [1.0, 0.0]


class _WithBoundaryText:
    """A case whose boundary.inp is written verbatim (what run_reference needs: write())."""
    def __init__(self, case, text):
        self.case, self.text = case, text

    def write(self, d):
        self.case.write(d)
        with open(os.path.join(d, "boundary.inp"), "w") as f:
            f.write(self.text)


def test_reference_scans_a_sloppy_boundary_file_like_we_do(tmp_path):
    from refrun import ref_available, run_reference
    if not ref_available("hypar_ref"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    case = cases.ns_channel((28, 24), "js")
    text = SLOPPY.replace("subsonic-outflow  0  -1  0.0 0.0  0.0 1.0\n", "subsonic-outflow  0  -1  0.0 0.0  0.0 1.0\n0.7142857142857143\n")
    p = str(tmp_path / "boundary.inp")
    with open(p, "w") as f:
        f.write(text)
    case.boundary = hypario.read_boundary_inp(p, 2, 4)           # what OUR reader makes of it
    o = run_reference(_WithBoundaryText(case, text), "rhs")       # what the reference makes of it
    S = hpo.Setup(case)
    O = hpo.Oracle(S)
    u = S.local_u0()
    rhs = O.rhs(u)
    assert np.array_equal(u, o["u"]["data"]) and np.array_equal(rhs, o["rhs"]["data"])


def test_solver_inp_io_modes_with_rank_count(tmp_path):
    case = cases.ns3d_density_wave((8, 8, 8), "js")
    d = str(tmp_path / "run")
    case.write(d)
    txt = open(os.path.join(d, "solver.inp")).read()
    txt = "\n".join(("  input_mode          parallel 4" if ln.split()[:1] == ["input_mode"] else
                     "  output_mode         parallel 4" if ln.split()[:1] == ["output_mode"] else ln) for ln in txt.splitlines())
    assert "parallel 4" in txt
    with open(os.path.join(d, "solver.inp"), "w") as f:
        f.write(txt + "\n")
    s = hypario.read_solver_inp(os.path.join(d, "solver.inp"))
    assert s["input_mode"] == "parallel" and s["output_mode"] == "parallel" and int(s["n_io_ranks"]) == 4
    assert s["model"] == "navierstokes3d" and s["size"] == [8, 8, 8]


# ---------------------------------------------------------------------------------------------------------------- output
# the text and Tecplot solution files (WriteText.c, WriteTecplot2D.c, WriteTecplot3D.c) next to the binary one.
def _ref_main(case, d):
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "hypar_ref_main")
    if not os.access(exe, os.X_OK):
        pytest.skip("oracle/_ref/hypar_ref_main not built (needs /root/reference)")
    case.write(d)
    p = subprocess.run([exe], cwd=d, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=600)
    assert p.returncode == 0 and "Finished." in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


def _fmt_case(case, fmt, overwrite):
    case.solver.update({"n_iter": 2, "file_op_iter": 1, "screen_op_iter": 1, "op_file_format": fmt,
                        "op_overwrite": "yes" if overwrite else "no"})
    return case


FORMATS = [
    (cases.linear_advection_sine(48, "js"), "text", False),
    (cases.euler1d_sod(61, "js"), "text", True),
    (cases.ns2d_vortex((14, 12), "z"), "text", False),
    (cases.ns2d_vortex((14, 12), "z"), "tecplot2d", False),
    (cases.ns3d_density_wave((6, 8, 10), "js"), "tecplot3d", True),
    (cases.ns3d_density_wave((6, 8, 10), "js"), "text", True),
    (cases.ns3d_density_wave((6, 8, 10), "js"), "binary", False),
]


@pytest.mark.parametrize("case,fmt,overwrite", FORMATS, ids=[f"{c.name}-{f}-{'ow' if o else 'idx'}" for c, f, o in FORMATS])
def test_solution_files_byte_identical_to_the_reference_writers(case, fmt, overwrite, tmp_path):
    """The reference's own main writes op*.dat / op*.bin after 0, 1 and 2 steps; the oracle's solutions (bit-identical to the
    reference's) through OUR writers give the same BYTES, under the same file names."""
    import filecmp
    case = _fmt_case(case, fmt, overwrite)
    dref, dnew = str(tmp_path / "ref"), str(tmp_path / "new")
    _ref_main(case, dref)
    os.makedirs(dnew)
    S = hpo.Setup(case, mpi_semantics=True)
    O = hpo.Oracle(S)
    u = S.local_u0()
    names = []
    for it in range(3):
        if it:
            O.time_step(u, float(case.solver["dt"]), hpo.rk_type_of(case))
        nm = hypario.solution_file_name(fmt, overwrite, it)
        hypario.write_solution(os.path.join(dnew, nm), case.x, S.interior(u), fmt)
        names.append(nm)
        if not overwrite or it == 2:
            assert os.path.exists(os.path.join(dref, nm)), (nm, sorted(os.listdir(dref)))
            assert filecmp.cmp(os.path.join(dref, nm), os.path.join(dnew, nm), shallow=False), f"{nm}: bytes differ"
    assert len(set(names)) == (1 if overwrite else 3)
    if fmt != "binary":
        x, v = hypario.read_op_text(os.path.join(dnew, names[-1]), case.ndims, case.nvars)
        assert np.array_equal(v, S.interior(u)) and all(np.array_equal(a, b) for a, b in zip(x, case.x))


def test_tecplot_writers_are_dimension_specific(tmp_path):
    c = cases.linear_advection_sine(16, "js")
    with pytest.raises(ValueError, match="WriteTecplot2D"):
        hypario.write_solution(str(tmp_path / "a.dat"), c.x, c.u0, "tecplot2d")
    with pytest.raises(ValueError, match="writes no file"):
        hypario.write_solution(str(tmp_path / "a.dat"), c.x, c.u0, "none")

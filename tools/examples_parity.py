"""The oracle (and the package's input readers) against the reference ON THE REFERENCE'S OWN EXAMPLES.

    python tools/examples_parity.py [/root/reference/Examples] [substring ...]        (authoring container only)

For every run directory the B200 path accepts (tools/examples_coverage.py) that ships an initial-solution generator
(aux/init.c, aux/exact.c ...): copy the directory to a scratch place, compile the generator with gcc and run it there (it
reads the directory's own solver.inp and writes initial.inp in the flavour that file names), set iproc to 1, then

  reference   oracle/_ref/hypar_ref_mpi1 rhs   -- the unmodified reference reads the directory with its own readers and dumps
                                                 u after the boundary conditions, hyp, par, source, rhs of one
                                                 TimeRHSFunctionExplicit
  here        hypar_b200.cases.from_directory  -- the package's readers -> oracle.hpo on the same arrays

and compare bit for bit. What this pins: the readers (ascii / binary initial.inp, the sloppy boundary.inp of several
examples, physics.inp defaults) and the oracle on the grids, boundary zones and parameter combinations the reference's
authors chose -- not only on the cases this repository made up. TEST INFRASTRUCTURE ONLY; writes profiles/examples_parity.txt.
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from examples_coverage import classify  # noqa: E402
from hypar_b200 import cases, hypario  # noqa: E402
from oracle import hpo  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "hypar_ref_mpi1")
MAX_POINTS = 3_000_000          # keep one directory to seconds
MAX_POINTS_STEPS = 300_000      # two full time steps as well below this size


def _generator(d):
    for name in ("init.c", "init.C", "init.cpp", "exact.c", "exact.C", "exact.cpp"):
        p = os.path.join(d, "aux", name)
        if os.path.exists(p):
            return p
    return None


def _set_iproc_one(path, nd, nsims):
    words = open(path).read().split()
    i = words.index("iproc")
    for k in range(nd * nsims):
        words[i + 1 + k] = "1"
    out, j = ["begin"], 1
    while j < len(words) and words[j] != "end":
        key = words[j]
        n = nd * nsims if key in ("size", "iproc", "size_exact") else 2 if (key in ("input_mode", "output_mode") and words[j + 1] != "serial") else 1
        out.append("  " + key + " " + " ".join(words[j + 1:j + 1 + n]))
        j += 1 + n
    out.append("end")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


def check(d):
    gen = _generator(d)
    if gen is None:
        return "skip: no initial-solution generator in aux/"
    nsims = hypario.read_simulation_inp(os.path.join(d, "simulation.inp"))
    if nsims > 1:
        return "skip: ensemble (tests/test_ensemble.py pins that driver)"
    s = hypario.read_solver_inp(os.path.join(d, "solver.inp"))
    if int(np.prod(s["size"])) > MAX_POINTS:
        return f"skip: {int(np.prod(s['size']))} points"
    if str(s.get("input_mode", "serial")) != "serial":
        return "skip: partitioned initial solution (tests/test_parallel_io.py pins those files)"
    w = tempfile.mkdtemp(prefix="hpb_ex_")
    try:
        for f in os.listdir(d):
            if os.path.isfile(os.path.join(d, f)) and f.endswith(".inp"):
                shutil.copy(os.path.join(d, f), w)
        exe = os.path.join(w, "gen")
        cc = ["g++", "-O1", "-w"] if gen.lower().endswith(("cpp", ".c")) and gen.endswith(("cpp", "C")) else ["gcc", "-O1", "-w", "-std=gnu99"]
        p = subprocess.run(cc + [gen, "-o", exe, "-lm"], capture_output=True, text=True)
        if p.returncode:
            return "skip: generator does not compile stand-alone"
        p = subprocess.run([exe], cwd=w, capture_output=True, text=True, timeout=300, input="\n")
        if not os.path.exists(os.path.join(w, "initial.inp")):
            return "skip: generator wrote no initial.inp"
        _set_iproc_one(os.path.join(w, "solver.inp"), int(s["ndims"]), 1)
        p = subprocess.run([EXE, "rhs"], cwd=w, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="1"))      # 1 thread: the reference's OpenMP build races
        if p.returncode:
            return "skip: reference harness failed: " + (p.stderr.strip().splitlines() or p.stdout.strip().splitlines() or ["?"])[-1][:80]
        ref = {os.path.basename(f)[4:-4]: hypario.read_ref_dump(f)["data"] for f in glob.glob(os.path.join(w, "ref_*.bin"))}
        case = cases.from_directory(w)
        S = hpo.Setup(case, mpi_semantics=True)
        O = hpo.Oracle(S)
        u = S.local_u0()
        rhs, hyp, par, src = O.rhs(u, parts=True)
        bad = []
        for k, a in (("u", u), ("hyp", hyp), ("par", par), ("source", src), ("rhs", rhs)):
            b = ref[k]
            m = np.isfinite(b)                                   # never-filled corner ghosts may hold NaN in the reference
            if a.shape != b.shape or not np.array_equal(a[m], b[m]):
                bad.append(k if a.shape != b.shape else f"{k} (max diff {np.abs(a[m] - b[m]).max():.2e})")
        if bad:
            return "DIFFERS: " + ", ".join(bad)
        # two steps of the reference's TimePreStep / TimeStep / TimePostStep loop with the directory's own integrator and dt
        if int(np.prod(s["size"])) <= MAX_POINTS_STEPS:
            p = subprocess.run([EXE, "steps", "2"], cwd=w, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="1"))
            if p.returncode:
                return "BIT-IDENTICAL (u, hyp, par, source, rhs); steps: reference harness failed"
            uf = hypario.read_ref_dump(os.path.join(w, "ref_ufinal.bin"))["data"]
            u = S.local_u0()
            for _ in range(2):
                O.time_step(u, float(case.solver["dt"]), hpo.rk_type_of(case))
            # equal_nan: a directory meant for an implicit integrator blows up under explicit RK in the reference and here alike
            if not np.array_equal(S.interior(u).ravel(), S.interior(uf).ravel(), equal_nan=True):
                return f"DIFFERS: u after 2 steps (max diff {np.abs(S.interior(u).ravel() - S.interior(uf).ravel()).max():.2e})"
            return "BIT-IDENTICAL (u, hyp, par, source, rhs; u after 2 steps)"
        return "BIT-IDENTICAL (u, hyp, par, source, rhs)"
    finally:
        shutil.rmtree(w, ignore_errors=True)


def main():
    args = [a for a in sys.argv[1:]]
    top = args[0] if args and os.path.isdir(args[0]) else "/root/reference/Examples"
    only = [a for a in args if not os.path.isdir(a)]
    dirs = sorted(os.path.dirname(os.path.join(r, f)) for r, _, fs in os.walk(top) for f in fs if f == "solver.inp")
    lines, n_ok, n_bad, n_skip = [], 0, 0, 0
    for d in dirs:
        rel = os.path.relpath(d, top)
        if only and not any(o in rel for o in only):
            continue
        try:
            if classify(d) is not None:
                continue
            r = check(d)
        except Exception as e:
            r = f"skip: {type(e).__name__}: {str(e)[:80]}"
        n_ok += r.startswith("BIT")
        n_bad += r.startswith("DIFF")
        n_skip += r.startswith("skip")
        lines.append(f"{rel:75s} {r}")
        print(lines[-1], flush=True)
    lines.append(f"\n{n_ok} bit-identical, {n_bad} differ, {n_skip} skipped, of {len(lines)} accepted run directories")
    print(lines[-1])
    if not only:
        with open(os.path.join(ROOT, "profiles", "examples_parity.txt"), "w") as f:
            f.write("tools/examples_parity.py: oracle + package readers vs the unmodified reference (hypar_ref_mpi1 rhs) on the reference's\n"
                    "own Examples/ run directories, initial solutions from the directories' own aux/ generators, iproc set to 1.\n\n")
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()

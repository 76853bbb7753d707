"""Run the reference executable (oracle/_ref) on a Case. TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations
import glob, os, subprocess, tempfile, shutil
import numpy as np
from hypar_b200 import hypario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def ref_available(exe: str = "hypar_ref") -> bool:
    return os.access(os.path.join(REFDIR, exe), os.X_OK)


def run_reference(case, mode: str = "rhs", args=(), exe: str = "hypar_ref", threads: int = 1,
                  keep: str | None = None, timeout: int = 3600):
    """Returns {name: ndarray} for every ref_*.bin the harness wrote, plus 'stdout'."""
    d = keep or tempfile.mkdtemp(prefix="hpbref_")
    try:
        case.write(d)
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        p = subprocess.run([os.path.join(REFDIR, exe), mode, *map(str, args)], cwd=d, env=env,
                           capture_output=True, text=True, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError(f"reference failed ({p.returncode}):\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
        out = {"stdout": p.stdout}
        for f in glob.glob(os.path.join(d, "ref_*.bin")):
            out[os.path.basename(f)[4:-4]] = hypario.read_ref_dump(f)
        for f in glob.glob(os.path.join(d, "ref_*.txt")):
            out[os.path.basename(f)[4:-4]] = float(open(f).read().split()[0])
        if os.path.exists(os.path.join(d, "op.bin")):
            out["op"] = hypario.read_op_bin(os.path.join(d, "op.bin"))
        return out
    finally:
        if keep is None:
            shutil.rmtree(d, ignore_errors=True)

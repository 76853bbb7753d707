"""Readers/writers for HyPar's own input and output files.

The B200 path keeps the reference's files byte-compatible so an existing case
directory (solver.inp, boundary.inp, physics.inp, weno.inp, initial.inp) can be
pointed at this library unchanged, and its solution files can be diffed with the
reference's own tools.

Formats follow the reference:
  * solver.inp / physics.inp / weno.inp : ``begin  <key> <value...>  end`` keyword
    files (src/Simulation/ReadInputs.c:93-420,
    src/PhysicalModels/NavierStokes3D/NavierStokes3DInitialize.c:79-94,
    src/InterpolationFunctions/WENOInitialize.c:62-96)
  * boundary.inp : zone count, then ``type dim face xmin0 xmax0 ...`` (+ wall
    velocity line for slip-wall) (src/Simulation/InitializeBoundaries.c:84-130)
  * initial.inp (binary): ``[x_0 | x_1 | ... | u AoS, dim 0 fastest]`` global, no
    ghosts (src/IOFunctions/ReadArray.c:225-256)
  * op.bin : ``int ndims, nvars, dim[ndims]; double x[sum dim]; double u[]``
    (src/IOFunctions/WriteBinary.c:34-90)
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Sequence

import numpy as np

# keys of solver.inp whose value is a vector of ndims integers
_VECTOR_KEYS = ("size", "iproc", "size_exact")


# --------------------------------------------------------------------------- keyword files
def write_keyword_file(path: str, entries: Dict[str, object]) -> None:
    with open(path, "w") as f:
        f.write("begin\n")
        for k, v in entries.items():
            if isinstance(v, (list, tuple, np.ndarray)):
                v = " ".join(_fmt(x) for x in v)
            else:
                v = _fmt(v)
            f.write(f"  {k:<20s}{v}\n")
        f.write("end\n")


def _fmt(x) -> str:
    if isinstance(x, float):
        return repr(x)
    return str(x)


def read_keyword_file(path: str, ndims_hint: int | None = None,
                      vector_keys: Dict[str, int] | None = None) -> Dict[str, object]:
    """Parse a ``begin ... end`` file into {key: str | [str,...]}.

    ``vector_keys`` maps a key to the number of values that follow it
    (e.g. ``{"size": ndims}``); all other keys take one value, like the
    reference's fscanf loop.
    """
    with open(path) as f:
        words = f.read().split()
    if not words or words[0] != "begin":
        raise ValueError(f"Error: Illegal format in file \"{os.path.basename(path)}\".")
    out: Dict[str, object] = {}
    i = 1
    vector_keys = dict(vector_keys or {})
    while i < len(words) and words[i] != "end":
        key = words[i]
        i += 1
        if key == "ndims" and ndims_hint is None:
            ndims_hint = int(words[i])
        n = vector_keys.get(key)
        if n is None and key in _VECTOR_KEYS and ndims_hint is not None:
            n = ndims_hint
        if n is None:
            out[key] = words[i]
            i += 1
        else:
            out[key] = words[i:i + n]
            i += n
    return out


def read_solver_inp(path: str) -> Dict[str, object]:
    raw = read_keyword_file(path)
    ndims = int(raw.get("ndims", 1))
    cfg: Dict[str, object] = {
        # defaults of src/Simulation/ReadInputs.c:112-146
        "ndims": ndims, "nvars": 1, "ghost": 1, "n_iter": 0, "restart_iter": 0,
        "time_scheme": "euler", "time_scheme_type": " ", "hyp_space_scheme": "1",
        "hyp_flux_split": "no", "hyp_interp_type": "characteristic",
        "par_space_type": "nonconservative-1stage", "par_space_scheme": "2",
        "dt": 0.0, "conservation_check": "no", "screen_op_iter": 1, "file_op_iter": 1000,
        "op_file_format": "text", "ip_file_type": "ascii", "input_mode": "serial",
        "output_mode": "serial", "op_overwrite": "no", "model": "none",
        "iproc": [1] * ndims,
    }
    for k, v in raw.items():
        if k in ("size", "iproc", "size_exact"):
            cfg[k] = [int(x) for x in v]
        elif k in ("ndims", "nvars", "ghost", "n_iter", "restart_iter", "screen_op_iter", "file_op_iter"):
            cfg[k] = int(v)
        elif k == "dt":
            cfg[k] = float(v)
        else:
            cfg[k] = v
    return cfg


# --------------------------------------------------------------------------- boundary.inp
def write_boundary_inp(path: str, zones: Sequence[dict]) -> None:
    """zones: dicts with type, dim, face, xmin[ndims], xmax[ndims] (+ wall_velocity)."""
    with open(path, "w") as f:
        f.write(f"{len(zones)}\n")
        for z in zones:
            ext = "  ".join(f"{a!r} {b!r}" for a, b in zip(z["xmin"], z["xmax"]))
            f.write(f"{z['type']}  {z['dim']}  {z['face']}  {ext}\n")
            if z["type"] in ("slip-wall", "noslip-wall"):
                f.write(" ".join(repr(float(v)) for v in z["wall_velocity"]) + "\n")


def read_boundary_inp(path: str, ndims: int, nvars: int) -> List[dict]:
    with open(path) as f:
        w = f.read().split()
    n = int(w[0])
    i = 1
    zones = []
    for _ in range(n):
        z = {"type": w[i], "dim": int(w[i + 1]), "face": int(w[i + 2])}
        i += 3
        z["xmin"] = [float(w[i + 2 * d]) for d in range(ndims)]
        z["xmax"] = [float(w[i + 2 * d + 1]) for d in range(ndims)]
        i += 2 * ndims
        if z["type"] in ("slip-wall", "noslip-wall"):
            z["wall_velocity"] = [float(x) for x in w[i:i + ndims]]
            i += ndims
        elif z["type"] in ("dirichlet", "sponge"):
            z["values"] = [float(x) for x in w[i:i + nvars]]
            i += nvars
        zones.append(z)
    return zones


# --------------------------------------------------------------------------- arrays
def write_initial_bin(path: str, x: Sequence[np.ndarray], u: np.ndarray) -> None:
    """x: list of 1-D coordinate arrays; u: shape (N_{nd-1},...,N_0, nvars) (C order,
    i.e. dim 0 fastest among the spatial axes, nvars innermost)."""
    with open(path, "wb") as f:
        for xd in x:
            np.ascontiguousarray(xd, dtype=np.float64).tofile(f)
        np.ascontiguousarray(u, dtype=np.float64).tofile(f)


def read_initial_bin(path: str, dims: Sequence[int], nvars: int):
    sz = int(np.sum(dims))
    n = int(np.prod(dims)) * nvars
    raw = np.fromfile(path, dtype=np.float64, count=sz + n)
    x, off = [], 0
    for d in dims:
        x.append(raw[off:off + d].copy())
        off += d
    u = raw[off:off + n].reshape(tuple(reversed(list(dims))) + (nvars,)).copy()
    return x, u


def write_op_bin(path: str, x: Sequence[np.ndarray], u: np.ndarray) -> None:
    """Solution file in the reference's binary format (WriteBinary.c:34-90)."""
    nvars = u.shape[-1]
    dims = [len(xd) for xd in x]
    with open(path, "wb") as f:
        f.write(struct.pack("i", len(dims)))
        f.write(struct.pack("i", nvars))
        f.write(struct.pack(f"{len(dims)}i", *dims))
        for xd in x:
            np.ascontiguousarray(xd, dtype=np.float64).tofile(f)
        np.ascontiguousarray(u, dtype=np.float64).tofile(f)


def read_op_bin(path: str):
    with open(path, "rb") as f:
        ndims, nvars = struct.unpack("2i", f.read(8))
        dims = struct.unpack(f"{ndims}i", f.read(4 * ndims))
        x = [np.fromfile(f, dtype=np.float64, count=d) for d in dims]
        u = np.fromfile(f, dtype=np.float64, count=int(np.prod(dims)) * nvars)
    return x, u.reshape(tuple(reversed(dims)) + (nvars,))


def read_ref_dump(path: str):
    """Dump written by oracle/ref_harness.cpp: {int ndims,nvars,ghosts,dim[]} + doubles."""
    with open(path, "rb") as f:
        ndims, nvars, ghosts = struct.unpack("3i", f.read(12))
        dims = list(struct.unpack(f"{ndims}i", f.read(4 * ndims)))
        a = np.fromfile(f, dtype=np.float64)
    return {"ndims": ndims, "nvars": nvars, "ghosts": ghosts, "dims": dims, "data": a}

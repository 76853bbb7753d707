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
            if k == "N_bv":
                continue            # written with HB 3 (NavierStokes3DInitialize.c:176-180: "HB 3 <N_bv>")
            if k == "HB" and int(v) == 3:
                v = [3, float(entries.get("N_bv", 0.0))]
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
                      vector_keys: Dict[str, int] | None = None, nsims: int = 1) -> Dict[str, object]:
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
            n = ndims_hint * nsims          # ensembles: one vector per simulation, in order (ReadInputs.c:186-250)
        if key == "HB" and words[i] == "3":      # "HB 3 <N_bv>"
            out["HB"], out["N_bv"] = "3", words[i + 1]
            i += 2
        elif key in ("input_mode", "output_mode") and words[i] != "serial":
            # ReadInputs.c:358-366: "parallel" / "mpi-io" are followed by the number of I/O ranks
            out[key], out["n_io_ranks"] = words[i], words[i + 1]
            i += 2
        elif n is None:
            out[key] = words[i]
            i += 1
        else:
            out[key] = words[i:i + n]
            i += n
    return out


def read_simulation_inp(path: str) -> int:
    """simulation.inp (src/main.cpp:181-240): ``nsims`` > 1 selects the ensemble driver. 1 when the file is absent."""
    if not os.path.exists(path):
        return 1
    raw = read_keyword_file(path)
    return int(raw.get("nsims", 1))


def read_ensemble_solver_inp(path: str, nsims: int) -> List[Dict[str, object]]:
    """solver.inp of an ensemble run: ``size`` / ``iproc`` / ``size_exact`` hold one vector per simulation, everything else is
    shared (ReadInputs.c:176-410 copies sim[0]'s value to the others)."""
    raw = read_keyword_file(path, nsims=nsims)
    nd = int(raw.get("ndims", 1))
    out = []
    for n in range(nsims):
        one = dict(raw)
        for k in _VECTOR_KEYS:
            if k in raw:
                one[k] = raw[k][n * nd:(n + 1) * nd]
        cfg = _solver_cfg(one)
        out.append(cfg)
    return out


def read_solver_inp(path: str) -> Dict[str, object]:
    return _solver_cfg(read_keyword_file(path))


def _solver_cfg(raw: Dict[str, object]) -> Dict[str, object]:
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
            # type-specific data, InitializeBoundaries.c:107-175
            t = z["type"]
            if t in ("slip-wall", "noslip-wall"):
                f.write(" ".join(repr(float(v)) for v in z["wall_velocity"]) + "\n")
            elif t in ("dirichlet", "sponge"):
                f.write(" ".join(repr(float(v)) for v in z["values"]) + "\n")
            elif t == "subsonic-inflow":
                f.write(" ".join(repr(float(v)) for v in [z["density"], *z["velocity"]]) + "\n")
            elif t == "subsonic-outflow":
                f.write(repr(float(z["pressure"])) + "\n")
            elif t in ("subsonic-ambivalent", "supersonic-inflow"):
                f.write(" ".join(repr(float(v)) for v in [z["density"], *z["velocity"], z["pressure"]]) + "\n")


def read_boundary_inp(path: str, ndims: int, nvars: int) -> List[dict]:
    """boundary.inp the way InitializeBoundaries.c:85-200 scans it: type, dim, face and the extents must be there; the
    zone-specific values are read with unchecked fscanf("%lf") calls, so a value that is missing (the next token is the
    next zone's type, or the file ends) stays 0 and consumes nothing -- the reference's own Examples rely on that
    (NavierStokes2D/1DSodShockTubeWithGravity gives one wall velocity in 2-D, FlatPlateSupersonic no outflow pressure)."""
    with open(path) as f:
        w = f.read().split()
    n = int(w[0])
    pos = [1]

    def lf() -> float:
        if pos[0] < len(w):
            try:
                v = float(w[pos[0]])
            except ValueError:
                return 0.0
            pos[0] += 1
            return v
        return 0.0

    zones = []
    for _ in range(n):
        i = pos[0]
        z = {"type": w[i], "dim": int(w[i + 1]), "face": int(w[i + 2])}
        i += 3
        z["xmin"] = [float(w[i + 2 * d]) for d in range(ndims)]
        z["xmax"] = [float(w[i + 2 * d + 1]) for d in range(ndims)]
        pos[0] = i + 2 * ndims
        if z["type"] in ("slip-wall", "noslip-wall"):
            z["wall_velocity"] = [lf() for _ in range(ndims)]
        elif z["type"] in ("dirichlet", "sponge"):
            z["values"] = [lf() for _ in range(nvars)]
        elif z["type"] == "subsonic-inflow":
            z["density"] = lf()
            z["velocity"] = [lf() for _ in range(ndims)]
        elif z["type"] == "subsonic-outflow":
            z["pressure"] = lf()
        elif z["type"] in ("subsonic-ambivalent", "supersonic-inflow"):
            z["density"] = lf()
            z["velocity"] = [lf() for _ in range(ndims)]
            z["pressure"] = lf()
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


def write_initial_ascii(path: str, x: Sequence[np.ndarray], u: np.ndarray) -> None:
    """The text flavour of initial.inp (``ip_file_type ascii``, ReadArray.c:173-217): the grid, dimension by dimension, then the
    field ONE VARIABLE AFTER THE OTHER, each over the whole grid with dimension 0 fastest. Written with repr(): 17
    significant digits, so the reference's fscanf("%lf") gets the same doubles back."""
    nvars = u.shape[-1]
    with open(path, "w") as f:
        for xd in x:
            f.write(" ".join(repr(float(v)) for v in np.asarray(xd, dtype=np.float64)) + "\n")
        flat = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, nvars)
        for v in range(nvars):
            f.write(" ".join(repr(float(t)) for t in flat[:, v]) + "\n")


def read_initial_ascii(path: str, dims: Sequence[int], nvars: int):
    dims = [int(d) for d in dims]
    sz, npts = int(np.sum(dims)), int(np.prod(dims))
    with open(path) as f:
        raw = np.array(f.read().split(), dtype=np.float64)
    if raw.size < sz + npts * nvars:
        raise ValueError(f"Error in ReadArraySerial(): unable to read data ({path}: {raw.size} values, "
                         f"{sz + npts * nvars} expected)")
    x, off = [], 0
    for d in dims:
        x.append(raw[off:off + d].copy())
        off += d
    u = np.ascontiguousarray(raw[off:off + npts * nvars].reshape(nvars, npts).T)
    return x, u.reshape(tuple(reversed(dims)) + (nvars,))


def read_initial(path: str, dims: Sequence[int], nvars: int, ip_file_type: str = "binary"):
    """initial.inp (or any array file of that layout) in the flavour solver.inp names (ReadArray.c:173, :218)."""
    if str(ip_file_type) == "ascii":
        return read_initial_ascii(path, dims, nvars)
    if str(ip_file_type) in ("bin", "binary"):
        return read_initial_bin(path, dims, nvars)
    raise ValueError(f"ip_file_type '{ip_file_type}' is neither ascii nor binary")


def write_initial(path: str, x: Sequence[np.ndarray], u: np.ndarray, ip_file_type: str = "binary") -> None:
    (write_initial_ascii if str(ip_file_type) == "ascii" else write_initial_bin)(path, x, u)


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


def _point_table(x: Sequence[np.ndarray], u: np.ndarray):
    """Rows of WriteText.c:54-66: the ndims indices, the ndims coordinates, the nvars values; points with dimension 0 fastest."""
    dims = [len(xd) for xd in x]
    nd, nvars = len(dims), u.shape[-1]
    grids = np.meshgrid(*[np.arange(n) for n in reversed(dims)], indexing="ij")      # axes: dim nd-1 ... dim 0
    idx = [grids[nd - 1 - d].reshape(-1) for d in range(nd)]
    cols = [i.astype(np.float64) for i in idx] + [np.asarray(x[d], dtype=np.float64)[idx[d]] for d in range(nd)]
    tab = np.column_stack(cols + [np.ascontiguousarray(u, dtype=np.float64).reshape(-1, nvars)])
    fmt = "%4d " * nd + "%+1.16E " * (nd + nvars)
    return tab, fmt


def write_op_text(path: str, x: Sequence[np.ndarray], u: np.ndarray) -> None:
    """``op_file_format text`` (WriteText.c:27-69): one line per grid point -- indices ``%4d``, coordinates and values
    ``%+1.16E``, each followed by a blank. Byte-identical to the reference's file for finite values."""
    tab, fmt = _point_table(x, u)
    with open(path, "w") as f:
        np.savetxt(f, tab, fmt=fmt, newline="\n")


def write_op_tecplot(path: str, x: Sequence[np.ndarray], u: np.ndarray) -> None:
    """``op_file_format tecplot2d`` / ``tecplot3d`` (WriteTecplot2D.c:48-72, WriteTecplot3D.c:49-73): the text rows under a
    Tecplot POINT-format header; variables are named "00", "01", ..."""
    dims = [len(xd) for xd in x]
    nd, nvars = len(dims), u.shape[-1]
    if nd not in (2, 3):
        raise ValueError(f"Error in WriteTecplot{nd}D(): hardcoded for 2- and 3-dimensional problems only")
    tab, fmt = _point_table(x, u)
    names = ["I", "J", "K"][:nd] + ["X", "Y", "Z"][:nd] + [f"{v // 10}{v % 10}" for v in range(nvars)]
    with open(path, "w") as f:
        f.write("VARIABLES=" + "".join(f'"{n}",' for n in names) + "\n")
        f.write("ZONE " + ",".join(f"{a}={n}" for a, n in zip("IJK", dims)) + ",F=POINT\n")
        np.savetxt(f, tab, fmt=fmt, newline="\n")


def read_op_text(path: str, ndims: int, nvars: int):
    """Back from a text / tecplot solution file: (x per dimension, u of shape (N_{nd-1}, ..., N_0, nvars))."""
    rows = []
    with open(path) as f:
        for ln in f:
            if ln.startswith(("VARIABLES", "ZONE")):
                continue
            rows.append(ln.split())
    a = np.array(rows, dtype=np.float64)
    idx = a[:, :ndims].astype(np.int64)
    dims = [int(idx[:, d].max()) + 1 for d in range(ndims)]
    x = []
    for d in range(ndims):
        xd = np.zeros(dims[d])
        xd[idx[:, d]] = a[:, ndims + d]
        x.append(xd)
    return x, a[:, 2 * ndims:].reshape(tuple(reversed(dims)) + (nvars,))


def solution_file_name(op_file_format: str, overwrite: bool, index: int = 0, root: str = "op", index_length: int = 5) -> str:
    """``<root>[_<index>].<dat|bin>`` as OutputSolution.cpp:100-118 / InitializeSolvers.c:416-431 form it."""
    fmt = str(op_file_format)
    if fmt in ("binary", "bin"):
        ext = ".bin"
    elif fmt in ("text", "tecplot2d", "tecplot3d"):
        ext = ".dat"
    else:
        raise ValueError(f"op_file_format '{fmt}' writes no file")
    return root + ("" if overwrite else f"_{index:0{index_length}d}") + ext


def write_solution(path: str, x: Sequence[np.ndarray], u: np.ndarray, op_file_format: str = "binary") -> None:
    fmt = str(op_file_format)
    if fmt in ("binary", "bin"):
        write_op_bin(path, x, u)
    elif fmt == "text":
        write_op_text(path, x, u)
    elif fmt in ("tecplot2d", "tecplot3d"):
        if len(x) != int(fmt[7]):
            raise ValueError(f"Error in WriteTecplot{fmt[7]}D(): hardcoded for {fmt[7]}-dimensional problems only")
        write_op_tecplot(path, x, u)
    else:
        raise ValueError(f"op_file_format '{fmt}' writes no file")


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


# --------------------------------------------------------------------------- partitioned ("parallel") and MPI-IO files
# SURVEY 8f rank 2. The reference's scalable I/O modes (solver.inp: input_mode / output_mode = "parallel n" | "mpi-io n"):
#   <root>_par.inp.<nnnn>   ReadArrayParallel  (ReadArray.c:293-505), made from <root>.inp by Extras/ParallelInput.c
#   <root>.bin.<nnnn>       WriteArrayParallel (WriteArray.c:139-323), stitched by Extras/ParallelOutput.c
#   <root>_mpi.inp          ReadArrayMPI_IO    (ReadArray.c:513-650)
# One file per I/O group; a file is the concatenation, in rank order, of one block per member rank:
#   [x_0 (dim_local[0]) | ... | x_{nd-1} | u AoS, no ghosts, dim 0 fastest]            (doubles, no header)
# With op_overwrite = no every output time appends another round of blocks to the same files.
# In the reference the group leader receives every member's block over MPI and writes / reads the file alone. Here
# every rank (= GPU) addresses ITS OWN block with a positional pread / pwrite at the offset the leader would have
# reached -- the files are byte-identical, no block is gathered through one rank, and all offsets are 64-bit.

def partition1d(nglobal: int, nproc: int, rank: int) -> int:
    """MPIPartition1D.c: nglobal/nproc points, the remainder on the last rank"""
    n = nglobal // nproc
    return nglobal - n * (nproc - 1) if rank == nproc - 1 else n


def rank_nd(iproc: Sequence[int], rank: int) -> List[int]:
    """MPIRanknD.c: rank = ip0 + iproc0*(ip1 + iproc1*ip2)"""
    ip = []
    for p in iproc:
        ip.append(rank % p)
        rank //= p
    return ip


def local_extent(dim_global: Sequence[int], iproc: Sequence[int], rank: int):
    """(is, ie) of this rank's block in global indices (MPILocalDomainLimits.c)"""
    ip = rank_nd(iproc, rank)
    is_ = [sum(partition1d(dim_global[d], iproc[d], r) for r in range(ip[d])) for d in range(len(iproc))]
    ie = [is_[d] + partition1d(dim_global[d], iproc[d], ip[d]) for d in range(len(iproc))]
    return is_, ie


def io_group(nproc: int, n_io: int, rank: int):
    """MPIIOGroups.c:45-72 -> (group index = file index, first rank, one-past-last rank); a rank count that is not a
    multiple of the number of I/O ranks falls back to ONE group, as the reference does"""
    if n_io < 1 or nproc % n_io != 0:
        n_io = 1
    gs = nproc // n_io
    g = rank // gs
    return g, g * gs, (g + 1) * gs


def block_doubles(dim_global, iproc, nvars: int, rank: int) -> int:
    is_, ie = local_extent(dim_global, iproc, rank)
    n = [b - a for a, b in zip(is_, ie)]
    return int(sum(n)) + int(nvars) * int(np.prod(n, dtype=np.int64))


def _block_offset(dim_global, iproc, nvars, first_rank: int, rank: int) -> int:
    return sum(block_doubles(dim_global, iproc, nvars, r) for r in range(first_rank, rank))


def _pack_block(x_local: Sequence[np.ndarray], u_local: np.ndarray) -> np.ndarray:
    return np.concatenate([np.ascontiguousarray(v, dtype=np.float64).reshape(-1) for v in list(x_local) + [u_local]])


def write_parallel_block(root_ext: str, rank: int, dim_global, iproc, nvars: int, n_io: int,
                         x_local: Sequence[np.ndarray], u_local: np.ndarray, record: int = 0, truncate: bool = False) -> str:
    """This rank's block of <root_ext>.<nnnn> (root_ext = "op.bin": WriteArrayParallel's filename_root).
    record = how many output times are already in the file (op_overwrite = no appends; yes -> always 0).
    u_local: (N_{nd-1},...,N_0,nvars), no ghosts. Returns the file name.
    truncate: cut the file at the end of this record. WriteArrayParallel opens the file "wb" on the first write of a run
    (and on every write with op_overwrite yes), so a longer file left by an earlier run loses its trailing records; here
    every rank writes on its own, so the caller asks for the cut -- and may only do so when no rank starts record r+1
    before all have finished record r (a barrier between output times), else a faster rank's next record would be cut."""
    nproc = int(np.prod(iproc))
    g, first, last = io_group(nproc, n_io, rank)
    fname = f"{root_ext}.{g:04d}"
    buf = _pack_block(x_local, u_local)
    assert buf.size == block_doubles(dim_global, iproc, nvars, rank), "block size does not match the decomposition"
    group_total = _block_offset(dim_global, iproc, nvars, first, last)
    off = 8 * (record * group_total + _block_offset(dim_global, iproc, nvars, first, rank))
    fd = os.open(fname, os.O_WRONLY | os.O_CREAT, 0o644)
    try:
        mv, done = memoryview(buf).cast("B"), 0
        while done < len(mv):                        # pwrite may be partial beyond 2 GiB
            done += os.pwrite(fd, mv[done:done + (1 << 30)], off + done)
        end = 8 * (record + 1) * group_total
        if truncate and os.fstat(fd).st_size > end:       # shrinks only: nothing of this or an earlier record lies beyond
            os.ftruncate(fd, end)
    finally:
        os.close(fd)
    return fname


def _read_block(fname: str, offset_doubles: int, dims: Sequence[int], nvars: int):
    n = int(sum(dims)) + nvars * int(np.prod(dims, dtype=np.int64))
    raw = np.fromfile(fname, dtype=np.float64, count=n, offset=8 * offset_doubles)
    if raw.size != n:
        raise IOError(f"{fname} contains insufficient data")            # as ReadArrayParallel's error
    x, o = [], 0
    for d in dims:
        x.append(raw[o:o + d].copy())
        o += d
    return x, raw[o:].reshape(tuple(reversed(list(dims))) + (nvars,))


def read_parallel_block(fname_root: str, rank: int, dim_global, iproc, nvars: int, n_io: int, record: int = 0,
                        suffix: str = "_par.inp"):
    """This rank's (x_local, u_local) out of <fname_root>_par.inp.<nnnn> (or, suffix=".bin", an output file)."""
    nproc = int(np.prod(iproc))
    g, first, last = io_group(nproc, n_io, rank)
    is_, ie = local_extent(dim_global, iproc, rank)
    group_total = _block_offset(dim_global, iproc, nvars, first, last)
    off = record * group_total + _block_offset(dim_global, iproc, nvars, first, rank)
    return _read_block(f"{fname_root}{suffix}.{g:04d}", off, [b - a for a, b in zip(is_, ie)], nvars)


def read_mpi_io_block(fname_root: str, rank: int, dim_global, iproc, nvars: int):
    """This rank's block of <fname_root>_mpi.inp: ONE file, the blocks of all ranks in rank order (ReadArray.c:558-572)"""
    is_, ie = local_extent(dim_global, iproc, rank)
    off = _block_offset(dim_global, iproc, nvars, 0, rank)
    return _read_block(f"{fname_root}_mpi.inp", off, [b - a for a, b in zip(is_, ie)], nvars)


def serial_to_parallel(path_serial: str, fname_root: str, dim_global, iproc, nvars: int, n_io: int, mpi_io: bool = False):
    """Extras/ParallelInput.c: split <root>.inp (global, read_initial_bin's layout) into <root>_par.inp.<nnnn>
    (or the single <root>_mpi.inp). Returns the files written."""
    x, u = read_initial_bin(path_serial, dim_global, nvars)
    nproc = int(np.prod(iproc))
    files = set()
    for r in range(nproc):
        is_, ie = local_extent(dim_global, iproc, r)
        xl = [x[d][is_[d]:ie[d]] for d in range(len(iproc))]
        ul = u[tuple(slice(is_[d], ie[d]) for d in reversed(range(len(iproc))))]
        if mpi_io:
            buf = _pack_block(xl, ul)
            fname = f"{fname_root}_mpi.inp"
            with open(fname, "r+b" if fname in files else "wb") as f:
                f.seek(8 * _block_offset(dim_global, iproc, nvars, 0, r))
                buf.tofile(f)
        else:
            fname = write_parallel_block(f"{fname_root}_par.inp", r, dim_global, iproc, nvars, n_io, xl, ul)
        files.add(fname)
    return sorted(files)


def parallel_to_serial(root_ext: str, dim_global, iproc, nvars: int, n_io: int, record: int = 0):
    """Extras/ParallelOutput.c: stitch the blocks of <root_ext>.<nnnn> (one output time) into the global (x, u) that
    WriteArraySerial would have written (write_op_bin's arguments)."""
    nd = len(iproc)
    x = [np.zeros(n) for n in dim_global]
    u = np.zeros(tuple(reversed(list(dim_global))) + (nvars,))
    root, ext = root_ext.rsplit(".", 1)
    for r in range(int(np.prod(iproc))):
        xl, ul = read_parallel_block(root, r, dim_global, iproc, nvars, n_io, record, suffix="." + ext)
        is_, ie = local_extent(dim_global, iproc, r)
        for d in range(nd):
            x[d][is_[d]:ie[d]] = xl[d]
        u[tuple(slice(is_[d], ie[d]) for d in reversed(range(nd)))] = ul
    return x, u

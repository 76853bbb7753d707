"""Partitioned I/O: the reference's parallel / MPI-IO file formats written and read block by
block, every rank addressing its own block (hypar_b200/hypario.py).

CPU: the format against the unmodified reference executable run live (it reads the files we split, we stitch the
files it writes), round trips with remainders, one group / several groups / a group count that does not divide the rank
count, appended output records, and world-4 concurrent writers (one process per rank, as on the GPUs).
GPU: Solver.write_solution_parallel / load_solution_parallel on decomposed device solutions.
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

from hypar_b200 import cases, hypario as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "hypar_main_mpi1")


def _field(dg, nv, seed=3):
    rng = np.random.RandomState(seed)
    x = [np.sort(rng.rand(n)) for n in dg]
    u = rng.standard_normal(tuple(reversed(dg)) + (nv,))
    return x, u


def _block(x, u, dg, ip, r):
    is_, ie = H.local_extent(dg, ip, r)
    nd = len(dg)
    return [x[d][is_[d]:ie[d]] for d in range(nd)], u[tuple(slice(is_[d], ie[d]) for d in reversed(range(nd)))]


def test_io_groups_follow_the_reference():
    """MPIIOGroups.c:45-72: equal groups of consecutive ranks; a non-divisor falls back to one group"""
    assert [H.io_group(8, 4, r) for r in (0, 1, 2, 7)] == [(0, 0, 2), (0, 0, 2), (1, 2, 4), (3, 6, 8)]
    assert H.io_group(6, 4, 5) == (0, 0, 6)
    assert H.io_group(1, 1, 0) == (0, 0, 1)


def test_partition_matches_the_library():
    from hypar_b200 import _lib
    L = _lib.load()
    for n, p in ((50, 2), (101, 3), (512, 8), (7, 7)):
        for r in range(p):
            assert H.partition1d(n, p, r) == L.hpb_partition1d(n, p, r)


@pytest.mark.parametrize("dg,ip,nv,nio", [([13, 10, 7], [2, 3, 2], 5, 4), ([13, 10, 7], [2, 3, 2], 5, 5),
                                           ([41, 9], [3, 2], 4, 2), ([101], [4], 3, 1), ([8, 8, 8], [1, 1, 1], 5, 1)])
def test_split_write_stitch_roundtrip(tmp_path, dg, ip, nv, nio):
    os.chdir(tmp_path)
    x, u = _field(dg, nv)
    H.write_initial_bin("initial.inp", x, u)
    nproc = int(np.prod(ip))
    files = H.serial_to_parallel("initial.inp", "initial", dg, ip, nv, nio)
    assert len(files) == (nio if nproc % nio == 0 else 1)
    assert sum(os.path.getsize(f) for f in files) == 8 * sum(H.block_doubles(dg, ip, nv, r) for r in range(nproc))
    H.serial_to_parallel("initial.inp", "initial", dg, ip, nv, 1, mpi_io=True)
    for r in range(nproc):
        xl, ul = H.read_parallel_block("initial", r, dg, ip, nv, nio)
        xr, ur = _block(x, u, dg, ip, r)
        assert np.array_equal(ul, ur) and all(np.array_equal(a, b) for a, b in zip(xl, xr))
        xm, um = H.read_mpi_io_block("initial", r, dg, ip, nv)
        assert np.array_equal(um, ur) and all(np.array_equal(a, b) for a, b in zip(xm, xr))
        # two appended output records (op_overwrite no), written in rank-reversed order to show the order is free
    for rec in (0, 1):
        for r in reversed(range(nproc)):
            xl, ul = _block(x, u, dg, ip, r)
            H.write_parallel_block("op.bin", r, dg, ip, nv, nio, xl, (rec + 1) * ul, record=rec)
    for rec in (0, 1):
        xs, us = H.parallel_to_serial("op.bin", dg, ip, nv, nio, record=rec)
        assert np.array_equal(us, (rec + 1) * u) and all(np.array_equal(a, b) for a, b in zip(xs, x))
    with pytest.raises(IOError):
        H.read_parallel_block("op", 0, dg, ip, nv, nio, record=2, suffix=".bin")


def _writer(args):
    d, r, dg, ip, nv, nio = args
    os.chdir(d)
    x, u = _field(dg, nv)
    xl, ul = _block(x, u, dg, ip, r)
    for rec in range(2):
        H.write_parallel_block("op.bin", r, dg, ip, nv, nio, xl, ul + rec, record=rec)
    return r


def test_concurrent_rank_writers(tmp_path):
    """one process per rank, all writing at once into the shared group files -- the layout the leader rank of
    WriteArrayParallel would have produced (blocks in rank order, records appended)"""
    dg, ip, nv, nio = [21, 17, 12], [2, 2, 2], 5, 2
    with mp.get_context("fork").Pool(8) as pool:
        assert sorted(pool.map(_writer, [(str(tmp_path), r, dg, ip, nv, nio) for r in range(8)])) == list(range(8))
    os.chdir(tmp_path)
    x, u = _field(dg, nv)
    # the reference's leader: sequential, rank order inside each group, one record after the other
    for g in range(nio):
        want = []
        for rec in range(2):
            for r in range(g * 4, (g + 1) * 4):
                xl, ul = _block(x, u, dg, ip, r)
                want.append(np.concatenate([np.concatenate(xl), (ul + rec).reshape(-1)]))
        assert np.array_equal(np.fromfile(f"op.bin.{g:04d}"), np.concatenate(want))


def test_stale_records_of_an_earlier_run_are_cut(tmp_path):
    """WriteArrayParallel opens "wb" on the first write of a run: a longer file from an earlier run must not keep its
    trailing records (stitching tools read to EOF). The rank-parallel writer cuts at the end of the record on request."""
    os.chdir(tmp_path)
    dg, ip, nv = [12, 10], [2, 1], 3
    x, u = _field(dg, nv)
    for rec in range(3):                                   # an earlier run: three output times
        for r in range(2):
            xl, ul = _block(x, u, dg, ip, r)
            H.write_parallel_block("op.bin", r, dg, ip, nv, 1, xl, ul + rec, record=rec)
    long_size = os.path.getsize("op.bin.0000")
    for r in range(2):                                     # this run: one output time
        xl, ul = _block(x, u, dg, ip, r)
        H.write_parallel_block("op.bin", r, dg, ip, nv, 1, xl, ul + 7, record=0, truncate=True)
    assert os.path.getsize("op.bin.0000") * 3 == long_size
    want = []
    for r in range(2):
        xl, ul = _block(x, u, dg, ip, r)
        want.append(np.concatenate([np.concatenate(xl), (ul + 7).reshape(-1)]))
    assert np.array_equal(np.fromfile("op.bin.0000"), np.concatenate(want))


@pytest.mark.parametrize("mode", ["parallel 1", "mpi-io 1"])
def test_formats_against_the_reference_live(tmp_path, mode):
    """the unmodified reference (1-rank MPI-semantics build) reads the partitioned input we split and writes
    partitioned output we stitch: both equal its own serial-mode files"""
    import subprocess
    if not os.access(REF_EXE, os.X_OK):
        pytest.skip("oracle/_ref/hypar_main_mpi1 not built (needs /root/reference)")
    dg = [14, 12, 10]
    c = cases.ns3d_density_wave(tuple(dg), "z")
    c.solver.update({"n_iter": 4, "file_op_iter": 2, "op_overwrite": "no"})
    ser, par = str(tmp_path / "ser"), str(tmp_path / "par")
    c.write(ser)
    c.solver.update({"input_mode": mode, "output_mode": "parallel 1"})
    c.write(par)
    os.chdir(par)
    H.serial_to_parallel("initial.inp", "initial", dg, [1, 1, 1], 5, 1, mpi_io=mode.startswith("mpi"))
    os.remove("initial.inp")
    for d in (ser, par):
        p = subprocess.run([REF_EXE], cwd=d, capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, OMP_NUM_THREADS="1"))
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert ("parallel mode" if mode.startswith("par") else "MPI-IO mode") in p.stdout
    for k in range(3):
        xs, us = H.read_op_bin(os.path.join(ser, f"op_{k:05d}.bin"))
        xp, up = H.parallel_to_serial(os.path.join(par, "op.bin"), dg, [1, 1, 1], 5, 1, record=k)
        assert np.array_equal(us, up) and all(np.array_equal(a, b) for a, b in zip(xs, xp)), f"record {k}"


@pytest.mark.gpu
def test_device_solution_through_partitioned_files(need_gpu, tmp_path):
    """8 ranks on one GPU: every rank loads its block of initial_par.inp.<nnnn>, steps, writes its block of
    op.bin.<nnnn>; the stitched file equals the decomposed solution gathered directly"""
    from _multirank import LocalRanks, MultiRankOracle
    os.chdir(tmp_path)
    dg, ip = [26, 25, 27], [2, 2, 2]
    case = cases.ns3d_turbulence(tuple(dg), "z", iproc=tuple(ip))
    H.write_initial_bin("initial.inp", case.x, case.u0)
    H.serial_to_parallel("initial.inp", "initial", dg, ip, 5, 4)
    H.serial_to_parallel("initial.inp", "initial", dg, ip, 5, 1, mpi_io=True)
    MO = MultiRankOracle(case)
    A, B, C = LocalRanks(case), LocalRanks(case), LocalRanks(case)
    A.set_solution(MO.local_u0())
    for sv in B.sv:
        sv.load_solution_parallel("initial", 4)
    for sv in C.sv:
        sv.load_solution_parallel("initial", 1, mode="mpi-io")
    for r in range(8):
        ua = A.sv[r].interior(A.sv[r].get_solution())
        assert np.array_equal(ua, B.sv[r].interior(B.sv[r].get_solution()))
        assert np.array_equal(ua, C.sv[r].interior(C.sv[r].get_solution()))
    for rec in range(2):
        B.time_step()
        A.time_step()
        for sv in B.sv:
            sv.write_solution_parallel("op.bin", 2, record=rec)
        xs, us = H.parallel_to_serial("op.bin", dg, ip, 5, 2, record=rec)
        assert all(np.array_equal(a, np.asarray(b)) for a, b in zip(xs, case.x))
        for r, sv in enumerate(A.sv):
            is_, ie = H.local_extent(dg, ip, r)
            assert np.array_equal(us[is_[2]:ie[2], is_[1]:ie[1], is_[0]:ie[0]], sv.interior(sv.get_solution()))
    for X in (A, B, C):
        X.close()

#!/usr/bin/env python
"""bench.py -- throughput of the explicit-RHS hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 512]

A "step" is one RK4 time step (TimePreStep BCs + 4 x [stage vector, BCs, halo, hyperbolic sweeps,
viscous terms] + step completion) of configuration C4 -- NavierStokes3D, WENO5 (mapped weights),
Rusanov upwinding, viscous terms (Re 333.33, Pr 0.72, Minf 0.3), periodic box -- on a synthetic
Taylor-Green + 16 solenoidal Fourier modes field, 512^3 points PER GPU (weak scaling: the global
grid is 512x512x512 / 512x512x1024 / 512x1024x1024 / 1024^3 at 1 / 2 / 4 / 8 GPUs).

Metric: Mpoint-RK-stage/s = global interior points x RK stages x steps / seconds.
  value : device-resident time loop (hpb_TimeStep), inputs already in HBM, CUDA events, max over ranks
  e2e   : the host-array entry point a HyPar caller binds (hpb_TimeIntegrate: H2D of u from pinned host
          memory + the step + D2H of u), every step
  roofline    : dominant kernel (one directional sweep), algorithmic bytes / CUDA-event time, vs the
                measured HBM copy bandwidth in MEASURED_PEAKS.json (see DESIGN.md for the byte counts)
  cpu_baseline: the reference's own CPU implementation (oracle/_ref, unmodified HyPar sources) on this
                box's host cores, same physics, 64^3 sample
--impl reference: only the reference CPU arm (bounded sample), same metric/unit/config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "3D NS WENO5 Mpoint-RK-stage/s"
UNIT = "Mpoint-RK-stage/s"
NSTAGES = 4
# algorithmic bytes per point (FP64, nvars = 5), DESIGN.md section "Roofline":
#   one directional sweep launch: read u (40 B) + read-modify-write rhs (80 B; the first direction only writes: 40 B)
#   whole RK stage (k never stored twice): 200 B  
# "sweep_fused": the last direction's sweep that also forms the next RK stage solution (stage fusion): read the
# stage solution (40) + read-modify-write of the right-hand side (80) + read u^n (40) + write the next stage solution (40) -- the
# whole-stage figure of SURVEY 8(d)
SWEEP_BYTES = {"sweep_x": 80.0, "sweep_y": 120.0, "sweep_z": 120.0, "sweep_fused": 200.0}
STAGE_BYTES = 200.0


def weak_grid(n: int, ngpus: int):
    """global size and iproc for `ngpus` blocks of n^3 (slowest dimensions split first)."""
    iproc = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[ngpus]
    return [n * iproc[0], n * iproc[1], n * iproc[2]], list(iproc)


def c4_inputs(size, iproc, dt=0.005, upwinding="rusanov", tstype="44", interp="components"):
    """solver.inp / boundary.inp / physics.inp / weno.inp of configuration C4 (hypar_b200.cases)."""
    from hypar_b200 import cases
    import numpy as np
    s = cases._solver(3, 5, size, "navierstokes3d", ts="rk", tstype=tstype, dt=dt, iproc=iproc,
                      par_type="nonconservative-2stage", par_scheme="4", interp=interp)
    b = cases._zones(3, "periodic", [-1e3] * 3, [1e3] * 3)
    ph = {"gamma": 1.4, "upwinding": upwinding, "Pr": 0.72, "Minf": 0.3, "Re": 333.333333333333333}
    w = cases.weno_inp("mapped")
    x = [np.arange(size[d], dtype=np.float64) * (2.0 * np.pi / size[d]) for d in range(3)]
    return s, b, ph, w, x


def c5_inputs(kind, size, iproc):
    """Configuration C5 (BASELINE.json configs[4]) at an arbitrary size, without materialising the global field:
    c5a = density sine wave (periodic, inviscid, RK4), c5b = rising thermal bubble (slip walls, gravity, HB 2, yc
    weights, SSPRK3) -- the dictionaries of hypar_b200.cases.ns3d_density_wave / ns3d_rising_bubble, dt scaled with
    the grid spacing (same CFL as the 64^3 cases)."""
    from hypar_b200 import cases
    import numpy as np
    small = (cases.ns3d_density_wave if kind == "c5a" else cases.ns3d_rising_bubble)((8, 8, 8))
    s = dict(small.solver)
    s["size"], s["iproc"] = list(size), list(iproc)
    s["dt"] = (1e-3 if kind == "c5a" else 0.01) * 64.0 / max(size)
    if kind == "c5a":
        x = [np.arange(size[d], dtype=np.float64) / size[d] for d in range(3)]
    else:
        x = [np.arange(size[d], dtype=np.float64) * (1000.0 / (size[d] - 1)) for d in range(3)]
    return s, small.boundary, small.physics, small.weno, x


def synth_field_c5(kind, x_loc, device):
    """This rank's block of the C5 initial fields (hypar_b200.cases: DensitySineWave exact.C:54-81,
    RisingThermalBubble_Config1 init.c:108-133), evaluated on the GPU. Returns (nz, ny, nx, 5) float64."""
    import math
    import torch
    gamma = 1.4
    X = torch.as_tensor(x_loc[0], device=device)[None, None, :]
    Y = torch.as_tensor(x_loc[1], device=device)[None, :, None]
    Z = torch.as_tensor(x_loc[2], device=device)[:, None, None]
    if kind == "c5a":
        rho = 1.0 + 0.1 * torch.sin(2 * math.pi * X) * torch.sin(2 * math.pi * Y) * torch.sin(2 * math.pi * Z)
        e = (1.0 / gamma) / (gamma - 1.0) + 0.5 * rho * 3.0
        return torch.stack([rho, rho, rho, rho, e], dim=-1)
    R, g, rho_ref, p_ref = 287.058, 9.8, 1.1612055171196529, 100000.0
    T_ref = p_ref / (R * rho_ref)
    Cp = gamma / (gamma - 1.0) * R
    r = torch.sqrt((X - 500.0) ** 2 + (Y - 260.0) ** 2 + (Z - 500.0) ** 2)
    dtheta = torch.where(r > 250.0, torch.zeros_like(r), 0.5 * (1.0 + torch.cos(math.pi * r / 250.0)))
    theta = T_ref + dtheta
    Pexner = (1.0 - (g / (Cp * T_ref)) * Y).expand_as(theta)
    rho = (p_ref / (R * theta)) * Pexner ** (1.0 / (gamma - 1.0))
    E = rho * (R / (gamma - 1.0)) * theta * Pexner
    zero = torch.zeros_like(rho)
    return torch.stack([rho, zero, zero, zero, E], dim=-1)


def synth_field_torch(x_loc, device, seed=20261017):
    """The C4 synthetic field of hypar_b200.cases.ns3d_turbulence evaluated on this rank's block, on the
    GPU (torch is plumbing here: it only creates the input). Returns (nz, ny, nx, 5) float64."""
    import numpy as np
    import torch
    gamma, Minf = 1.4, 0.3
    X = torch.as_tensor(x_loc[0], device=device)[None, None, :]
    Y = torch.as_tensor(x_loc[1], device=device)[None, :, None]
    Z = torch.as_tensor(x_loc[2], device=device)[:, None, None]
    vx = Minf * torch.sin(X) * torch.cos(Y) * torch.cos(Z)
    vy = -Minf * torch.cos(X) * torch.sin(Y) * torch.cos(Z)
    vz = torch.zeros_like(vx)
    rng = np.random.RandomState(seed)
    for _ in range(16):
        k = rng.randint(-4, 5, size=3).astype(np.float64)
        if not k.any():
            k[0] = 1.0
        a = rng.standard_normal(3)
        a -= k * (a @ k) / (k @ k)
        nrm = np.linalg.norm(a)
        if nrm < 1e-12:
            continue
        a *= 0.1 * Minf / (4.0 * nrm)
        ph = rng.uniform(0.0, 2.0 * np.pi)
        s = torch.sin(k[0] * X + k[1] * Y + k[2] * Z + ph)
        vx += a[0] * s
        vy += a[1] * s
        vz += a[2] * s
        del s
    rho = torch.ones_like(vx)
    e = (1.0 / gamma) / (gamma - 1.0) + 0.5 * rho * (vx * vx + vy * vy + vz * vz)
    return torch.stack([rho, rho * vx, rho * vy, rho * vz, e], dim=-1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ reference CPU arm
def run_reference_cpu(n: int, nsteps: int, warmup: int, threads: int):
    """Times the UNMODIFIED reference (oracle/_ref/hypar_ref: HyPar's own TimePreStep/TimeStep/TimePostStep
    loop, its own per-step wctime) on a n^3 sample of the C4 configuration. Returns (Mpt-stage/s, seconds/step)."""
    from hypar_b200 import cases
    from refrun import run_reference, ref_available
    if not ref_available("hypar_ref"):
        raise RuntimeError("oracle/_ref/hypar_ref is missing (build it where /root/reference exists: make -C oracle ref)")
    case = cases.ns3d_turbulence((n, n, n), "mapped")
    out = run_reference(case, "steps", [warmup + nsteps], exe="hypar_ref", threads=threads, timeout=3000)
    wct = [float(ln.split()[3]) for ln in out["stdout"].splitlines() if ln.startswith("STEP ")]
    wct = wct[warmup:]
    sec = sum(wct) / len(wct)
    return n ** 3 * NSTAGES / sec / 1e6, sec


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.ref_n
    t0 = time.time()
    val, sec = run_reference_cpu(n, args.steps, max(args.warmup, 1), threads)
    size, iproc = weak_grid(args.n, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C4 NavierStokes3D WENO5(mapped)+Rusanov+viscous RK4, {size[0]}x{size[1]}x{size[2]} periodic",
                   "sample": f"{n}^3 grid, same physics and scheme, one RK4 step per bench step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"{n}^3, {args.steps} RK4 steps after {max(args.warmup, 1)} warm-up, OMP_NUM_THREADS={threads}, "
                                   "HyPar's own per-step wctime"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
STRONG_IPROC = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}   # configs[3]: 512^3 over 1/2/4/8 GPUs


def workload_inputs(workload, size, iproc):
    if workload == "c4":
        s, b, ph, w, x = c4_inputs(size, iproc)
        return s, b, ph, w, x, "C4 NavierStokes3D WENO5(mapped)+Rusanov+viscous RK4", "periodic"
    if workload == "c4roe":
        # the configuration of the reference's own flagship CUDA run (Examples/3D/NavierStokes3D/DNS_IsotropicTurbulenceDecay_CUDA:
        # weno5 mapped, components, Roe, viscous, SSPRK3; its out.log: 256^3 on 64 V100, 82 ms per step)
        s, b, ph, w, x = c4_inputs(size, iproc, upwinding="roe", tstype="ssprk3")
        return s, b, ph, w, x, "C4-Roe NavierStokes3D WENO5(mapped)+Roe+viscous SSPRK3 (the reference's DNS_IsotropicTurbulenceDecay_CUDA setup)", "periodic"
    if workload == "c4char":
        # the same with characteristic-wise reconstruction (HyPar's default hyp_interp_type, ReadInputs.c:136)
        s, b, ph, w, x = c4_inputs(size, iproc, upwinding="roe", tstype="ssprk3", interp="characteristic")
        return s, b, ph, w, x, "C4-char NavierStokes3D characteristic WENO5(mapped)+Roe+viscous SSPRK3", "periodic"
    s, b, ph, w, x = c5_inputs(workload, size, iproc)
    label = ("C5a NavierStokes3D density sine wave, WENO5(JS)+Rusanov, inviscid, RK4" if workload == "c5a" else
             "C5b NavierStokes3D rising thermal bubble, WENO5(YC)+Rusanov, gravity (HB 2) source, SSPRK3")
    return s, b, ph, w, x, label, ("periodic" if workload == "c5a" else "slip walls")


class Run:
    """one workload on this rank's GPU: solver (+ NCCL transport when decomposed), synthetic field, device-timed steps"""

    def __init__(self, workload, size, iproc, rank, local_rank, world, overlap=True, use_fused=True):
        import torch
        from hypar_b200.solver import Solver
        self.torch, self.world, self.rank = torch, world, rank
        self.dev = torch.device("cuda", local_rank)
        self.size, self.iproc, self.workload = list(size), list(iproc), workload
        s, b, ph, w, x, self.label, self.bc = workload_inputs(workload, size, iproc)
        if world == 1:
            self.sv, self.stepper = Solver(s, b, ph, w, x, rank=0, device=local_rank, use_fused=use_fused), None
        else:
            from hypar_b200.multigpu import DistributedSolver
            self.stepper = DistributedSolver(s, b, ph, w, x, rank=rank, device=local_rank, overlap=overlap, use_fused=use_fused)
            self.sv = self.stepper.solver
        sv = self.sv
        self.g, self.nloc = sv.ghosts, sv.dim_local
        x_loc = [x[d][sv.is_global[d]:sv.is_global[d] + self.nloc[d]] for d in range(3)]
        # synthetic input, created on the device, staged into a pinned host array in HyPar's own layout
        self.u_host_t = torch.zeros(sv.npoints_local_wghosts * 5, dtype=torch.float64).pin_memory()
        self.u_host = self.u_host_t.numpy()
        fld = synth_field_torch(x_loc, self.dev) if workload in ("c4", "c4roe", "c4char") else synth_field_c5(workload, x_loc, self.dev)
        g, n = self.g, self.nloc
        self.u_host_t.view(n[2] + 2 * g, n[1] + 2 * g, n[0] + 2 * g, 5)[g:-g, g:-g, g:-g, :].copy_(fld)
        del fld
        torch.cuda.empty_cache()
        sv.set_solution(self.u_host)
        self.stream = torch.cuda.ExternalStream(sv.stream, device=self.dev)
        self.npts_global = float(size[0]) * size[1] * size[2]
        self.nstages = sv.nstages

    def barrier(self):
        torch = self.torch
        torch.cuda.synchronize()
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def one_step(self):
        if self.stepper is None:
            self.sv.TimeStep()
        else:
            self.stepper.time_step()

    def timed(self, fn, nsteps):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(nsteps):
            fn()
        e1.record(self.stream)
        self.sv.synchronize()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def device_loop(self, steps, warmup):
        """W untimed + K timed device-resident steps; returns (ms total, Mpoint-RK-stage/s, CFL afterwards)"""
        import numpy as np
        for _ in range(warmup):
            self.one_step()
        self.sv.synchronize()
        ms = self.timed(self.one_step, steps)
        cfl = self.sv.dev_ComputeCFL()
        if not np.isfinite(cfl) or cfl <= 0 or cfl > 10:
            raise RuntimeError(f"solution blew up during the bench of {self.workload} (CFL = {cfl})")
        return ms, self.npts_global * self.nstages * steps / (ms * 1e-3) / 1e6, cfl

    def close(self):
        self.sv.close()
        del self.u_host, self.u_host_t
        self.torch.cuda.empty_cache()


def sub_record(workload, size, iproc, rank, local_rank, world, steps, warmup, overlap, use_fused=True):
    """a second workload measured in the same process after the main one (device-resident loop only)"""
    R = Run(workload, size, iproc, rank, local_rank, world, overlap=overlap, use_fused=use_fused)
    ms, val, cfl = R.device_loop(steps, warmup)
    rec = {"workload": f"{R.label}, {size[0]}x{size[1]}x{size[2]} {R.bc}", "iproc": list(iproc),
           "points_per_gpu": "x".join(str(v) for v in R.nloc), "value": val, "unit": UNIT, "ms_per_step": ms / steps,
           "steps": steps, "warmup": warmup, "rk_stages_per_step": R.nstages, "cfl": cfl}
    R.close()
    return rec


def gpu_arm(args):
    import numpy as np
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # one process per GPU: keep the host staging buffers on the GPU's own NUMA node (matters for `e2e` at 4-8 GPUs)
    from hypar_b200.multigpu import bind_to_gpu_numa
    numa_cpus = bind_to_gpu_numa(local_rank) if not args.no_numa_bind else None

    overlap = not args.serial_halo
    size, iproc = weak_grid(args.n, world)
    R = Run(args.workload, size, iproc, rank, local_rank, world, overlap=overlap)
    sv, stepper = R.sv, R.stepper
    wl_label, wl_bc = R.label, R.bc
    g, nloc = R.g, R.nloc
    u_host_t, u_host = R.u_host_t, R.u_host
    stream, npts_global, nstages = R.stream, R.npts_global, R.nstages
    barrier, one_step, timed = R.barrier, R.one_step, R.timed

    # ---- device-resident loop
    for _ in range(args.warmup):
        one_step()
    sv.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sv.profile_enable(True)
    l0 = sv.kernel_launches
    ms = timed(one_step, args.steps)
    launches = sv.kernel_launches - l0
    prof = sv.profile_query()
    sv.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    value = npts_global * nstages * args.steps / (ms * 1e-3) / 1e6

    # sanity: the field must still be finite after the timed steps
    cfl = sv.dev_ComputeCFL()
    if not np.isfinite(cfl) or cfl <= 0 or cfl > 10:
        raise RuntimeError(f"solution blew up during the bench (CFL = {cfl})")

    # ---- end to end through the host-array entry point (single GPU: hpb_TimeIntegrate; multi GPU: the
    # distributed stepper's host-array step): H2D of u + step + D2H of u inside the timed region
    # Two figures: (i) `sync`: the blocking call, copy - step - copy one after the other; (ii) the headline `value`: the
    # same call in its enqueue-only form (hpb_TimeIntegrateAsync / hpb_pipe_*) over a sequence of independent fields
    # (an ensemble: every step takes a field from pinned host memory and returns its result to pinned host memory) --
    # the H2D copy of field k+1 and the D2H copy of field k-1 run on copy streams under the step of field k.
    nbytes = u_host.nbytes
    e2e_val = ms_e2e = sync_val = ms_sync = None
    e2e_steps = sync_steps = 0
    if not args.no_e2e:
        sync_steps = max(1, min(args.steps, 3))

        def e2e_step():
            if stepper is None:
                sv.TimeIntegrate(u_host, 1, sv.time)
            else:
                stepper.time_integrate_host(u_host, 1)

        u_in_t = torch.empty_like(u_host_t).pin_memory()     # the synthetic input, kept: every pipelined step consumes it
        u_in_t.copy_(u_host_t)
        u_in = u_in_t.numpy()
        e2e_step()
        ms_sync = timed(e2e_step, sync_steps)
        sync_val = npts_global * nstages * sync_steps / (ms_sync * 1e-3) / 1e6
        if not np.isfinite(u_host).all():
            raise RuntimeError("non-finite values in the host solution after the end-to-end steps")

        # the pipeline has a fixed fill / drain cost (the first H2D and the last D2H, ~100 ms each at 512^3, cannot hide behind a
        # step): 40 steps amortise it to ~5 ms per step (16 steps of round 1 left 12.5 ms per step in the figure)
        e2e_steps = max(args.steps, 40)
        u_out = u_host                               # results land in the first pinned array

        def e2e_submit():
            if stepper is None:
                sv.TimeIntegrateAsync(u_in, u_out, 1, 0.0)
            else:
                stepper.time_integrate_host_async(u_in, u_out, 1)

        def e2e_all():
            for _ in range(e2e_steps):
                e2e_submit()
            sv.pipe_join()                           # the closing event (on the solver's stream) comes after the last D2H

        # the pipelined result must be the blocking call's result for the same input, bit for bit
        u_host_t.copy_(u_in_t)
        e2e_step()
        u_chk = sv.interior(u_host).copy()
        u_host_t.zero_()
        e2e_submit(); e2e_submit()
        sv.pipe_wait()
        if not np.array_equal(sv.interior(u_out), u_chk):
            raise RuntimeError("pipelined end-to-end step differs from the blocking hpb_TimeIntegrate result")
        del u_chk
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        e2e_all()
        e1.record(stream)
        sv.pipe_wait()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
        if dist is not None:
            tt = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_e2e = float(tt.item())
        e2e_val = npts_global * nstages * e2e_steps / (ms_e2e * 1e-3) / 1e6
        if not np.isfinite(u_out).all():
            raise RuntimeError("non-finite values in the host solution after the pipelined end-to-end steps")

    # ---- sub-records (device-resident loop only), measured by all ranks after the main workload has released its memory:
    #   strong : configs[3] as BASELINE.json states it -- C4 at 512^3 GLOBAL, split over the N GPUs like HyPar would
    #            (iproc (1,1,2) / (1,2,2) / (2,2,2), Initialize.c:66-98, MPIPartition1D.c) -- with its efficiency against
    #            the 1-GPU rate of this run's per-GPU block size
    #   c5b    : configs[4] -- rising thermal bubble with the gravity source, WENO5-YC, SSPRK3, 512^3 per GPU (1024^3 at 8)
    tma_launches = sv.tma_launches
    stage_fusion = sv.stage_fusion_active          # (the solver is closed before the line is assembled)
    fp64_peak = sv.fp64_issue_peak() if rank == 0 else None      # measured live on this GPU (hpb_fp64_issue_peak)
    comm_msgs, comm_bytes = sv.comm_stats() if stepper is not None else (0, 0)
    sub = {}
    if not args.no_sub and args.workload == "c4":
        R.close()
        del u_host_t, u_host
        if not args.no_e2e:
            del u_in_t, u_in, u_out
        torch.cuda.empty_cache()
        if world > 1:
            sub["strong"] = sub_record("c4", [args.n] * 3, STRONG_IPROC[world], rank, local_rank, world,
                                       max(args.steps, 10), args.warmup, overlap)
            sub["strong"]["scaling"] = "strong"
            sub["strong"]["note"] = (f"{args.n}^3 global on {world} GPUs; compare value with the N=1 line's (same global grid): "
                                     "strong-scaling efficiency = value / (N x value at N=1)")
        sub["c5b"] = sub_record("c5b", size, iproc, rank, local_rank, world, args.steps, args.warmup, overlap)
        sub["c5b"]["scaling"] = "weak"
        if world == 1:
            sub["c4roe"] = sub_record("c4roe", size, iproc, rank, local_rank, world, args.steps, args.warmup, overlap)
            sub["c4char"] = sub_record("c4char", size, iproc, rank, local_rank, world, args.steps, args.warmup, overlap)
            # the reference-exact path (use_fused = 0: one thread per interface, no FMA contraction, bit-identical to the
            # reference) on the same physics, 256^3 (its interface / weight scratch is sized for parity runs, not for 512^3)
            sub["c4_exact_path"] = sub_record("c4", [256] * 3, iproc, rank, local_rank, world, 3, 2, overlap, use_fused=False)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per-launch CUDA-event time, this rank)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    npts_local = float(nloc[0]) * nloc[1] * nloc[2]
    sweeps = {k: v for k, v in prof.items() if k.startswith("sweep") and v[1] > 0}
    dom = max(sweeps, key=lambda k: sweeps[k][0] / sweeps[k][1])
    dom_ms = sweeps[dom][0] / sweeps[dom][1]
    achieved = SWEEP_BYTES[dom] * npts_local / (dom_ms * 1e-3) / 1e9
    total_prof = sum(v[0] for v in prof.values())
    # DRAM traffic and FP64-pipe activity of the same kernel from the committed ncu --set full capture (512^3 per launch)
    traffic, fp64_pct, ncu_src, fp64_instr = None, None, None, None
    try:
        prof_ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_sweep_512.json")))
        if tuple(nloc) == (512, 512, 512) and dom in prof_ncu and tma_launches > 0:
            traffic = float(prof_ncu[dom]["dram_bytes_read"] + prof_ncu[dom]["dram_bytes_write"]) / 1e9
            fp64_pct = prof_ncu[dom]["fp64_pipe_active_pct"]
            fp64_instr = prof_ncu[dom].get("fp64_thread_instr")
            ncu_src = "profiles/ncu_sweep_512.json"
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_unit": "GB per launch (dram read + write, ncu)", "fp64_pipe_active_pct_ncu": fp64_pct,
        "ncu_source": ncu_src, "algorithmic_GB_per_launch": SWEEP_BYTES[dom] * npts_local / 1e9,
        "peak_source": peak_src, "ms_per_launch": dom_ms,
        "bytes_per_point": SWEEP_BYTES[dom],
        "share_of_step": {k: (v[0] / total_prof if total_prof > 0 else None) for k, v in prof.items() if v[1] > 0},
        "whole_stage": {"bytes_per_point_stage": STAGE_BYTES,
                        "achieved": STAGE_BYTES * npts_local * nstages * args.steps / (ms * 1e-3) / 1e9,
                        "frac": STAGE_BYTES * npts_local * nstages * args.steps / (ms * 1e-3) / 1e9 / peak},
        # the binding roofline: FP64 thread-instructions of the same launch (DFMA + DMUL + DADD executed, from the committed
        # ncu source-page capture of this kernel at 512^3: profiles/ncu_sweep_512.json) over its live CUDA-event time, against
        # the FP64 issue peak measured live on this GPU
        "fp64": (None if not (fp64_instr and fp64_peak) else {
            "bound": "fp64 issue", "achieved": fp64_instr / (dom_ms * 1e-3) / 1e12, "peak": fp64_peak / 1e12, "unit": "T thread-instr/s",
            "frac": fp64_instr / (dom_ms * 1e-3) / fp64_peak, "fp64_thread_instr_per_launch": fp64_instr,
            "fp64_thread_instr_per_cell": fp64_instr / npts_local,
            "peak_source": "hpb_fp64_issue_peak: DMUL chains, 8 CTAs x 256 threads per SM, measured in this run"}),
        "note": "FP64-issue-bound kernel (DESIGN.md): the HBM fraction is reported as the metric demands; the FP64 pipe "
                "is the binding unit (fp64_pipe_active_pct_ncu); traffic exceeds the algorithmic bytes by the 8 "
                "derivative scalars (64 B/point) the viscous flux reads. sweep_fused = the last direction's sweep that also "
                "forms the next RK stage solution: 200 B/point, the whole-stage figure of SURVEY 8(d)",
    }

    # ---- CPU baseline: the reference itself on a bounded sample
    cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "unavailable"}
    if not args.no_cpu and args.workload == "c4":
        try:
            threads = os.cpu_count() or 1
            v, sec = run_reference_cpu(args.cpu_n, 2, 1, threads)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"{args.cpu_n}^3 grid, same physics/scheme, 2 RK4 steps after 1 warm-up, "
                             f"{sec:.2f} s/step, OMP_NUM_THREADS={threads} (unmodified HyPar sources, oracle/_ref)"}
        except Exception as ex:  # the baseline is a report, not a dependency of the GPU number
            cpu["sample"] = f"failed: {ex}"

    # ---- second baseline: the reference's OWN CUDA path (its *_GPU.cu kernels, unmodified, re-targeted from sm_70 to
    # sm_100a: oracle/Makefile `refgpu`) on this GPU, HyPar's own per-iteration wall clock; 128^3 (its weight arrays alone
    # are 12 x 3 x 5 doubles per point: 512^3 does not fit one GPU, and its 32-bit indices overflow beyond ~329^3)
    ref_gpu = None
    if not args.no_cpu and args.workload == "c4" and world == 1 and sub:
        try:
            import numpy as _np
            from hypar_b200 import cases as _cases
            from refgpu_bench import EXE as _REFGPU, run_ref_gpu
            if os.access(_REFGPU, os.X_OK):
                n_rg = args.ref_gpu_n
                rr = run_ref_gpu(_cases.ns3d_turbulence((n_rg, n_rg, n_rg), "mapped"), 8, timeout=600)
                w = rr["wctime"][2:] if len(rr["wctime"]) > 4 else rr["wctime"]
                sec_rg = float(_np.median(w))
                ref_gpu = {"value": n_rg ** 3 * NSTAGES / sec_rg / 1e6, "unit": UNIT, "grid": f"{n_rg}^3", "s_per_step": sec_rg,
                           "kind": "reference CUDA path (use_gpu yes; unmodified sources built with -DHAVE_CUDA for sm_100a, "
                                   "oracle/_ref/hypar_ref_gpu), same physics and scheme, HyPar's own wctime, median of 6 steps"}
        except Exception as ex:
            ref_gpu = {"value": None, "error": str(ex)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{wl_label}, {size[0]}x{size[1]}x{size[2]} {wl_bc}",
                   "points_per_gpu": f"{nloc[0]}x{nloc[1]}x{nloc[2]}", "iproc": iproc, "rk_stages_per_step": nstages,
                   "halo": (None if stepper is None else (
                       "in-library ncclSend/ncclRecv on a communication stream, overlapped: u faces under the full-array RK "
                       "update, Q-derivative faces of dims 1.. under the x-sweep (hpb_TimeStepsDistributed, no Python in the step)"
                       if stepper.overlap else "in-library ncclSend/ncclRecv, serial (hpb_TimeStepsDistributed)")),
                   "stage_fusion": ("the last sweep of an RK stage also writes the next stage solution (bit-identical to the unfused "
                                    "schedule; HPB_STAGE_FUSION=0 turns it off)" if stage_fusion else "off"),
                   "l2": "working set (5.6 GB per array) >> L2, no flush needed",
                   "host_numa_bind": (f"{len(numa_cpus)} GPU-local CPUs" if numa_cpus else "none")},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                "ms_per_step": (ms_e2e / e2e_steps if e2e_steps else None), "steps": e2e_steps,
                "api": ("hpb_TimeIntegrateAsync(pinned host u_in -> pinned host u_out, 1 step)" if stepper is None
                        else "DistributedSolver.time_integrate_host_async") + ", one call per step over a sequence of "
                       "independent fields; the copies of neighbouring steps overlap the step (copy streams)",
                "sync": {"value": sync_val, "ms_per_step": (ms_sync / sync_steps if sync_steps else None), "steps": sync_steps,
                         "api": "hpb_TimeIntegrate(host u, 1 step), blocking: copy, step, copy in sequence"
                                if stepper is None else "DistributedSolver.time_integrate_host, blocking"}},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "cfl": cfl,
        "halo_traffic": (None if stepper is None else {"messages_sent_rank0": comm_msgs, "bytes_sent_rank0": comm_bytes}),
        "strong": sub.get("strong"), "c5b": sub.get("c5b"), "c4roe": sub.get("c4roe"), "c4char": sub.get("c4char"), "c4_exact_path": sub.get("c4_exact_path"), "ref_gpu_baseline": ref_gpu,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=512, help="points per dimension per GPU")
    ap.add_argument("--cpu-n", type=int, default=64, help="grid of the cpu_baseline sample")
    ap.add_argument("--ref-n", type=int, default=64, help="grid of the --impl reference sample")
    ap.add_argument("--ref-gpu-n", type=int, default=128, help="grid of the ref_gpu_baseline run (the reference's own CUDA path)")
    ap.add_argument("--workload", default="c4", choices=["c4", "c4roe", "c4char", "c5a", "c5b"],
                    help="c4 (default): the configuration BASELINE.json's metric is quoted on; c5a / c5b: configs[4] "
                         "(density sine wave / rising thermal bubble with gravity), 1024^3 at 8 GPUs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end legs (profiler runs of the device loop only)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the GPU's NUMA-local CPUs")
    ap.add_argument("--serial-halo", action="store_true", help="multi-GPU: pack - exchange - unpack in sequence instead of "
                    "the overlapped schedule (hpb_set_overlap 0); same results bit for bit")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (C4 512^3 strong scaling at N > 1, C5b)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()

"""Synthetic cases for the five BASELINE.json configurations.

Each builder returns a ``Case``: the same information a HyPar run directory holds
(solver.inp / boundary.inp / physics.inp / weno.inp / initial.inp), as Python data.
``Case.write(dir)`` materialises the directory in the reference's own formats, so the
reference executable and this library consume identical inputs.

Initial conditions follow the reference's example generators:
  C1  Examples/1D/LinearAdvection/SineWave/aux/init.c
  C2  Examples/1D/Euler1D/SodShockTube/aux/init.c
  C3  Examples/2D/NavierStokes2D/InviscidVortexConvection/aux/exact.c:65-92
  C4  Taylor-Green vortex + deterministic solenoidal Fourier modes (no FFTW available)
  C5a Examples/3D/NavierStokes3D/DensitySineWave/aux/exact.C:54-81
  C5b Examples/3D/NavierStokes3D/RisingThermalBubble_Config1/aux/init.c:108-133
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import hypario


@dataclasses.dataclass
class Case:
    name: str
    solver: Dict[str, object]
    boundary: List[dict]
    physics: Dict[str, object]
    weno: Optional[Dict[str, object]]
    x: List[np.ndarray]            # global coordinates per dimension
    u0: np.ndarray                 # shape (N_{nd-1}, ..., N_0, nvars)
    muscl: Optional[Dict[str, object]] = None      # muscl.inp (epsilon, limiter) for muscl2 / muscl3
    advection_field: Optional[np.ndarray] = None   # LinearADR `advection_filename advection`: shape (N_{nd-1},...,N_0, ndims*nvars)
    glm_gee: Optional[Dict[str, object]] = None    # glm_gee.inp (ee_mode yeps | yyt) for time_scheme glm-gee
    lusolver: Optional[Dict[str, object]] = None   # lusolver.inp (maxiter, atol, rtol, evaluate_norm): compact schemes across ranks

    @property
    def ndims(self) -> int:
        return int(self.solver["ndims"])

    @property
    def nvars(self) -> int:
        return int(self.solver["nvars"])

    @property
    def dims(self) -> List[int]:
        return [int(v) for v in self.solver["size"]]

    def write(self, d: str) -> None:
        os.makedirs(d, exist_ok=True)
        hypario.write_keyword_file(os.path.join(d, "solver.inp"), self.solver)
        hypario.write_boundary_inp(os.path.join(d, "boundary.inp"), self.boundary)
        hypario.write_keyword_file(os.path.join(d, "physics.inp"), self.physics)
        if self.weno is not None:
            hypario.write_keyword_file(os.path.join(d, "weno.inp"), self.weno)
        if self.muscl is not None:
            hypario.write_keyword_file(os.path.join(d, "muscl.inp"), self.muscl)
        if self.glm_gee is not None:
            hypario.write_keyword_file(os.path.join(d, "glm_gee.inp"), self.glm_gee)
        if self.lusolver is not None:
            hypario.write_keyword_file(os.path.join(d, "lusolver.inp"), self.lusolver)
        ipt = str(self.solver.get("ip_file_type", "binary"))
        if self.advection_field is not None:           # same layout and flavour as initial.inp (ReadArray.c:173-256)
            hypario.write_initial(os.path.join(d, "advection.inp"), self.x, self.advection_field, ipt)
        hypario.write_initial(os.path.join(d, "initial.inp"), self.x, self.u0, ipt)


def from_directory(path: str) -> Case:
    """An existing HyPar run directory (solver.inp, boundary.inp, [physics.inp], [weno.inp], [muscl.inp], initial.inp in the
    flavour solver.inp names, [<advection_filename>.inp]) as a Case -- the same reading Solver.from_directory does."""
    s = hypario.read_solver_inp(os.path.join(path, "solver.inp"))
    nd, nv = int(s["ndims"]), int(s["nvars"])
    b = hypario.read_boundary_inp(os.path.join(path, "boundary.inp"), nd, nv)
    vk = {"gravity": 3 if s["model"] == "navierstokes3d" else 2 if s["model"] == "navierstokes2d" else 1,
          "advection": nd * nv, "diffusion": nd * nv}
    pf = os.path.join(path, "physics.inp")
    ph = hypario.read_keyword_file(pf, vector_keys=vk) if os.path.exists(pf) else {}
    for k in ("advection", "diffusion", "gravity"):
        if k in ph:
            ph[k] = [float(v) for v in (ph[k] if isinstance(ph[k], (list, tuple)) else [ph[k]])]
    wf, mf, gf = os.path.join(path, "weno.inp"), os.path.join(path, "muscl.inp"), os.path.join(path, "glm_gee.inp")
    w = hypario.read_keyword_file(wf) if os.path.exists(wf) else None
    mu = hypario.read_keyword_file(mf) if os.path.exists(mf) else None
    gg = hypario.read_keyword_file(gf) if os.path.exists(gf) else None
    lf = os.path.join(path, "lusolver.inp")
    lu = hypario.read_keyword_file(lf) if os.path.exists(lf) else None
    ipt = str(s.get("ip_file_type", "ascii"))
    x, u0 = hypario.read_initial(os.path.join(path, "initial.inp"), s["size"], nv, ipt)
    case = Case(name=os.path.basename(os.path.normpath(path)), solver=s, boundary=b, physics=ph, weno=w, x=x, u0=u0, muscl=mu,
                glm_gee=gg, lusolver=lu)
    if str(ph.get("advection_filename", "none")) != "none":
        fn = os.path.join(path, str(ph["advection_filename"]) + ".inp")
        if os.path.exists(fn):
            case.advection_field = hypario.read_initial(fn, s["size"], nd * nv, ipt)[1]
    return case


def write_ensemble(d: str, sims: Sequence[Case]) -> None:
    """A run directory of the reference's ensemble driver (simulation.inp + one solver.inp whose size / iproc hold one vector
    per simulation, initial_<n>.inp, shared boundary / physics / weno files; InitialSolution.c:36-43)."""
    os.makedirs(d, exist_ok=True)
    nsims = len(sims)
    width = len(str(nsims - 1)) if nsims > 1 else 1       # (int) log10(nsims) + 1 digits
    s = dict(sims[0].solver)
    s["size"] = [int(v) for c in sims for v in c.solver["size"]]
    s["iproc"] = [int(v) for c in sims for v in c.solver.get("iproc", [1] * c.ndims)]
    hypario.write_keyword_file(os.path.join(d, "simulation.inp"), {"nsims": nsims})
    hypario.write_keyword_file(os.path.join(d, "solver.inp"), s)
    hypario.write_boundary_inp(os.path.join(d, "boundary.inp"), sims[0].boundary)
    hypario.write_keyword_file(os.path.join(d, "physics.inp"), sims[0].physics)
    if sims[0].weno is not None:
        hypario.write_keyword_file(os.path.join(d, "weno.inp"), sims[0].weno)
    if sims[0].muscl is not None:
        hypario.write_keyword_file(os.path.join(d, "muscl.inp"), sims[0].muscl)
    ipt = str(s.get("ip_file_type", "binary"))
    for n, c in enumerate(sims):
        tag = f"_{n:0{width}d}" if nsims > 1 else ""
        hypario.write_initial(os.path.join(d, f"initial{tag}.inp"), c.x, c.u0, ipt)
        if c.advection_field is not None:
            hypario.write_initial(os.path.join(d, f"advection{tag}.inp"), c.x, c.advection_field, ipt)


def ensemble(name: str, n_iter: int = 4) -> List[Case]:
    """Named ensembles (the reference's Examples/*_Ensemble and LaSDI/*/training_data): the same equations on several grids,
    one time step for all -- simulation 0's (ReadInputs.c:321)."""
    if name == "vortex3":
        sims = [ns2d_vortex((32, 16), "js"), ns2d_vortex((16, 32), "js"), ns2d_vortex((24, 24), "js")]
    elif name == "burgers2":
        sims = [burgers_nd((96,), "mapped"), burgers_nd((64,), "mapped")]
    elif name == "linadvvar2":
        sims = [linear_advection_varying((24, 20), "z"), linear_advection_varying((16, 28), "z")]
    elif name == "sod2":
        sims = [euler1d_sod(101, "js"), euler1d_sod(81, "js")]
    elif name == "turb12":                # 12 simulations: two-digit file indices
        sims = [ns3d_turbulence((8 + (k % 3) * 2, 8, 8 + (k % 2) * 2), "mapped") for k in range(12)]
    else:
        raise KeyError(name)
    for c in sims:
        c.solver.update({"dt": sims[0].solver["dt"], "n_iter": n_iter, "screen_op_iter": 2, "file_op_iter": n_iter,
                         "op_overwrite": "yes", "op_file_format": "binary"})
        c.name = f"ens_{name}_" + c.name
    return sims


def _solver(ndims, nvars, size, model, *, iproc=None, ts="rk", tstype="44", dt=1e-3,
            interp="components", par_type="nonconservative-1stage", par_scheme="2",
            n_iter=1, scheme="weno5") -> Dict[str, object]:
    return {
        "ndims": ndims, "nvars": nvars, "size": list(size),
        "iproc": list(iproc) if iproc is not None else [1] * ndims,
        "ghost": 3, "n_iter": n_iter, "restart_iter": 0,
        "time_scheme": ts, "time_scheme_type": tstype,
        "hyp_space_scheme": scheme, "hyp_flux_split": "no", "hyp_interp_type": interp,
        "par_space_type": par_type, "par_space_scheme": par_scheme,
        "dt": float(dt), "conservation_check": "no",
        "screen_op_iter": 1, "file_op_iter": 1000000,
        "ip_file_type": "binary", "input_mode": "serial", "output_mode": "serial",
        "op_file_format": "binary", "op_overwrite": "yes", "model": model,
    }


def with_time_scheme(case: "Case", time_scheme: str, tstype: str = " ") -> "Case":
    """the same case with another time integrator (solver.inp time_scheme / time_scheme_type)"""
    case.solver["time_scheme"], case.solver["time_scheme_type"] = time_scheme, tstype
    if time_scheme == "euler":
        del case.solver["time_scheme_type"]          # no type keyword for forward Euler (ReadInputs.c:131)
    case.name += f"_{time_scheme}{'' if time_scheme == 'euler' else tstype}"
    return case


def with_glmgee(case: "Case", tstype: str, ee_mode: Optional[str] = None) -> "Case":
    """the same case advanced by a general linear method with global error estimation (solver.inp time_scheme glm-gee,
    time_scheme_type 23 | 24 | 25i | 35 | exrk2a | rk32g1 | rk285ex; glm_gee.inp ee_mode yeps (default) | yyt)"""
    case = with_time_scheme(case, "glm-gee", tstype)
    if ee_mode is not None:
        case.glm_gee = {"ee_mode": ee_mode}
        case.name += "_" + ee_mode
    return case


def glmgee(base: str, tstype: str, ee_mode: Optional[str] = None, **kwargs) -> "Case":
    """builder form of with_glmgee (fixtures name a builder and its keyword arguments): base = name of a builder here"""
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in kwargs.items()}
    return with_glmgee(globals()[base](**kw), tstype, ee_mode)


def with_characteristic(case: "Case") -> "Case":
    """the same case with characteristic-based reconstruction (solver.inp hyp_interp_type)"""
    case.solver["hyp_interp_type"] = "characteristic"
    case.name += "_char"
    return case


def with_muscl(case: "Case", scheme: str, epsilon: float = 1e-3, limiter: str = "gmm") -> "Case":
    """the same case with a MUSCL reconstruction (muscl2: limiter; muscl3: epsilon) and its muscl.inp"""
    case.solver["hyp_space_scheme"] = scheme
    case.muscl = {"epsilon": float(epsilon), "limiter": limiter}
    case.name += f"_{scheme}_{limiter if scheme == 'muscl2' else epsilon}"
    return case


def with_sponge(case: "Case", dim: int, face: int, xstart: float, xend: float, values) -> "Case":
    """the same case with a sponge zone (BCSponge.c): an interior box spanning [xstart, xend] along `dim` (the whole
    domain in the other dimensions) in which the solution is relaxed towards `values`"""
    nd = case.ndims
    z = {"type": "sponge", "dim": dim, "face": face,
         "xmin": [float(xstart) if k == dim else -1e3 for k in range(nd)],
         "xmax": [float(xend) if k == dim else 1e3 for k in range(nd)],
         "values": [float(v) for v in values]}
    case.boundary = list(case.boundary) + [z]
    case.name += f"_sponge{dim}{'p' if face > 0 else 'm'}"
    return case


def _sfx(scheme: str) -> str:
    return "" if scheme == "weno5" else "_" + scheme


def weno_inp(kind: str = "js", eps: float = 1e-6, no_limiting: int = 0) -> Dict[str, object]:
    """kind in {js, mapped, z, yc} (WENOInitialize.c:62-96); the hybridisation parameters of hcweno5 ride on the
    kind as "+rc<value>" / "+xi<value>" (e.g. "mapped+rc0.2+xi0.01"; defaults 0.3, 0.001, WENOInitialize.c:57-58)."""
    kind, *opts = kind.split("+")
    w = {
        "mapped": int(kind == "mapped"), "borges": int(kind == "z"), "yc": int(kind == "yc"),
        "no_limiting": no_limiting, "epsilon": float(eps), "p": 2.0, "rc": 0.3, "xi": 0.001,
    }
    for o in opts:
        w[o[:2]] = float(o[2:])
    return w


def _zones(ndims, kind_per_face, lo, hi, wall_velocity=None):
    """One zone per (dim, face); extents as in the reference's examples: the normal
    extent is degenerate (0 0), tangential extents span the domain."""
    zones = []
    for d in range(ndims):
        for face in (1, -1):
            kind = kind_per_face[(d, face)] if isinstance(kind_per_face, dict) else kind_per_face
            xmin = [0.0 if k == d else float(lo[k]) for k in range(ndims)]
            xmax = [0.0 if k == d else float(hi[k]) for k in range(ndims)]
            z = {"type": kind, "dim": d, "face": face, "xmin": xmin, "xmax": xmax}
            if kind in ("slip-wall", "noslip-wall"):
                z["wall_velocity"] = list(wall_velocity or [0.0] * ndims)
            zones.append(z)
    return zones


# ------------------------------------------------------------------------------------- C1
def linear_advection_sine(n: int = 1024, weno: str = "js", dt: float = 5e-4,
                          diffusion: float = 0.0, par_scheme: str = "2", scheme: str = "weno5") -> Case:
    x = np.arange(n, dtype=np.float64) / n
    u = np.sin(2.0 * np.pi * x).reshape(n, 1)
    phys: Dict[str, object] = {"advection": 1.0}
    if diffusion != 0.0:
        phys["diffusion"] = float(diffusion)
    return Case(
        name=f"c1_linadv_{n}_{weno}" + _sfx(scheme),
        solver=_solver(1, 1, [n], "linear-advection-diffusion-reaction", dt=dt, par_scheme=par_scheme, scheme=scheme),
        boundary=_zones(1, "periodic", [-1e3], [1e3]),
        physics=phys, weno=weno_inp(weno), x=[x], u0=u)


def linear_advection_nd(n: Sequence[int] = (32, 24), weno: str = "js", advection=None, diffusion=None,
                        par_scheme: str = "2", tstype: str = "44", scheme: str = "weno5", iproc=None) -> Case:
    """Scalar linear advection(-diffusion) of a smooth periodic field in 2-D or 3-D
    (Examples/2D/LinearAdvection/SineWave, Examples/3D/LinearAdvection): u = prod_d sin(2 pi x_d) + 0.5."""
    nd = len(n)
    xs = [np.arange(n[d], dtype=np.float64) / n[d] for d in range(nd)]
    grids = np.meshgrid(*[xs[d] for d in reversed(range(nd))], indexing="ij")
    X = list(reversed(grids))
    u = 0.5 + np.prod([np.sin(2.0 * np.pi * X[d] + 0.3 * d) for d in range(nd)], axis=0)
    adv = list(advection) if advection is not None else [1.0, -0.5, 0.25][:nd]
    phys: Dict[str, object] = {"advection": adv}
    if diffusion is not None:
        phys["diffusion"] = list(diffusion)
    return Case(
        name=f"linadv{nd}d_{'x'.join(str(v) for v in n)}_{weno}" + ("_diff" if diffusion is not None else "") + _sfx(scheme),
        solver=_solver(nd, 1, n, "linear-advection-diffusion-reaction", ts="rk", tstype=tstype, dt=0.2 / max(n),
                       par_scheme=par_scheme, scheme=scheme, iproc=iproc),
        boundary=_zones(nd, "periodic", [-1e3] * nd, [1e3] * nd),
        physics=phys, weno=weno_inp(weno), x=xs, u0=u[..., None])


def linear_advection_varying(n: Sequence[int] = (96,), weno: str = "js", tstype: str = "44", scheme: str = "weno5",
                             iproc=None, periodic: bool = True) -> Case:
    """Scalar advection by a spatially varying velocity field read from a file (physics.inp `advection_filename`,
    Examples/1D/LinearAdvection/SineWave_NonConstantAdvection, Examples/2D/.../SineWave_NonConstantAdvection): the field
    changes sign, so LinearADRUpwind.c:56-82 takes all three of its branches."""
    base = linear_advection_nd(n, weno, tstype=tstype, scheme=scheme, iproc=iproc) if len(n) > 1 else None
    nd = len(n)
    xs = [np.arange(n[d], dtype=np.float64) / n[d] for d in range(nd)]
    grids = np.meshgrid(*[xs[d] for d in reversed(range(nd))], indexing="ij")
    X = list(reversed(grids))
    u = 0.5 + np.prod([np.sin(2.0 * np.pi * X[d] + 0.3 * d) for d in range(nd)], axis=0)
    a = np.stack([0.3 + np.cos(2.0 * np.pi * X[d] + 0.7 * d) * np.prod([np.cos(2.0 * np.pi * X[k]) for k in range(nd) if k != d] or [1.0], axis=0)
                  for d in range(nd)], axis=-1)
    kind = "periodic" if periodic else "extrapolate"
    case = Case(
        name=f"linadvvar{nd}d_{'x'.join(str(v) for v in n)}_{weno}_{kind[:3]}" + _sfx(scheme),
        solver=_solver(nd, 1, n, "linear-advection-diffusion-reaction", ts="rk", tstype=tstype, dt=0.1 / max(n),
                       scheme=scheme, iproc=iproc),
        boundary=_zones(nd, kind, [-1e3] * nd, [1e3] * nd),
        physics={"advection_filename": "advection"}, weno=weno_inp(weno), x=xs, u0=u[..., None])
    case.advection_field = np.ascontiguousarray(a)
    del base
    return case


def burgers_nd(n: Sequence[int] = (64,), weno: str = "js", tstype: str = "ssprk3", scheme: str = "weno5", iproc=None) -> Case:
    """Inviscid Burgers equation (model `burgers`, Examples/1D/Burgers/SineWave, Examples/2D/Burgers, 3D/Burgers): a
    periodic sine field with sign changes (both upwind branches and the local Lax-Friedrichs one are taken)."""
    nd = len(n)
    xs = [np.arange(n[d], dtype=np.float64) / n[d] for d in range(nd)]
    grids = np.meshgrid(*[xs[d] for d in reversed(range(nd))], indexing="ij")
    X = list(reversed(grids))
    u = 0.2 + np.prod([np.sin(2.0 * np.pi * X[d] + 0.4 * d) for d in range(nd)], axis=0)
    return Case(
        name=f"burgers{nd}d_{'x'.join(str(v) for v in n)}_{weno}" + _sfx(scheme),
        solver=_solver(nd, 1, n, "burgers", ts="rk", tstype=tstype, dt=0.2 / max(n), scheme=scheme, iproc=iproc),
        boundary=_zones(nd, "periodic", [-1e3] * nd, [1e3] * nd),
        physics={}, weno=weno_inp(weno), x=xs, u0=u[..., None])


# ------------------------------------------------------------------------------------- C2
def euler1d_sod(n: int = 201, weno: str = "js", interp: str = "characteristic",
                upwinding: str = "roe", tstype: str = "ssprk3", scheme: str = "weno5",
                gravity: float = 0.0, gravity_type: int = 0) -> Case:
    x = np.arange(n, dtype=np.float64) / (n - 1)
    gamma = 1.4
    rho = np.where(x < 0.5, 1.0, 0.125)
    p = np.where(x < 0.5, 1.0, 0.1)
    v = np.zeros_like(x)
    u = np.stack([rho, rho * v, p / (gamma - 1.0) + 0.5 * rho * v * v], axis=-1)
    return Case(
        name=f"c2_sod_{n}_{weno}_{interp}_{upwinding}" + _sfx(scheme)
             + (f"_grav{gravity_type}" if gravity != 0.0 else ""),
        solver=_solver(1, 3, [n], "euler1d", ts="rk", tstype=tstype, dt=2.5e-3 * (201.0 / n),
                       interp=interp, scheme=scheme),
        boundary=_zones(1, "extrapolate", [-1e3], [1e3]),
        physics=({"gamma": gamma, "upwinding": upwinding} if gravity == 0.0 else
                 {"gamma": gamma, "upwinding": upwinding, "gravity": float(gravity), "gravity_type": int(gravity_type)}),
        weno=weno_inp(weno), x=[x], u0=u)


# ------------------------------------------------------------------------------------- C3
def ns2d_vortex(n: Sequence[int] = (1024, 1024), weno: str = "js", tstype: str = "ssprk3",
                iproc=None, scheme: str = "weno5", upwinding: str = "rusanov", interp: str = "components") -> Case:
    nx, ny = n
    L = 10.0
    x = np.arange(nx, dtype=np.float64) * (L / nx)
    y = np.arange(ny, dtype=np.float64) * (L / ny)
    X, Y = np.meshgrid(x, y, indexing="xy")          # shape (ny, nx): dim 0 fastest
    gamma, b, x0, y0, uinf, vinf = 1.4, 0.5, 5.0, 5.0, 0.5, 0.0
    rx, ry = X - x0, Y - y0
    rsq = rx * rx + ry * ry
    rho = (1.0 - ((gamma - 1.0) * b * b) / (8.0 * gamma * np.pi * np.pi) * np.exp(1.0 - rsq)) ** (1.0 / (gamma - 1.0))
    du = -b / (2.0 * np.pi) * np.exp(0.5 * (1.0 - rsq)) * ry
    dv = b / (2.0 * np.pi) * np.exp(0.5 * (1.0 - rsq)) * rx
    vx, vy = uinf + du, vinf + dv
    p = rho ** gamma
    u = np.stack([rho, rho * vx, rho * vy, p / (gamma - 1.0) + 0.5 * rho * (vx * vx + vy * vy)], axis=-1)
    return Case(
        name=f"c3_vortex_{nx}x{ny}_{weno}" + ("" if upwinding == "rusanov" else "_" + upwinding)
             + ("" if interp == "components" else "_char") + _sfx(scheme),
        solver=_solver(2, 4, [nx, ny], "navierstokes2d", ts="rk", tstype=tstype, dt=0.005 * 1024.0 / max(nx, ny),
                       iproc=iproc, scheme=scheme, interp=interp),
        boundary=_zones(2, "periodic", [-1e3, -1e3], [1e3, 1e3]),
        physics={"gamma": gamma, "upwinding": upwinding}, weno=weno_inp(weno), x=[x, y], u0=u)


def ns2d_rising_bubble(n: Sequence[int] = (64, 64), weno: str = "js", tstype: str = "ssprk3", dt: float = 0.01,
                       iproc=None, hb: int = 2, upwinding: str = "rusanov", scheme: str = "weno5",
                       interp: str = "components") -> Case:
    """2-D rising thermal bubble (Examples/2D/NavierStokes2D/RisingThermalBubble/aux/init.c:80-115): slip walls,
    gravity (0, 9.8), HB = 2 hydrostatic balance, well-balanced source term."""
    gamma, R, g = 1.4, 287.058, 9.8
    rho_ref, p_ref = 1.1612055171196529, 100000.0
    L = 1000.0
    xs = [np.arange(n[d], dtype=np.float64) * (L / (n[d] - 1)) for d in range(2)]
    X, Y = np.meshgrid(xs[0], xs[1], indexing="xy")          # shape (ny, nx): dim 0 fastest
    T_ref = p_ref / (R * rho_ref)
    Cp = gamma / (gamma - 1.0) * R
    tc, xc, yc, rc = 0.5, 500.0, 350.0, 250.0
    r = np.sqrt((X - xc) ** 2 + (Y - yc) ** 2)
    dtheta = np.where(r > rc, 0.0, 0.5 * tc * (1.0 + np.cos(np.pi * r / rc)))
    theta = T_ref + dtheta
    Pexner = 1.0 - (g * Y) / (Cp * T_ref)
    rho = (p_ref / (R * theta)) * Pexner ** (1.0 / (gamma - 1.0))
    E = rho * (R / (gamma - 1.0)) * theta * Pexner
    zero = np.zeros_like(rho)
    u = np.stack([rho, zero, zero, E], axis=-1)
    return Case(
        name=f"c3g_bubble2d_{n[0]}x{n[1]}_{weno}_hb{hb}" + ("" if upwinding == "rusanov" else "_" + upwinding)
             + ("" if interp == "components" else "_char") + _sfx(scheme),
        solver=_solver(2, 4, n, "navierstokes2d", ts="rk", tstype=tstype, dt=dt, iproc=iproc,
                       par_type="nonconservative-2stage", par_scheme="4", scheme=scheme, interp=interp),
        boundary=_zones(2, "slip-wall", [0.0] * 2, [L] * 2, wall_velocity=[0.0, 0.0]),
        physics={"gamma": gamma, "upwinding": upwinding, "gravity": [0.0, g],
                 "rho_ref": rho_ref, "p_ref": p_ref, "R": R, "HB": hb},
        weno=weno_inp(weno), x=xs, u0=u)


# ------------------------------------------------------------------------------------- open / wall boundaries
def _flow_zone(kind, ndims, d, face, lo, hi, rho, vel, p, nvars, gamma):
    """one boundary.inp zone of the given type with the data InitializeBoundaries.c:107-175 reads for it"""
    z = {"type": kind, "dim": d, "face": face,
         "xmin": [0.0 if k == d else float(lo[k]) for k in range(ndims)],
         "xmax": [0.0 if k == d else float(hi[k]) for k in range(ndims)]}
    if kind in ("slip-wall", "noslip-wall"):
        z["wall_velocity"] = [0.0] * ndims
    elif kind == "dirichlet":
        z["values"] = [rho] + [rho * v for v in vel] + [p / (gamma - 1.0) + 0.5 * rho * sum(v * v for v in vel)]
    elif kind == "subsonic-inflow":
        z["density"], z["velocity"] = rho, list(vel)
    elif kind == "subsonic-outflow":
        z["pressure"] = p
    elif kind in ("subsonic-ambivalent", "supersonic-inflow"):
        z["density"], z["velocity"], z["pressure"] = rho, list(vel), p
    return z


BC_SETS = {
    "sup": {(0, 1): "supersonic-inflow", (0, -1): "supersonic-outflow", (1, 1): "dirichlet", (1, -1): "extrapolate"},
    "amb2": {(0, 1): "subsonic-ambivalent", (0, -1): "subsonic-ambivalent", (1, 1): "subsonic-ambivalent",
             (1, -1): "subsonic-ambivalent"},
    "sup3": {(0, 1): "supersonic-inflow", (0, -1): "supersonic-outflow", (1, 1): "dirichlet", (1, -1): "extrapolate",
             (2, 1): "subsonic-ambivalent", (2, -1): "subsonic-ambivalent"},
    "amb3": {(2, 1): "subsonic-ambivalent", (2, -1): "noslip-wall"},
}


def ns_channel(n: Sequence[int] = (32, 24), weno: str = "js", bcs: Optional[Dict] = None, mach: float = 0.5,
               upwinding: str = "rusanov", tstype: str = "ssprk3", viscous: bool = False, iproc=None,
               scheme: str = "weno5") -> Case:
    """Uniform stream (Mach `mach` along x, a small cross flow) with a smooth pressure / velocity disturbance in a box
    [0,1]^nd, 2-D (navierstokes2d) or 3-D (navierstokes3d) by len(n). `bcs` maps (dim, face) to a boundary type among
    slip-wall, noslip-wall, dirichlet, extrapolate, subsonic-inflow, subsonic-outflow, subsonic-ambivalent,
    supersonic-inflow, supersonic-outflow; default: inflow / outflow along x, walls on the other faces."""
    nd = len(n)
    if isinstance(bcs, str):            # named sets (fixtures store their arguments as JSON)
        bcs = BC_SETS[bcs]
    gamma = 1.4
    rho0, p0 = 1.0, 1.0 / gamma
    vel0 = [mach, 0.05 * mach, -0.03 * mach][:nd]
    xs = [(np.arange(n[d], dtype=np.float64) + 0.5) / n[d] for d in range(nd)]
    grids = np.meshgrid(*[xs[d] for d in reversed(range(nd))], indexing="ij")     # slowest dimension first
    X = list(reversed(grids))
    bump = np.ones_like(X[0])
    for d in range(nd):
        bump = bump * np.sin(np.pi * X[d]) ** 2
    rho = rho0 * (1.0 + 0.05 * bump)
    p = p0 * (1.0 + 0.08 * bump)
    vel = [vel0[d] * (1.0 + 0.1 * bump * np.cos(2.0 * np.pi * X[(d + 1) % nd])) for d in range(nd)]
    ke = 0.5 * rho * sum(v * v for v in vel)
    u = np.stack([rho] + [rho * v for v in vel] + [p / (gamma - 1.0) + ke], axis=-1)
    default = {(0, 1): "subsonic-inflow", (0, -1): "subsonic-outflow"}
    for d in range(1, nd):
        default[(d, 1)], default[(d, -1)] = "noslip-wall", "slip-wall"
    kinds = dict(default)
    kinds.update(bcs or {})
    zones = [_flow_zone(kinds[(d, f)], nd, d, f, [0.0] * nd, [1.0] * nd, rho0, vel0, p0, nd + 2, gamma)
             for d in range(nd) for f in (1, -1)]
    tag = "-".join(kinds[(d, f)].replace("subsonic-", "sub").replace("supersonic-", "sup")[:8] for d in range(nd) for f in (1, -1))
    phys = {"gamma": gamma, "upwinding": upwinding}
    if viscous:
        phys.update({"Pr": 0.72, "Minf": mach, "Re": 100.0})
    return Case(
        name=f"chan{nd}d_{'x'.join(str(v) for v in n)}_{weno}_{tag}" + ("" if upwinding == "rusanov" else "_" + upwinding)
             + ("_visc" if viscous else "") + _sfx(scheme),
        solver=_solver(nd, nd + 2, n, "navierstokes2d" if nd == 2 else "navierstokes3d", ts="rk", tstype=tstype,
                       dt=0.2 / max(n) / (1.0 + mach), iproc=iproc, par_type="nonconservative-2stage", par_scheme="4",
                       scheme=scheme),
        boundary=zones, physics=phys, weno=weno_inp(weno), x=xs, u0=u)


# ------------------------------------------------------------------------------------- C4
def _grid3(n, L):
    xs = [np.arange(n[d], dtype=np.float64) * (L[d] / n[d]) for d in range(3)]
    Z, Y, X = np.meshgrid(xs[2], xs[1], xs[0], indexing="ij")   # shape (nz, ny, nx)
    return xs, X, Y, Z


def ns3d_turbulence(n: Sequence[int] = (512, 512, 512), weno: str = "mapped", viscous: bool = True,
                    upwinding: str = "rusanov", tstype: str = "44", dt: float = 0.005,
                    iproc=None, seed: int = 20261017, interp: str = "components",
                    scheme: str = "weno5") -> Case:
    """Taylor-Green vortex plus 16 solenoidal Fourier modes |k|<=4 (deterministic phases),
    rho = 1, p = 1/gamma, Minf = 0.3: smooth, fully 3-D, every term of the RHS active."""
    gamma, Minf = 1.4, 0.3
    xs, X, Y, Z = _grid3(n, [2.0 * np.pi] * 3)
    vx = Minf * np.sin(X) * np.cos(Y) * np.cos(Z)
    vy = -Minf * np.cos(X) * np.sin(Y) * np.cos(Z)
    vz = np.zeros_like(vx)
    rng = np.random.RandomState(seed)
    for _ in range(16):
        k = rng.randint(-4, 5, size=3).astype(np.float64)
        if not k.any():
            k[0] = 1.0
        a = rng.standard_normal(3)
        a -= k * (a @ k) / (k @ k)                    # a . k = 0 -> divergence free
        nrm = np.linalg.norm(a)
        if nrm < 1e-12:
            continue
        a *= 0.1 * Minf / (4.0 * nrm)
        ph = rng.uniform(0.0, 2.0 * np.pi)
        s = np.sin(k[0] * X + k[1] * Y + k[2] * Z + ph)
        vx += a[0] * s
        vy += a[1] * s
        vz += a[2] * s
    rho = np.ones_like(vx)
    p = np.full_like(vx, 1.0 / gamma)
    u = np.stack([rho, rho * vx, rho * vy, rho * vz,
                  p / (gamma - 1.0) + 0.5 * rho * (vx * vx + vy * vy + vz * vz)], axis=-1)
    phys: Dict[str, object] = {"gamma": gamma, "upwinding": upwinding, "Pr": 0.72, "Minf": Minf}
    phys["Re"] = 333.333333333333333 if viscous else -1.0
    return Case(
        name=f"c4_turb_{n[0]}x{n[1]}x{n[2]}_{weno}_{'visc' if viscous else 'inv'}"
             + ("" if upwinding == "rusanov" else "_" + upwinding) + ("" if interp == "components" else "_char") + _sfx(scheme),
        solver=_solver(3, 5, n, "navierstokes3d", ts="rk", tstype=tstype, dt=dt, iproc=iproc, interp=interp,
                       par_type="nonconservative-2stage", par_scheme="4", scheme=scheme),
        boundary=_zones(3, "periodic", [-1e3] * 3, [1e3] * 3),
        physics=phys, weno=weno_inp(weno), x=xs, u0=u)


# ------------------------------------------------------------------------------------- C5a
def ns3d_density_wave(n: Sequence[int] = (64, 64, 64), weno: str = "js", tstype: str = "44",
                      dt: float = 1e-3, iproc=None, scheme: str = "weno5", upwinding: str = "rusanov") -> Case:
    gamma = 1.4
    xs, X, Y, Z = _grid3(n, [1.0] * 3)
    rho = 1.0 + 0.1 * np.sin(2 * np.pi * X) * np.sin(2 * np.pi * Y) * np.sin(2 * np.pi * Z)
    v = np.ones_like(rho)
    p = np.full_like(rho, 1.0 / gamma)
    u = np.stack([rho, rho * v, rho * v, rho * v, p / (gamma - 1.0) + 0.5 * rho * 3.0 * v * v], axis=-1)
    return Case(
        name=f"c5a_denswave_{n[0]}x{n[1]}x{n[2]}_{weno}" + _sfx(scheme),
        solver=_solver(3, 5, n, "navierstokes3d", ts="rk", tstype=tstype, dt=dt, iproc=iproc,
                       par_type="nonconservative-2stage", par_scheme="4", scheme=scheme),
        boundary=_zones(3, "periodic", [-1e3] * 3, [1e3] * 3),
        physics={"gamma": gamma, "upwinding": upwinding}, weno=weno_inp(weno), x=xs, u0=u)


# ------------------------------------------------------------------------------------- C5b
def ns3d_rising_bubble(n: Sequence[int] = (64, 64, 64), weno: str = "yc", tstype: str = "ssprk3",
                       dt: float = 0.01, iproc=None, hb: int = 2, scheme: str = "weno5",
                       upwinding: str = "rusanov") -> Case:
    """Rising thermal bubble: slip walls, gravity (0, 9.8, 0), HB = 2 hydrostatic balance."""
    gamma, R, g = 1.4, 287.058, 9.8
    rho_ref, p_ref = 1.1612055171196529, 100000.0
    L = 1000.0
    # cell-centred-at-nodes grid of the example: x_i = i * L/(N-1)
    xs = [np.arange(n[d], dtype=np.float64) * (L / (n[d] - 1)) for d in range(3)]
    Z, Y, X = np.meshgrid(xs[2], xs[1], xs[0], indexing="ij")
    T_ref = p_ref / (R * rho_ref)
    Cp = gamma / (gamma - 1.0) * R
    tc, xc, yc, zc, rc = 1.0, 500.0, 260.0, 500.0, 250.0
    r = np.sqrt((X - xc) ** 2 + (Y - yc) ** 2 + (Z - zc) ** 2)
    dtheta = np.where(r > rc, 0.0, 0.5 * tc * (1.0 + np.cos(np.pi * r / rc)))
    theta = T_ref + dtheta
    Pexner = 1.0 - (g / (Cp * T_ref)) * Y
    rho = (p_ref / (R * theta)) * Pexner ** (1.0 / (gamma - 1.0))
    E = rho * (R / (gamma - 1.0)) * theta * Pexner
    zero = np.zeros_like(rho)
    u = np.stack([rho, zero, zero, zero, E], axis=-1)
    return Case(
        name=f"c5b_bubble_{n[0]}x{n[1]}x{n[2]}_{weno}" + _sfx(scheme),
        solver=_solver(3, 5, n, "navierstokes3d", ts="rk", tstype=tstype, dt=dt, iproc=iproc,
                       par_type="nonconservative-2stage", par_scheme="4", scheme=scheme),
        boundary=_zones(3, "slip-wall", [0.0] * 3, [L] * 3, wall_velocity=[0.0, 0.0, 0.0]),
        physics={"gamma": gamma, "upwinding": upwinding, "gravity": [0.0, g, 0.0],
                 "rho_ref": rho_ref, "p_ref": p_ref, "R": R, "HB": hb},
        weno=weno_inp(weno), x=xs, u0=u)

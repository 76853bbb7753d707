"""Host-side mirror of HyPar's solver object for the explicit-RHS path.

``Solver`` wraps one ``hpb_solver`` (one rank / one GPU) of libhypar_b200.so. Its methods carry the
names and argument meaning of the function pointers in the reference's ``struct HyPar``
(include/hypar.h:211-359): ``ApplyBoundaryConditions``, ``HyperbolicFunction``, ``ParabolicFunction``,
``SourceFunction``, ``FFunction``, ``UFunction``, ``SetInterpLimiterVar``,
``InterpolateInterfacesHyp``, ``Upwind``, ``FirstDerivativePar``, ``SecondDerivativePar``,
``ComputeCFL``, and ``RHSFunction`` / ``TimeIntegrate`` of ``struct TimeIntegration``. Arrays are
numpy float64 in HyPar's own layout (ghost-padded AoS, flattened). Every call goes through the C ABI
into CUDA kernels; errors raise ``HyParB200Error`` (the C layer also keeps a sticky error state).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib, hypario

MODELS = {"linear-advection-diffusion-reaction": 0, "euler1d": 1, "navierstokes2d": 2, "navierstokes3d": 3, "burgers": 4}
BCTYPES = {"periodic": 0, "extrapolate": 1, "slip-wall": 2, "noslip-wall": 3, "dirichlet": 4, "subsonic-inflow": 5,
           "subsonic-outflow": 6, "subsonic-ambivalent": 7, "supersonic-inflow": 8, "supersonic-outflow": 9, "sponge": 10}
PAR_TYPES = {"nonconservative-1stage": 0, "nonconservative-1.5stage": 1, "nonconservative-2stage": 2, "conservative-1stage": 3}
UPWINDS = {"roe": 1, "rusanov": 2, "rf-char": 3, "llf-char": 4}
RK_TYPES = {"44": 0, "ssprk3": 1, "tvdrk3": 1, "1fe": 2, "22": 3, "33": 4}
GLMGEE_TYPES = {"23": 16, "24": 17, "25i": 18, "35": 19, "exrk2a": 20, "rk32g1": 21, "rk285ex": 22}
GLMGEE_MODES = {"yeps": 0, "yyt": 1}
SCHEMES = {"weno5": 0, "crweno5": 1, "cupw5": 2, "upw5": 3, "1": 4, "2": 5, "4": 6, "muscl2": 7, "muscl3": 8, "hcweno5": 9}
LIMITERS = {"gmm": 0, "minmod": 1, "vanleer": 2, "superbee": 3}
FIELD_U, FIELD_QDERIVX, FIELD_QDERIVY = 0, 1, 2


class HyParB200Error(RuntimeError):
    pass


def _dp(a: np.ndarray):
    if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
        raise HyParB200Error("arrays must be C-contiguous float64")
    return a.ctypes.data_as(C.POINTER(C.c_double))


def config_from_inputs(solver: Dict[str, object], boundary: Sequence[dict], physics: Dict[str, object],
                       weno: Optional[Dict[str, object]], x: Sequence[np.ndarray], rank: int = 0,
                       device: int = -1, use_fused: bool = True, muscl: Optional[Dict[str, object]] = None,
                       advection_field: Optional[np.ndarray] = None, glm_gee: Optional[Dict[str, object]] = None,
                       lusolver: Optional[Dict[str, object]] = None):
    """Translate the contents of solver.inp / boundary.inp / physics.inp / weno.inp / muscl.inp (as parsed
    dictionaries) into an ``hpb_config``. Unsupported choices raise here or in ``hpb_create``."""
    L = _lib.load()
    c = _lib.Config()
    L.hpb_config_defaults(C.byref(c))
    nd = int(solver["ndims"])
    c.ndims, c.nvars, c.ghosts = nd, int(solver["nvars"]), int(solver.get("ghost", 1))
    size = solver["size"]
    iproc = solver.get("iproc", [1] * nd)
    for d in range(nd):
        c.dim_global[d] = int(size[d])
        c.iproc[d] = int(iproc[d])
    c.rank = rank
    model = str(solver.get("model", "none"))
    if model not in MODELS:
        raise HyParB200Error(f"model '{model}' is not on the B200 path (supported: {sorted(MODELS)})")
    c.model = MODELS[model]
    scheme = str(solver.get("hyp_space_scheme", "1"))
    if scheme not in SCHEMES:
        raise HyParB200Error(f"hyp_space_scheme '{scheme}' is not on the B200 path (weno5, crweno5, hcweno5, cupw5, upw5, 1, 2, 4, muscl2, muscl3)")
    c.hyp_scheme = SCHEMES[scheme]
    mu = muscl or {}
    c.muscl_eps = float(mu.get("epsilon", 1e-3))                         # MUSCLInitialize.c:26-27
    c.muscl_limiter = LIMITERS.get(str(mu.get("limiter", "gmm")), 0)     # :72-75: unknown names fall back to gmm
    ts = str(solver.get("time_scheme", "euler"))
    if ts not in ("rk", "euler", "glm-gee"):
        raise HyParB200Error(f"time_scheme '{ts}' is not on the B200 path (rk, euler, glm-gee)")
    if ts == "glm-gee":                                   # TimeGLMGEEInitialize.c:41-70, glm_gee.inp :431-470
        tst = str(solver.get("time_scheme_type", " "))
        if tst not in GLMGEE_TYPES:
            raise HyParB200Error(f"time_scheme_type '{tst}' is not a glm-gee method (23, 24, 25i, 35, exrk2a, rk32g1, rk285ex)")
        c.rk_type = GLMGEE_TYPES[tst]
        mode = str((glm_gee or {}).get("ee_mode", "yeps"))
        if mode not in GLMGEE_MODES:
            raise HyParB200Error(f"glm_gee.inp: ee_mode '{mode}' (yeps, yyt)")
        c.glm_ee_mode = GLMGEE_MODES[mode]
    else:
        tst = str(solver.get("time_scheme_type", " ")) if ts == "rk" else "1fe"     # TimeForwardEuler.c = RK "1fe"
        if tst not in RK_TYPES:
            raise HyParB200Error(f"time_scheme_type '{tst}' is not on the B200 path (1fe, 22, 33, 44, ssprk3, tvdrk3)")
        c.rk_type = RK_TYPES[tst]
    lu = lusolver or {}                                   # lusolver.inp (tridiagLUInit.c:54-90): compact schemes across ranks
    c.lu_maxiter, c.lu_evaluate_norm = int(lu.get("maxiter", 10)), int(lu.get("evaluate_norm", 1))
    c.lu_atol, c.lu_rtol = float(lu.get("atol", 1e-12)), float(lu.get("rtol", 1e-10))
    rst = str(lu.get("reducedsolvetype", "jacobi"))
    if rst not in ("jacobi", "gather-and-solve"):
        raise HyParB200Error(f"lusolver.inp: reducedsolvetype '{rst}' (jacobi, gather-and-solve)")
    c.lu_gather_and_solve = int(rst == "gather-and-solve")
    if str(solver.get("immersed_body", "none")) != "none":
        raise HyParB200Error("immersed bodies are not on the B200 path")
    if str(solver.get("hyp_flux_split", "no")) != "no":
        raise HyParB200Error("hyp_flux_split yes is not on the B200 path")
    it = str(solver.get("hyp_interp_type", "characteristic"))
    if it not in ("characteristic", "components"):
        raise HyParB200Error(f"{it} is not a supported interpolation type")
    c.interp_char = int(it == "characteristic")
    c.par_scheme = int(solver.get("par_space_scheme", "2"))
    pst = str(solver.get("par_space_type", "nonconservative-1stage"))
    if pst not in PAR_TYPES:          # InitializeSolvers.c:177-181
        raise HyParB200Error(f"{pst} is not a supported spatial discretization type for the parabolic terms")
    c.par_space_type = PAR_TYPES[pst]
    c.dt = float(solver.get("dt", 0.0))
    w = weno or {}
    c.weno_type = 3 if int(w.get("yc", 0)) else 2 if int(w.get("borges", 0)) else 1 if int(w.get("mapped", 0)) else 0
    c.no_limiting = int(w.get("no_limiting", 0))
    c.weno_eps = float(w.get("epsilon", 1e-6))
    c.weno_rc, c.weno_xi = float(w.get("rc", 0.3)), float(w.get("xi", 0.001))     # WENOInitialize.c:57-58
    ph = physics or {}
    if c.model == 0:
        c.upwind = 0
        if str(ph.get("centered_flux", "no")) != "no":
            raise HyParB200Error("LinearADR centered_flux is not on the B200 path")
        if c.nvars != 1:
            raise HyParB200Error("LinearADR: nvars must be 1 on the B200 path")
        if "advection_filename" in ph:
            # LinearADRInitialize.c:98-124: `advection` and `advection_filename` exclude each other; a field file that is
            # not there leaves the field at zero (LinearADRAdvectionField.c:117-120)
            if "advection" in ph:
                raise HyParB200Error("LinearADR: both advection and advection_filename are specified")
            if str(ph["advection_filename"]) != "none":
                npts = int(np.prod([int(v) for v in size[:nd]]))
                af = (np.zeros(npts * nd * c.nvars) if advection_field is None
                      else np.ascontiguousarray(advection_field, dtype=np.float64).reshape(-1))
                if af.size != npts * nd * c.nvars:
                    raise HyParB200Error("LinearADR: the advection field must have ndims*nvars components on the global grid")
                c.advection_field = _dp(af)
                c._advf_keep = af
    elif c.model == 4:
        c.upwind = 0                 # BurgersUpwind: the model's only upwinding
        if c.nvars != 1:
            raise HyParB200Error("burgers: nvars must be 1")
    else:
        up = str(ph.get("upwinding", "roe"))
        if up not in UPWINDS:
            raise HyParB200Error(f"upwinding '{up}' is not on the B200 path (roe, rusanov, rf-char, llf-char)")
        c.upwind = UPWINDS[up]
    c.gamma = float(ph.get("gamma", 1.4))
    c.Re, c.Pr, c.Minf = float(ph.get("Re", -1.0)), float(ph.get("Pr", 0.72)), float(ph.get("Minf", 1.0))
    if c.model in (2, 3) and c.Re > 0 and str(solver.get("par_space_type")) != "nonconservative-2stage":
        raise HyParB200Error('Parabolic term spatial discretization must be "nonconservative-2stage"')
    grav = ph.get("gravity", [0.0] * 3)
    grav = list(grav) if isinstance(grav, (list, tuple)) else [grav]
    if c.model == 2:
        grav = (grav + [0.0, 0.0])[:2] + [0.0]
    elif c.model == 1:
        g1 = ph.get("gravity", 0.0)
        grav = [float(g1[0] if isinstance(g1, (list, tuple)) else g1), 0.0, 0.0]
        c.gravity_type = int(ph.get("gravity_type", 0))
    for d in range(min(3, len(grav))):
        c.gravity[d] = float(grav[d])
    c.rho_ref, c.p_ref, c.R = float(ph.get("rho_ref", 1.0)), float(ph.get("p_ref", 1.0)), float(ph.get("R", 1.0))
    c.HB, c.N_bv = int(ph.get("HB", 1)), float(ph.get("N_bv", 0.0))
    adv = ph.get("advection", [])
    adv = list(adv) if isinstance(adv, (list, tuple)) else [adv]
    for i, a in enumerate(adv):
        c.advection[i] = float(a)
    dif = ph.get("diffusion", [])
    dif = list(dif) if isinstance(dif, (list, tuple)) else [dif]
    for i, a in enumerate(dif):
        c.diffusion[i] = float(a)
    if len(boundary) > _lib.MAX_ZONES:
        raise HyParB200Error("too many boundary zones")
    c.nzones = len(boundary)
    for n, z in enumerate(boundary):
        if z["type"] not in BCTYPES:
            raise HyParB200Error(f"boundary type '{z['type']}' is not on the B200 path ({', '.join(BCTYPES)})")
        cz = c.zones[n]
        cz.type, cz.dim, cz.face = BCTYPES[z["type"]], int(z["dim"]), int(z["face"])
        for d in range(nd):
            cz.xmin[d], cz.xmax[d] = float(z["xmin"][d]), float(z["xmax"][d])
            cz.wall_velocity[d] = float(z.get("wall_velocity", z.get("velocity", [0.0] * nd))[d])
        cz.flow_density, cz.flow_pressure = float(z.get("density", 0.0)), float(z.get("pressure", 0.0))
        for v, val in enumerate(z.get("values", [])):
            cz.dirichlet[v] = float(val)
    xg = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.float64) for v in x]))
    c.x_global = _dp(xg)
    c.device = device
    c.use_fused = int(use_fused)
    c.conservation_check = int(str(solver.get("conservation_check", "no")) == "yes")
    return c, xg


class Solver:
    """One rank of the B200 explicit-RHS path."""

    def __init__(self, solver: Dict[str, object], boundary, physics, weno, x, rank: int = 0,
                 device: int = -1, use_fused: bool = True, muscl=None, advection_field=None, glm_gee=None, lusolver=None):
        self.L = _lib.load()
        self.inputs = {"solver": solver, "boundary": boundary, "physics": physics, "weno": weno, "muscl": muscl,
                       "advection_field": advection_field, "glm_gee": glm_gee, "lusolver": lusolver}
        cfg, self._xg = config_from_inputs(solver, boundary, physics, weno, x, rank, device, use_fused, muscl,
                                           advection_field, glm_gee, lusolver)
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.L.hpb_create(C.byref(cfg), C.byref(self.h))
        if rc:
            raise HyParB200Error(self.L.hpb_last_error().decode())
        self.ndims, self.nvars, self.ghosts = cfg.ndims, cfg.nvars, cfg.ghosts
        dl, isg = (C.c_int * 3)(), (C.c_int * 3)()
        self.L.hpb_get_local_dims(self.h, dl, isg)
        self.dim_local = [dl[d] for d in range(self.ndims)]
        self.is_global = [isg[d] for d in range(self.ndims)]
        self.dim_global = [cfg.dim_global[d] for d in range(self.ndims)]
        self.iproc = [cfg.iproc[d] for d in range(self.ndims)]
        self.rank = rank
        self.npoints_local_wghosts = int(self.L.hpb_npoints_local_wghosts(self.h))
        self.dt = cfg.dt
        nb = (C.c_int * 6)()
        self.L.hpb_get_neighbors(self.h, nb)
        self.neighbors = [nb[k] for k in range(2 * self.ndims)]

    # -- construction helpers
    @classmethod
    def from_case(cls, case, rank: int = 0, device: int = -1, use_fused: bool = True) -> "Solver":
        return cls(case.solver, case.boundary, case.physics, case.weno, case.x, rank, device, use_fused,
                   muscl=getattr(case, "muscl", None), advection_field=getattr(case, "advection_field", None),
                   glm_gee=getattr(case, "glm_gee", None), lusolver=getattr(case, "lusolver", None))

    @classmethod
    def from_directory(cls, path: str, rank: int = 0, device: int = -1, use_fused: bool = True) -> "Solver":
        """Attach to an existing HyPar run directory (solver.inp, boundary.inp, physics.inp,
        [weno.inp], initial.inp), unchanged."""
        s = hypario.read_solver_inp(os.path.join(path, "solver.inp"))
        nd, nv = int(s["ndims"]), int(s["nvars"])
        b = hypario.read_boundary_inp(os.path.join(path, "boundary.inp"), nd, nv)
        vk = {"gravity": 3 if s["model"] == "navierstokes3d" else 2 if s["model"] == "navierstokes2d" else 1,
              "advection": nd * nv, "diffusion": nd * nv}
        pf = os.path.join(path, "physics.inp")
        ph = hypario.read_keyword_file(pf, vector_keys=vk) if os.path.exists(pf) else {}
        for k in ("advection", "diffusion", "gravity"):
            if k in ph:
                ph[k] = [float(v) for v in ph[k]]
        wf = os.path.join(path, "weno.inp")
        w = hypario.read_keyword_file(wf) if os.path.exists(wf) else None
        ipt = str(s.get("ip_file_type", "ascii"))
        if ipt not in ("ascii", "binary", "bin"):
            raise HyParB200Error(f"ip_file_type '{ipt}' is neither ascii nor binary")
        x, u0 = hypario.read_initial(os.path.join(path, "initial.inp"), s["size"], nv, ipt)
        mf = os.path.join(path, "muscl.inp")
        mu = hypario.read_keyword_file(mf) if os.path.exists(mf) else None
        af = None
        if str(ph.get("advection_filename", "none")) != "none":        # same layout and flavour as initial.inp (ReadArray.c)
            fn = os.path.join(path, str(ph["advection_filename"]) + ".inp")
            if os.path.exists(fn):
                af = hypario.read_initial(fn, s["size"], nd * nv, ipt)[1]
        gf = os.path.join(path, "glm_gee.inp")
        gg = hypario.read_keyword_file(gf) if os.path.exists(gf) else None
        lf = os.path.join(path, "lusolver.inp")
        lu = hypario.read_keyword_file(lf) if os.path.exists(lf) else None
        obj = cls(s, b, ph, w, x, rank, device, use_fused, muscl=mu, advection_field=af, glm_gee=gg, lusolver=lu)
        obj.u0_global = u0
        return obj

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.hpb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise HyParB200Error(self.L.hpb_last_error().decode())

    # -- sizes / set-up products
    def zeros(self) -> np.ndarray:
        return np.zeros(self.npoints_local_wghosts * self.nvars)

    def ninterfaces(self, d: int) -> int:
        return int(self.L.hpb_ninterfaces(self.h, d))

    def shape_g(self):
        g = self.ghosts
        return tuple(n + 2 * g for n in reversed(self.dim_local)) + (self.nvars,)

    def interior(self, a: np.ndarray) -> np.ndarray:
        g = self.ghosts
        sl = tuple(slice(g, g + n) for n in reversed(self.dim_local))
        return a.reshape(self.shape_g())[sl]

    def grid(self):
        n = sum(d + 2 * self.ghosts for d in self.dim_local)
        x, dxinv = np.zeros(n), np.zeros(n)
        self.L.hpb_get_grid(self.h, _dp(x), _dp(dxinv))
        return x, dxinv

    def zone_extent(self, n: int):
        a, b, on = (C.c_int * 3)(), (C.c_int * 3)(), C.c_int()
        self._ck(self.L.hpb_get_zone_extent(self.h, n, a, b, C.byref(on)))
        return [a[d] for d in range(self.ndims)], [b[d] for d in range(self.ndims)], on.value

    def gravity_field(self):
        f, g = np.zeros(self.npoints_local_wghosts), np.zeros(self.npoints_local_wghosts)
        self.L.hpb_get_gravity_field(self.h, _dp(f), _dp(g))
        return f, g

    def advection_field(self) -> np.ndarray:
        """LinearADR::a of a spatially varying advection: [point with ghosts][ndims*nvars]."""
        a = np.zeros(self.npoints_local_wghosts * self.ndims * self.nvars)
        self._ck(self.L.hpb_get_advection_field(self.h, _dp(a)))
        return a

    def local_from_global(self, ug: np.ndarray) -> np.ndarray:
        """This rank's ghost-padded AoS block (ghosts zero) out of a global (N_{nd-1},...,N_0,nvars) array."""
        g = self.ghosts
        u = np.zeros(self.shape_g())
        src = tuple(slice(self.is_global[d], self.is_global[d] + self.dim_local[d]) for d in reversed(range(self.ndims)))
        dst = tuple(slice(g, g + n) for n in reversed(self.dim_local))
        u[dst] = ug[src]
        return np.ascontiguousarray(u).reshape(-1)

    # -- the reference's function-pointer surface (host arrays)
    def ApplyBoundaryConditions(self, u: np.ndarray, t: float = 0.0) -> np.ndarray:
        self._ck(self.L.hpb_ApplyBoundaryConditions(self.h, _dp(u), t))
        return u

    def HyperbolicFunction(self, u: np.ndarray, t: float = 0.0, LimFlag: int = 1) -> np.ndarray:
        hyp = self.zeros()
        self._ck(self.L.hpb_HyperbolicFunction(self.h, _dp(hyp), _dp(u), t, LimFlag))
        return hyp

    def ParabolicFunction(self, u: np.ndarray, t: float = 0.0) -> np.ndarray:
        par = self.zeros()
        self._ck(self.L.hpb_ParabolicFunction(self.h, _dp(par), _dp(u), t))
        return par

    def SourceFunction(self, u: np.ndarray, t: float = 0.0) -> np.ndarray:
        src = self.zeros()
        self._ck(self.L.hpb_SourceFunction(self.h, _dp(src), _dp(u), t))
        return src

    def RHSFunction(self, u: np.ndarray, t: float = 0.0) -> np.ndarray:
        rhs = self.zeros()
        self._ck(self.L.hpb_RHSFunction(self.h, _dp(rhs), _dp(u), t))
        return rhs

    def FFunction(self, u: np.ndarray, d: int, t: float = 0.0) -> np.ndarray:
        f = self.zeros()
        self._ck(self.L.hpb_FFunction(self.h, _dp(f), _dp(u), d, t))
        return f

    def UFunction(self, u: np.ndarray, d: int = 0, t: float = 0.0) -> np.ndarray:
        uC = self.zeros()
        self._ck(self.L.hpb_UFunction(self.h, _dp(uC), _dp(u), d, t))
        return uC

    def SetInterpLimiterVar(self, fC: np.ndarray, u: np.ndarray, d: int) -> None:
        self._ck(self.L.hpb_SetInterpLimiterVar(self.h, _dp(fC), _dp(u), d))

    def GetInterpWeights(self, d: int) -> np.ndarray:
        w = np.zeros(12 * self.ninterfaces(d) * self.nvars)
        self._ck(self.L.hpb_GetInterpWeights(self.h, d, _dp(w)))
        return w

    def InterpolateInterfacesHyp(self, fC: np.ndarray, u: np.ndarray, upw: int, d: int, uflag: int) -> np.ndarray:
        fI = np.zeros(self.ninterfaces(d) * self.nvars)
        self._ck(self.L.hpb_InterpolateInterfacesHyp(self.h, _dp(fI), _dp(fC), _dp(u), upw, d, uflag))
        return fI

    def Upwind(self, fL, fR, uL, uR, u, d: int, t: float = 0.0) -> np.ndarray:
        fI = np.zeros_like(fL)
        self._ck(self.L.hpb_Upwind(self.h, _dp(fI), _dp(fL), _dp(fR), _dp(uL), _dp(uR), _dp(u), d, t))
        return fI

    def FirstDerivativePar(self, f: np.ndarray, d: int, bias: int = 1) -> np.ndarray:
        Df = self.zeros()
        self._ck(self.L.hpb_FirstDerivativePar(self.h, _dp(Df), _dp(f), d, bias))
        return Df

    def SecondDerivativePar(self, f: np.ndarray, d: int) -> np.ndarray:
        D2 = self.zeros()
        self._ck(self.L.hpb_SecondDerivativePar(self.h, _dp(D2), _dp(f), d))
        return D2

    def ComputeCFL(self, u: np.ndarray, dt: Optional[float] = None, t: float = 0.0) -> float:
        out = C.c_double()
        self._ck(self.L.hpb_ComputeCFL(self.h, _dp(u), self.dt if dt is None else dt, t, C.byref(out)))
        return out.value

    def TimeIntegrate(self, u: np.ndarray, nsteps: int = 1, t0: float = 0.0) -> np.ndarray:
        self._ck(self.L.hpb_TimeIntegrate(self.h, _dp(u), nsteps, t0))
        return u

    # -- pipelined host-array stepping over a sequence of independent fields (ensembles): calls only enqueue
    def TimeIntegrateAsync(self, u_in: np.ndarray, u_out: np.ndarray, nsteps: int = 1, t0: float = 0.0) -> None:
        self._ck(self.L.hpb_TimeIntegrateAsync(self.h, _dp(u_in), _dp(u_out), nsteps, t0))

    def pipe_upload(self, u_in: np.ndarray, t0: float = 0.0) -> None:
        self._ck(self.L.hpb_pipe_upload(self.h, _dp(u_in), t0))

    def pipe_download(self, u_out: np.ndarray) -> None:
        self._ck(self.L.hpb_pipe_download(self.h, _dp(u_out)))

    def pipe_join(self) -> None:
        self._ck(self.L.hpb_pipe_join(self.h))

    def pipe_wait(self) -> None:
        self._ck(self.L.hpb_pipe_wait(self.h))

    # -- device-resident path
    def set_solution(self, u: np.ndarray) -> None:
        self._ck(self.L.hpb_dev_set_solution(self.h, _dp(u)))

    def get_solution(self) -> np.ndarray:
        u = self.zeros()
        self._ck(self.L.hpb_dev_get_solution(self.h, _dp(u)))
        return u

    def interior_grid(self) -> List[np.ndarray]:
        """This rank's coordinates without ghosts, one array per dimension."""
        x, _ = self.grid()
        g, out, off = self.ghosts, [], 0
        for d in range(self.ndims):
            out.append(x[off + g: off + g + self.dim_local[d]].copy())
            off += self.dim_local[d] + 2 * g
        return out

    def write_solution(self, directory: str, index: int = 0, root: str = "op") -> str:
        """The device solution as the reference's OutputSolution would write it on one rank: ``op_file_format`` text /
        tecplot2d / tecplot3d / binary and ``op_overwrite`` of solver.inp decide format and name (hypario.write_solution:
        byte-identical to WriteText.c / WriteTecplot*.c / WriteBinary.c). Decomposed runs use write_solution_parallel."""
        if any(p != 1 for p in self.iproc):
            raise HyParB200Error("write_solution gathers nothing: one rank only (use write_solution_parallel)")
        s = self.inputs["solver"]
        fmt = str(s.get("op_file_format", "text"))
        name = hypario.solution_file_name(fmt, str(s.get("op_overwrite", "no")) == "yes", index, root)
        hypario.write_solution(os.path.join(directory, name), self.interior_grid(), self.interior(self.get_solution()), fmt)
        return name

    def TimeStep(self) -> None:
        self._ck(self.L.hpb_TimeStep(self.h))

    def TimeSteps(self, n: int) -> None:
        self._ck(self.L.hpb_TimeSteps(self.h, n))

    def dev_RHS(self, t: float = 0.0, want: bool = True) -> Optional[np.ndarray]:
        rhs = self.zeros() if want else None
        self._ck(self.L.hpb_dev_RHS(self.h, t, _dp(rhs) if want else None))
        return rhs

    def get_stage_rhs(self, stage: int = 0) -> np.ndarray:
        rhs = self.zeros()
        self._ck(self.L.hpb_dev_get_stage_rhs(self.h, stage, _dp(rhs)))
        return rhs

    def dev_ComputeCFL(self) -> float:
        out = C.c_double()
        self._ck(self.L.hpb_dev_ComputeCFL(self.h, C.byref(out)))
        return out.value

    def dev_StepNormSumSq(self) -> float:
        out = C.c_double()
        self._ck(self.L.hpb_dev_StepNormSumSq(self.h, C.byref(out)))
        return out.value

    # -- partitioned I/O: this rank's block of the reference's parallel / MPI-IO files, addressed
    #    directly (hypario: no gather through a leader rank, 64-bit offsets)
    def local_grid(self):
        x, _ = self.grid()
        g, out, off = self.ghosts, [], 0
        for n in self.dim_local:
            out.append(x[off + g: off + g + n].copy())
            off += n + 2 * g
        return out

    def write_solution_parallel(self, root_ext: str = "op.bin", n_io: int = 1, record: int = 0, truncate: bool = False) -> str:
        """WriteArrayParallel (WriteArray.c:139-323): the device solution of this rank into <root_ext>.<nnnn>
        (truncate: see hypario.write_parallel_block -- only with a barrier between output times)"""
        u = np.ascontiguousarray(self.interior(self.get_solution()))
        return hypario.write_parallel_block(root_ext, self.rank, self.dim_global, self.iproc, self.nvars, n_io,
                                            self.local_grid(), u, record, truncate)

    def load_solution_parallel(self, fname_root: str = "initial", n_io: int = 1, mode: str = "parallel") -> None:
        """ReadArrayParallel / ReadArrayMPI_IO (ReadArray.c:293-650): this rank's block of <fname_root>_par.inp.<nnnn>
        (mode "parallel") or <fname_root>_mpi.inp (mode "mpi-io") onto the device; the grid in the file must be the
        one the solver was created with"""
        if mode == "parallel":
            xl, ul = hypario.read_parallel_block(fname_root, self.rank, self.dim_global, self.iproc, self.nvars, n_io)
        elif mode == "mpi-io":
            xl, ul = hypario.read_mpi_io_block(fname_root, self.rank, self.dim_global, self.iproc, self.nvars)
        else:
            raise HyParB200Error(f"input mode '{mode}' (parallel, mpi-io)")
        for a, b in zip(xl, self.local_grid()):
            if not np.array_equal(a, b):
                raise HyParB200Error(f"{fname_root}: the grid in the file differs from the solver's grid")
        g = self.ghosts
        u = np.zeros(self.shape_g())
        u[tuple(slice(g, g + n) for n in reversed(self.dim_local))] = ul
        self.set_solution(np.ascontiguousarray(u).reshape(-1))

    # -- conservation / error diagnostics (this rank's parts; see include/hypar_b200.h)
    def dev_VolumeIntegral(self) -> np.ndarray:
        out = np.zeros(self.nvars)
        self._ck(self.L.hpb_dev_VolumeIntegral(self.h, _dp(out)))
        return out

    def dev_StageBoundaryIntegral(self, slot: int = -1) -> np.ndarray:
        out = np.zeros(2 * self.ndims * self.nvars)
        self._ck(self.L.hpb_dev_StageBoundaryIntegral(self.h, slot, _dp(out)))
        return out

    def dev_StepBoundaryIntegral(self) -> np.ndarray:
        out = np.zeros(2 * self.ndims * self.nvars)
        self._ck(self.L.hpb_dev_StepBoundaryIntegral(self.h, _dp(out)))
        return out

    def BoundaryIntegral(self, step_bi: np.ndarray) -> np.ndarray:
        out = np.zeros(self.nvars)
        self._ck(self.L.hpb_BoundaryIntegral(self.h, _dp(np.ascontiguousarray(step_bi, dtype=np.float64)), _dp(out)))
        return out

    def CalculateConservationError(self, vol, vol0, total_bi) -> np.ndarray:
        err = np.zeros(self.nvars)
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (vol, vol0, total_bi)]
        self._ck(self.L.hpb_CalculateConservationError(self.nvars, _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(err)))
        return err

    def dev_ErrorSums(self, uex: np.ndarray) -> np.ndarray:
        out = np.zeros(6)
        self._ck(self.L.hpb_dev_ErrorSums(self.h, _dp(uex), _dp(out)))
        return out

    # -- GLM-GEE (time_scheme glm-gee): the auxiliary solution and the estimated global error
    def get_aux_solution(self) -> np.ndarray:
        """TimeGetAuxSolutions.c: the auxiliary solution the GLM-GEE method propagates (HyPar layout, with ghosts)"""
        out = np.zeros(self.npoints_local_wghosts * self.nvars)
        self._ck(self.L.hpb_dev_get_aux_solution(self.h, _dp(out)))
        return out

    def set_aux_solution(self, uaux: np.ndarray) -> None:
        self._ck(self.L.hpb_dev_set_aux_solution(self.h, _dp(np.ascontiguousarray(uaux, dtype=np.float64).ravel())))

    def dev_GLMGEEErrorSums(self, uex: Optional[np.ndarray] = None) -> np.ndarray:
        out = np.zeros(9)
        self._ck(self.L.hpb_dev_GLMGEEErrorSums(self.h, _dp(uex) if uex is not None else None, _dp(out)))
        return out

    def glmgee_error(self, uex: Optional[np.ndarray] = None, sums: Optional[np.ndarray] = None) -> np.ndarray:
        """the six numbers TimeError.c:43-127 writes to glm_err.dat after dt (L1, L2, Linf of the estimated error, then of
        (u - uex) - estimate, or -1 without uex). `sums`: the nine sums already reduced over the ranks (sum, sum, max)."""
        s = self.dev_GLMGEEErrorSums(uex) if sums is None else np.asarray(sums, dtype=np.float64)
        npg = float(np.prod(self.dim_global))
        norm = lambda t: np.array([t[0] / npg, np.sqrt(t[1] / npg), t[2]])
        sol, err = norm(s[0:3]), np.empty(6)
        err[0:3] = norm(s[3:6])
        err[3:6] = norm(s[6:9]) if uex is not None else -1.0
        if (sol > 1e-15).all():
            err[0:3] /= sol
            if uex is not None:
                err[3:6] /= sol
        return err

    def synchronize(self) -> None:
        self._ck(self.L.hpb_synchronize(self.h))

    @property
    def stream(self) -> int:
        return int(self.L.hpb_stream(self.h) or 0)

    @property
    def kernel_launches(self) -> int:
        return int(self.L.hpb_kernel_launch_count(self.h))

    @property
    def tma_launches(self) -> int:
        """launches of the TMA-fed fused sweep (sweep_tma.cuh) so far"""
        return int(self.L.hpb_tma_launch_count(self.h))

    @property
    def nstages(self) -> int:
        return int(self.L.hpb_nstages(self.h))

    @property
    def time(self) -> float:
        return float(self.L.hpb_current_time(self.h))

    PROF = {"sweep_x": 0, "sweep_y": 1, "sweep_z": 2, "viscous": 3, "rk": 4, "bc": 5, "halo": 6, "other": 7, "sweep_fused": 8}

    def fp64_issue_peak(self) -> float:
        """FP64 thread-instructions per second this device issues at most (measured live)"""
        v = C.c_double()
        self._ck(self.L.hpb_fp64_issue_peak(self.h, C.byref(v)))
        return v.value

    def profile_enable(self, on: bool = True) -> None:
        self._ck(self.L.hpb_profile_enable(self.h, int(on)))

    def profile_query(self) -> Dict[str, tuple]:
        """{category: (total ms, launch groups)} since profile_enable(True)."""
        out = {}
        for name, cat in self.PROF.items():
            ms, n = C.c_double(), C.c_longlong()
            self._ck(self.L.hpb_profile_query(self.h, cat, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    # ---- in-library halo exchange (include/hypar_b200.h: hpb_comm_*, hpb_*Distributed)
    def exchange_plan(self, slot: int = 0):
        """[("send" | "recv", face 2*d + side, peer rank, doubles)] in issue order (host logic: needs no device)"""
        ops, counts, n = (C.c_int * 72)(), (C.c_longlong * 24)(), C.c_int()
        self._ck(self.L.hpb_exchange_plan(self.h, slot, ops, counts, C.byref(n)))
        return [("recv" if ops[3 * i] else "send", ops[3 * i + 1], ops[3 * i + 2], counts[i]) for i in range(n.value)]

    def comm_init_nccl(self, unique_id: bytes, nranks: int) -> None:
        self._ck(self.L.hpb_comm_init_nccl(self.h, unique_id, nranks))

    def set_overlap(self, on: bool) -> None:
        self._ck(self.L.hpb_set_overlap(self.h, int(bool(on))))

    def set_stage_fusion(self, on: bool) -> None:
        """the last sweep of an RK stage also writes the next stage solution (default on where it applies)"""
        self._ck(self.L.hpb_set_stage_fusion(self.h, int(bool(on))))

    @property
    def stage_fusion_active(self) -> bool:
        return bool(self.L.hpb_stage_fusion_active(self.h))

    def TimeStepsDistributed(self, n: int = 1) -> None:
        self._ck(self.L.hpb_TimeStepsDistributed(self.h, n))

    def RHSFunctionDistributed(self) -> None:
        self._ck(self.L.hpb_RHSFunctionDistributed(self.h))

    def comm_allreduce(self, values, op: str = "sum") -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        self._ck(self.L.hpb_comm_allreduce(self.h, _dp(v), v.size, 1 if op == "max" else 0))
        return v

    def comm_stats(self):
        m, b = C.c_longlong(), C.c_longlong()
        self._ck(self.L.hpb_comm_stats(self.h, C.byref(m), C.byref(b)))
        return m.value, b.value


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it and broadcasts the 128 bytes)"""
    L = _lib.load()
    buf = C.create_string_buffer(128)
    if L.hpb_comm_get_unique_id(buf) != 0:
        raise HyParB200Error(L.hpb_last_error().decode())
    return buf.raw


def comm_init_local(solvers) -> None:
    """in-process transport: `solvers` = every rank of the decomposition, indexed by rank"""
    L = _lib.load()
    arr = (C.c_void_p * len(solvers))(*[sv.h for sv in solvers])
    if L.hpb_comm_init_local(arr, len(solvers)) != 0:
        raise HyParB200Error(L.hpb_last_error().decode())

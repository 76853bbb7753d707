"""ctypes front-end of the CPU oracle (oracle/hypar_oracle.c). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. It also restates, in numpy, the reference's *set-up* logic that feeds the
hot path (domain partition, ghost coordinates, dxinv, boundary-zone extents), so that the
product's own C host code can be checked against an independent implementation:

  partition      src/MPIFunctions/MPIPartition1D.c, MPIRanknD.c / MPIRank1D.c
  x ghosts       src/IOFunctions/ReadArray.c:60-100
  dxinv          src/Simulation/InitialSolution.c:74-119
  zone extents   src/Simulation/InitializeBoundaries.c:380-440, src/MathFunctions/FindInterval.c
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAXD, MAXV, MAXZ = 3, 5, 16

MODELS = {"linear-advection-diffusion-reaction": 0, "euler1d": 1, "navierstokes2d": 2, "navierstokes3d": 3, "burgers": 4}
BCTYPES = {"periodic": 0, "extrapolate": 1, "slip-wall": 2, "noslip-wall": 3, "dirichlet": 4, "subsonic-inflow": 5,
           "subsonic-outflow": 6, "subsonic-ambivalent": 7, "supersonic-inflow": 8, "supersonic-outflow": 9, "sponge": 10}
UPWINDS = {"default": 0, "roe": 1, "rusanov": 2, "rf-char": 3, "llf-char": 4}
SCHEMES = {"weno5": 0, "crweno5": 1, "cupw5": 2, "upw5": 3, "1": 4, "2": 5, "4": 6, "muscl2": 7, "muscl3": 8, "hcweno5": 9}
LIMITERS = {"gmm": 0, "minmod": 1, "vanleer": 2, "superbee": 3}


class Zone(C.Structure):
    _fields_ = [("type", C.c_int), ("dim", C.c_int), ("face", C.c_int),
                ("is_", C.c_int * MAXD), ("ie", C.c_int * MAXD), ("on_this_proc", C.c_int),
                ("wall_vel", C.c_double * MAXD), ("rho", C.c_double), ("pressure", C.c_double),
                ("values", C.c_double * MAXV), ("xstart", C.c_double), ("xend", C.c_double)]


class Ctx(C.Structure):
    _fields_ = [("ndims", C.c_int), ("nvars", C.c_int), ("ghosts", C.c_int),
                ("dim", C.c_int * MAXD), ("iproc", C.c_int * MAXD), ("ip", C.c_int * MAXD),
                ("periodic", C.c_int * MAXD),
                ("model", C.c_int), ("weno_type", C.c_int), ("no_limiting", C.c_int),
                ("interp_char", C.c_int), ("upwind", C.c_int), ("par_order", C.c_int),
                ("HB", C.c_int), ("nzones", C.c_int),
                ("weno_eps", C.c_double), ("gamma", C.c_double), ("Re", C.c_double), ("Pr", C.c_double),
                ("grav", C.c_double * MAXD), ("rho0", C.c_double), ("p0", C.c_double), ("R", C.c_double),
                ("N_bv", C.c_double),
                ("adv", C.c_double * (MAXD * MAXV)), ("diff", C.c_double * (MAXD * MAXV)),
                ("zones", Zone * MAXZ),
                ("x", C.POINTER(C.c_double)), ("dxinv", C.POINTER(C.c_double)),
                ("grav_f", C.POINTER(C.c_double)), ("grav_g", C.POINTER(C.c_double)),
                ("scheme", C.c_int), ("muscl_limiter", C.c_int), ("muscl_eps", C.c_double), ("grav_type", C.c_int),
                ("adv_field", C.POINTER(C.c_double)), ("weno_rc", C.c_double), ("weno_xi", C.c_double)]


_lib = None


def build() -> str:
    so = os.path.join(HERE, "libhypar_oracle.so")
    src = os.path.join(HERE, "hypar_oracle.c")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        assert _lib.hpo_sizeof_ctx() == C.sizeof(Ctx), "oracle ctx layout mismatch"
        dp, cp = C.POINTER(C.c_double), C.POINTER(Ctx)
        for name in ("hpo_npoints_wghosts", "hpo_ninterfaces", "hpo_face_count"):
            getattr(_lib, name).restype = C.c_long
        _lib.hpo_cfl.restype = C.c_double
        _lib.hpo_sumsq_diff.restype = C.c_double
        _lib.hpo_cfl.argtypes = [cp, dp, C.c_double]
        _lib.hpo_time_step.argtypes = [cp, dp, C.c_double, C.c_int, C.c_int]
        _lib.hpo_time_step_cons.argtypes = [cp, dp, C.c_double, C.c_int, C.c_int, dp]
        _lib.hpo_volume_integral.argtypes = [cp, dp, dp]
        _lib.hpo_boundary_integral.argtypes = [cp, dp, dp]
        _lib.hpo_conservation_error.argtypes = [C.c_int, dp, dp, dp, dp]
        _lib.hpo_set_boundary_flux_sink.argtypes = [dp]
        _lib.hpo_norm_sums.argtypes = [cp, dp, dp, dp]
        _lib.hpo_time_step_glmgee.argtypes = [cp, dp, dp, C.c_double, C.c_int, C.c_int, C.c_int]
        _lib.hpo_glmgee_error.argtypes = [cp, dp, dp, C.c_int, C.c_int, dp, dp]
        _lib.hpo_glmgee_gamma.restype = C.c_double
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ----------------------------------------------------------------------------- set-up logic
def partition1d(nglobal: int, nproc: int, rank: int) -> int:
    """MPIPartition1D.c: nglobal/nproc, remainder on the last rank."""
    n = nglobal // nproc
    if rank == nproc - 1:
        n = nglobal - n * (nproc - 1)
    return n


def rank_nd(iproc: Sequence[int], rank: int) -> List[int]:
    """MPIRanknD.c: rank = ip0 + iproc0*(ip1 + iproc1*ip2)."""
    ip = []
    for n in iproc:
        ip.append(rank % n)
        rank //= n
    return ip


def rank_1d(iproc: Sequence[int], ip: Sequence[int]) -> int:
    r, f = 0, 1
    for n, i in zip(iproc, ip):
        r += f * i
        f *= n
    return r


def local_start(nglobal: int, nproc: int, rank: int) -> int:
    return (nglobal // nproc) * rank


def find_interval(a: float, b: float, x: np.ndarray):
    """FindInterval.c"""
    n = len(x)
    dxs = np.diff(x)
    tol = 1e-10 * (dxs.min() if len(dxs) else 1.0)
    imax, imin = -1, n
    for i in range(n):
        if x[i] <= b + tol:
            imax = i + 1
    for i in range(n - 1, -1, -1):
        if x[i] >= a - tol:
            imin = i
    return imin, imax


class Setup:
    """Everything one rank needs: local sizes, x and dxinv with ghosts, zone extents, gravity
    field, and the oracle context."""

    def __init__(self, case, rank: int = 0, mpi_semantics: bool = True):
        s = case.solver
        self.case = case
        self.ndims, self.nvars, self.ghosts = int(s["ndims"]), int(s["nvars"]), int(s["ghost"])
        nd, g = self.ndims, self.ghosts
        self.dim_global = [int(v) for v in s["size"]]
        self.iproc = [int(v) for v in s["iproc"]]
        self.rank = rank
        self.ip = rank_nd(self.iproc, rank)
        self.dim = [partition1d(self.dim_global[d], self.iproc[d], self.ip[d]) for d in range(nd)]
        self.is_ = [local_start(self.dim_global[d], self.iproc[d], self.ip[d]) for d in range(nd)]
        self.mpi_semantics = mpi_semantics
        self.periodic = [0] * nd
        for z in case.boundary:
            # mpi->bcperiodic is set ONLY when iproc[dim] > 1 (InitializeBoundaries.c:199-201); with one
            # rank along dim, periodicity of u is BCPeriodicU's job and NOTHING makes the viscous
            # derivative arrays periodic -- the serial and 1-rank MPI builds therefore agree
            if z["type"] == "periodic" and self.iproc[z["dim"]] > 1:
                self.periodic[z["dim"]] = 1

        # x with ghosts (ReadArray.c:60-100): interior from the global grid, ghosts from the
        # neighbour rank if there is one, else linear extrapolation (also on periodic faces)
        xs, dxs = [], []
        for d in range(nd):
            n, i0, xg = self.dim[d], self.is_[d], np.asarray(case.x[d], dtype=np.float64)
            xl = np.zeros(n + 2 * g)
            xl[g:g + n] = xg[i0:i0 + n]
            if self.ip[d] == 0:
                for i in range(g):
                    delta = g - i
                    xl[i] = xl[g] + float(delta) * (xl[g] - xl[g + 1])
            else:
                xl[:g] = xg[i0 - g:i0]
            if self.ip[d] == self.iproc[d] - 1:
                for i in range(n + g, n + 2 * g):
                    delta = i - (n + g - 1)
                    xl[i] = xl[n + g - 1] + float(delta) * (xl[n + g - 1] - xl[n + g - 2])
            else:
                xl[n + g:] = xg[i0 + n:i0 + n + g]
            xs.append(xl)
        # dxinv (InitialSolution.c:74-119): interior 2/(x[i+1]-x[i-1]); internal-face ghosts are the
        # neighbour's interior values; physical-face ghosts copy the first/last interior value
        for d in range(nd):
            n = self.dim[d]
            # neighbour values: recompute from the global grid with the same formula the owner uses
            xg = np.asarray(case.x[d], dtype=np.float64)
            ng = len(xg)
            xge = np.zeros(ng + 2 * g)
            xge[g:g + ng] = xg
            for i in range(g):
                xge[i] = xg[0] + float(g - i) * (xg[0] - xg[1])
                xge[ng + g + i] = xg[-1] + float(i + 1) * (xg[-1] - xg[-2])
            dg = np.zeros(ng + 2 * g)
            # each owner computes dxinv from ITS local x (whose ghosts are the true neighbours'
            # coordinates on internal faces and extrapolations on physical faces)
            dg[g:g + ng] = 2.0 / (xge[g + 1:g + ng + 1] - xge[g - 1:g + ng - 1])
            dl = np.zeros(n + 2 * g)
            i0 = self.is_[d]
            dl[:] = dg[i0:i0 + n + 2 * g]
            if self.ip[d] == 0:
                dl[:g] = dl[g]
            if self.ip[d] == self.iproc[d] - 1:
                dl[n + g:] = dl[n + g - 1]
            dxs.append(dl)
        self.x = np.ascontiguousarray(np.concatenate(xs))
        self.dxinv = np.ascontiguousarray(np.concatenate(dxs))
        self.xs, self.dxs = xs, dxs

        # boundary zone extents (InitializeBoundaries.c:380-440)
        self.zones = []
        for z in case.boundary:
            dim, face = int(z["dim"]), int(z["face"])
            zi = {"type": z["type"], "dim": dim, "face": face, "is": [0] * nd, "ie": [0] * nd,
                  "on": 0, "wall_vel": list(z.get("wall_velocity", z.get("velocity", [0.0] * nd))),
                  "rho": float(z.get("density", 0.0)), "pressure": float(z.get("pressure", 0.0)),
                  "values": list(z.get("values", []))}
            zi["xstart"], zi["xend"] = float(z["xmin"][dim]), float(z["xmax"][dim])
            edge = (self.ip[dim] == 0) if face == 1 else (self.ip[dim] == self.iproc[dim] - 1)
            if z["type"] == "sponge":       # InitializeBoundaries.c:381-395: an interior box, FindInterval in every dimension
                zi["on"] = 1
                for d in range(nd):
                    a, b = find_interval(z["xmin"][d], z["xmax"][d], xs[d][g:g + self.dim[d]])
                    zi["is"][d], zi["ie"][d] = a, b
                    if b - a <= 0:
                        zi["on"] = 0
            elif edge:
                zi["on"] = 1
                for d in range(nd):
                    if d == dim:
                        if face == 1:
                            zi["is"][d], zi["ie"][d] = -g, 0
                        else:
                            zi["is"][d], zi["ie"][d] = self.dim[d], self.dim[d] + g
                    else:
                        a, b = find_interval(z["xmin"][d], z["xmax"][d], xs[d][g:g + self.dim[d]])
                        zi["is"][d], zi["ie"][d] = a, b
                        if b - a <= 0:
                            zi["on"] = 0
            self.zones.append(zi)

        self.ctx = self._make_ctx()

    # -- sizes
    @property
    def npoints_g(self) -> int:
        return int(np.prod([n + 2 * self.ghosts for n in self.dim]))

    def shape_g(self):
        """numpy shape of a ghost-padded cell array: (N_{nd-1}+2g, ..., N_0+2g, nvars)."""
        return tuple(n + 2 * self.ghosts for n in reversed(self.dim)) + (self.nvars,)

    def interior(self, a: np.ndarray) -> np.ndarray:
        g = self.ghosts
        sl = tuple(slice(g, g + n) for n in reversed(self.dim))
        return a.reshape(self.shape_g())[sl]

    def ninterfaces(self, d: int) -> int:
        return int(np.prod([n + (1 if k == d else 0) for k, n in enumerate(self.dim)]))

    def local_u0(self) -> np.ndarray:
        """Initial solution on this rank, ghost-padded AoS (ghosts zero)."""
        g = self.ghosts
        u = np.zeros(self.shape_g())
        src = tuple(slice(self.is_[d], self.is_[d] + self.dim[d]) for d in reversed(range(self.ndims)))
        dst = tuple(slice(g, g + n) for n in reversed(self.dim))
        u[dst] = self.case.u0[src]
        return np.ascontiguousarray(u).reshape(-1)

    def _local_adv_field(self, af: np.ndarray) -> np.ndarray:
        nd, g, nc = self.ndims, self.ghosts, self.ndims * self.nvars
        shp = tuple(n + 2 * g for n in reversed(self.dim)) + (nc,)
        a = np.zeros(shp)
        axis = lambda d: nd - 1 - d                                        # numpy axis of dimension d
        inner = [slice(g, g + self.dim[d]) for d in range(nd)]            # per dimension
        glob = [slice(self.is_[d], self.is_[d] + self.dim[d]) for d in range(nd)]
        put = lambda loc, gl: a.__setitem__(tuple(loc[d] for d in reversed(range(nd))),
                                            af[tuple(gl[d] for d in reversed(range(nd)))])
        put(inner, glob)
        # internal (MPI) faces, periodic wrap included when iproc > 1
        for d in range(nd):
            N, n, i0 = self.dim_global[d], self.dim[d], self.is_[d]
            for side in (0, 1):
                if side == 0:
                    has = self.ip[d] > 0 or (self.periodic[d] and self.iproc[d] > 1)
                    gidx = [(i0 - g + k) % N for k in range(g)]
                    loc_sl = slice(0, g)
                else:
                    has = self.ip[d] < self.iproc[d] - 1 or (self.periodic[d] and self.iproc[d] > 1)
                    gidx = [(i0 + n + k) % N for k in range(g)]
                    loc_sl = slice(g + n, g + n + g)
                if not has:
                    continue
                loc = list(inner); loc[d] = loc_sl
                gl = list(glob); gl[d] = gidx
                put(loc, gl)
        # physical faces (LinearADRAdvectionField.c:118-190), dimension by dimension
        per = {int(z["dim"]) for z in self.case.boundary if z["type"] == "periodic"}
        for d in range(nd):
            n = self.dim[d]
            ax = axis(d)
            def view(lo, hi, rev=False):
                idx = [slice(g, g + self.dim[k]) for k in reversed(range(nd))] + [slice(None)]
                idx[ax] = slice(lo, hi)
                v = a[tuple(idx)]
                return np.flip(v, axis=ax) if rev else v
            def assign(lo, hi, val):
                idx = [slice(g, g + self.dim[k]) for k in reversed(range(nd))] + [slice(None)]
                idx[ax] = slice(lo, hi)
                a[tuple(idx)] = val
            if d in per and self.iproc[d] == 1:
                assign(0, g, view(n, n + g).copy())              # left ghosts <- last g interior
                assign(g + n, g + n + g, view(g, 2 * g).copy())  # right ghosts <- first g interior
            else:
                if self.ip[d] == 0:
                    assign(0, g, view(g, 2 * g, rev=True).copy())                 # mirror: ghost -1-k <- interior k
                if self.ip[d] == self.iproc[d] - 1:
                    assign(g + n, g + n + g, view(n, n + g, rev=True).copy())
        return np.ascontiguousarray(a).reshape(-1)

    def _make_ctx(self) -> Ctx:
        case, s, ph = self.case, self.case.solver, self.case.physics
        w = case.weno or {}
        c = Ctx()
        c.ndims, c.nvars, c.ghosts = self.ndims, self.nvars, self.ghosts
        for d in range(self.ndims):
            c.dim[d], c.iproc[d], c.ip[d], c.periodic[d] = self.dim[d], self.iproc[d], self.ip[d], self.periodic[d]
        c.model = MODELS[s["model"]]
        c.scheme = SCHEMES[str(s.get("hyp_space_scheme", "weno5"))]
        mu = getattr(case, "muscl", None) or {}
        c.muscl_eps = float(mu.get("epsilon", 1e-3))                       # MUSCLInitialize.c:26-27 defaults
        c.muscl_limiter = LIMITERS.get(str(mu.get("limiter", "gmm")), 0)   # :72-75: unknown names fall back to gmm
        c.weno_type = 3 if int(w.get("yc", 0)) else 2 if int(w.get("borges", 0)) else 1 if int(w.get("mapped", 0)) else 0
        c.no_limiting = int(w.get("no_limiting", 0))
        c.weno_eps = float(w.get("epsilon", 1e-6))
        c.weno_rc, c.weno_xi = float(w.get("rc", 0.3)), float(w.get("xi", 0.001))     # WENOInitialize.c:57-58
        c.interp_char = int(s.get("hyp_interp_type", "characteristic") == "characteristic")
        up = ph.get("upwinding", "roe" if c.model in (1, 3) else "default")
        c.upwind = UPWINDS.get(up, 0) if c.model != 0 else 0
        c.par_order = int(s.get("par_space_scheme", "2"))
        c.gamma = float(ph.get("gamma", 1.4))
        Re, Minf = float(ph.get("Re", -1.0)), float(ph.get("Minf", 1.0))
        c.Re = Re / Minf
        c.Pr = float(ph.get("Pr", 0.72))
        grav = ph.get("gravity", [0.0, 0.0, 0.0])
        if not isinstance(grav, (list, tuple)):
            grav = [float(grav), 0.0, 0.0]
        for d in range(3):
            c.grav[d] = float(grav[d]) if d < len(grav) else 0.0
        c.rho0, c.p0, c.R = float(ph.get("rho_ref", 1.0)), float(ph.get("p_ref", 1.0)), float(ph.get("R", 1.0))
        c.HB, c.N_bv = int(ph.get("HB", 1)), float(ph.get("N_bv", 0.0))
        adv = ph.get("advection", [])
        adv = list(adv) if isinstance(adv, (list, tuple)) else [adv]
        for i, a in enumerate(adv):
            c.adv[i] = float(a)
        dif = ph.get("diffusion", [])
        if c.model == 0 and str(s.get("par_space_type", "nonconservative-1stage")) != "nonconservative-1stage" and \
                any(float(v) != 0.0 for v in (dif if isinstance(dif, (list, tuple)) else [dif])):
            raise NotImplementedError("oracle: LinearADR diffusion is restated as ParabolicFunctionNC1Stage only")
        dif = list(dif) if isinstance(dif, (list, tuple)) else [dif]
        for i, a in enumerate(dif):
            c.diff[i] = float(a)
        c.nzones = len(self.zones)
        for n, z in enumerate(self.zones):
            cz = c.zones[n]
            cz.type, cz.dim, cz.face, cz.on_this_proc = BCTYPES[z["type"]], z["dim"], z["face"], z["on"]
            for d in range(self.ndims):
                cz.is_[d], cz.ie[d], cz.wall_vel[d] = z["is"][d], z["ie"][d], z["wall_vel"][d]
            cz.rho, cz.pressure = z["rho"], z["pressure"]
            cz.xstart, cz.xend = z["xstart"], z["xend"]
            for v, val in enumerate(z["values"]):
                cz.values[v] = float(val)
        c.x, c.dxinv = _p(self.x), _p(self.dxinv)
        # gravity field
        n = self.npoints_g
        self.grav_f, self.grav_g = np.ones(n), np.ones(n)
        c.grav_f, c.grav_g = _p(self.grav_f), _p(self.grav_g)
        c.grav_type = int(ph.get("gravity_type", 0))
        # LinearADR spatially varying advection (LinearADRAdvectionField.c:25-193): this rank's block of the field with
        # ghosts -- interior from the file; internal faces from the neighbour (MPIExchangeBoundariesnD = the global field,
        # wrapped where the dimension is periodic and split); physical faces: periodic copy (one rank along the dimension)
        # or mirror extrapolation; edges / corners stay zero
        self.adv_field = None
        af = getattr(case, "advection_field", None)
        if c.model == 0 and af is not None:
            self.adv_field = self._local_adv_field(np.asarray(af, dtype=np.float64))
            c.adv_field = _p(self.adv_field)
        if c.model in (2, 3):       # NavierStokes2D / 3D gravity field (identical to 1 without gravity, HB 1)
            lib().hpo_ns3d_gravity_field(C.byref(c), _p(self.grav_f), _p(self.grav_g))
        elif c.model == 1:          # Euler1D: one field (exp(0) = 1 without gravity)
            lib().hpo_e1d_gravity_field(C.byref(c), _p(self.grav_f), _p(self.grav_g))
        return c


# ----------------------------------------------------------------------------- thin call wrappers
class Oracle:
    def __init__(self, setup: Setup):
        self.s, self.c, self.L = setup, C.byref(setup.ctx), lib()

    def zeros(self):
        return np.zeros(self.s.npoints_g * self.s.nvars)

    def weights_size(self, d=None):
        if d is None:
            return sum(12 * self.s.ninterfaces(k) * self.s.nvars for k in range(self.s.ndims))
        return 12 * self.s.ninterfaces(d) * self.s.nvars

    def apply_bc(self, u):
        self.L.hpo_apply_bc(self.c, _p(u))
        return u

    def exchange_self(self, a, nvars=None):
        self.L.hpo_exchange_self(self.c, _p(a), C.c_int(nvars or self.s.nvars))
        return a

    def flux(self, u, d):
        f = self.zeros(); self.L.hpo_flux(self.c, _p(u), _p(f), C.c_int(d)); return f

    def modified_solution(self, u):
        f = self.zeros(); self.L.hpo_modified_solution(self.c, _p(u), _p(f)); return f

    def weno_weights(self, fC, u, d):
        w = np.zeros(self.weights_size(d)); self.L.hpo_weno_weights(self.c, _p(fC), _p(u), C.c_int(d), _p(w)); return w

    def interp(self, fC, u, w, upw, d, uflag):
        fI = np.zeros(self.s.ninterfaces(d) * self.s.nvars)
        self.L.hpo_interp(self.c, _p(fI), _p(fC), _p(u), _p(w), C.c_int(upw), C.c_int(d), C.c_int(uflag)); return fI

    def upwind(self, fL, fR, uL, uR, u, d):
        fI = np.zeros_like(fL)
        self.L.hpo_upwind(self.c, _p(fI), _p(fL), _p(fR), _p(uL), _p(uR), _p(u), C.c_int(d)); return fI

    def hyperbolic(self, u, want_weights=False):
        hyp = self.zeros(); w = np.zeros(self.weights_size())
        self.L.hpo_hyperbolic(self.c, _p(hyp), _p(u), _p(w))
        return (hyp, w) if want_weights else hyp

    def first_derivative(self, f, d):
        Df = self.zeros(); self.L.hpo_first_derivative(self.c, _p(Df), _p(f), C.c_int(d)); return Df

    def second_derivative(self, f, d, order):
        Df = self.zeros(); self.L.hpo_second_derivative(self.c, _p(Df), _p(f), C.c_int(d), C.c_int(order)); return Df

    def parabolic(self, u, mpi_semantics=None):
        ms = self.s.mpi_semantics if mpi_semantics is None else mpi_semantics
        par = self.zeros()
        if self.s.ctx.model in (2, 3):
            self.L.hpo_ns_parabolic(self.c, _p(par), _p(u), C.c_int(int(ms)))
        else:
            self.L.hpo_parabolic_nc1(self.c, _p(par), _p(u))
        return par

    def parabolic_p1(self, u):
        """NavierStokes parabolic term up to the halo exchange: Q and the un-scaled derivatives."""
        Q, QDx, QDy, QDz = self.zeros(), self.zeros(), self.zeros(), self.zeros()
        self.L.hpo_ns_parabolic_p1(self.c, _p(u), _p(Q), _p(QDx), _p(QDy), _p(QDz))
        return Q, QDx, QDy, QDz

    def parabolic_p2(self, Q, QDx, QDy, QDz):
        """... and after it (QDx/QDy/QDz are scaled in place)."""
        par = self.zeros()
        self.L.hpo_ns_parabolic_p2(self.c, _p(Q), _p(QDx), _p(QDy), _p(QDz), _p(par))
        return par

    def source(self, u, w):
        src = self.zeros()
        if self.s.ctx.model in (1, 2, 3):
            self.L.hpo_ns3d_source(self.c, _p(src), _p(u), _p(w))
        self.L.hpo_sponge_source(self.c, _p(src), _p(u))
        return src

    def rhs(self, u, parts=False, mpi_semantics=None):
        ms = self.s.mpi_semantics if mpi_semantics is None else mpi_semantics
        rhs, hyp, par, src = self.zeros(), self.zeros(), self.zeros(), self.zeros()
        self.L.hpo_rhs(self.c, _p(rhs), _p(u), _p(hyp), _p(par), _p(src), C.c_int(int(ms)))
        return (rhs, hyp, par, src) if parts else rhs

    def time_step(self, u, dt, rk_type, mpi_semantics=None):
        ms = self.s.mpi_semantics if mpi_semantics is None else mpi_semantics
        self.L.hpo_time_step(self.c, _p(u), C.c_double(dt), C.c_int(rk_type), C.c_int(int(ms)))
        return u

    def cfl(self, u, dt):
        return float(self.L.hpo_cfl(self.c, _p(u), C.c_double(dt)))

    # ---- GLM-GEE (TimeGLMGEE.c): the solution and one auxiliary solution
    def glmgee_aux0(self, u, mode):
        """TimeInitialize.c:156-169: the auxiliary solution before the first step"""
        return u.copy() if mode == 1 else np.zeros_like(u)

    def time_step_glmgee(self, u, uaux, dt, method, mode, mpi_semantics=None):
        ms = self.s.mpi_semantics if mpi_semantics is None else mpi_semantics
        self.L.hpo_time_step_glmgee(self.c, _p(u), _p(uaux), C.c_double(dt), C.c_int(method), C.c_int(mode), C.c_int(int(ms)))
        return u, uaux

    def glmgee_error(self, u, uaux, method, mode, uex=None):
        """the six numbers TimeError.c writes to glm_err.dat (after dt)"""
        out = np.zeros(6)
        self.L.hpo_glmgee_error(self.c, _p(u), _p(uaux), C.c_int(method), C.c_int(mode), _p(uex) if uex is not None else None, _p(out))
        return out

    # ---- conservation diagnostics (VolumeIntegral.c, BoundaryIntegral.c, CalculateConservationError.c, the
    #      StageBoundaryIntegral bookkeeping of HyperbolicFunction.c:103-106 / TimeRK.c:172-193); local (this rank's) parts
    def stage_boundary_integral(self, u, mpi_semantics=None):
        """StageBoundaryIntegral left by one TimeRHSFunctionExplicit(u): [(2d+face)*nvars+v]"""
        sbi = np.zeros(2 * self.s.ndims * self.s.nvars)
        self.L.hpo_set_boundary_flux_sink(_p(sbi))
        try:
            self.rhs(u, mpi_semantics=mpi_semantics)
        finally:
            self.L.hpo_set_boundary_flux_sink(None)
        return sbi

    def time_step_cons(self, u, dt, rk_type, mpi_semantics=None):
        """one step; returns StepBoundaryIntegral of that step"""
        ms = self.s.mpi_semantics if mpi_semantics is None else mpi_semantics
        sbi = np.zeros(2 * self.s.ndims * self.s.nvars)
        self.L.hpo_time_step_cons(self.c, _p(u), C.c_double(dt), C.c_int(rk_type), C.c_int(int(ms)), _p(sbi))
        return sbi

    def volume_integral(self, u):
        out = np.zeros(self.s.nvars); self.L.hpo_volume_integral(self.c, _p(u), _p(out)); return out

    def boundary_integral(self, step_bi):
        out = np.zeros(self.s.nvars); self.L.hpo_boundary_integral(self.c, _p(np.ascontiguousarray(step_bi)), _p(out)); return out

    def conservation_error(self, vol, vol0, total_bi):
        err = np.zeros(self.s.nvars)
        self.L.hpo_conservation_error(C.c_int(self.s.nvars), _p(np.ascontiguousarray(vol)), _p(np.ascontiguousarray(vol0)),
                                      _p(np.ascontiguousarray(total_bi)), _p(err))
        return err

    def norm_sums(self, a, b=None):
        """(sum |a-b|, sum (a-b)^2, max |a-b|) over the interior (CalculateError.c:68-104, this rank's part)"""
        out = np.zeros(3)
        self.L.hpo_norm_sums(self.c, _p(a), _p(b) if b is not None else None, _p(out))
        return out

    def pack(self, a, d, side, nvars=None):
        nv = nvars or self.s.nvars
        n = self.L.hpo_face_count(self.c, C.c_int(d), C.c_int(nv))
        buf = np.zeros(n)
        self.L.hpo_pack(self.c, _p(a), C.c_int(nv), C.c_int(d), C.c_int(side), _p(buf)); return buf

    def unpack(self, a, d, side, buf, nvars=None):
        nv = nvars or self.s.nvars
        self.L.hpo_unpack(self.c, _p(a), C.c_int(nv), C.c_int(d), C.c_int(side), _p(buf)); return a


RK_TYPES = {"44": 0, "ssprk3": 1, "tvdrk3": 1, "1fe": 2, "22": 3, "33": 4}


GLMGEE_METHODS = {"23": 0, "24": 1, "25i": 2, "35": 3, "exrk2a": 4, "rk32g1": 5, "rk285ex": 6}
GLMGEE_MODES = {"yeps": 0, "yyt": 1}


def glmgee_of(case):
    """(method, mode) of a case advanced by time_scheme glm-gee, else None"""
    if str(case.solver.get("time_scheme", "rk")) != "glm-gee":
        return None
    gg = getattr(case, "glm_gee", None) or {}
    return GLMGEE_METHODS[str(case.solver["time_scheme_type"])], GLMGEE_MODES[str(gg.get("ee_mode", "yeps"))]


def glmgee_table(method: int, mode: int):
    """the coefficient table the oracle uses (oracle/glmgee_tables.h) as numpy arrays"""
    L = lib()
    s = L.hpo_glmgee_nstages(C.c_int(method))
    t = {"s": s, "A": np.zeros(s * s), "B": np.zeros(2 * s), "C": np.zeros(2 * s), "D": np.zeros(4)}
    L.hpo_glmgee_coefficients(C.c_int(method), C.c_int(mode), _p(t["A"]), _p(t["B"]), _p(t["C"]), _p(t["D"]))
    return t


def rk_type_of(case) -> int:
    """oracle tableau id of a case's time integrator; time_scheme "euler" (TimeForwardEuler.c) is RK "1fe" """
    if str(case.solver.get("time_scheme", "rk")) == "euler":
        return RK_TYPES["1fe"]
    return RK_TYPES[str(case.solver["time_scheme_type"])]

"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

    python tools/make_golden.py [fixture names ...]

Each fixture holds, for one small case of a BASELINE.json configuration family, the outputs of the
reference's own code on that case's input files:
  rhs mode   : u after ApplyBoundaryConditions, hyp, par, source, rhs of ONE TimeRHSFunctionExplicit
  steps mode : u after 3 time steps of the reference's TimePreStep/TimeStep/TimePostStep loop, and (run with
               conservation_check yes) the volume integral, boundary-flux integrals and conservation error of every step
  pieces mode: (selected cases) FFunction, WENO weights, uL/uR/fL/fR, Upwind result per direction
The case itself is re-created from hypar_b200.cases by name + arguments (stored in the fixture), so the
inputs are not duplicated. TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from hypar_b200 import cases  # noqa: E402
from refrun import run_reference  # noqa: E402

# (fixture name, builder, kwargs, exe, with pieces)
GOLDEN = [
    ("c1_linadv_js", "linear_advection_sine", dict(n=64, weno="js"), "hypar_ref", True),
    ("c1_linadv_mapped_diff4", "linear_advection_sine", dict(n=64, weno="mapped", diffusion=0.01, par_scheme="4"), "hypar_ref", False),
    ("c1_linadv_1024_mapped", "linear_advection_sine", dict(n=1024, weno="mapped"), "hypar_ref", False),
    ("c2_sod_js_char_roe", "euler1d_sod", dict(n=101, weno="js"), "hypar_ref", True),
    ("c2_sod_201_mapped_char_roe", "euler1d_sod", dict(n=201, weno="mapped"), "hypar_ref", False),
    ("c2_sod_z_comp_rusanov", "euler1d_sod", dict(n=101, weno="z", interp="components", upwinding="rusanov"), "hypar_ref", False),
    ("c3_vortex_yc", "ns2d_vortex", dict(n=(20, 16), weno="yc"), "hypar_ref", True),
    ("c3_vortex_mapped", "ns2d_vortex", dict(n=(28, 36), weno="mapped"), "hypar_ref", False),
    ("c4_turb_mapped_visc", "ns3d_turbulence", dict(n=(10, 8, 8), weno="mapped"), "hypar_ref_mpi1", True),
    ("c4_turb_js_roe_inv", "ns3d_turbulence", dict(n=(12, 10, 14), weno="js", upwinding="roe", viscous=False), "hypar_ref", False),
    ("c4_turb_z_char_inv", "ns3d_turbulence", dict(n=(12, 12, 10), weno="z", interp="characteristic", viscous=False), "hypar_ref", False),
    ("c2_sod_js_char_rf", "euler1d_sod", dict(n=101, weno="js", upwinding="rf-char"), "hypar_ref", True),
    ("c2_sod_yc_comp_llf", "euler1d_sod", dict(n=101, weno="yc", interp="components", upwinding="llf-char"), "hypar_ref", False),
    ("c4_turb_z_rf_inv", "ns3d_turbulence", dict(n=(12, 10, 14), weno="z", upwinding="rf-char", viscous=False), "hypar_ref", False),
    ("c4_turb_mapped_char_llf_visc", "ns3d_turbulence", dict(n=(10, 8, 8), weno="mapped", interp="characteristic", upwinding="llf-char"), "hypar_ref_mpi1", False),
    ("c5a_denswave_js", "ns3d_density_wave", dict(n=(12, 10, 8), weno="js"), "hypar_ref", False),
    ("c5b_bubble_yc", "ns3d_rising_bubble", dict(n=(12, 16, 10), weno="yc"), "hypar_ref_mpi1", True),
    ("c5b_bubble_mapped_hb1", "ns3d_rising_bubble", dict(n=(10, 12, 14), weno="mapped", hb=1), "hypar_ref", False),
    # compact schemes (one tridiagonal system per line and component) and the linear fifth-order upwind
    ("c1_linadv_crweno_mapped", "linear_advection_sine", dict(n=64, weno="mapped", scheme="crweno5"), "hypar_ref", True),
    ("c2_sod_crweno_js_comp_roe", "euler1d_sod", dict(n=101, weno="js", interp="components", scheme="crweno5"), "hypar_ref", False),
    ("c3_vortex_crweno_z", "ns2d_vortex", dict(n=(20, 16), weno="z", scheme="crweno5"), "hypar_ref_mpi1", True),
    ("c4_turb_crweno_mapped_visc", "ns3d_turbulence", dict(n=(10, 8, 8), weno="mapped", scheme="crweno5"), "hypar_ref_mpi1", False),
    ("c5b_bubble_crweno_yc", "ns3d_rising_bubble", dict(n=(12, 16, 10), weno="yc", scheme="crweno5"), "hypar_ref_mpi1", True),
    ("c5a_denswave_cupw5", "ns3d_density_wave", dict(n=(12, 10, 8), weno="js", scheme="cupw5"), "hypar_ref", True),
    ("c3_vortex_cupw5", "ns2d_vortex", dict(n=(28, 36), weno="js", scheme="cupw5"), "hypar_ref", False),
    # NavierStokes2D: Roe (+ Harten fix), characteristic Roe-fixed / local Lax-Friedrichs, characteristic WENO5
    ("c3_vortex_js_roe", "ns2d_vortex", dict(n=(20, 16), weno="js", upwinding="roe"), "hypar_ref", True),
    ("c3_vortex_mapped_char_rf", "ns2d_vortex", dict(n=(24, 20), weno="mapped", upwinding="rf-char", interp="characteristic"), "hypar_ref_mpi1", True),
    ("c3_vortex_z_llf", "ns2d_vortex", dict(n=(20, 24), weno="z", upwinding="llf-char"), "hypar_ref", False),
    # NavierStokes2D with gravity: well-balanced source, HB 2 / HB 3, Rusanov / llf-char / Roe
    ("c3g_bubble2d_js_hb2", "ns2d_rising_bubble", dict(n=(20, 24), weno="js"), "hypar_ref_mpi1", True),
    ("c3g_bubble2d_yc_hb1_llf", "ns2d_rising_bubble", dict(n=(24, 20), weno="yc", hb=1, upwinding="llf-char"), "hypar_ref", False),
    ("c3g_bubble2d_mapped_roe_crweno", "ns2d_rising_bubble", dict(n=(20, 20), weno="mapped", upwinding="roe", scheme="crweno5"), "hypar_ref", False),
    # inflow / outflow / wall / Dirichlet boundary zones
    ("chan2d_js_inflow_outflow_walls", "ns_channel", dict(n=(24, 20), weno="js"), "hypar_ref", False),
    ("chan2d_mapped_supersonic_dirichlet", "ns_channel", dict(n=(20, 24), weno="mapped", mach=1.6, bcs="sup"), "hypar_ref_mpi1", False),
    ("chan3d_z_ambivalent_visc", "ns_channel", dict(n=(12, 10, 10), weno="z", viscous=True, bcs="amb3"), "hypar_ref_mpi1", False),
    ("c2_sod_js_char_roe_gravity", "euler1d_sod", dict(n=101, weno="js", gravity=1.0), "hypar_ref", True),
    ("c2_sod_crweno_js_char_roe", "euler1d_sod", dict(n=101, weno="js", scheme="crweno5"), "hypar_ref", True),
    ("c3_vortex_crweno_z_char_roe", "ns2d_vortex", dict(n=(20, 16), weno="z", upwinding="roe", interp="characteristic", scheme="crweno5"), "hypar_ref_mpi1", False),
    ("burgers2d_z", "burgers_nd", dict(n=(24, 20), weno="z"), "hypar_ref", True),
    ("linadvvar2d_js", "linear_advection_varying", dict(n=(24, 20), weno="js"), "hypar_ref", True),
    ("linadvvar1d_mapped_ext", "linear_advection_varying", dict(n=(80,), weno="mapped", periodic=False, tstype="ssprk3"), "hypar_ref_mpi1", True),
    # hybrid compact-WENO5 (Interp1PrimFifthOrderHCWENO.c / ...HCWENOChar.c), default and non-default rc / xi
    ("c1_linadv_hcweno_z", "linear_advection_sine", dict(n=64, weno="z", scheme="hcweno5"), "hypar_ref", True),
    ("c3_vortex_hcweno_js_char_roe", "ns2d_vortex", dict(n=(20, 16), weno="js+rc0.5+xi0.01", upwinding="roe", interp="characteristic", scheme="hcweno5"), "hypar_ref_mpi1", True),
    ("c5b_bubble_hcweno_mapped", "ns3d_rising_bubble", dict(n=(12, 16, 10), weno="mapped+rc0.2", scheme="hcweno5"), "hypar_ref_mpi1", False),
    ("c2_sod_hcweno_yc_char_llf_gravity", "euler1d_sod", dict(n=101, weno="yc", upwinding="llf-char", gravity=1.0, scheme="hcweno5"), "hypar_ref", False),
    # GLM-GEE time integrators (TimeGLMGEE.c): the solution, the auxiliary solution and TimeError's norms after 3 steps
    ("glm_linadv_23_yeps", "glmgee", dict(base="linear_advection_sine", tstype="23", n=64, weno="js"), "hypar_ref", False),
    ("glm_vortex_exrk2a_yyt", "glmgee", dict(base="ns2d_vortex", tstype="exrk2a", ee_mode="yyt", n=(20, 16), weno="yc"), "hypar_ref_mpi1", False),
    ("glm_turb_rk32g1_yeps_visc", "glmgee", dict(base="ns3d_turbulence", tstype="rk32g1", n=(10, 8, 8), weno="mapped"), "hypar_ref_mpi1", False),
    ("glm_bubble_rk285ex_yyt", "glmgee", dict(base="ns3d_rising_bubble", tstype="rk285ex", ee_mode="yyt", n=(10, 12, 8), weno="js"), "hypar_ref", False),
    ("glm_sod_35_yyt", "glmgee", dict(base="euler1d_sod", tstype="35", ee_mode="yyt", n=101, weno="z"), "hypar_ref", False),
    ("c2_sod_upw5_comp_rusanov", "euler1d_sod", dict(n=101, weno="js", interp="components", upwinding="rusanov", scheme="upw5"), "hypar_ref", True),
]


def parse_conservation(stdout):
    """CONS0 / CONS / STEPBI lines of oracle/ref_harness.cpp (steps mode, conservation_check yes)"""
    r = {"vol0": None, "cons": [], "stepbi": []}
    for ln in stdout.splitlines():
        t = ln.split()
        if ln.startswith("CONS0"):
            r["vol0"] = np.array([float(x) for x in t[1:]])
        elif ln.startswith("CONS "):
            r["cons"].append([float(x) for x in t[2:]])
        elif ln.startswith("STEPBI "):
            r["stepbi"].append([float(x) for x in t[2:]])
    return r


def build_case(builder, kwargs):
    kw = dict(kwargs)
    if "n" in kw and isinstance(kw["n"], list):
        kw["n"] = tuple(kw["n"])
    return getattr(cases, builder)(**kw)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = set(sys.argv[1:])          # optional: fixture names to (re)generate; default all
    for name, builder, kwargs, exe, pieces in GOLDEN:
        if only and name not in only:
            continue
        case = build_case(builder, kwargs)
        data = {}
        o = run_reference(case, "rhs", exe=exe)
        for k in ("u", "hyp", "par", "source", "rhs", "x", "dxinv"):
            data["rhs_" + k] = o[k]["data"]
        case.solver["conservation_check"] = "yes"          # diagnostics only: the solution is unaffected
        o = run_reference(case, "steps", [3], exe=exe)
        case.solver["conservation_check"] = "no"
        data["steps3_u"] = o["ufinal"]["data"]
        if "uaux" in o:                                      # time_scheme glm-gee
            import re
            data["steps3_uaux"] = o["uaux"]["data"]
            data["steps3_glmerr"] = np.array([float(v) for v in re.search(r"GLMERR (.*)", o["stdout"]).group(1).split()])
        cons = parse_conservation(o["stdout"])
        data["cons_vol0"] = cons["vol0"]
        data["cons_steps"] = np.array(cons["cons"])          # per step: VolumeIntegral | TotalBoundaryIntegral | ConservationError
        data["cons_stepbi"] = np.array(cons["stepbi"])       # per step: StepBoundaryIntegral [(2d+face)*nvars+v]
        if pieces:
            o = run_reference(case, "pieces", exe=exe)
            for k, v in o.items():
                if isinstance(v, dict):
                    data["pieces_" + k] = v["data"]
                elif k == "cfl":
                    data["pieces_cfl"] = np.array([v])
        meta = {"builder": builder, "kwargs": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kwargs.items()},
                "exe": exe, "reference": "debog/hypar sources under /root/reference, gcc -O3 -std=c99 (oracle/Makefile)"}
        data["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        sub = os.path.join(out_dir, "glmgee") if builder == "glmgee" else out_dir      # own directory, own tests
        os.makedirs(sub, exist_ok=True)
        np.savez_compressed(os.path.join(sub, name + ".npz"), **data)
        print(name, {k: v.shape for k, v in data.items() if k != "meta"} if False else len(data))


if __name__ == "__main__":
    main()

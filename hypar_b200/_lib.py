"""ctypes binding of the C ABI (include/hypar_b200.h -> hypar_b200/libhypar_b200.so).

The shared library is the product; this module only declares its entry points. It fails
loudly when the library is missing: there is no Python or CPU fallback for any compute call.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# HYPAR_B200_LIB: another build of the same library (kernel-variant A/B measurements: make -C hypar_b200/csrc OUT=... BUILD=...)
LIB_PATH = os.environ.get("HYPAR_B200_LIB") or os.path.join(HERE, "libhypar_b200.so")

MAX_NDIMS, MAX_NVARS, MAX_ZONES = 3, 5, 16


class BoundaryZone(C.Structure):
    _fields_ = [("type", C.c_int), ("dim", C.c_int), ("face", C.c_int),
                ("xmin", C.c_double * MAX_NDIMS), ("xmax", C.c_double * MAX_NDIMS),
                ("wall_velocity", C.c_double * MAX_NDIMS),
                ("flow_density", C.c_double), ("flow_pressure", C.c_double), ("dirichlet", C.c_double * MAX_NVARS)]


class Config(C.Structure):
    _fields_ = [
        ("ndims", C.c_int), ("nvars", C.c_int), ("ghosts", C.c_int),
        ("dim_global", C.c_int * MAX_NDIMS), ("iproc", C.c_int * MAX_NDIMS), ("rank", C.c_int),
        ("model", C.c_int), ("interp_char", C.c_int), ("par_scheme", C.c_int), ("rk_type", C.c_int),
        ("dt", C.c_double),
        ("weno_type", C.c_int), ("no_limiting", C.c_int), ("weno_eps", C.c_double),
        ("upwind", C.c_int),
        ("gamma", C.c_double), ("Re", C.c_double), ("Pr", C.c_double), ("Minf", C.c_double),
        ("gravity", C.c_double * MAX_NDIMS), ("rho_ref", C.c_double), ("p_ref", C.c_double),
        ("R", C.c_double), ("N_bv", C.c_double), ("HB", C.c_int),
        ("advection", C.c_double * (MAX_NDIMS * MAX_NVARS)),
        ("diffusion", C.c_double * (MAX_NDIMS * MAX_NVARS)),
        ("nzones", C.c_int), ("zones", BoundaryZone * MAX_ZONES),
        ("x_global", C.POINTER(C.c_double)),
        ("device", C.c_int), ("use_fused", C.c_int), ("conservation_check", C.c_int),
        ("hyp_scheme", C.c_int), ("muscl_limiter", C.c_int), ("muscl_eps", C.c_double),
        ("gravity_type", C.c_int), ("advection_field", C.POINTER(C.c_double)),
        ("weno_rc", C.c_double), ("weno_xi", C.c_double),
        ("par_space_type", C.c_int), ("glm_ee_mode", C.c_int),
        ("lu_maxiter", C.c_int), ("lu_evaluate_norm", C.c_int), ("lu_atol", C.c_double), ("lu_rtol", C.c_double),
        ("lu_gather_and_solve", C.c_int),
    ]


# every symbol include/hypar_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "hpb_config_defaults", "hpb_create", "hpb_destroy", "hpb_last_error", "hpb_error_state", "hpb_clear_error",
    "hpb_device_count", "hpb_version", "hpb_sizeof_config",
    "hpb_partition1d", "hpb_rank1d", "hpb_ranknd", "hpb_get_local_dims", "hpb_npoints_local_wghosts",
    "hpb_ninterfaces", "hpb_get_grid", "hpb_get_neighbors", "hpb_get_zone_extent", "hpb_get_gravity_field",
    "hpb_get_advection_field",
    "hpb_ApplyBoundaryConditions", "hpb_HyperbolicFunction", "hpb_ParabolicFunction", "hpb_SourceFunction",
    "hpb_RHSFunction", "hpb_FFunction", "hpb_UFunction", "hpb_SetInterpLimiterVar", "hpb_GetInterpWeights",
    "hpb_InterpolateInterfacesHyp", "hpb_Upwind", "hpb_FirstDerivativePar", "hpb_SecondDerivativePar",
    "hpb_ComputeCFL", "hpb_TimeIntegrate", "hpb_pipe_upload", "hpb_pipe_download", "hpb_pipe_join", "hpb_pipe_wait", "hpb_TimeIntegrateAsync",
    "hpb_dev_set_solution", "hpb_dev_get_solution", "hpb_dev_fill_solution_from_global", "hpb_TimeStep",
    "hpb_TimeSteps", "hpb_current_time", "hpb_dev_ComputeCFL", "hpb_dev_StepNormSumSq", "hpb_dev_RHS",
    "hpb_comm_get_unique_id", "hpb_comm_nccl_version", "hpb_comm_init_nccl", "hpb_comm_init_local", "hpb_comm_finalize",
    "hpb_comm_kind", "hpb_comm_allreduce", "hpb_comm_stats", "hpb_exchange_plan", "hpb_ExchangeBoundariesnD", "hpb_ExchangeBoundariesLocal",
    "hpb_TimeStepDistributed", "hpb_TimeStepsDistributed", "hpb_RHSFunctionDistributed", "hpb_TimeStepsLocal",
    "hpb_RHSFunctionLocal", "hpb_set_overlap", "hpb_stage_overlap_supported", "hpb_set_stage_fusion", "hpb_stage_fusion_active",
    "hpb_dev_get_stage_rhs", "hpb_nstages", "hpb_needs_viscous_exchange",
    "hpb_dev_VolumeIntegral", "hpb_dev_StageBoundaryIntegral", "hpb_dev_StepBoundaryIntegral", "hpb_BoundaryIntegral",
    "hpb_CalculateConservationError", "hpb_dev_ErrorSums",
    "hpb_dev_get_aux_solution", "hpb_dev_set_aux_solution", "hpb_dev_GLMGEEErrorSums", "hpb_glmgee_gamma",
    "hpb_stream", "hpb_synchronize", "hpb_kernel_launch_count", "hpb_tma_launch_count", "hpb_profile_enable", "hpb_profile_query", "hpb_fp64_issue_peak",
]

_lib = None


def load():
    """Load libhypar_b200.so (raises if it has not been built: run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C hypar_b200/csrc` "
                           "(or __graft_entry__.build()); hypar_b200 has no fallback path")
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    L.hpb_sizeof_config.restype = C.c_size_t
    if L.hpb_sizeof_config() != C.sizeof(Config):
        raise RuntimeError(f"hpb_config layout mismatch: library {L.hpb_sizeof_config()} bytes, binding {C.sizeof(Config)} "
                           "(rebuild hypar_b200/libhypar_b200.so: make -C hypar_b200/csrc)")
    L.hpb_last_error.restype = C.c_char_p
    L.hpb_version.restype = C.c_char_p
    L.hpb_config_defaults.argtypes = [C.POINTER(Config)]
    L.hpb_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.hpb_destroy.argtypes = [vp]
    L.hpb_partition1d.argtypes = [C.c_int, C.c_int, C.c_int]
    L.hpb_rank1d.argtypes = [C.c_int, ip, ip]
    L.hpb_ranknd.argtypes = [C.c_int, C.c_int, ip, ip]
    L.hpb_ranknd.restype = None
    L.hpb_get_local_dims.argtypes = [vp, ip, ip]
    L.hpb_npoints_local_wghosts.argtypes = [vp]
    L.hpb_npoints_local_wghosts.restype = C.c_longlong
    L.hpb_ninterfaces.argtypes = [vp, C.c_int]
    L.hpb_ninterfaces.restype = C.c_longlong
    L.hpb_get_grid.argtypes = [vp, dp, dp]
    L.hpb_get_neighbors.argtypes = [vp, ip]
    L.hpb_get_zone_extent.argtypes = [vp, C.c_int, ip, ip, ip]
    L.hpb_get_gravity_field.argtypes = [vp, dp, dp]
    L.hpb_get_advection_field.argtypes = [vp, dp]
    L.hpb_ApplyBoundaryConditions.argtypes = [vp, dp, C.c_double]
    L.hpb_HyperbolicFunction.argtypes = [vp, dp, dp, C.c_double, C.c_int]
    L.hpb_ParabolicFunction.argtypes = [vp, dp, dp, C.c_double]
    L.hpb_SourceFunction.argtypes = [vp, dp, dp, C.c_double]
    L.hpb_RHSFunction.argtypes = [vp, dp, dp, C.c_double]
    L.hpb_FFunction.argtypes = [vp, dp, dp, C.c_int, C.c_double]
    L.hpb_UFunction.argtypes = [vp, dp, dp, C.c_int, C.c_double]
    L.hpb_SetInterpLimiterVar.argtypes = [vp, dp, dp, C.c_int]
    L.hpb_GetInterpWeights.argtypes = [vp, C.c_int, dp]
    L.hpb_InterpolateInterfacesHyp.argtypes = [vp, dp, dp, dp, C.c_int, C.c_int, C.c_int]
    L.hpb_Upwind.argtypes = [vp, dp, dp, dp, dp, dp, dp, C.c_int, C.c_double]
    L.hpb_FirstDerivativePar.argtypes = [vp, dp, dp, C.c_int, C.c_int]
    L.hpb_SecondDerivativePar.argtypes = [vp, dp, dp, C.c_int]
    L.hpb_ComputeCFL.argtypes = [vp, dp, C.c_double, C.c_double, dp]
    L.hpb_TimeIntegrate.argtypes = [vp, dp, C.c_int, C.c_double]
    L.hpb_pipe_upload.argtypes = [vp, dp, C.c_double]
    L.hpb_pipe_download.argtypes = [vp, dp]
    L.hpb_pipe_wait.argtypes = [vp]
    L.hpb_pipe_join.argtypes = [vp]
    L.hpb_TimeIntegrateAsync.argtypes = [vp, dp, dp, C.c_int, C.c_double]
    L.hpb_dev_set_solution.argtypes = [vp, dp]
    L.hpb_dev_get_solution.argtypes = [vp, dp]
    L.hpb_dev_fill_solution_from_global.argtypes = [vp, dp]
    L.hpb_TimeStep.argtypes = [vp]
    L.hpb_TimeSteps.argtypes = [vp, C.c_int]
    L.hpb_current_time.argtypes = [vp]
    L.hpb_current_time.restype = C.c_double
    L.hpb_dev_ComputeCFL.argtypes = [vp, dp]
    L.hpb_dev_StepNormSumSq.argtypes = [vp, dp]
    L.hpb_dev_RHS.argtypes = [vp, C.c_double, dp]
    for name in ("hpb_nstages", "hpb_needs_viscous_exchange", "hpb_synchronize", "hpb_comm_finalize", "hpb_comm_kind",
                 "hpb_ExchangeBoundariesnD", "hpb_TimeStepDistributed", "hpb_RHSFunctionDistributed",
                 "hpb_stage_overlap_supported"):
        getattr(L, name).argtypes = [vp]
    L.hpb_comm_get_unique_id.argtypes = [C.c_char_p]
    L.hpb_comm_init_nccl.argtypes = [vp, C.c_char_p, C.c_int]
    L.hpb_comm_init_local.argtypes = [C.POINTER(vp), C.c_int]
    L.hpb_comm_allreduce.argtypes = [vp, dp, C.c_int, C.c_int]
    L.hpb_comm_stats.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.hpb_exchange_plan.argtypes = [vp, C.c_int, ip, C.POINTER(C.c_longlong), ip]
    L.hpb_TimeStepsDistributed.argtypes = [vp, C.c_int]
    L.hpb_TimeStepsLocal.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    L.hpb_RHSFunctionLocal.argtypes = [C.POINTER(vp), C.c_int]
    L.hpb_ExchangeBoundariesLocal.argtypes = [C.POINTER(vp), C.c_int]
    L.hpb_set_overlap.argtypes = [vp, C.c_int]
    L.hpb_set_stage_fusion.argtypes = [vp, C.c_int]
    L.hpb_stage_fusion_active.argtypes = [vp]
    L.hpb_dev_get_stage_rhs.argtypes = [vp, C.c_int, dp]
    L.hpb_dev_VolumeIntegral.argtypes = [vp, dp]
    L.hpb_dev_StageBoundaryIntegral.argtypes = [vp, C.c_int, dp]
    L.hpb_dev_StepBoundaryIntegral.argtypes = [vp, dp]
    L.hpb_BoundaryIntegral.argtypes = [vp, dp, dp]
    L.hpb_CalculateConservationError.argtypes = [C.c_int, dp, dp, dp, dp]
    L.hpb_dev_ErrorSums.argtypes = [vp, dp, dp]
    L.hpb_dev_get_aux_solution.argtypes = [vp, dp]
    L.hpb_dev_set_aux_solution.argtypes = [vp, dp]
    L.hpb_dev_GLMGEEErrorSums.argtypes = [vp, dp, dp]
    L.hpb_glmgee_gamma.argtypes = [vp]
    L.hpb_glmgee_gamma.restype = C.c_double
    L.hpb_stream.argtypes = [vp]
    L.hpb_stream.restype = vp
    L.hpb_kernel_launch_count.argtypes = [vp]
    L.hpb_kernel_launch_count.restype = C.c_longlong
    L.hpb_tma_launch_count.argtypes = [vp]
    L.hpb_tma_launch_count.restype = C.c_longlong
    L.hpb_fp64_issue_peak.argtypes = [vp, dp]
    L.hpb_profile_enable.argtypes = [vp, C.c_int]
    L.hpb_profile_query.argtypes = [vp, C.c_int, dp, C.POINTER(C.c_longlong)]
    _lib = L
    return L

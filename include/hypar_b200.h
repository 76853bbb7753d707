/* hypar_b200.h -- C ABI of the B200-native explicit-RHS path for HyPar.
 *
 * Drop-in boundary: HyPar has no plugin loader; its "operator API" is the set of function
 * pointers in `struct HyPar` (reference include/hypar.h:211-359) and
 * `TimeIntegration::RHSFunction / TimeIntegrate` (include/timeintegration_struct.h), assigned in
 * src/Simulation/InitializeSolvers.c:71-393 and <Model>Initialize. Every entry point below
 * replaces one of those pointers (cited per function); `void *s, void *m` (HyPar*, MPIVariables*)
 * of the reference collapse into the opaque `hpb_solver*` created from the same inputs HyPar
 * reads (solver.inp / boundary.inp / physics.inp / weno.inp / the grid). INTEGRATION.md shows the
 * ~100-line C glue (`hyparb200_attach`) a HyPar maintainer adds to install them.
 *
 * Conventions
 *   - plain C, `extern "C"`, pointers + sizes only; return 0 = ok, non-zero = error
 *     (message via hpb_last_error()); like the reference, callers may drop the return value, so
 *     every error is also made sticky: hpb_error_state() stays non-zero until cleared.
 *   - "HOST" entry points take host arrays in HyPar's own layout (ghost-padded AoS, nvars
 *     innermost, dim 0 fastest: include/arrayfunctions.h:40-47); the library moves them to the
 *     GPU, runs the CUDA kernels and copies results back. There is NO CPU fallback: without a
 *     CUDA device every compute call fails with HPB_ERR_NO_DEVICE.
 *   - "DEVICE-RESIDENT" entry points (hpb_dev_*, hpb_Time*) keep the state on the GPU between
 *     calls (device layout: ghost-padded SoA, component-major) -- this is the production path
 *     the time loop uses.
 *   - one hpb_solver per rank/GPU; calls on one solver are serialized by the caller
 *     (same contract as HyPar: one thread per rank).
 */
#ifndef HYPAR_B200_H
#define HYPAR_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPB_MAX_NDIMS 3
#define HPB_MAX_NVARS 5
#define HPB_MAX_ZONES 16

/* error codes */
#define HPB_OK             0
#define HPB_ERR_INVALID    1   /* bad / unsupported configuration (scheme, model, BC type ...) */
#define HPB_ERR_NO_DEVICE  2   /* no CUDA device: the product has no CPU path                  */
#define HPB_ERR_CUDA       3   /* a CUDA runtime call or kernel failed                         */
#define HPB_ERR_ALLOC      4

/* `model` (solver.inp) -- reference src/Simulation/InitializePhysics.c:84-157 */
#define HPB_MODEL_LINEAR_ADR 0  /* "linear-advection-diffusion-reaction" */
#define HPB_MODEL_EULER1D    1  /* "euler1d"        */
#define HPB_MODEL_NS2D       2  /* "navierstokes2d" */
#define HPB_MODEL_NS3D       3  /* "navierstokes3d" */
#define HPB_MODEL_BURGERS    4  /* "burgers": inviscid Burgers equation, nvars = 1, 1-3 dimensions (src/PhysicalModels/Burgers) */

/* weno.inp -- reference src/InterpolationFunctions/WENOFifthOrderCalculateWeights.c:40-61 */
#define HPB_WENO_JS 0
#define HPB_WENO_M  1   /* mapped  */
#define HPB_WENO_Z  2   /* borges  */
#define HPB_WENO_YC 3   /* yc      */

/* physics.inp `upwinding` */
#define HPB_UPWIND_DEFAULT 0    /* LinearADRUpwind */
#define HPB_UPWIND_ROE     1
#define HPB_UPWIND_RUSANOV 2
#define HPB_UPWIND_RF      3    /* "rf-char":  characteristic-based Roe-fixed (Euler1D, NavierStokes3D)          */
#define HPB_UPWIND_LLF     4    /* "llf-char": characteristic-based local Lax-Friedrichs (Euler1D, NavierStokes3D) */

/* solver.inp `hyp_space_scheme` -- reference src/Simulation/InitializeSolvers.c:232-346. The compact schemes
   solve one tridiagonal system per grid line and component (Interp1PrimFifthOrderCRWENO.c:80,
   Interp1PrimFifthOrderCompactUpwind.c:73, Interp1PrimFifthOrderHCWENO.c:65; TridiagLU/tridiagLU.c:84) or, with
   hyp_interp_type characteristic, one block tridiagonal system per grid line (...Char.c; blocktridiagLU.c:103);
   the grid lines may be split among ranks (all four stages of the reference's solve with its Jacobi reduced
   system). Anything but WENO5 runs on the reference-exact kernels. */
#define HPB_SCHEME_WENO5   0    /* "weno5"   */
#define HPB_SCHEME_CRWENO5 1    /* "crweno5" */
#define HPB_SCHEME_CUPW5   2    /* "cupw5": fifth-order compact upwind */
#define HPB_SCHEME_UPW5    3    /* "upw5":  fifth-order upwind         */
#define HPB_SCHEME_FIRST   4    /* "1": first-order upwind  (Interp1PrimFirstOrderUpwind.c)   */
#define HPB_SCHEME_SECOND  5    /* "2": second-order central (Interp1PrimSecondOrderCentral.c) */
#define HPB_SCHEME_FOURTH  6    /* "4": fourth-order central (Interp1PrimFourthOrderCentral.c) */
#define HPB_SCHEME_MUSCL2  7    /* "muscl2": Interp1PrimSecondOrderMUSCL.c, limiter from muscl.inp */
#define HPB_SCHEME_MUSCL3  8    /* "muscl3": Interp1PrimThirdOrderMUSCL.c (Koren), epsilon from muscl.inp */
#define HPB_SCHEME_HCWENO5 9    /* "hcweno5": hybrid compact-WENO5 (Interp1PrimFifthOrderHCWENO.c / ...HCWENOChar.c); rc, xi from weno.inp */
/* muscl.inp `limiter` -- MUSCLInitialize.c:62-75, src/LimiterFunctions/ */
#define HPB_LIMITER_GMM      0
#define HPB_LIMITER_MINMOD   1
#define HPB_LIMITER_VANLEER  2
#define HPB_LIMITER_SUPERBEE 3

/* boundary.inp zone types implemented on the device (reference: 17 types, src/BoundaryConditions/BCInitialize.c; the first
   three are the ones the BASELINE configurations use, the others cover the reference's Navier-Stokes examples;
   types 3, 5-9 exist for 2-D and 3-D models only, as in the reference) */
#define HPB_BC_PERIODIC            0
#define HPB_BC_EXTRAPOLATE         1
#define HPB_BC_SLIP_WALL           2
#define HPB_BC_NOSLIP_WALL         3   /* BCNoslipWall.c: wall_velocity                                  */
#define HPB_BC_DIRICHLET           4   /* BCDirichlet.c: dirichlet[nvars]                                */
#define HPB_BC_SUBSONIC_INFLOW     5   /* BCSubsonicInflow.c: flow_density, wall_velocity = flow velocity */
#define HPB_BC_SUBSONIC_OUTFLOW    6   /* BCSubsonicOutflow.c: flow_pressure                             */
#define HPB_BC_SUBSONIC_AMBIVALENT 7   /* BCSubsonicAmbivalent.c: density, velocity, pressure            */
#define HPB_BC_SUPERSONIC_INFLOW   8   /* BCSupersonicInflow.c: density, velocity, pressure              */
#define HPB_BC_SUPERSONIC_OUTFLOW  9   /* BCSupersonicOutflow.c                                          */
#define HPB_BC_SPONGE             10   /* BCSponge.c: an interior box [xmin, xmax] in which the solution is relaxed towards
                                          dirichlet[] (SpongeValue) by a source term, SourceFunction.c:52-75; no ghost fill */

/* time_scheme_type for time_scheme "rk" -- reference TimeExplicitRKInitialize.c:58-79 */
#define HPB_RK_44     0
#define HPB_RK_SSPRK3 1   /* "ssprk3" = "tvdrk3" */
#define HPB_RK_1FE    2   /* "1fe"; also time_scheme "euler" (TimeForwardEuler.c: the same update) */
#define HPB_RK_22     3
#define HPB_RK_33     4
/* time_scheme "glm-gee" (general linear methods with global error estimation, TimeGLMGEE.c:45-155): rk_type holds the
   method -- reference TimeGLMGEEInitialize.c:41-70, include/timeintegration_struct.h:209-257 */
#define HPB_GLMGEE_23      16
#define HPB_GLMGEE_24      17
#define HPB_GLMGEE_25I     18
#define HPB_GLMGEE_35      19
#define HPB_GLMGEE_EXRK2A  20
#define HPB_GLMGEE_RK32G1  21
#define HPB_GLMGEE_RK285EX 22
/* glm_gee.inp `ee_mode` (TimeGLMGEEInitialize.c:431-470; default yeps) */
#define HPB_GLM_YEPS 0
#define HPB_GLM_YYT  1

/* par_space_type -- reference InitializeSolvers.c:107-176 (the form of the parabolic term) */
#define HPB_PAR_NC_1STAGE   0   /* "nonconservative-1stage" (default, ReadInputs.c:134): ParabolicFunctionNC1Stage  */
#define HPB_PAR_NC_1_5STAGE 1   /* "nonconservative-1.5stage"                                                        */
#define HPB_PAR_NC_2STAGE   2   /* "nonconservative-2stage": what the Navier-Stokes models need                      */
#define HPB_PAR_CONS_1STAGE 3   /* "conservative-1stage"                                                             */

/* fields that have a halo exchange (reference MPIExchangeBoundariesnD call sites) */
#define HPB_FIELD_U       0   /* TimeRHSFunctionExplicit.c:60, TimePreStep.c:71         */
#define HPB_FIELD_QDERIVX 1   /* NavierStokes3DParabolicFunction.c:125                  */
#define HPB_FIELD_QDERIVY 2   /* NavierStokes3DParabolicFunction.c:127 (and :129, sic)  */

typedef struct hpb_boundary_zone {
  int    type;                         /* HPB_BC_*                                       */
  int    dim, face;                    /* face = +1 (low side) / -1 (high side)          */
  double xmin[HPB_MAX_NDIMS], xmax[HPB_MAX_NDIMS];
  double wall_velocity[HPB_MAX_NDIMS]; /* DomainBoundary::FlowVelocity: wall velocity (slip / no-slip wall) or the
                                          flow velocity of an inflow zone                  */
  double flow_density, flow_pressure;  /* DomainBoundary::FlowDensity, FlowPressure      */
  double dirichlet[HPB_MAX_NVARS];     /* DomainBoundary::DirichletValue                 */
} hpb_boundary_zone;

typedef struct hpb_config {
  /* --- solver.inp --- */
  int    ndims, nvars, ghosts;
  int    dim_global[HPB_MAX_NDIMS];
  int    iproc[HPB_MAX_NDIMS];         /* ranks per dimension                            */
  int    rank;                         /* this rank: ip0 + iproc0*(ip1 + iproc1*ip2)     */
  int    model;                        /* HPB_MODEL_*                                    */
  int    interp_char;                  /* hyp_interp_type: 1 characteristic, 0 components*/
  int    par_scheme;                   /* par_space_scheme: 2 or 4                       */
  int    rk_type;                      /* HPB_RK_* (time_scheme rk / euler) or HPB_GLMGEE_* (time_scheme glm-gee) */
  double dt;
  /* --- weno.inp --- */
  int    weno_type;                    /* HPB_WENO_*                                     */
  int    no_limiting;
  double weno_eps;
  /* --- physics.inp --- */
  int    upwind;                       /* HPB_UPWIND_*                                   */
  double gamma, Re, Pr, Minf;          /* Re as given in physics.inp (divided by Minf inside) */
  double gravity[HPB_MAX_NDIMS], rho_ref, p_ref, R, N_bv;
  int    HB;
  double advection[HPB_MAX_NDIMS*HPB_MAX_NVARS];  /* LinearADR a[nvars*dir+v]              */
  double diffusion[HPB_MAX_NDIMS*HPB_MAX_NVARS];  /* LinearADR nu[nvars*dir+v]             */
  /* --- boundary.inp --- */
  int    nzones;
  hpb_boundary_zone zones[HPB_MAX_ZONES];
  /* --- grid: GLOBAL coordinates, concatenated per dimension (as in initial.inp) --- */
  const double* x_global;
  /* --- device --- */
  int    device;                       /* CUDA device ordinal; -1 = current              */
  int    use_fused;                    /* 1 (default): fused sweep kernels where available (TMA-fed variant
                                             when the padded row length is even), 2: fused sweeps without
                                             the TMA variant, 0: generic per-interface kernels only */
  /* --- solver.inp (cont.) --- */
  int    conservation_check;           /* ConservationCheck "yes": keep the boundary-flux bookkeeping of
                                          HyperbolicFunction.c:103-106 / TimeRK.c:172-193 on the device       */
  int    hyp_scheme;                   /* hyp_space_scheme: HPB_SCHEME_* (default WENO5)                      */
  /* --- muscl.inp --- */
  int    muscl_limiter;                /* HPB_LIMITER_* (default gmm)                                         */
  double muscl_eps;                    /* default 1e-3                                                        */
  /* --- physics.inp (cont.) --- */
  int    gravity_type;                 /* Euler1D `gravity_type` (Euler1DGravityField.c:44-52): 0 exp(-g x), 1 sinusoidal
                                          potential; the 1-D gravity is gravity[0]                             */
  const double* advection_field;       /* LinearADR `advection_filename` (LinearADRAdvectionField.c): the spatially varying
                                          advection field on the GLOBAL grid, [point][ndims*nvars], points ordered like the
                                          solution in initial.inp (no ghosts); NULL = constant advection[]. Copied by
                                          hpb_create.                                                          */
  double weno_rc, weno_xi;             /* weno.inp `rc`, `xi` (WENOInitialize.c:53-54: 0.3, 0.001): hybridisation parameters of hcweno5 */
  int    par_space_type;               /* HPB_PAR_*: LinearADR with a non-zero diffusion coefficient is on the device only
                                          as nonconservative-1stage -- the other forms are different arithmetic
                                          (ParabolicFunctionNC2Stage.c / NC1_5Stage / Cons1Stage) and fail in hpb_create */
  int    glm_ee_mode;                  /* HPB_GLM_*: glm_gee.inp `ee_mode` of the GLM-GEE methods                */
  /* --- lusolver.inp (tridiagLUInit.c:54-90): the reduced system of the compact schemes when a grid line is split among ranks
     (tridiagIterJacobi.c / blocktridiagIterJacobi.c). reducedsolvetype "gather-and-solve" is not built: hpb_create fails. */
  int    lu_maxiter;                   /* default 10                                                              */
  int    lu_evaluate_norm;             /* default 1: stop on atol / rtol as well; 0: exactly maxiter iterations    */
  double lu_atol, lu_rtol;             /* defaults 1e-12, 1e-10                                                    */
  int    lu_gather_and_solve;          /* 1 = reducedsolvetype gather-and-solve (refused)                          */
} hpb_config;

typedef struct hpb_solver hpb_solver;

/* ------------------------------------------------------------------ life cycle / errors */
void        hpb_config_defaults(hpb_config* cfg);   /* the defaults of ReadInputs.c:112-146 etc. */
int         hpb_create(const hpb_config* cfg, hpb_solver** out);
int         hpb_destroy(hpb_solver* h);
const char* hpb_last_error(void);
int         hpb_error_state(void);
void        hpb_clear_error(void);
int         hpb_device_count(void);                 /* 0 when no CUDA device is visible  */
const char* hpb_version(void);
size_t      hpb_sizeof_config(void);                /* sizeof(hpb_config) of the library: bindings check their own layout */

/* ------------------------------------------------------------------ host set-up queries
 * (pure host code, usable without a GPU: partitioning as MPIPartition1D.c / MPIRanknD.c, ghost
 * coordinates as ReadArray.c:60-100, dxinv as InitialSolution.c:74-119, zone extents as
 * InitializeBoundaries.c:380-440, gravity field as NavierStokes3DGravityField.c:33-152) */
int  hpb_partition1d(int nglobal, int nproc, int rank);
int  hpb_rank1d(int ndims, const int* iproc, const int* ip);
void hpb_ranknd(int ndims, int rank, const int* iproc, int* ip);
int  hpb_get_local_dims(const hpb_solver* h, int* dim_local, int* is_global);
long long hpb_npoints_local_wghosts(const hpb_solver* h);
long long hpb_ninterfaces(const hpb_solver* h, int dir);
int  hpb_get_grid(const hpb_solver* h, double* x_wghosts, double* dxinv_wghosts); /* size sum(dim+2g) */
int  hpb_get_neighbors(const hpb_solver* h, int* neighbor_rank /* [2*ndims], -1 = none */);
int  hpb_get_zone_extent(const hpb_solver* h, int zone, int* is, int* ie, int* on_this_proc);
int  hpb_get_gravity_field(const hpb_solver* h, double* grav_f, double* grav_g);
/* LinearADR varying advection field of this rank (LinearADR::a after LinearADRAdvectionField.c), HyPar layout
   [point with ghosts][ndims*nvars]; HPB_ERR_INVALID when the advection is constant */
int  hpb_get_advection_field(const hpb_solver* h, double* a);

/* ------------------------------------------------------------------ HOST entry points
 * (the reference's function-pointer surface; arrays = host, HyPar layout, local + ghosts) */

/* solver->ApplyBoundaryConditions (hypar.h:214; ApplyBoundaryConditions.c) : fills face ghosts of u */
int hpb_ApplyBoundaryConditions(hpb_solver* h, double* u, double t);
/* solver->HyperbolicFunction (hypar.h:250-253; HyperbolicFunction.c:31) with the model's
   FFunction/Upwind; LimFlag as in the reference */
int hpb_HyperbolicFunction(hpb_solver* h, double* hyp, const double* u, double t, int LimFlag);
/* solver->ParabolicFunction (hypar.h:256; NavierStokes3DParabolicFunction.c:50,
   NavierStokes2DParabolicFunction.c:38, ParabolicFunctionNC1Stage.c) -- single rank */
int hpb_ParabolicFunction(hpb_solver* h, double* par, const double* u, double t);
/* solver->SourceFunction (hypar.h:259; SourceFunction.c + NavierStokes3DSource.c:38): uses the
   flux weights of the LAST hpb_HyperbolicFunction call, like the reference */
int hpb_SourceFunction(hpb_solver* h, double* source, const double* u, double t);
/* TimeIntegration::RHSFunction = TimeRHSFunctionExplicit (TimeRHSFunctionExplicit.c:30).
   u is modified (boundary conditions), as in the reference. Single rank. */
int hpb_RHSFunction(hpb_solver* h, double* rhs, double* u, double t);
/* fine-grained pointers: FFunction (hypar.h:276), UFunction (:321), SetInterpLimiterVar (:234),
   InterpolateInterfacesHyp (:224), Upwind (:295), FirstDerivativePar (:243),
   SecondDerivativePar (:247), ComputeCFL (:269) */
int hpb_FFunction(hpb_solver* h, double* f, const double* u, int dir, double t);
int hpb_UFunction(hpb_solver* h, double* uC, const double* u, int dir, double t);
int hpb_SetInterpLimiterVar(hpb_solver* h, const double* fC, const double* u, int dir);
int hpb_GetInterpWeights(hpb_solver* h, int dir, double* w /* [12*ninterfaces*nvars]: LF,LU,RF,RU x w1..w3 */);
int hpb_InterpolateInterfacesHyp(hpb_solver* h, double* fI, const double* fC, const double* u,
                                 int upw, int dir, int uflag);
int hpb_Upwind(hpb_solver* h, double* fI, const double* fL, const double* fR, const double* uL,
               const double* uR, const double* u, int dir, double t);
int hpb_FirstDerivativePar(hpb_solver* h, double* Df, const double* f, int dir, int bias);
int hpb_SecondDerivativePar(hpb_solver* h, double* D2f, const double* f, int dir);
int hpb_ComputeCFL(hpb_solver* h, const double* u, double dt, double t, double* cfl);
/* TimeIntegration::TimeIntegrate = TimeRK (TimeRK.c:35) preceded by TimePreStep's BC/halo
   (TimePreStep.c:50-76): advances host u by nsteps steps of size cfg.dt. Single rank. */
int hpb_TimeIntegrate(hpb_solver* h, double* u, int nsteps, double t0);
/* The same for a SEQUENCE of independent fields on one solver -- HyPar's ensemble runs (nsims > 1: Solve.cpp:37-180
   advances every SimulationObject each step; include/ensemble_simulations.h) or any caller that keeps its fields in
   host memory: the calls only ENQUEUE work, so the host->device copy of a field overlaps the steps of the previous
   one and the device->host copy of the one before (two copy streams + the solver's stream, ordered by events).
     hpb_pipe_upload    H2D of u_in (HyPar layout; pinned memory for a true overlap) + transposition into the device
                        solution, after the previous field's steps
     ... steps: hpb_TimeStep / hpb_TimeSteps (single rank) or hpb_TimeStepsDistributed (decomposed) ...
     hpb_pipe_download  transposition + D2H of the device solution into u_out
     hpb_pipe_wait      blocks until everything enqueued has finished (u_in may be reused after the upload's copy:
                        conservatively, after hpb_pipe_wait)
     hpb_TimeIntegrateAsync = upload + nsteps steps + download, single rank. */
int hpb_pipe_upload(hpb_solver* h, const double* u_in, double t0);
int hpb_pipe_download(hpb_solver* h, double* u_out);
int hpb_pipe_join(hpb_solver* h);   /* orders the solver's stream (hpb_stream) after every copy enqueued so far, without blocking */
int hpb_pipe_wait(hpb_solver* h);
int hpb_TimeIntegrateAsync(hpb_solver* h, const double* u_in, double* u_out, int nsteps, double t0);

/* ------------------------------------------------------------------ DEVICE-RESIDENT path */
int hpb_dev_set_solution(hpb_solver* h, const double* u_host);     /* H2D + AoS->SoA */
int hpb_dev_get_solution(hpb_solver* h, double* u_host);           /* SoA->AoS + D2H */
int hpb_dev_fill_solution_from_global(hpb_solver* h, const double* u_global); /* global AoS, no ghosts */
/* one full time step on the device (TimePreStep BC/halo + TimeRK + time update). Single rank. */
int hpb_TimeStep(hpb_solver* h);
int hpb_TimeSteps(hpb_solver* h, int nsteps);
double hpb_current_time(const hpb_solver* h);
/* reductions on the device solution: CFL (ComputeCFL) and the step norm of TimePostStep.c:44-63
   (returns the local max CFL and the LOCAL sum of squares of u^{n+1} - u^n of the last step, formed from the stage
   right-hand sides the step combined -- valid until the next step starts; callers all-reduce) */
int hpb_dev_ComputeCFL(hpb_solver* h, double* cfl_local_max);
int hpb_dev_StepNormSumSq(hpb_solver* h, double* sumsq_local);
/* ---- conservation and error diagnostics, evaluated on the device solution; only scalars
 * cross the bus. Reductions are deterministic (fixed launch shape and summation tree); they differ from the
 * reference's serial sums by rounding only. Each call returns THIS RANK'S part; the caller sums / maxes over
 * ranks exactly where the reference calls MPISum_double / MPIMax_double.
 *   solver->VolumeIntegralFunction   (hypar.h; VolumeIntegral.c:14-48)          hpb_dev_VolumeIntegral
 *   solver->StageBoundaryIntegral    (HyperbolicFunction.c:65,103-106)          hpb_dev_StageBoundaryIntegral
 *   solver->StepBoundaryIntegral     (TimePreStep.c:117, TimeRK.c:172-193)      hpb_dev_StepBoundaryIntegral
 *   solver->BoundaryIntegralFunction (BoundaryIntegral.c:20-60, local part)     hpb_BoundaryIntegral
 *   solver->CalculateConservationError (CalculateConservationError.c:14-39)     hpb_CalculateConservationError
 *   CalculateError (CalculateError.c:26-124, the six local sums it all-reduces) hpb_dev_ErrorSums
 * The boundary-flux bookkeeping needs cfg.conservation_check = 1 (as ConservationCheck "yes" in solver.inp). */
int hpb_dev_VolumeIntegral(hpb_solver* h, double* vol_local /* [nvars] */);
/* slot: 0..nstages-1 = BoundaryFlux[stage] of the last step; -1 = StageBoundaryIntegral left by the last
   hpb_HyperbolicFunction / hpb_RHSFunction / hpb_dev_RHS call. sbi[(2*d+face)*nvars + v], face 0 = low. */
int hpb_dev_StageBoundaryIntegral(hpb_solver* h, int slot, double* sbi /* [2*ndims*nvars] */);
int hpb_dev_StepBoundaryIntegral(hpb_solver* h, double* step_bi /* [2*ndims*nvars] */);
int hpb_BoundaryIntegral(const hpb_solver* h, const double* step_bi, double* bi_local /* [nvars] */);
int hpb_CalculateConservationError(int nvars, const double* vol, const double* vol_initial,
                                   const double* total_boundary_integral, double* err /* [nvars] */);
/* sums[0..2] = (sum |uex|, sum uex^2, max |uex|), sums[3..5] = the same of (u_dev - uex), over this rank's
   interior points and all components; uex_host: exact solution, HyPar layout (local + ghosts) */
int hpb_dev_ErrorSums(hpb_solver* h, const double* uex_host, double* sums /* [6] */);
/* ---- GLM-GEE: time_scheme glm-gee advances the solution and ONE auxiliary solution (r = 2) that
 * carries the global-error estimate. hpb_TimeStep(s) / hpb_TimeStepsDistributed / hpb_TimeStepsLocal run
 *   TimeGLMGEE           src/TimeIntegration/TimeGLMGEE.c:45-155 (stage values :66-113, step completion :116-151)
 * when cfg.rk_type is one of HPB_GLMGEE_*. The auxiliary solution starts as TimeInitialize.c:156-169 sets it (a copy of
 * the solution for ee_mode yyt, zero for yeps) at the first step after hpb_dev_set_solution.
 *   TimeGetAuxSolutions  src/TimeIntegration/TimeGetAuxSolutions.c:30-63 (op_aux files)    hpb_dev_get_aux_solution
 *   TimeError            src/TimeIntegration/TimeError.c:43-127 (glm_err.dat)              hpb_dev_GLMGEEErrorSums
 * sums[0..2] = (sum |u|, sum u^2, max |u|), sums[3..5] = the same of the estimated error (the auxiliary solution, or
 * (u - aux)/(1 - gamma) for yyt), sums[6..8] = the same of (u - uex) - estimate (uex_host may be NULL: zeros), over this
 * rank's interior points; the caller reduces over ranks and normalises as TimeError.c:53-118 does. */
int hpb_dev_get_aux_solution(hpb_solver* h, double* uaux_host);
int hpb_dev_set_aux_solution(hpb_solver* h, const double* uaux_host);
int hpb_dev_GLMGEEErrorSums(hpb_solver* h, const double* uex_host, double* sums /* [9] */);
double hpb_glmgee_gamma(const hpb_solver* h);       /* GLMGEEParameters::gamma of the method (0 for explicit RK) */
/* evaluate rhs(u_dev) once into an internal buffer and copy it to the host (HyPar layout) */
int hpb_dev_RHS(hpb_solver* h, double t, double* rhs_host /* may be NULL */);

/* ---- multi-GPU: one solver per rank / GPU, HyPar's Cartesian blocks. The halo exchange lives inside the
 * library (csrc/comm.cu): it replaces
 *   MPIExchangeBoundariesnD            src/MPIFunctions/MPIExchangeBoundariesnD.c:42-173
 *   gpuMPIExchangeBoundariesnD         src/MPIFunctions/MPIExchangeBoundariesnD_GPU.cu:435-558
 * at their call sites on the explicit path: TimePreStep.c:57-76 (u), TimeRHSFunctionExplicit.c:60 (stage solution),
 * NavierStokes3DParabolicFunction.c:125-130 (QDerivX, QDerivY; QDerivZ is not exchanged there, nor here).
 * Transport: NCCL point-to-point over NVLink (ncclSend / ncclRecv, grouped, on a communication stream of the library;
 * libnccl.so.2 is dlopen'ed, there is no link-time dependency), or an in-process transport when one process owns every
 * rank (device-to-device copies; the single-GPU tests, or one process driving several GPUs).
 *
 *   rank 0:  hpb_comm_get_unique_id(id)  ->  the caller broadcasts the 128 bytes (HyPar: MPI_Bcast on mpi.world)
 *   all:     hpb_comm_init_nccl(h, id, nranks)          collective (ncclCommInitRank; rank = cfg.rank)
 *            hpb_dev_set_solution(h, u) ... hpb_TimeStepsDistributed(h, n) ... hpb_dev_get_solution(h, u)
 *            hpb_comm_allreduce(h, v, n, op)            MPISum_double / MPIMax_double for CFL, norms, integrals
 *            hpb_comm_finalize(h)                       (hpb_destroy does it too)
 * hpb_TimeStepsDistributed only ENQUEUES (no host synchronisation between stages or steps: compute and communication
 * streams are ordered by CUDA events); hpb_synchronize / any download waits for it. */
#define HPB_COMM_ID_BYTES 128
#define HPB_SLOT_U   0   /* exchange of the solution, all dimensions */
#define HPB_SLOT_Q0  1   /* exchange of the Q-derivatives across the faces of dimension 0 */
#define HPB_SLOT_Q12 2   /* ... of the other dimensions */
int hpb_comm_get_unique_id(void* id /* [HPB_COMM_ID_BYTES] */);
int hpb_comm_nccl_version(void);                          /* e.g. 22809; 0 = libnccl could not be loaded */
int hpb_comm_init_nccl(hpb_solver* h, const void* id, int nranks);
int hpb_comm_init_local(hpb_solver** ranks, int nranks);  /* every rank of the decomposition, indexed by rank, one process */
int hpb_comm_finalize(hpb_solver* h);
int hpb_comm_kind(const hpb_solver* h);                   /* 0 none, 1 NCCL, 2 in-process */
int hpb_comm_allreduce(hpb_solver* h, double* v, int n /* <= 8 */, int op /* 0 sum, 1 max */);
int hpb_comm_stats(const hpb_solver* h, long long* messages_sent, long long* bytes_sent);
/* the ordered point-to-point operations of one exchange on this rank (host logic, needs no device): ops[3*i] = 0 send /
   1 recv, ops[3*i+1] = face 2*d + side (side 0 = low), ops[3*i+2] = peer rank, counts[i] = doubles; at most 24 entries.
   Per dimension: send(low), send(high), recv(high), recv(low) -- between one pair of ranks messages match in issue order
   (the role of the reference's tags 1630 / 1631 when both neighbours are the same rank). */
int hpb_exchange_plan(const hpb_solver* h, int slot, int* ops, long long* counts, int* nops);
/* MPIExchangeBoundariesnD on the device solution: ghost faces of u <- the neighbours' interior layers. Blocking. */
int hpb_ExchangeBoundariesnD(hpb_solver* h);
int hpb_ExchangeBoundariesLocal(hpb_solver** ranks, int nranks);      /* the same for in-process ranks */
/* TimePreStep (BCs + halo of u) + TimeRK + step completion, nsteps times; this rank's part (NCCL transport) */
int hpb_TimeStepDistributed(hpb_solver* h);
int hpb_TimeStepsDistributed(hpb_solver* h, int nsteps);
/* one TimeRHSFunctionExplicit of the device solution into the stage-0 right-hand side (hpb_dev_get_stage_rhs(h, 0, .)) */
int hpb_RHSFunctionDistributed(hpb_solver* h);
/* the same for in-process ranks: all of them advance in lock step */
int hpb_TimeStepsLocal(hpb_solver** ranks, int nranks, int nsteps);
int hpb_RHSFunctionLocal(hpb_solver** ranks, int nranks);
/* Schedule of the distributed step. 1 (default): the exchange of u travels under the full-array RK update (the face
 * layers of the stage vector are evaluated first, straight into the send buffers; the step completion does the same,
 * which makes the TimePreStep exchange of the next step free), the Q-derivative exchange of dimensions 1.. travels
 * under the x-sweep (sweep d reads the halos of dimension d only: faces, never edges or corners). 0: pack - exchange -
 * unpack in sequence at the reference's call sites. Results are bit-identical. */
int hpb_set_overlap(hpb_solver* h, int on);
int hpb_stage_overlap_supported(const hpb_solver* h);     /* 1: this configuration is driven sweep by sweep when overlapped */
/* Stage fusion, 1 (default): where row s+1 of the explicit RK tableau has the single entry a_{s+1,s} (every row of RK4, the
 * second of SSPRK3: TimeExplicitRKInitialize.c:58-79) the last directional sweep of stage s also writes the stage solution
 * U_{s+1} = u + a dt k_s of TimeRK.c:131-141 -- same two roundings, bit-identical results, one pass over memory less per
 * stage. 0: every stage solution by its own kernel. Applies to the TMA-fed sweeps (NavierStokes2D / 3D, component-wise,
 * no sponge); hpb_stage_fusion_active says whether this solver uses it. */
int hpb_set_stage_fusion(hpb_solver* h, int on);
int hpb_stage_fusion_active(const hpb_solver* h);
int hpb_dev_get_stage_rhs(hpb_solver* h, int stage, double* rhs_host);   /* Udot[stage] -> host (HyPar layout) */
int hpb_nstages(const hpb_solver* h);
int hpb_needs_viscous_exchange(const hpb_solver* h);
void* hpb_stream(hpb_solver* h);                      /* cudaStream_t the kernels run on */
int hpb_synchronize(hpb_solver* h);

/* ------------------------------------------------------------------ instrumentation */
long long hpb_kernel_launch_count(const hpb_solver* h);   /* kernels launched by this solver so far */
long long hpb_tma_launch_count(const hpb_solver* h);      /* of those, launches of the TMA-fed fused sweep */
/* Device timing per kernel category (the analogue of the reference's GPU_STAT cudaEvent timers,
   HyperbolicFunction.c:69-121): when enabled, every launch group of a category is bracketed by a
   CUDA event pair on the solver's stream. hpb_profile_query synchronises and returns the summed
   milliseconds and the number of launch groups recorded since hpb_profile_enable(h, 1). */
#define HPB_PROF_SWEEP_X  0   /* hyperbolic sweep kernels, direction 0 / 1 / 2 */
#define HPB_PROF_SWEEP_Y  1
#define HPB_PROF_SWEEP_Z  2
#define HPB_PROF_VISCOUS  3   /* parabolic kernels */
#define HPB_PROF_RK       4   /* stage-vector and step-completion kernels */
#define HPB_PROF_BC       5   /* boundary-condition kernels */
#define HPB_PROF_HALO     6   /* pack / unpack kernels */
#define HPB_PROF_OTHER    7
#define HPB_PROF_SWEEP_FUSED 8   /* the last direction's sweep when it also writes the next RK stage solution (stage fusion) */
#define HPB_PROF_NCAT     9
/* FP64 issue peak of the solver's device in thread-instructions per second, measured live (the denominator of the FP64
   roofline of the sweeps: bench.py roofline.fp64) */
int hpb_fp64_issue_peak(hpb_solver* h, double* thread_instr_per_s);
int hpb_profile_enable(hpb_solver* h, int on);
int hpb_profile_query(hpb_solver* h, int category, double* total_ms, long long* count);

#ifdef __cplusplus
}
#endif
#endif /* HYPAR_B200_H */

/* hyparb200_attach.c -- the reference-side binding of libhypar_b200.so: plain C99 compiled INTO HyPar
 * (against HyPar's own headers) and linked with -lhypar_b200. It only uses the C ABI of include/hypar_b200.h.
 *
 * HyPar has no plugin loader: its operator API is the set of function pointers in `struct HyPar`
 * (include/hypar.h:211-359) that InitializeSolvers (src/Simulation/InitializeSolvers.c:71-393) and
 * <Model>Initialize assign; the reference's own accelerator path (`use_gpu yes`) swaps the same pointers for
 * gpu... twins. hyparb200_attach() does the same swap, ONE call after InitializePhysicsData
 * (src/main.cpp:403-420) and before Solve -- no other source change (integration/hypar_b200_main.cpp is
 * src/main.cpp's single-simulation sequence with that one call added).
 *
 * Two modes (environment variable HYPARB200_MODE):
 *   resident (default)  the solution lives on the GPU; HyPar::TimeIntegrate = one hpb_TimeStep; the host mirror
 *                       solver->u is refreshed only when HyPar is about to read it (screen norm, file output,
 *                       CalculateError, the end of the run); CFL, VolumeIntegral and the boundary-flux
 *                       bookkeeping are device reductions.
 *   host                every pointer goes through the host-array entry points (H2D, kernels, D2H per call);
 *                       TimeIntegrate = hpb_TimeIntegrate(u, 1 step). Slower; exercises the fine-grained ABI.
 * Both give the files HyPar itself writes (op.bin, conservation.dat, errors.dat) with HyPar's own writers.
 *
 * Several ranks (one GPU per rank, resident mode): rank 0 asks the library for an ncclUniqueId, HyPar's own
 * MPIBroadcast_character carries it over mpi->world, hpb_comm_init_nccl brings the library's NCCL transport up and
 * HyPar::TimeIntegrate becomes hpb_TimeStepsDistributed -- the halo exchange (MPIExchangeBoundariesnD's job) happens
 * inside the library, device to device over NVLink. Ensembles (nsims > 1, TimeRK.c:50-93): one hpb_solver per
 * SimulationObject.
 *
 * Resident mode keeps HyPar's host code off the hot path: solver->PreStep (NavierStokes3DPreStep.c:55-99 rebuilds a
 * 3 x 25-double Jacobian per point that the explicit path never reads) is cleared, and the direct call of
 * MPIExchangeBoundariesnD on the stale host mirror in TimePreStep.c:57-76 is answered by
 * __wrap_MPIExchangeBoundariesnD below (link with -Wl,--wrap=MPIExchangeBoundariesnD; a maintainer would rather put
 * `if (solver->use_b200) return 0;` at the top of that call site).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <basic.h>
#include <arrayfunctions.h>
#include <mpivars.h>
#include <hypar.h>
#include <simulation_object.h>
#include <timeintegration.h>
#include <boundaryconditions.h>
#include <interpolation.h>
#include <limiters.h>
#include <tridiagLU.h>
#include <physicalmodels/linearadr.h>
#include <physicalmodels/euler1d.h>
#include <physicalmodels/navierstokes2d.h>
#include <physicalmodels/navierstokes3d.h>
#include "hypar_b200.h"

#define B200_MAXSIM 256
typedef struct { hpb_solver *h; HyPar *solver; MPIVariables *mpi; } B200Sim;
static struct {
  int         nsims;
  B200Sim     sim[B200_MAXSIM];
  int         resident;
  long long   steps;
} g;

/* the library solver behind a HyPar solver object (the `void *s` every function pointer receives) */
static hpb_solver* H(void *s)
{
  for (int n = 0; n < g.nsims; n++) if ((void*) g.sim[n].solver == s) return g.sim[n].h;
  fprintf(stderr, "hypar_b200: called with a solver object that was not attached\n");
  exit(1);
}
static int is_resident_solution(const double *u)
{
  if (!g.resident) return 0;
  for (int n = 0; n < g.nsims; n++) if (u == g.sim[n].solver->u) return 1;
  return 0;
}

static void die_on_error(const char *where)
{
  if (hpb_error_state()) {
    fprintf(stderr, "hypar_b200 (%s): %s\n", where, hpb_last_error());
    exit(1);                       /* HyPar drops return values (basic.h:15-23): fail loudly here */
  }
}

/* ---- HyPar::ApplyBoundaryConditions (hypar.h:214). Resident mode: the device step applies the boundary
   conditions to the device solution itself (TimePreStep.c:50 acts on the stale host mirror: nothing to do). */
static int B200_BC(void *s, void *m, double *u, double *xref, double t)
{
  (void)m; (void)xref;
  if (g.resident && u == ((HyPar*)s)->u) return 0;
  hpb_ApplyBoundaryConditions(H(s), u, t); die_on_error("ApplyBoundaryConditions");
  return 0;
}

/* ---- TimePreStep.c:57-76 calls MPIExchangeBoundariesnD directly (no function pointer). In resident mode the array it
   names is the stale host mirror: the device step exchanges the device solution itself (hpb_TimeStepsDistributed).
   -Wl,--wrap=MPIExchangeBoundariesnD routes the call here; every other array goes to the reference's function. */
int __real_MPIExchangeBoundariesnD(int, int, int*, int, void*, double*);
int __wrap_MPIExchangeBoundariesnD(int ndims, int nvars, int *dim, int ghosts, void *m, double *var)
{
  if (is_resident_solution(var)) return 0;
  return __real_MPIExchangeBoundariesnD(ndims, nvars, dim, ghosts, m, var);
}

/* ---- HyPar::HyperbolicFunction (hypar.h:250-253); the FFunction/Upwind arguments are the model's own, fixed
   when the hpb_solver was created */
static int B200_Hyp(double *hyp, double *u, void *s, void *m, double t, int LimFlag,
                    int (*F)(double*, double*, int, void*, double),
                    int (*U)(double*, double*, double*, double*, double*, double*, int, void*, double))
{
  HyPar *solver = (HyPar*) s;
  (void)m; (void)F; (void)U;
  hpb_HyperbolicFunction(H(s), hyp, u, t, LimFlag); die_on_error("HyperbolicFunction");
  if (!strcmp(solver->ConservationCheck, "yes"))
    hpb_dev_StageBoundaryIntegral(H(s), -1, solver->StageBoundaryIntegral);
  return 0;
}
static int B200_Par(double *par, double *u, void *s, void *m, double t)
{ (void)s; (void)m; hpb_ParabolicFunction(H(s), par, u, t); die_on_error("ParabolicFunction"); return 0; }
static int B200_Src(double *src, double *u, void *s, void *m, double t)
{ (void)s; (void)m; hpb_SourceFunction(H(s), src, u, t); die_on_error("SourceFunction"); return 0; }
static int B200_FFunction(double *f, double *u, int dir, void *s, double t)
{ (void)s; hpb_FFunction(H(s), f, u, dir, t); die_on_error("FFunction"); return 0; }
static int B200_UFunction(double *uC, double *u, int dir, void *s, void *m, double t)
{ (void)s; (void)m; hpb_UFunction(H(s), uC, u, dir, t); die_on_error("UFunction"); return 0; }
static int B200_Upwind(double *fI, double *fL, double *fR, double *uL, double *uR, double *u, int dir, void *s, double t)
{ (void)s; hpb_Upwind(H(s), fI, fL, fR, uL, uR, u, dir, t); die_on_error("Upwind"); return 0; }
static int B200_SetInterpLimiterVar(double *fC, double *u, double *x, int dir, void *s, void *m)
{ (void)x; (void)s; (void)m; hpb_SetInterpLimiterVar(H(s), fC, u, dir); die_on_error("SetInterpLimiterVar"); return 0; }
static int B200_InterpolateInterfacesHyp(double *fI, double *fC, double *u, double *x, int upw, int dir, void *s, void *m, int uflag)
{ (void)x; (void)s; (void)m; hpb_InterpolateInterfacesHyp(H(s), fI, fC, u, upw, dir, uflag); die_on_error("InterpolateInterfacesHyp"); return 0; }
static int B200_FirstDerivativePar(double *Df, double *f, int dir, int bias, void *s, void *m)
{ (void)s; (void)m; hpb_FirstDerivativePar(H(s), Df, f, dir, bias); die_on_error("FirstDerivativePar"); return 0; }
static int B200_SecondDerivativePar(double *D2f, double *f, int dir, void *s, void *m)
{ (void)s; (void)m; hpb_SecondDerivativePar(H(s), D2f, f, dir); die_on_error("SecondDerivativePar"); return 0; }

/* ---- HyPar::ComputeCFL (hypar.h:269), called by TimePreStep.c:93 every screen_op_iter steps */
static double B200_ComputeCFL(void *s, void *m, double dt, double t)
{
  double cfl = -1.0;
  (void)m;
  if (g.resident) hpb_dev_ComputeCFL(H(s), &cfl);
  else            hpb_ComputeCFL(H(s), ((HyPar*)s)->u, dt, t, &cfl);
  die_on_error("ComputeCFL");
  return cfl;
}

/* ---- HyPar::VolumeIntegralFunction (VolumeIntegral.c), called by TimePostStep.c:83 when ConservationCheck = yes */
static int (*ref_VolumeIntegral)(double*, double*, void*, void*);
static int B200_VolumeIntegral(double *VolumeIntegral, double *u, void *s, void *m)
{
  HyPar *solver = (HyPar*) s;
  if (g.resident && u == solver->u) {
    double local[HPB_MAX_NVARS];
    hpb_dev_VolumeIntegral(H(s), local); die_on_error("VolumeIntegral");
    return MPISum_double(VolumeIntegral, local, solver->nvars, &((MPIVariables*)m)->world);
  }
  return ref_VolumeIntegral(VolumeIntegral, u, s, m);     /* some other host array: HyPar's own host code */
}

/* ---- HyPar::TimeIntegrate (hypar.h:220) = TimeRK (TimeRK.c:35): every simulation of the ensemble advances by one step
   (with explicit RK they never exchange data: TimeRK.c:50-93 walks one concatenated vector, simulation after simulation) */
static int B200_TimeRK(void *ts)
{
  TimeIntegration  *TS  = (TimeIntegration*) ts;
  SimulationObject *sim = (SimulationObject*) TS->simulation;
  for (int ns = 0; ns < g.nsims; ns++) {
    HyPar *solver = &sim[ns].solver;
    hpb_solver *h = H(solver);
    const int cons = !strcmp(solver->ConservationCheck, "yes");
    if (!g.resident) {
      hpb_TimeIntegrate(h, solver->u, 1, TS->waqt); die_on_error("TimeIntegrate");
    } else {
      if (sim[ns].mpi.nproc > 1) { hpb_TimeStepsDistributed(h, 1); die_on_error("TimeStepsDistributed"); }
      else                       { hpb_TimeStep(h); die_on_error("TimeStep"); }
      /* the host mirror is read by: TimePreStep.c:84 + TimePostStep.c:44-63 (screen norm, steps with
         (iter+1) % screen_op_iter == 0: the copy is taken BEFORE that step, so the step before it refreshes too),
         OutputSolution (file_op_iter), CalculateError and the final output */
      const int it = TS->iter + 1;
      const int refresh = (it % solver->screen_op_iter == 0) || ((it + 1) % solver->screen_op_iter == 0)
                       || (it % solver->file_op_iter == 0) || (it == TS->n_iter);
      if (refresh) { hpb_dev_get_solution(h, solver->u); die_on_error("get_solution"); }
    }
    if (cons) {    /* TimeRK.c:182-193 leaves the step's flux integrals in solver->StepBoundaryIntegral */
      hpb_dev_StepBoundaryIntegral(h, solver->StepBoundaryIntegral); die_on_error("StepBoundaryIntegral");
    }
  }
  g.steps++;
  return 0;
}

/* ---- HyPar::TimeIntegrate = TimeGLMGEE (TimeGLMGEE.c:45): the solution and the auxiliary solution TS->U[r] advance
   together. TimeGetAuxSolutions (the op_aux files of OutputSolution.cpp:42-79) and TimeError (glm_err.dat) read TS->U[r],
   so it is refreshed from the device whenever the solution's host mirror is. */
static int B200_TimeGLMGEE(void *ts)
{
  TimeIntegration  *TS  = (TimeIntegration*) ts;
  SimulationObject *sim = (SimulationObject*) TS->simulation;
  for (int ns = 0; ns < g.nsims; ns++) {
    HyPar *solver = &sim[ns].solver;
    hpb_solver *h = H(solver);
    GLMGEEParameters *p = (GLMGEEParameters*) solver->msti;
    double *aux = TS->U[p->r] + TS->u_offsets[ns];
    const int cons = !strcmp(solver->ConservationCheck, "yes");
    if (!g.resident) {         /* the host arrays are the state: both go up, one step, both come back */
      hpb_dev_set_solution(h, solver->u);  hpb_dev_set_aux_solution(h, aux);  die_on_error("set_solution");
      hpb_TimeStep(h);                                                         die_on_error("TimeStep");
      hpb_dev_get_solution(h, solver->u);  hpb_dev_get_aux_solution(h, aux);  die_on_error("get_solution");
    } else {
      if (sim[ns].mpi.nproc > 1) { hpb_TimeStepsDistributed(h, 1); die_on_error("TimeStepsDistributed"); }
      else                       { hpb_TimeStep(h); die_on_error("TimeStep"); }
      const int it = TS->iter + 1;
      const int refresh = (it % solver->screen_op_iter == 0) || ((it + 1) % solver->screen_op_iter == 0)
                       || (it % solver->file_op_iter == 0) || (it == TS->n_iter);
      if (refresh) { hpb_dev_get_solution(h, solver->u); hpb_dev_get_aux_solution(h, aux); die_on_error("get_solution"); }
    }
    if (cons) { hpb_dev_StepBoundaryIntegral(h, solver->StepBoundaryIntegral); die_on_error("StepBoundaryIntegral"); }
  }
  g.steps++;
  return 0;
}

static int upwind_choice(const char *name)
{
  if (!strcmp(name, _RUSANOV_)) return HPB_UPWIND_RUSANOV;
  if (!strcmp(name, _ROE_))     return HPB_UPWIND_ROE;
  if (!strcmp(name, _RF_))      return HPB_UPWIND_RF;
  if (!strcmp(name, _LLF_))     return HPB_UPWIND_LLF;
  return -1;
}

static int bc_type(const char *name)
{
  if (!strcmp(name, _PERIODIC_))    return HPB_BC_PERIODIC;
  if (!strcmp(name, _EXTRAPOLATE_)) return HPB_BC_EXTRAPOLATE;
  if (!strcmp(name, _SLIP_WALL_))   return HPB_BC_SLIP_WALL;
  if (!strcmp(name, _NOSLIP_WALL_))          return HPB_BC_NOSLIP_WALL;
  if (!strcmp(name, _DIRICHLET_))            return HPB_BC_DIRICHLET;
  if (!strcmp(name, _SUBSONIC_INFLOW_))      return HPB_BC_SUBSONIC_INFLOW;
  if (!strcmp(name, _SUBSONIC_OUTFLOW_))     return HPB_BC_SUBSONIC_OUTFLOW;
  if (!strcmp(name, _SUBSONIC_AMBIVALENT_))  return HPB_BC_SUBSONIC_AMBIVALENT;
  if (!strcmp(name, _SUPERSONIC_INFLOW_))    return HPB_BC_SUPERSONIC_INFLOW;
  if (!strcmp(name, _SUPERSONIC_OUTFLOW_))   return HPB_BC_SUPERSONIC_OUTFLOW;
  if (!strcmp(name, _SPONGE_))               return HPB_BC_SPONGE;
  return -1;
}

static int attach_one(SimulationObject *sim, int n, int nsims)
{
  HyPar *s = &sim[n].solver;
  MPIVariables *mpi = &sim[n].mpi;
  hpb_solver *h = NULL;
  if (mpi->nproc != 1 && !g.resident) {
    fprintf(stderr, "hyparb200_attach: several ranks need the resident mode (the host-array entry points are single-rank)\n");
    return 1;
  }
  int scheme = -1;
  if      (!strcmp(s->spatial_scheme_hyp, _FIFTH_ORDER_WENO_))           scheme = HPB_SCHEME_WENO5;
  else if (!strcmp(s->spatial_scheme_hyp, _FIFTH_ORDER_CRWENO_))         scheme = HPB_SCHEME_CRWENO5;
  else if (!strcmp(s->spatial_scheme_hyp, _FIFTH_ORDER_HCWENO_))         scheme = HPB_SCHEME_HCWENO5;
  else if (!strcmp(s->spatial_scheme_hyp, _FIFTH_ORDER_COMPACT_UPWIND_)) scheme = HPB_SCHEME_CUPW5;
  else if (!strcmp(s->spatial_scheme_hyp, _FIFTH_ORDER_UPWIND_))         scheme = HPB_SCHEME_UPW5;
  else if (!strcmp(s->spatial_scheme_hyp, _FIRST_ORDER_UPWIND_))         scheme = HPB_SCHEME_FIRST;
  else if (!strcmp(s->spatial_scheme_hyp, _SECOND_ORDER_CENTRAL_))       scheme = HPB_SCHEME_SECOND;
  else if (!strcmp(s->spatial_scheme_hyp, _FOURTH_ORDER_CENTRAL_))       scheme = HPB_SCHEME_FOURTH;
  else if (!strcmp(s->spatial_scheme_hyp, _SECOND_ORDER_MUSCL_))         scheme = HPB_SCHEME_MUSCL2;
  else if (!strcmp(s->spatial_scheme_hyp, _THIRD_ORDER_MUSCL_))          scheme = HPB_SCHEME_MUSCL3;
  const int glm = !strcmp(s->time_scheme, _GLM_GEE_);
  if (scheme < 0 || (strcmp(s->time_scheme, _RK_) && !glm) || strcmp(s->SplitHyperbolicFlux, "no") || s->flag_ib) {
    fprintf(stderr, "hyparb200_attach: only weno5 / crweno5 / hcweno5 / cupw5 / upw5 / 1 / 2 / 4 / muscl2 / muscl3 + explicit RK / GLM-GEE without flux splitting / "
                    "immersed boundaries is on the B200 path\n");
    return 1;
  }

  hpb_config c;
  double *advf = NULL;              /* LinearADR varying advection field, global, freed after hpb_create */
  if (hpb_sizeof_config() != sizeof(hpb_config)) {
    fprintf(stderr, "hyparb200_attach: include/hypar_b200.h does not match libhypar_b200.so (hpb_config layout)\n");
    return 1;
  }
  hpb_config_defaults(&c);
  c.ndims = s->ndims;  c.nvars = s->nvars;  c.ghosts = s->ghosts;  c.rank = mpi->rank;  c.dt = s->dt;
  for (int d = 0; d < s->ndims; d++) { c.dim_global[d] = s->dim_global[d]; c.iproc[d] = mpi->iproc[d]; }
  c.interp_char = !strcmp(s->interp_type, _CHARACTERISTIC_);
  if (glm) {
    const char *names[] = { _GLM_GEE_23_, _GLM_GEE_24_, _GLM_GEE_25I_, _GLM_GEE_35_, _GLM_GEE_EXRK2A_, _GLM_GEE_RK32G1_, _GLM_GEE_RK285EX_ };
    c.rk_type = -1;
    for (int k = 0; k < 7; k++) if (!strcmp(s->time_scheme_type, names[k])) c.rk_type = HPB_GLMGEE_23 + k;
    c.glm_ee_mode = !strcmp(((GLMGEEParameters*) s->msti)->ee_mode, _GLM_GEE_YYT_) ? HPB_GLM_YYT : HPB_GLM_YEPS;
  }
  else if (!strcmp(s->time_scheme_type, _RK_44_))     c.rk_type = HPB_RK_44;
  else if (!strcmp(s->time_scheme_type, _RK_SSP3_) || !strcmp(s->time_scheme_type, _RK_TVD3_)) c.rk_type = HPB_RK_SSPRK3;
  else if (!strcmp(s->time_scheme_type, _RK_1FE_))    c.rk_type = HPB_RK_1FE;
  else if (!strcmp(s->time_scheme_type, _RK_22_))     c.rk_type = HPB_RK_22;
  else if (!strcmp(s->time_scheme_type, _RK_33_))     c.rk_type = HPB_RK_33;
  else { fprintf(stderr, "hyparb200_attach: rk type %s is not on the B200 path (44, ssprk3)\n", s->time_scheme_type); return 1; }
  c.par_scheme = atoi(s->spatial_scheme_par);
  c.par_space_type = !strcmp(s->spatial_type_par, _NC_1STAGE_) ? HPB_PAR_NC_1STAGE : !strcmp(s->spatial_type_par, _NC_1_5STAGE_) ? HPB_PAR_NC_1_5STAGE
                   : !strcmp(s->spatial_type_par, _NC_2STAGE_) ? HPB_PAR_NC_2STAGE : HPB_PAR_CONS_1STAGE;
  c.conservation_check = !strcmp(s->ConservationCheck, "yes");
  c.hyp_scheme = scheme;
  if (scheme == HPB_SCHEME_MUSCL2 || scheme == HPB_SCHEME_MUSCL3) {
    MUSCLParameters *mp = (MUSCLParameters*) s->interp;
    c.muscl_eps = mp->eps;
    c.muscl_limiter = !strcmp(mp->limiter_type, _LIM_MM_) ? HPB_LIMITER_MINMOD : !strcmp(mp->limiter_type, _LIM_VANLEER_) ? HPB_LIMITER_VANLEER
                    : !strcmp(mp->limiter_type, _LIM_SUPERBEE_) ? HPB_LIMITER_SUPERBEE : HPB_LIMITER_GMM;
  }
  if (s->lusolver) {               /* compact schemes: lusolver.inp as tridiagLUInit read it (reduced system across ranks) */
    TridiagLU *lu = (TridiagLU*) s->lusolver;
    c.lu_maxiter = lu->maxiter;  c.lu_evaluate_norm = lu->evaluate_norm;  c.lu_atol = lu->atol;  c.lu_rtol = lu->rtol;
    c.lu_gather_and_solve = !strcmp(lu->reducedsolvetype, _TRIDIAG_GS_);
  }
  if (scheme == HPB_SCHEME_WENO5 || scheme == HPB_SCHEME_CRWENO5 || scheme == HPB_SCHEME_HCWENO5) {   /* s->interp is NULL for the linear schemes */
    WENOParameters *w = (WENOParameters*) s->interp;
    c.weno_type = w->yc ? HPB_WENO_YC : w->borges ? HPB_WENO_Z : w->mapped ? HPB_WENO_M : HPB_WENO_JS;
    c.no_limiting = w->no_limiting;  c.weno_eps = w->eps;  c.weno_rc = w->rc;  c.weno_xi = w->xi;
  }

  if (!strcmp(s->model, _NAVIER_STOKES_3D_)) {
    NavierStokes3D *p = (NavierStokes3D*) s->physics;
    /* HyPar already divided Re by Minf (NavierStokes3DInitialize.c:368); Minf enters the path only through that quotient:
       pass the scaled value with Minf = 1 so that it is not multiplied and divided again (1 ulp when Minf is not 2^k) */
    c.model = HPB_MODEL_NS3D;  c.gamma = p->gamma;  c.Pr = p->Pr;  c.Minf = 1.0;
    c.Re = p->Re;
    c.upwind = upwind_choice(p->upw_choice);
    c.gravity[0] = p->grav_x; c.gravity[1] = p->grav_y; c.gravity[2] = p->grav_z;
    c.rho_ref = p->rho0; c.p_ref = p->p0; c.R = p->R; c.HB = p->HB; c.N_bv = p->N_bv;
    if (c.Re > 0 && strcmp(s->spatial_type_par, _NC_2STAGE_)) {
      fprintf(stderr, "hyparb200_attach: viscous NavierStokes3D needs par_space_type %s (got %s)\n", _NC_2STAGE_, s->spatial_type_par);
      return 1;
    }
  } else if (!strcmp(s->model, _NAVIER_STOKES_2D_)) {
    NavierStokes2D *p = (NavierStokes2D*) s->physics;
    c.model = HPB_MODEL_NS2D;  c.gamma = p->gamma;  c.Pr = p->Pr;  c.Minf = 1.0;
    c.Re = p->Re;                                   /* already / Minf: NavierStokes2DInitialize.c:205 */
    c.upwind = upwind_choice(p->upw_choice);
    c.gravity[0] = p->grav_x; c.gravity[1] = p->grav_y;
    c.rho_ref = p->rho0; c.p_ref = p->p0; c.R = p->R; c.HB = p->HB; c.N_bv = p->N_bv;
  } else if (!strcmp(s->model, _EULER_1D_)) {
    Euler1D *p = (Euler1D*) s->physics;
    c.model = HPB_MODEL_EULER1D;  c.gamma = p->gamma;
    c.upwind = upwind_choice(p->upw_choice);
    c.gravity[0] = p->grav;  c.gravity_type = p->grav_type;
  } else if (!strcmp(s->model, "burgers")) {            /* _BURGERS_ (physicalmodels/burgers.h): no parameters */
    c.model = HPB_MODEL_BURGERS;  c.upwind = HPB_UPWIND_DEFAULT;
  } else if (!strcmp(s->model, _LINEAR_ADVECTION_DIFFUSION_REACTION_)) {
    LinearADR *p = (LinearADR*) s->physics;
    c.model = HPB_MODEL_LINEAR_ADR;  c.upwind = HPB_UPWIND_DEFAULT;
    if (strcmp(p->centered_flux, "no") || s->nvars != 1) {
      fprintf(stderr, "hyparb200_attach: LinearADR needs upwinded fluxes and nvars = 1 on the B200 path\n"); return 1;
    }
    for (int i = 0; i < s->ndims * s->nvars; i++) c.diffusion[i] = p->d[i];
    if (p->constant_advection == 1) {
      for (int i = 0; i < s->ndims * s->nvars; i++) c.advection[i] = p->a[i];
    } else if (p->constant_advection == 0 && mpi->nproc != 1) {
      fprintf(stderr, "hyparb200_attach: a spatially varying advection field with several ranks is not wired up in this glue\n");
      return 1;
    } else if (p->constant_advection == 0) {
      /* spatially varying field (LinearADRAdvectionField.c): p->a is this rank's ghost-padded block, [point][ndims*nvars];
         one rank, so its interior IS the global field the C-ABI takes (hpb_create rebuilds the ghosts) */
      const int nc = s->ndims * s->nvars;
      advf = (double*) calloc((size_t) s->npoints_global * nc, sizeof(double));
      int idx[3] = {0, 0, 0}, done = 0;  long q = 0;
      while (!done) {
        int pg; _ArrayIndex1D_(s->ndims, s->dim_local, idx, s->ghosts, pg);
        for (int k = 0; k < nc; k++) advf[q * nc + k] = p->a[(long) pg * nc + k];
        q++;
        _ArrayIncrementIndex_(s->ndims, s->dim_local, idx, done);
      }
      c.advection_field = advf;
    }
  } else {
    fprintf(stderr, "hyparb200_attach: model %s is not on the B200 path\n", s->model); return 1;
  }
  if (c.upwind < 0) { fprintf(stderr, "hyparb200_attach: upwinding scheme is not on the B200 path (roe, rusanov, rf-char, llf-char)\n"); return 1; }

  DomainBoundary *b = (DomainBoundary*) s->boundary;
  if (s->nBoundaryZones > HPB_MAX_ZONES) {
    fprintf(stderr, "hyparb200_attach: %d boundary zones (at most %d)\n", s->nBoundaryZones, HPB_MAX_ZONES);
    return 1;
  }
  c.nzones = s->nBoundaryZones;
  for (int n = 0; n < c.nzones; n++) {
    c.zones[n].type = bc_type(b[n].bctype);
    if (c.zones[n].type < 0) { fprintf(stderr, "hyparb200_attach: boundary type %s is not on the B200 path\n", b[n].bctype); return 1; }
    c.zones[n].dim = b[n].dim;  c.zones[n].face = b[n].face;
    for (int d = 0; d < s->ndims; d++) {
      c.zones[n].xmin[d] = b[n].xmin[d];  c.zones[n].xmax[d] = b[n].xmax[d];
      c.zones[n].wall_velocity[d] = b[n].FlowVelocity ? b[n].FlowVelocity[d] : 0.0;   /* allocated only where read */
    }
    c.zones[n].flow_density = b[n].FlowDensity;  c.zones[n].flow_pressure = b[n].FlowPressure;
    if (b[n].DirichletValue) for (int v = 0; v < s->nvars; v++) c.zones[n].dirichlet[v] = b[n].DirichletValue[v];
    if (b[n].SpongeValue)    for (int v = 0; v < s->nvars; v++) c.zones[n].dirichlet[v] = b[n].SpongeValue[v];
  }

  /* global grid, concatenated per dimension as in initial.inp (ReadArray.c:225-256): the interior parts of the ranks'
     solver->x (layout include/basic.h:31-36), gathered on rank 0 and broadcast with HyPar's own helpers */
  int ntot = 0;
  for (int d = 0; d < s->ndims; d++) ntot += s->dim_global[d];
  double *xg = (double*) calloc(ntot, sizeof(double));
  for (int d = 0, off = 0, offg = 0; d < s->ndims; d++) {
    if (mpi->nproc == 1) {
      for (int i = 0; i < s->dim_local[d]; i++) xg[offg + i] = s->x[off + s->ghosts + i];
    } else {
      MPIGatherArray1D(mpi, (mpi->rank ? NULL : xg + offg), s->x + off, mpi->is[d], mpi->ie[d], s->dim_local[d], s->ghosts);
      MPIBroadcast_double(xg + offg, s->dim_global[d], 0, &mpi->world);
    }
    off += s->dim_local[d] + 2 * s->ghosts;  offg += s->dim_global[d];
  }
  c.x_global = xg;
  /* one GPU per rank (as the reference's gpu_device_no + rank would); HYPARB200_DEVICE pins a device for one-rank runs */
  const int ndev = hpb_device_count();
  c.device = getenv("HYPARB200_DEVICE") ? atoi(getenv("HYPARB200_DEVICE")) : (ndev > 0 ? mpi->rank % ndev : 0);
  const char *fz = getenv("HYPARB200_USE_FUSED");
  if (fz) c.use_fused = atoi(fz);
  int rc = hpb_create(&c, &h);
  free(xg);
  if (advf) free(advf);
  if (rc) { fprintf(stderr, "hyparb200_attach: %s\n", hpb_last_error()); return 1; }   /* unsupported choices fail here */

  if (mpi->nproc > 1) {
    /* the library's NCCL transport: the communicator id travels over HyPar's own MPI world */
    char id[HPB_COMM_ID_BYTES];
    memset(id, 0, sizeof(id));
    if (!mpi->rank && hpb_comm_get_unique_id(id)) { fprintf(stderr, "hyparb200_attach: %s\n", hpb_last_error()); return 1; }
    MPIBroadcast_character(id, HPB_COMM_ID_BYTES, 0, &mpi->world);
    if (hpb_comm_init_nccl(h, id, mpi->nproc)) { fprintf(stderr, "hyparb200_attach: %s\n", hpb_last_error()); return 1; }
    const char *ov = getenv("HYPARB200_OVERLAP");
    if (ov) hpb_set_overlap(h, atoi(ov));
  }
  g.sim[n].h = h;  g.sim[n].solver = s;  g.sim[n].mpi = mpi;
  if (g.resident && hpb_dev_set_solution(h, s->u)) { fprintf(stderr, "hyparb200_attach: %s\n", hpb_last_error()); return 1; }
  if (g.resident) s->PreStep = NULL;     /* see the header comment: host work on the stale mirror that nothing reads */

  /* the pointer swap (what InitializeSolvers.c:71-109 / NavierStokes3DInitialize.c:389-446 do for use_gpu) */
  s->ApplyBoundaryConditions  = B200_BC;
  s->HyperbolicFunction       = B200_Hyp;
  s->ParabolicFunction        = B200_Par;
  s->SourceFunction           = B200_Src;
  s->FFunction                = B200_FFunction;
  if (s->UFunction) s->UFunction = B200_UFunction;
  s->Upwind                   = B200_Upwind;
  if (s->SetInterpLimiterVar) s->SetInterpLimiterVar = B200_SetInterpLimiterVar;   /* NULL for the linear schemes */
  s->InterpolateInterfacesHyp = B200_InterpolateInterfacesHyp;
  s->FirstDerivativePar       = B200_FirstDerivativePar;
  s->SecondDerivativePar      = B200_SecondDerivativePar;
  s->ComputeCFL               = B200_ComputeCFL;
  ref_VolumeIntegral          = s->VolumeIntegralFunction;
  s->VolumeIntegralFunction   = B200_VolumeIntegral;
  s->TimeIntegrate            = glm ? B200_TimeGLMGEE : B200_TimeRK;        /* picked up by TimeInitialize.c:55 */
  if (!mpi->rank) printf("hypar_b200 attached%s: %s, %s mode, device %d%s\n", (nsims > 1 ? " (one solver per simulation)" : ""),
                         hpb_version(), g.resident ? "resident" : "host", c.device,
                         mpi->nproc > 1 ? ", in-library NCCL halo exchange" : "");
  return 0;
}

int hyparb200_attach(void *sims, int nsims)
{
  SimulationObject *sim = (SimulationObject*) sims;
  if (nsims < 1 || nsims > B200_MAXSIM) { fprintf(stderr, "hyparb200_attach: %d simulations (1..%d)\n", nsims, B200_MAXSIM); return 1; }
  const char *mode = getenv("HYPARB200_MODE");
  g.resident = !(mode && !strcmp(mode, "host"));
  g.nsims = 0;
  for (int n = 0; n < nsims; n++) {
    if (attach_one(sim, n, nsims)) return 1;
    g.nsims = n + 1;
  }
  return 0;
}

int hyparb200_detach(void)
{
  long long launches = 0;
  for (int n = 0; n < g.nsims; n++) if (g.sim[n].h) launches += hpb_kernel_launch_count(g.sim[n].h);
  if (g.nsims && !g.sim[0].mpi->rank) printf("hypar_b200: %lld steps, %lld kernel launches\n", g.steps, launches);
  for (int n = 0; n < g.nsims; n++) { if (g.sim[n].h) hpb_destroy(g.sim[n].h); g.sim[n].h = NULL; }
  g.nsims = 0;
  return 0;
}

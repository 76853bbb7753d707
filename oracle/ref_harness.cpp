/* ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A small driver of our own that links against the UNMODIFIED reference objects
 * compiled in place from /root/reference (see oracle/Makefile) and runs the
 * reference's own initialisation sequence (the one in src/main.cpp:334-420),
 * then, instead of only calling Solve(), exposes the hot path piecewise so the
 * parity tests can compare every function on the solver function-pointer
 * surface:
 *
 *   hypar_ref run            -> exactly what HyPar does: Solve() + output files
 *   hypar_ref rhs [t]        -> TimeRHSFunctionExplicit once (TimeRHSFunctionExplicit.c:30)
 *                               dumps ref_u.bin (post-BC), ref_hyp.bin, ref_par.bin,
 *                               ref_source.bin, ref_rhs.bin
 *   hypar_ref pieces         -> per-direction dumps of FFunction, UFunction, the WENO
 *                               weights, uL/uR/fL/fR, Upwind result (HyperbolicFunction.c:167-222)
 *                               + FirstDerivativePar / SecondDerivativePar / ComputeCFL
 *   hypar_ref steps N        -> TimeInitialize + N x (TimePreStep, TimeStep, TimePostStep);
 *                               dumps ref_ufinal.bin (with ghosts) and prints per-step wctime; with
 *                               `conservation_check yes` also the volume / boundary-flux integrals and the
 *                               conservation error of every step; with an `exact.inp` in the directory also
 *                               CalculateError's three norms (errors.dat); with `time_scheme glm-gee` also
 *                               ref_uaux.bin (the auxiliary solution TS->U[r]) and TimeError's six norms (glm_err.dat)
 *   hypar_ref glmgee         -> the coefficient tables TimeGLMGEEInitialize.c builds for the method and ee_mode of the
 *                               run directory (solver.inp time_scheme_type, glm_gee.inp), as hexadecimal floats:
 *                               tools/make_glmgee_tables.py turns them into the tables of the oracle and of the library
 *
 * All dumps: header {int ndims, nvars, ghosts, dim[ndims]} then raw doubles in the
 * reference's own layout (ghost-padded AoS for cell arrays).
 * Works in the current directory, which must hold solver.inp, boundary.inp,
 * physics.inp, initial.inp (+ optional weno.inp) like any HyPar run.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <vector>
#include <string>

#ifndef serial
#include <mpi.h>
#endif

#include <basic.h>
#include <simulation.h>          /* extern "C" ReadInputs/Initialize/...; Solve() (C++) */
#include <timeintegration_cpp.h>
#include <mpivars_cpp.h>
#include <interpolation.h>

extern "C" int TimeRHSFunctionExplicit(double*, double*, void*, void*, double);
extern "C" int CalculateError(void*, void*);
extern "C" int TimeError(void*, void*, double*);

static int g_rank = 0, g_nproc = 1;
static void dump(const char* name_in, const HyPar* s, const double* a, long n)
{
  /* several ranks (the multi-process shim): ref_<what>.bin -> ref_<what>.r<rank>.bin, every rank dumps its own block */
  char name[512];
  if (g_nproc > 1) {
    const char* dot = strrchr(name_in, '.');
    snprintf(name, sizeof(name), "%.*s.r%04d%s", (int)(dot - name_in), name_in, g_rank, dot);
  } else snprintf(name, sizeof(name), "%s", name_in);
  FILE* f = fopen(name, "wb");
  if (!f) { fprintf(stderr, "cannot write %s\n", name); exit(2); }
  int hdr[3] = { s->ndims, s->nvars, s->ghosts };
  fwrite(hdr, sizeof(int), 3, f);
  fwrite(s->dim_local, sizeof(int), s->ndims, f);
  fwrite(a, sizeof(double), n, f);
  fclose(f);
}

static long ncells_g(const HyPar* s)
{
  long n = 1;
  for (int d = 0; d < s->ndims; d++) n *= (s->dim_local[d] + 2*s->ghosts);
  return n;
}

static long ninterfaces(const HyPar* s, int dir)
{
  long n = 1;
  for (int d = 0; d < s->ndims; d++) n *= (s->dim_local[d] + (d == dir ? 1 : 0));
  return n;
}

int main(int argc, char** argv)
{
  const char* mode = (argc > 1 ? argv[1] : "run");
  int rank = 0, nproc = 1, ierr = 0;
#ifndef serial
  MPI_Init(&argc, &argv);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &nproc);
#endif
  g_rank = rank; g_nproc = nproc;

  SimulationObject* sim = new SimulationObject;
  memset(sim, 0, sizeof(SimulationObject));
  sim->solver.my_idx = 0;
  sim->solver.nsims  = 1;
  sim->mpi.rank  = rank;
  sim->mpi.nproc = nproc;
#ifndef serial
  MPI_Comm_dup(MPI_COMM_WORLD, &sim->mpi.world);
#endif

  /* the reference's own start-up sequence (src/main.cpp:361-420) */
  ierr = ReadInputs(sim, 1, rank);                     if (ierr) return ierr;
  ierr = Initialize(sim, 1);                           if (ierr) return ierr;
  ierr = InitialSolution(sim, 1);                      if (ierr) return ierr;
  ierr = InitializeBoundaries(sim, 1);                 if (ierr) return ierr;
  ierr = InitializeImmersedBoundaries(sim, 1);         if (ierr) return ierr;
  ierr = InitializeSolvers(sim, 1);                    if (ierr) return ierr;
  ierr = InitializePhysics(sim, 1);                    if (ierr) return ierr;
  ierr = InitializePhysicsData(sim, 0, 1, NULL);       if (ierr) return ierr;

  HyPar*        solver = &sim->solver;
  MPIVariables* mpi    = &sim->mpi;
  const long nc = ncells_g(solver) * solver->nvars;

  if (!strcmp(mode, "run")) {

    ierr = Solve(sim, 1, rank, nproc); if (ierr) return ierr;

  } else if (!strcmp(mode, "rhs")) {

    double t = (argc > 2 ? atof(argv[2]) : 0.0);
    std::vector<double> rhs(nc, 0.0);
    TimeRHSFunctionExplicit(rhs.data(), solver->u, solver, mpi, t);
    dump("ref_u.bin",      solver, solver->u,      nc);
    dump("ref_hyp.bin",    solver, solver->hyp,    nc);
    dump("ref_par.bin",    solver, solver->par,    nc);
    dump("ref_source.bin", solver, solver->source, nc);
    dump("ref_rhs.bin",    solver, rhs.data(),     nc);
    dump("ref_dxinv.bin",  solver, solver->dxinv,  solver->size_x);
    dump("ref_x.bin",      solver, solver->x,      solver->size_x);

  } else if (!strcmp(mode, "pieces")) {

    /* boundary conditions + halo, as TimeRHSFunctionExplicit.c:46-60 */
    solver->ApplyBoundaryConditions(solver, mpi, solver->u, NULL, 0.0);
    MPIExchangeBoundariesnD(solver->ndims, solver->nvars, solver->dim_local,
                            solver->ghosts, mpi, solver->u);
    dump("ref_u.bin", solver, solver->u, nc);

    int offset = 0;
    for (int d = 0; d < solver->ndims; d++) {
      char fn[256];
      double* x = solver->x + offset;
      const long ni = ninterfaces(solver, d) * solver->nvars;

      /* the sequence of HyperbolicFunction.c:81-83 and ReconstructHyperbolic :201-218 */
      solver->FFunction(solver->fluxC, solver->u, d, solver, 0.0);
      snprintf(fn, 256, "ref_fluxC_%d.bin", d); dump(fn, solver, solver->fluxC, nc);

      if (solver->SetInterpLimiterVar)
        solver->SetInterpLimiterVar(solver->fluxC, solver->u, x, d, solver, mpi);
      if (solver->interp && solver->SetInterpLimiterVar) {
        WENOParameters* weno = (WENOParameters*) solver->interp;
        /* blocks [LF | LU | RF | RU], each weno->size, direction slice at offset[d] */
        std::vector<double> w(12 * ni);
        for (int blk = 0; blk < 4; blk++) {
          memcpy(&w[(3*blk+0)*ni], weno->w1 + blk*weno->size + weno->offset[d], ni*sizeof(double));
          memcpy(&w[(3*blk+1)*ni], weno->w2 + blk*weno->size + weno->offset[d], ni*sizeof(double));
          memcpy(&w[(3*blk+2)*ni], weno->w3 + blk*weno->size + weno->offset[d], ni*sizeof(double));
        }
        snprintf(fn, 256, "ref_weights_%d.bin", d); dump(fn, solver, w.data(), 12*ni);
      }

      double* uC = solver->u;
      if (solver->UFunction) {
        uC = solver->uC;
        solver->UFunction(uC, solver->u, d, solver, mpi, 0.0);
      }
      snprintf(fn, 256, "ref_uC_%d.bin", d); dump(fn, solver, uC, nc);

      solver->InterpolateInterfacesHyp(solver->uL, uC,            solver->u, x,  1, d, solver, mpi, 1);
      solver->InterpolateInterfacesHyp(solver->uR, uC,            solver->u, x, -1, d, solver, mpi, 1);
      solver->InterpolateInterfacesHyp(solver->fL, solver->fluxC, solver->u, x,  1, d, solver, mpi, 0);
      solver->InterpolateInterfacesHyp(solver->fR, solver->fluxC, solver->u, x, -1, d, solver, mpi, 0);
      snprintf(fn, 256, "ref_uL_%d.bin", d); dump(fn, solver, solver->uL, ni);
      snprintf(fn, 256, "ref_uR_%d.bin", d); dump(fn, solver, solver->uR, ni);
      snprintf(fn, 256, "ref_fL_%d.bin", d); dump(fn, solver, solver->fL, ni);
      snprintf(fn, 256, "ref_fR_%d.bin", d); dump(fn, solver, solver->fR, ni);

      solver->Upwind(solver->fluxI, solver->fL, solver->fR, solver->uL, solver->uR,
                     solver->u, d, solver, 0.0);
      snprintf(fn, 256, "ref_fluxI_%d.bin", d); dump(fn, solver, solver->fluxI, ni);

      /* derivative operators on u itself (un-scaled differences) */
      if (solver->FirstDerivativePar) {
        std::vector<double> D1(nc, 0.0);
        solver->FirstDerivativePar(D1.data(), solver->u, d, 1, solver, mpi);
        snprintf(fn, 256, "ref_D1_%d.bin", d); dump(fn, solver, D1.data(), nc);
      }
      if (solver->SecondDerivativePar) {
        std::vector<double> D2(nc, 0.0);
        solver->SecondDerivativePar(D2.data(), solver->u, d, solver, mpi);
        snprintf(fn, 256, "ref_D2_%d.bin", d); dump(fn, solver, D2.data(), nc);
      }
      offset += solver->dim_local[d] + 2*solver->ghosts;
    }
    if (solver->ComputeCFL) {
      double cfl = solver->ComputeCFL(solver, mpi, solver->dt, 0.0);
      FILE* f = fopen("ref_cfl.txt", "w"); fprintf(f, "%.17e\n", cfl); fclose(f);
    }

  } else if (!strcmp(mode, "steps")) {

    int nsteps = (argc > 2 ? atoi(argv[2]) : 1);
    TimeIntegration TS;
    TimeInitialize(sim, 1, rank, nproc, &TS);
    if (!strcmp(solver->ConservationCheck, "yes")) {
      printf("CONS0");
      for (int v = 0; v < solver->nvars; v++) printf(" %.17e", solver->VolumeIntegralInitial[v]);
      printf("\n");
    }
    double total = 0.0;
    for (TS.iter = TS.restart_iter; TS.iter < TS.restart_iter + nsteps; TS.iter++) {
      TimePreStep(&TS);
      TimeStep(&TS);
      TimePostStep(&TS);
      if (!rank) printf("STEP %d wctime %.6e norm %.17e maxcfl %.17e\n", TS.iter+1, TS.iter_wctime, TS.norm, TS.max_cfl);
      if (!strcmp(solver->ConservationCheck, "yes")) {
        /* conservation diagnostics of TimePostStep.c:81-93 (VolumeIntegral.c, BoundaryIntegral.c,
           CalculateConservationError.c) and the per-face flux integrals TimeRK.c:182-193 accumulates */
        printf("CONS %d", TS.iter+1);
        for (int v = 0; v < solver->nvars; v++) printf(" %.17e", solver->VolumeIntegral[v]);
        for (int v = 0; v < solver->nvars; v++) printf(" %.17e", solver->TotalBoundaryIntegral[v]);
        for (int v = 0; v < solver->nvars; v++) printf(" %.17e", solver->ConservationError[v]);
        printf("\nSTEPBI %d", TS.iter+1);
        for (int k = 0; k < 2*solver->ndims*solver->nvars; k++) printf(" %.17e", solver->StepBoundaryIntegral[k]);
        printf("\n");
      }
      total += TS.iter_wctime;
    }
    printf("TOTAL_WCTIME %.6e NSTEPS %d\n", total, nsteps);
    dump("ref_ufinal.bin", solver, solver->u, nc);
    const bool glmgee = !strcmp(solver->time_scheme, _GLM_GEE_);
    if (glmgee) dump("ref_uaux.bin", solver, TS.U[((GLMGEEParameters*) solver->msti)->r], nc);
    { /* CalculateError.c (errors.dat): only when the run directory holds exact.inp */
      FILE* fe = fopen("exact.inp", "rb");
      if (fe) {
        fclose(fe);
        CalculateError(solver, mpi);
        printf("ERRORS %.17e %.17e %.17e\n", solver->error[0], solver->error[1], solver->error[2]);
      } else if (glmgee) TimeError(solver, mpi, NULL);      /* what CalculateError.c:60 does without an exact solution */
      if (glmgee && !rank) {                                /* TimeError.c:121-127 wrote glm_err.dat */
        FILE* fg = fopen("glm_err.dat", "r");
        double v[7] = {0,0,0,0,0,0,0};
        if (fg) { for (int k = 0; k < 7; k++) if (fscanf(fg, "%lf", &v[k]) != 1) break; fclose(fg); }
        printf("GLMERR %.17e %.17e %.17e %.17e %.17e %.17e\n", v[1], v[2], v[3], v[4], v[5], v[6]);
      }
    }
    TimeCleanup(&TS);

  } else if (!strcmp(mode, "glmgee")) {

    if (strcmp(solver->time_scheme, _GLM_GEE_)) { fprintf(stderr, "time_scheme is not glm-gee\n"); return 4; }
    GLMGEEParameters* p = (GLMGEEParameters*) solver->msti;
    const int s = p->nstages, r = p->r;
    printf("GLMGEE %s %s %d %d %a\n", solver->time_scheme_type, p->ee_mode, s, r, p->gamma);
    printf("A");  for (int k = 0; k < s*s; k++) printf(" %a", p->A[k]);  printf("\n");
    printf("B");  for (int k = 0; k < r*s; k++) printf(" %a", p->B[k]);  printf("\n");
    printf("C");  for (int k = 0; k < s*r; k++) printf(" %a", p->C[k]);  printf("\n");
    printf("D");  for (int k = 0; k < r*r; k++) printf(" %a", p->D[k]);  printf("\n");
    printf("c");  for (int k = 0; k < s;   k++) printf(" %a", p->c[k]);  printf("\n");

  } else {
    fprintf(stderr, "unknown mode %s\n", mode);
    return 3;
  }

#ifndef serial
  MPI_Finalize();
#endif
  return 0;
}

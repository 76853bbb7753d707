/* hypar_b200_main.cpp -- HyPar's own main() sequence for a single simulation (reference src/main.cpp:330-470,
 * through its SingleSimulation class) with ONE added call: hyparb200_attach() after the initialisations and
 * before Solve(). Everything else -- input parsing, Solve()'s time loop (TimePreStep / TimeStep / TimePostStep /
 * TimePrintStep), OutputSolution and its writers, CalculateError, SimWriteErrors -- is the unmodified reference,
 * linked from the objects oracle/Makefile builds in place from /root/reference. The result: HyPar's executable
 * with the explicit-RHS path running on the B200, writing the same op.bin / errors.dat / conservation.dat.
 *
 * Built by integration/Makefile into oracle/_ref/hypar_b200_dropin (test infrastructure for the drop-in claim;
 * tests/test_gpu_dropin.py compares its output files with the reference executable's). The same file compiled with
 * -DHYPARB200_NO_ATTACH and without the library is oracle/_ref/hypar_main_mpi1: the plain reference executable.
 */
#include <stdio.h>
#include <string.h>
#include <sys/time.h>
#ifndef serial
#include <mpi.h>
#endif
#include <basic.h>
#include <mpivars_cpp.h>
#include <simulation_library.h>

extern "C" int hyparb200_attach(void*, int);
extern "C" int hyparb200_detach(void);

class B200Simulation : public SingleSimulation {
  public:
    int attach() { return hyparb200_attach((void*) m_sim, 1); }
};
/* ensembles (simulation.inp present: src/main.cpp:278-318): every SimulationObject gets its own library solver */
class B200Ensemble : public EnsembleSimulation {
  public:
    int attach() { return hyparb200_attach((void*) m_sims.data(), m_nsims); }
};

int main(int argc, char** argv)
{
  int rank = 0, nproc = 1, ierr = 0;
  struct timeval main_start, solve_start, solve_end, main_end;
#ifndef serial
  MPI_Comm world;
  MPI_Init(&argc, &argv);
  MPI_Comm_dup(MPI_COMM_WORLD, &world);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &nproc);
#endif
  gettimeofday(&main_start, NULL);
  /* single or ensemble simulation, chosen as src/main.cpp:273-318 does (sparse grids are not on this path) */
  int ensemble = 0;
  if (!rank) { FILE* f = fopen(_ENSEMBLE_SIM_INP_FNAME_, "r"); if (f) { ensemble = 1; fclose(f); } }
#ifndef serial
  MPI_Bcast(&ensemble, 1, MPI_INT, 0, MPI_COMM_WORLD);
#endif
  B200Simulation* single = ensemble ? NULL : new B200Simulation;
  B200Ensemble*   many   = ensemble ? new B200Ensemble : NULL;
  Simulation* sim = ensemble ? (Simulation*) many : (Simulation*) single;
  if (ensemble && !rank) printf("-- Ensemble Simulation --\n");
  ierr = sim->define(rank, nproc);                 if (ierr) return ierr;
#ifndef serial
  ierr = sim->mpiCommDup();                        if (ierr) return ierr;
#endif
  ierr = sim->ReadInputs();                        if (ierr) return ierr;
  ierr = sim->Initialize();                        if (ierr) return ierr;
  ierr = sim->InitialSolution();                   if (ierr) return ierr;
  ierr = sim->InitializeBoundaries();              if (ierr) return ierr;
  ierr = sim->InitializeImmersedBoundaries();      if (ierr) return ierr;
  ierr = sim->InitializeSolvers();                 if (ierr) return ierr;
  ierr = sim->InitializePhysics();                 if (ierr) return ierr;
  ierr = sim->InitializePhysicsData();             if (ierr) return ierr;
  ierr = sim->InitializationWrapup();              if (ierr) return ierr;

#ifndef HYPARB200_NO_ATTACH
  ierr = ensemble ? many->attach() : single->attach();     /* <-- the one added call */
  if (ierr) { fprintf(stderr, "hyparb200_attach failed on process %d\n", rank); return ierr; }
#endif

  gettimeofday(&solve_start, NULL);
  ierr = sim->Solve();                             if (ierr) return ierr;
  gettimeofday(&solve_end, NULL);
  gettimeofday(&main_end, NULL);
  double main_runtime   = ((main_end.tv_sec  - main_start.tv_sec)  * 1000000LL + (main_end.tv_usec  - main_start.tv_usec))  / 1.0e6;
  double solver_runtime = ((solve_end.tv_sec - solve_start.tv_sec) * 1000000LL + (solve_end.tv_usec - solve_start.tv_usec)) / 1.0e6;
  sim->WriteErrors(solver_runtime, main_runtime);
#ifndef HYPARB200_NO_ATTACH
  hyparb200_detach();
#endif
  delete sim;
  if (!rank) printf("Finished.\n");
#ifndef serial
  MPI_Comm_free(&world);
  MPI_Finalize();
#endif
  return 0;
}

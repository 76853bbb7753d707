// hpb_internal.h -- internal types shared by the host layer and the CUDA kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/hypar_b200.h"

#define HPB_G 3   // ghost layers required by WENO5 (_MINIMUM_GHOSTS_ 3)

// ---------------------------------------------------------------------------------------------
// Device-side description of one rank's block. All index arithmetic is 64-bit (the reference's
// 32-bit int products overflow beyond ~329^3 points per rank).
// Device layout of a cell array: SoA, a[v*npg + p], p = (i0+g) + P0*((i1+g) + P1*(i2+g)),
// P_d = N_d + 2g -- the reference's ghost-padded index (arrayfunctions.h:40-47) per component.
// Interface array for sweep d: a[v*ni + q], q = i0 + M0*(i1 + M1*i2), M_k = N_k + (k==d).
struct Geom {
  int ndims, nvars, g;
  int N[3];            // local interior size (1 for unused dims)
  int P[3];            // padded size N+2g (1 for unused dims)
  long long st[3];     // stride of dim d in cells
  long long npg;       // points with ghosts
  int xoff[3];         // offset of dim d in the concatenated x/dxinv arrays
  int lo_phys[3], hi_phys[3];   // this block's low / high face of dim d lies on the domain boundary (ip == 0 / ip == iproc-1):
                                // where the compact schemes close their systems (Interp1PrimFifthOrderCRWENO.c:158-160)
};

// the schemes that solve a (block) tridiagonal system per grid line / that keep nonlinear weights (InitializeSolvers.c:262-346)
static inline bool hpb_scheme_is_compact(int s) { return s == HPB_SCHEME_CRWENO5 || s == HPB_SCHEME_CUPW5 || s == HPB_SCHEME_HCWENO5; }
static inline bool hpb_scheme_has_weights(int s) { return s == HPB_SCHEME_WENO5 || s == HPB_SCHEME_CRWENO5 || s == HPB_SCHEME_HCWENO5; }

struct Phys {
  int model, weno, no_limiting, interp_char, upwind, par_scheme, has_grav;
  int scheme;                        // HPB_SCHEME_*
  int muscl_limiter;                 // HPB_LIMITER_* (muscl2)
  double muscl_eps;                  // muscl3
  double hc_rc, hc_xi;               // hcweno5: weno.inp rc, xi
  double eps, gamma, Re, Pr, RT;     // Re already / Minf ; RT = p0/rho0
  double grav[3];
  double adv[15], diff[15];
  const double* advf;                // LinearADR spatially varying advection on the device, [dir*advf_npg + p]; nullptr = constant
  long long advf_npg;
};

struct ZoneDev {
  int type, dim, face, on;
  int is[3], ie[3];
  double wall[3];                 // wall / inflow velocity
  double rho, pressure;           // inflow density, inflow / outflow pressure
  double val[HPB_MAX_NVARS];      // Dirichlet / sponge values
  double xs, xe;                  // sponge: start and end coordinate along dim
};

// explicit RK (TimeExplicitRKInitialize.c) or GLM-GEE (glm = 1; TimeGLMGEEInitialize.c: A = stage coefficients, b = the
// first row of B, b1 = its second row, C (s x 2) and D (2 x 2) weight the solution and the auxiliary solution)
#define HPB_MAX_STAGES 9
struct RKTableau {
  int ns; double A[HPB_MAX_STAGES * HPB_MAX_STAGES], b[HPB_MAX_STAGES], c[HPB_MAX_STAGES];
  int glm, mode; double b1[HPB_MAX_STAGES], C[HPB_MAX_STAGES * 2], D[4], gamma;
};

// ---------------------------------------------------------------------------------------------
struct hpb_solver {
  hpb_config cfg;
  Geom geo;
  Phys phys;
  RKTableau rk;
  int ip[3], is_global[3];
  int neighbor[6];                 // rank of the neighbour across face 2d (low) / 2d+1 (high), -1 none
  int bcperiodic[3];               // mpi->bcperiodic: periodic AND iproc>1
  std::vector<ZoneDev> zones;
  std::vector<double> x_h, dxinv_h, gravf_h, gravg_h;   // host copies (set-up products)
  std::vector<double> advf_h;      // LinearADR varying advection field of this rank, ghost-padded, [dir*npg + p]
  double* d_advf = nullptr;
  bool device_ready = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  long long tma_launches = 0;
  double t = 0.0;

  // device arrays
  double *d_x = nullptr, *d_dxinv = nullptr, *d_gravf = nullptr, *d_gravg = nullptr;
  double *d_u = nullptr;           // solution (SoA, ghosts)
  double *d_U = nullptr;           // stage solution
  double *d_U2 = nullptr;          // second stage-solution array: the last sweep of a stage writes the NEXT stage solution while
                                   // it still reads the current one (sweep_tma.cuh, RKF)
  double *U_pre = nullptr;         // the next stage solution if the sweeps of the current stage have formed it, else nullptr
  int stage_fusion = 1;            // hpb_set_stage_fusion: allow that (default); 0 = every stage vector by k_rk_combine
  double *U_cur = nullptr;         // distributed step: the array holding the current stage solution (d_u for stage 0)
  double *d_Udot[HPB_MAX_STAGES] = {};
  double *d_aux = nullptr, *d_aux2 = nullptr;   // GLM-GEE: the auxiliary solution (TS->U[r]) and its next value
  bool aux_valid = false;          // d_aux holds the auxiliary solution of the current d_u (else: TimeInitialize.c:156-169 at the next step)
  double *d_fI = nullptr;          // interface flux (generic path), max over dirs
  double *d_sI = nullptr;          // interface gravity-source function (generic path)
  double *d_QD[3] = {nullptr, nullptr, nullptr};   // scaled primitive derivatives (viscous)
  double *d_FV = nullptr;          // viscous flux scratch
  double *d_qd4 = nullptr;         // fused path: scaled derivatives of (u,v,w,T), [dir][comp][npg]
  double *d_par = nullptr, *d_src = nullptr;   // exact path: separate accumulators of par and source
  double *d_stage_aos = nullptr;   // AoS staging for host<->device transposes
  double *d_tmp[4] = {nullptr, nullptr, nullptr, nullptr};   // scratch cell arrays for the fine-grained API
  double *d_w = nullptr;           // stored WENO weights of the fine-grained API (all dirs)
  double *d_iface[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // interface scratch (fine-grained API)
  // piecewise path of the schemes other than WENO5 (crweno5, cupw5, upw5): cell scratch (flux / source function,
  // modified solution), tridiagonal rows (sub-diagonal, diagonal, super-diagonal; the right-hand side is solved in
  // place in the output interface array), pivot-error flag
  double *d_cell[2] = {nullptr, nullptr};
  double *d_tri[3] = {nullptr, nullptr, nullptr};
  double *d_bx = nullptr;          // characteristic compact schemes: right-hand side / solution of the block systems
  double *d_mr = nullptr;          // compact schemes across ranks (compact_mr.cu): exchange rows, Jacobi scratch; 16 x systems
  double *d_bmr = nullptr;         // the same for the block systems of the characteristic compact schemes
  double *h_mr = nullptr;          // pinned: norms of the reduced-system iteration
  int *d_err = nullptr;
  // pipelined host-array stepping (hpb_pipe_*): copy streams, AoS staging of the incoming / outgoing field, events
  // [in ready, in free, out ready, out free]
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  double *d_pipe_in = nullptr, *d_pipe_out = nullptr;
  cudaEvent_t ev_pipe[4] = {nullptr, nullptr, nullptr, nullptr};
  double *d_red = nullptr;         // reduction scratch
  // conservation diagnostics (cfg.conservation_check): boundary-flux bookkeeping of HyperbolicFunction.c:103-106 /
  // TimeRK.c:172-193. d_cons = [slot][2*ndims*nvars]: slots 0..3 = BoundaryFlux[stage], 4 = StageBoundaryIntegral of
  // the last host-facing call, 5 = StepBoundaryIntegral
  double *d_cons = nullptr;
  double *d_face = nullptr;        // compact face array of interface fluxes (nvars * largest face)
  double *d_part = nullptr;        // per-block partials of the deterministic reductions
  double *h_red = nullptr;         // pinned
  // halo buffers per field: send/recv per face
  double *d_send[3][6] = {}, *d_recv[3][6] = {};
  size_t face_bytes[6] = {};
  bool halo_nccl_mem = false;      // the halo buffers were re-allocated with ncclMemAlloc (comm.cu)
  // in-library halo exchange (comm.cu): transport, communication stream, one event pair per exchange slot
  // ([0] = send buffers filled, recorded on `stream`; [1] = receive buffers filled, recorded on `s_comm`)
  struct HpbComm* comm = nullptr;
  cudaStream_t s_comm = nullptr;
  cudaEvent_t ev_x[3][2] = {};
  bool u_halo_valid = false;       // the face ghosts of d_u already hold the neighbours' values of the current u
  int overlap = 1;                 // distributed step: 1 = exchanges hidden behind producers / sweeps, 0 = serial
  long long xchg_count = 0, xchg_bytes = 0;   // messages sent / bytes sent by this rank (instrumentation)
  bool w_valid = false;
  // optional per-category device timing (hpb_profile_*): CUDA event pairs on h->stream
  bool prof_on = false;
  struct ProfRec { int cat; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> prof_pool;
  // cached TMA descriptors of the fused sweeps (sweep_fused.cu): one per (array, field count, box kind)
  struct alignas(64) TmaBlob { unsigned char b[128]; };
  struct TmaEntry { const void* ptr; int nf, kind; TmaBlob map; };
  std::vector<TmaEntry> tma_cache;
};

// RAII scope: when profiling is on, brackets the kernels launched inside it with an event pair
struct ProfScope {
  hpb_solver* h; int idx;
  ProfScope(hpb_solver* h_, int cat);
  ~ProfScope();
};

// error plumbing (capi.cu)
int hpb_fail(int code, const char* fmt, ...);
#define HPB_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return hpb_fail(HPB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// host set-up (host_setup.cpp)
int hpb_setup_host(hpb_solver* h);

// kernel launchers (kernels.cu); all asynchronous on h->stream
namespace hpbk {
void aos_to_soa(hpb_solver* h, const double* aos, double* soa, long long npts, int nv);
void soa_to_aos(hpb_solver* h, const double* soa, double* aos, long long npts, int nv);
void apply_bc(hpb_solver* h, double* u);
// sponge zones (BCSponge.c): out -= sigma (u - u_ref) inside every sponge box; true if the configuration has one
bool has_sponge(const hpb_solver* h);
void sponge_source(hpb_solver* h, const double* u, double* out);
// hyperbolic term. negate=true: out = -sum_d dxinv*(fhat_{j+1}-fhat_j) (the first direction overwrites the
// interior, i.e. includes the zeroing of TimeRHSFunctionExplicit.c:89); negate=false: out = +hyp.
// with_source: gravity-source contribution of each gravity direction is ADDED to src (quirk Q5).
void hyperbolic(hpb_solver* h, const double* u, double* out, bool negate, bool with_source, double* src);
void hyperbolic_generic(hpb_solver* h, const double* u, double* out, bool negate, bool with_source, double* src);
// schemes other than WENO5: the reference's own sequence of pieces (HyperbolicFunction.c:167-222) with stored weights
void hyperbolic_pieces(hpb_solver* h, const double* u, double* out, bool negate, bool with_source, double* src);
// the same for a group of ranks (hpbc, below): the reconstructions of a compact scheme couple the ranks of a grid line
int  hyperbolic_pieces_group(hpb_solver** hs, int n, const double* const* u, double* const* out, bool negate, bool with_source,
                             double* const* src);
int  tridiag_error(hpb_solver* h);      // 1 if a tridiagonal solve met a zero pivot since the last call (synchronises)
// fused sweeps (sweep_fused.cu); qd != nullptr: the NavierStokes3D viscous terms are evaluated inside the sweeps
bool fused_available(const hpb_solver* h);
// unext != nullptr: the sweep of the last direction also writes the next RK stage solution ubase + adt * out there (only
// when stage_fusion_available(h); out = the complete right-hand side: nothing may be added to it afterwards; ubase = u^n)
bool hyperbolic_fused(hpb_solver* h, const double* u, double* out, bool negate, bool with_source, double* src,
                      const double* qd, int only_dir = -1, double* unext = nullptr, double adt = 0.0,
                      const double* ubase = nullptr);
bool stage_fusion_available(const hpb_solver* h);
// fused viscous path (viscous_fused.cu)
int qderiv_fused(hpb_solver* h, const double* u, int part = 0);
// exact path: rhs = (rhs + par) + src in the reference's order (TimeRHSFunctionExplicit.c:89-92)
void combine_rhs(hpb_solver* h, double* rhs, const double* par, const double* src);
void parabolic_phase1(hpb_solver* h, const double* u);
void parabolic_phase2(hpb_solver* h, const double* u, double* out, bool accumulate);
void parabolic_nc1(hpb_solver* h, const double* u, double* out, bool accumulate);
void set_zero(hpb_solver* h, double* a, long long n);
void rk_stage(hpb_solver* h, int stage);
void rk_finish(hpb_solver* h);
void copy(hpb_solver* h, double* dst, const double* src, long long n);
void cfl(hpb_solver* h, const double* u, double dt, double* out_host);
void sumsq_diff(hpb_solver* h, const double* a, const double* b, double* out_host);
void step_norm_sumsq(hpb_solver* h, double* out_host);   // from the stage right-hand sides of the last step
// conservation / error diagnostics (deterministic reductions)
void boundary_flux(hpb_solver* h, const double* u, int d, double* sbi);
void step_boundary_integral(hpb_solver* h, const double* bf, double* step_bi);
// GLM-GEE (TimeGLMGEE.c): stage value j into d_U, step completion (d_u and d_aux advance; d_U keeps the previous u
// for the step norm), the auxiliary solution's initial value, the estimated error / error of the estimate
void glm_aux_init(hpb_solver* h);
void glm_stage(hpb_solver* h, int j);
void glm_finish(hpb_solver* h);
void glm_error_fields(hpb_solver* h, const double* uex, double* est, double* dif);
void volume_integral(hpb_solver* h, const double* u, double* out_host);
void diff_norm_sums(hpb_solver* h, const double* a, const double* b, double* out_host);
int diag_partial_size();
// fine-grained API kernels
void flux(hpb_solver* h, const double* u, double* f, int dir);
void modified_solution(hpb_solver* h, const double* u, double* uC);
void weno_weights(hpb_solver* h, const double* fC, const double* u, int dir, double* w);
void weno_interp(hpb_solver* h, double* fI, const double* fC, const double* u, const double* w, int upw, int dir, int uflag);
void upwind(hpb_solver* h, double* fI, const double* fL, const double* fR, const double* uL, const double* uR,
            const double* u, int dir);
void first_derivative(hpb_solver* h, double* Df, const double* f, int dir, int nv);
void second_derivative(hpb_solver* h, double* D2f, const double* f, int dir, int nv, int order);
}

// in-library halo exchange (comm.cu). The primitives act on a GROUP of solvers: one element with the NCCL transport
// (one process per GPU), every rank of the decomposition with the in-process transport (all ranks in one process:
// tests on one GPU, or one process driving several GPUs) -- the step is written once, each primitive loops over the
// group, so that the in-process ranks advance in lock step.
namespace hpbc {
enum { BUF_U = 0, BUF_QD = 1 };                 // what travels: the solution / the Q-derivative arrays of the viscous term
enum { SLOT_U = 0, SLOT_Q0 = 1, SLOT_Q12 = 2 }; // exchange slots: u (all dimensions), Q-derivatives of dimension 0 / of the others
struct RKCoef { const double* k[4]; double a[4]; int n; };
int  slot_bufset(int slot);
int  slot_dimmask(const hpb_solver* h, int slot);
// before the send buffers of `slot` are rewritten (in-process transport: the neighbours' copies out of them must be done)
int  fill_begin(hpb_solver** hs, int n, int slot);
// send buffers <- face layers of array `a` (BUF_U: nvars components x g layers; BUF_QD: see comm.cu)
void pack_faces(hpb_solver* h, int slot, const double* a);
// send buffers of SLOT_U <- u + sum a_s k_s on the face layers (the stage vector / the step completion, TimeRK.c:131-141,
// :182-193, evaluated where it is sent from: the exchange then runs under the full-array update)
void rk_faces(hpb_solver* h, const double* u, const RKCoef& c);
int  xchg_start(hpb_solver** hs, int n, int slot);   // after the fill: exchange on the communication stream
int  xchg_wait(hpb_solver** hs, int n, int slot);    // the compute stream waits for the receive buffers
void unpack_faces(hpb_solver* h, int slot, double* a);
int  comm_free(hpb_solver* h);
bool comm_ready(const hpb_solver* h);
// small messages along one dimension's line of ranks (compact schemes across ranks, compact_mr.cu), on the compute
// streams. active[r] = 0 skips group member r.
int  line_rank(const hpb_solver* h, int dir, int k);            // world rank of the block with ip[dir] = k on this rank's line
// every member sends `send` (count doubles) to ip + step and receives into `recv` from ip - step (no wrap-around; a member
// without the source keeps its recv buffer as it is)
int  line_shift(hpb_solver** hs, int n, int dir, int step, double* const* send, double* const* recv, const long long* counts, const int* active);
// both directions at once: recv_lo <- the low neighbour's send_hi, recv_hi <- the high neighbour's send_lo
int  line_swap(hpb_solver** hs, int n, int dir, double* const* send_lo, double* const* send_hi, double* const* recv_lo,
               double* const* recv_hi, const long long* counts, const int* active);
// one device double per member -> on the host the values of all members of the member's line, ordered by ip[dir]
int  line_gather(hpb_solver** hs, int n, int dir, double* const* d_val, double* const* d_scratch, double (*h_out)[64], const int* active);
}

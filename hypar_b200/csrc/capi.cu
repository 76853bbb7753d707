// capi.cu -- the extern "C" boundary (include/hypar_b200.h): solver life cycle, the host-array
// entry points that mirror HyPar's function-pointer surface, the device-resident time loop and the
// staged multi-GPU step. No CPU fallback: every compute entry point requires a CUDA device.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include "hpb_internal.h"

// ------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static thread_local int g_err_state = 0;   // per host thread, like the message (one thread per GPU must not see another's failure)

int hpb_fail(int code, const char* fmt, ...)
{
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  g_err_state = code;
  fprintf(stderr, "hypar_b200: error %d: %s\n", code, g_err);     // callers may drop return values (basic.h:15-23)
  return code;
}

extern "C" const char* hpb_last_error(void) { return g_err; }
extern "C" int hpb_error_state(void) { return g_err_state; }
extern "C" void hpb_clear_error(void) { g_err_state = 0; g_err[0] = 0; }
extern "C" const char* hpb_version(void) { return "hypar_b200 0.1 (sm_100a)"; }

extern "C" size_t hpb_sizeof_config(void) { return sizeof(hpb_config); }

extern "C" int hpb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" void hpb_config_defaults(hpb_config* c)
{
  memset(c, 0, sizeof(*c));
  c->ndims = 1; c->nvars = 1; c->ghosts = 1;                 // ReadInputs.c:112-115
  for (int d = 0; d < HPB_MAX_NDIMS; d++) { c->iproc[d] = 1; c->dim_global[d] = 1; }
  c->interp_char = 1;                                         // ReadInputs.c:136 "characteristic"
  c->par_scheme = 2;                                          // ReadInputs.c:135
  c->rk_type = HPB_RK_44;
  c->weno_type = HPB_WENO_JS; c->no_limiting = 0; c->weno_eps = 1e-6;   // WENOInitialize.c:51-60
  c->weno_rc = 0.3; c->weno_xi = 0.001;
  c->lu_maxiter = 10; c->lu_evaluate_norm = 1; c->lu_atol = 1e-12; c->lu_rtol = 1e-10;   // tridiagLUInit.c:54-61
  c->upwind = HPB_UPWIND_ROE;
  c->gamma = 1.4; c->Re = -1.0; c->Pr = 0.72; c->Minf = 1.0;   // NavierStokes3DInitialize.c:78-91
  c->rho_ref = 1.0; c->p_ref = 1.0; c->R = 1.0; c->HB = 1; c->N_bv = 0.0;
  c->device = -1;
  c->use_fused = 1;
  c->conservation_check = 0;
  c->hyp_scheme = HPB_SCHEME_WENO5;
  c->muscl_limiter = HPB_LIMITER_GMM; c->muscl_eps = 1e-3;     // MUSCLInitialize.c:26-27
}

// ------------------------------------------------------------------------------------ helpers
static long long ncell(const hpb_solver* h) { return h->geo.npg * h->geo.nvars; }

static long long nif(const hpb_solver* h, int dir)
{
  const Geom& G = h->geo;
  return (long long)(G.N[0] + (dir == 0)) * (G.N[1] + (dir == 1)) * (G.N[2] + (dir == 2));
}
static long long nif_max(const hpb_solver* h)
{
  long long m = 0;
  for (int d = 0; d < h->geo.ndims; d++) { long long k = nif(h, d); if (k > m) m = k; }
  return m;
}

static int dalloc(double** p, long long n)
{
  if (*p) return HPB_OK;
  cudaError_t e = cudaMalloc((void**)p, (size_t)n * sizeof(double));
  if (e != cudaSuccess) return hpb_fail(HPB_ERR_ALLOC, "cudaMalloc of %lld doubles failed: %s", n, cudaGetErrorString(e));
  // the solver's stream is non-blocking: a memset on the legacy stream is NOT ordered before later work on it, so
  // wait for it here (allocation time only) -- otherwise it can land on top of the first data written to the buffer
  e = cudaMemset(*p, 0, (size_t)n * sizeof(double));
  if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
  if (e != cudaSuccess) return hpb_fail(HPB_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
  return HPB_OK;
}
#define TRY(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

static int need_device(hpb_solver* h)
{
  if (!h) return hpb_fail(HPB_ERR_INVALID, "null solver");
  if (!h->device_ready) return hpb_fail(HPB_ERR_NO_DEVICE, "no CUDA device: hypar_b200 has no CPU path");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return hpb_fail(HPB_ERR_CUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e));
  return HPB_OK;
}

static int check_async(hpb_solver* h, const char* what)
{
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return hpb_fail(HPB_ERR_CUDA, "%s: kernel launch failed: %s", what, cudaGetErrorString(e));
  (void)h;
  return HPB_OK;
}

static int sync_check(hpb_solver* h, const char* what)
{
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return hpb_fail(HPB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  if (h->d_err && hpbk::tridiag_error(h))      // tridiagLU.c returns -1 on a zero pivot
    return hpb_fail(HPB_ERR_INVALID, "%s: singular tridiagonal system in a compact-scheme solve", what);
  return HPB_OK;
}

// host AoS (HyPar layout) -> device SoA and back, n points, nv components
static int upload(hpb_solver* h, const double* host_aos, double* dev_soa, long long npts, int nv)
{
  TRY(dalloc(&h->d_stage_aos, (h->geo.npg > nif_max(h) ? h->geo.npg : nif_max(h)) * h->geo.nvars * 3));
  HPB_CUDA(cudaMemcpyAsync(h->d_stage_aos, host_aos, (size_t)npts * nv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  hpbk::aos_to_soa(h, h->d_stage_aos, dev_soa, npts, nv);
  return check_async(h, "upload");
}
static int download(hpb_solver* h, const double* dev_soa, double* host_aos, long long npts, int nv)
{
  TRY(dalloc(&h->d_stage_aos, (h->geo.npg > nif_max(h) ? h->geo.npg : nif_max(h)) * h->geo.nvars * 3));
  hpbk::soa_to_aos(h, dev_soa, h->d_stage_aos, npts, nv);
  HPB_CUDA(cudaMemcpyAsync(host_aos, h->d_stage_aos, (size_t)npts * nv * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return sync_check(h, "download");
}

static bool viscous_on(const hpb_solver* h)
{
  return (h->cfg.model == HPB_MODEL_NS3D || h->cfg.model == HPB_MODEL_NS2D) && h->phys.Re > 0;
}

// production path = fused sweeps; NavierStokes3D viscous terms inside the sweeps
static bool fused_path(const hpb_solver* h) { return hpbk::fused_available(h); }
static bool fused_visc(const hpb_solver* h) { return fused_path(h) && h->cfg.model == HPB_MODEL_NS3D && viscous_on(h); }

// scratch of the piecewise path (schemes other than WENO5) and of the compact schemes' tridiagonal systems
static long long woff(const hpb_solver* h, int dir);
static int ensure_pieces(hpb_solver* h)
{
  if (h->cfg.hyp_scheme == HPB_SCHEME_WENO5) return HPB_OK;
  const long long n = ncell(h);
  TRY(dalloc(&h->d_fI, nif_max(h) * h->geo.nvars));
  for (int k = 0; k < 2; k++) TRY(dalloc(&h->d_cell[k], n));
  for (int k = 0; k < 5; k++) TRY(dalloc(&h->d_iface[k], nif_max(h) * h->geo.nvars));
  TRY(dalloc(&h->d_w, woff(h, h->geo.ndims)));
  if (hpb_scheme_is_compact(h->cfg.hyp_scheme)) {
    // component-wise: one scalar row per interface and component; characteristic: one nvars x nvars block per interface
    const long long blk = (h->phys.interp_char ? (long long)h->geo.nvars * h->geo.nvars : h->geo.nvars);
    for (int k = 0; k < 3; k++) TRY(dalloc(&h->d_tri[k], nif_max(h) * blk));
    if (h->phys.interp_char) TRY(dalloc(&h->d_bx, nif_max(h) * h->geo.nvars));
    if (!h->d_err) {
      HPB_CUDA(cudaMalloc((void**)&h->d_err, sizeof(int)));
      HPB_CUDA(cudaMemset(h->d_err, 0, sizeof(int)));
      HPB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    }
  }
  return HPB_OK;
}

// buffers of the exact (generic) kernels; allocated on first use so that a production run does not carry them
static int ensure_generic(hpb_solver* h)
{
  const long long n = ncell(h);
  TRY(ensure_pieces(h));
  TRY(dalloc(&h->d_fI, nif_max(h) * h->geo.nvars));
  if (h->phys.has_grav) { TRY(dalloc(&h->d_sI, nif_max(h) * 2)); TRY(dalloc(&h->d_src, n)); }
  if (hpbk::has_sponge(h)) TRY(dalloc(&h->d_src, n));
  if (viscous_on(h)) {
    for (int d = 0; d < h->geo.ndims; d++) TRY(dalloc(&h->d_QD[d], n));
    TRY(dalloc(&h->d_FV, h->geo.npg * (h->geo.nvars - 1)));
  }
  if (viscous_on(h) || h->cfg.model == HPB_MODEL_LINEAR_ADR) TRY(dalloc(&h->d_par, n));
  return HPB_OK;
}

// boundary-flux bookkeeping (conservation_check): 6 slots of 2*ndims*nvars sums + the compact face scratch
static constexpr int CONS_SLOT_LAST = HPB_MAX_STAGES, CONS_SLOT_STEP = HPB_MAX_STAGES + 1, CONS_NSLOT = HPB_MAX_STAGES + 2;
static int nbf(const hpb_solver* h) { return 2 * h->geo.ndims * h->geo.nvars; }
static double* cons_slot(hpb_solver* h, int slot) { return h->d_cons + (size_t)slot * nbf(h); }
static int ensure_conservation(hpb_solver* h)
{
  long long nface = 1;
  for (int d = 0; d < h->geo.ndims; d++) { long long k = (long long)h->geo.N[0] * h->geo.N[1] * h->geo.N[2] / h->geo.N[d]; if (k > nface) nface = k; }
  TRY(dalloc(&h->d_face, nface * h->geo.nvars));
  TRY(dalloc(&h->d_cons, (long long)CONS_NSLOT * nbf(h)));
  return HPB_OK;
}
// StageBoundaryIntegral of the state U (all directions, or one) into slot `slot`
static void stage_boundary_flux(hpb_solver* h, const double* U, int slot, int only_dir = -1)
{
  if (!h->cfg.conservation_check) return;
  for (int d = 0; d < h->geo.ndims; d++) if (only_dir < 0 || d == only_dir) hpbk::boundary_flux(h, U, d, cons_slot(h, slot));
}

static bool stage_fusion_on(const hpb_solver* h);
static int alloc_main(hpb_solver* h)
{
  const long long n = ncell(h);
  TRY(dalloc(&h->d_u, n)); TRY(dalloc(&h->d_U, n));
  for (int s = 0; s < h->rk.ns; s++) TRY(dalloc(&h->d_Udot[s], n));
  if (h->rk.glm) { TRY(dalloc(&h->d_aux, n)); TRY(dalloc(&h->d_aux2, n)); }
  if (stage_fusion_on(h) && !h->rk.glm && h->rk.ns > 1) TRY(dalloc(&h->d_U2, n));
  if (fused_visc(h)) TRY(dalloc(&h->d_qd4, 12 * h->geo.npg));
  if (!fused_path(h) || (viscous_on(h) && !fused_visc(h))) TRY(ensure_generic(h));
  TRY(dalloc(&h->d_part, hpbk::diag_partial_size()));
  if (h->cfg.conservation_check) TRY(ensure_conservation(h));
  return HPB_OK;
}

static void prof_release(hpb_solver* h);
// ------------------------------------------------------------------------------------ life cycle
extern "C" int hpb_create(const hpb_config* cfg, hpb_solver** out)
{
  if (!cfg || !out) return hpb_fail(HPB_ERR_INVALID, "hpb_create: null argument");
  *out = nullptr;
  hpb_solver* h = new (std::nothrow) hpb_solver();
  if (!h) return hpb_fail(HPB_ERR_ALLOC, "out of host memory");
  h->cfg = *cfg;
  int rc = hpb_setup_host(h);
  if (rc) { delete h; return rc; }
  { const char* e = getenv("HPB_STAGE_FUSION"); if (e && e[0] == '0') h->stage_fusion = 0; }     // A/B measurements (bench.py)
  // keep a private copy of the global grid (the caller's pointer need not outlive this call)
  h->cfg.x_global = nullptr;

  int ndev = hpb_device_count();
  if (ndev > 0) {
    int dev = cfg->device;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
    if (dev >= ndev) { delete h; return hpb_fail(HPB_ERR_INVALID, "device %d not present (%d devices)", dev, ndev); }
    h->device = dev;
    cudaError_t e = cudaSetDevice(dev);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return hpb_fail(HPB_ERR_CUDA, "device init: %s", cudaGetErrorString(e)); }
    const Geom& G = h->geo;
    const size_t nx = h->x_h.size();
    rc = dalloc(&h->d_x, (long long)nx);           if (rc) { hpb_destroy(h); return rc; }
    rc = dalloc(&h->d_dxinv, (long long)nx);       if (rc) { hpb_destroy(h); return rc; }
    rc = dalloc(&h->d_gravf, G.npg);               if (rc) { hpb_destroy(h); return rc; }
    rc = dalloc(&h->d_gravg, G.npg);               if (rc) { hpb_destroy(h); return rc; }
    rc = dalloc(&h->d_red, 8);                     if (rc) { hpb_destroy(h); return rc; }
    cudaMemcpy(h->d_x, h->x_h.data(), nx * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_dxinv, h->dxinv_h.data(), nx * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_gravf, h->gravf_h.data(), (size_t)G.npg * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_gravg, h->gravg_h.data(), (size_t)G.npg * sizeof(double), cudaMemcpyHostToDevice);
    h->phys.advf = nullptr; h->phys.advf_npg = G.npg;
    if (!h->advf_h.empty()) {
      rc = dalloc(&h->d_advf, (long long)h->advf_h.size());   if (rc) { hpb_destroy(h); return rc; }
      cudaMemcpy(h->d_advf, h->advf_h.data(), h->advf_h.size() * sizeof(double), cudaMemcpyHostToDevice);
      h->phys.advf = h->d_advf;
    }
    cudaStreamSynchronize(cudaStreamLegacy);   // pageable H2D may still be in flight on return; h->stream is non-blocking
    if (cudaMallocHost((void**)&h->h_red, 8 * sizeof(double)) != cudaSuccess) { hpb_destroy(h); return hpb_fail(HPB_ERR_ALLOC, "pinned alloc"); }
    // halo buffers (only for faces that have a neighbour)
    for (int d = 0; d < G.ndims; d++) {
      long long nf = (long long)G.nvars * G.g;
      for (int k = 0; k < G.ndims; k++) if (k != d) nf *= G.N[k];
      h->face_bytes[2*d] = h->face_bytes[2*d+1] = (size_t)nf * sizeof(double);
      const int nfields = viscous_on(h) ? 3 : 1;
      for (int f = 0; f < nfields; f++) for (int s = 0; s < 2; s++) if (h->neighbor[2*d+s] >= 0) {
        rc = dalloc(&h->d_send[f][2*d+s], nf); if (rc) { hpb_destroy(h); return rc; }
        rc = dalloc(&h->d_recv[f][2*d+s], nf); if (rc) { hpb_destroy(h); return rc; }
      }
    }
    rc = alloc_main(h);
    if (rc) { hpb_destroy(h); return rc; }
    h->device_ready = true;
  }
  *out = h;
  return HPB_OK;
}

extern "C" int hpb_destroy(hpb_solver* h)
{
  if (!h) return HPB_OK;
  if (h->stream || h->d_x) cudaSetDevice(h->device);
  hpbc::comm_free(h);          // before the halo buffers go (it frees them itself when NCCL allocated them)
  double** ptrs[] = { &h->d_x, &h->d_dxinv, &h->d_gravf, &h->d_gravg, &h->d_u, &h->d_U, &h->d_fI, &h->d_sI,
                      &h->d_FV, &h->d_stage_aos, &h->d_w, &h->d_red, &h->d_qd4, &h->d_par, &h->d_src,
                      &h->d_cons, &h->d_face, &h->d_part, &h->d_aux, &h->d_aux2, &h->d_U2 };
  for (double** p : ptrs) if (*p) { cudaFree(*p); *p = nullptr; }
  for (int i = 0; i < HPB_MAX_STAGES; i++) if (h->d_Udot[i]) cudaFree(h->d_Udot[i]);
  for (int i = 0; i < 4; i++) if (h->d_tmp[i]) cudaFree(h->d_tmp[i]);
  for (int i = 0; i < 3; i++) if (h->d_QD[i]) cudaFree(h->d_QD[i]);
  for (int i = 0; i < 5; i++) if (h->d_iface[i]) cudaFree(h->d_iface[i]);
  for (int i = 0; i < 2; i++) if (h->d_cell[i]) cudaFree(h->d_cell[i]);
  for (int i = 0; i < 3; i++) if (h->d_tri[i]) cudaFree(h->d_tri[i]);
  if (h->d_err) cudaFree(h->d_err);
  if (h->d_mr) cudaFree(h->d_mr);
  if (h->d_bmr) cudaFree(h->d_bmr);
  if (h->h_mr) cudaFreeHost(h->h_mr);
  if (h->d_bx) cudaFree(h->d_bx);
  if (h->d_advf) cudaFree(h->d_advf);
  if (h->d_pipe_in) cudaFree(h->d_pipe_in);
  if (h->d_pipe_out) cudaFree(h->d_pipe_out);
  for (int k = 0; k < 4; k++) if (h->ev_pipe[k]) cudaEventDestroy(h->ev_pipe[k]);
  if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
  if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
  for (int f = 0; f < 3; f++) for (int k = 0; k < 6; k++) {
    if (h->d_send[f][k]) cudaFree(h->d_send[f][k]);
    if (h->d_recv[f][k]) cudaFree(h->d_recv[f][k]);
  }
  if (h->h_red) cudaFreeHost(h->h_red);
  prof_release(h);
  for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return HPB_OK;
}

// ------------------------------------------------------------------------------------ queries
extern "C" int hpb_get_local_dims(const hpb_solver* h, int* dim_local, int* is_global)
{
  for (int d = 0; d < h->geo.ndims; d++) { if (dim_local) dim_local[d] = h->geo.N[d]; if (is_global) is_global[d] = h->is_global[d]; }
  return HPB_OK;
}
extern "C" long long hpb_npoints_local_wghosts(const hpb_solver* h) { return h->geo.npg; }
extern "C" long long hpb_ninterfaces(const hpb_solver* h, int dir) { return nif(h, dir); }
extern "C" int hpb_get_grid(const hpb_solver* h, double* x, double* dxinv)
{
  if (x) memcpy(x, h->x_h.data(), h->x_h.size() * sizeof(double));
  if (dxinv) memcpy(dxinv, h->dxinv_h.data(), h->dxinv_h.size() * sizeof(double));
  return HPB_OK;
}
extern "C" int hpb_get_neighbors(const hpb_solver* h, int* nb)
{
  for (int k = 0; k < 2 * h->geo.ndims; k++) nb[k] = h->neighbor[k];
  return HPB_OK;
}
extern "C" int hpb_get_zone_extent(const hpb_solver* h, int zone, int* is, int* ie, int* on)
{
  if (zone < 0 || zone >= (int)h->zones.size()) return hpb_fail(HPB_ERR_INVALID, "zone index out of range");
  for (int d = 0; d < h->geo.ndims; d++) { is[d] = h->zones[zone].is[d]; ie[d] = h->zones[zone].ie[d]; }
  *on = h->zones[zone].on;
  return HPB_OK;
}
extern "C" int hpb_get_gravity_field(const hpb_solver* h, double* f, double* g)
{
  if (f) memcpy(f, h->gravf_h.data(), h->gravf_h.size() * sizeof(double));
  if (g) memcpy(g, h->gravg_h.data(), h->gravg_h.size() * sizeof(double));
  return HPB_OK;
}
extern "C" int hpb_get_advection_field(const hpb_solver* h, double* a)
{
  if (h->advf_h.empty() || !a) return hpb_fail(HPB_ERR_INVALID, "no spatially varying advection field");
  const int nd = h->cfg.ndims;
  const size_t npg = h->advf_h.size() / nd;
  for (size_t p = 0; p < npg; p++) for (int d = 0; d < nd; d++) a[p * nd + d] = h->advf_h[(size_t)d * npg + p];
  return HPB_OK;
}
extern "C" long long hpb_kernel_launch_count(const hpb_solver* h) { return h->launches; }
extern "C" long long hpb_tma_launch_count(const hpb_solver* h) { return h->tma_launches; }
extern "C" void* hpb_stream(hpb_solver* h) { return (void*)h->stream; }
extern "C" int hpb_synchronize(hpb_solver* h) { TRY(need_device(h)); return sync_check(h, "synchronize"); }
extern "C" double hpb_current_time(const hpb_solver* h) { return h->t; }
extern "C" int hpb_nstages(const hpb_solver* h) { return h->rk.ns; }
extern "C" int hpb_needs_viscous_exchange(const hpb_solver* h) { return viscous_on(h) ? 1 : 0; }

// ------------------------------------------------------------------------------------ device timing
ProfScope::ProfScope(hpb_solver* h_, int cat) : h(h_), idx(-1)
{
  if (!h->prof_on || h->prof.size() >= 16384) return;
  hpb_solver::ProfRec r; r.cat = cat;
  cudaEvent_t* e[2] = { &r.a, &r.b };
  for (int k = 0; k < 2; k++) {
    if (!h->prof_pool.empty()) { *e[k] = h->prof_pool.back(); h->prof_pool.pop_back(); }
    else if (cudaEventCreate(e[k]) != cudaSuccess) { cudaGetLastError(); return; }
  }
  cudaEventRecord(r.a, h->stream);
  idx = (int)h->prof.size();
  h->prof.push_back(r);
}
ProfScope::~ProfScope() { if (idx >= 0) cudaEventRecord(h->prof[idx].b, h->stream); }

static void prof_release(hpb_solver* h)
{
  for (auto& r : h->prof) { h->prof_pool.push_back(r.a); h->prof_pool.push_back(r.b); }
  h->prof.clear();
}

extern "C" int hpb_profile_enable(hpb_solver* h, int on)
{
  TRY(need_device(h));
  TRY(sync_check(h, "profile_enable"));
  prof_release(h);
  h->prof_on = (on != 0);
  return HPB_OK;
}

extern "C" int hpb_profile_query(hpb_solver* h, int category, double* total_ms, long long* count)
{
  TRY(need_device(h));
  TRY(sync_check(h, "profile_query"));
  double t = 0.0; long long n = 0;
  for (auto& r : h->prof) if (r.cat == category) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { t += ms; n++; } else cudaGetLastError();
  }
  if (total_ms) *total_ms = t;
  if (count) *count = n;
  return HPB_OK;
}

// ------------------------------------------------------------------------------------ FP64 issue peak, measured live
// The production sweeps are bound by the FP64 pipe, not by HBM: this is the denominator of that
// roofline, measured on the device the solver runs on -- 8 independent DMUL chains per thread, 8 CTAs of 256 threads per
// SM, best of 4 launches (DMUL and DADD issue at the same rate; DFMA is ~8 % slower on the B200: profiles/r01_microbench.txt).
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a)
{
  double x0 = threadIdx.x * 1e-9 + 1.0, x1 = x0 + 0.1, x2 = x0 + 0.2, x3 = x0 + 0.3, x4 = x0 + 0.4, x5 = x0 + 0.5, x6 = x0 + 0.6, x7 = x0 + 0.7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) { x0 *= a; x1 *= a; x2 *= a; x3 *= a; x4 *= a; x5 *= a; x6 *= a; x7 *= a; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int hpb_fp64_issue_peak(hpb_solver* h, double* thread_instr_per_s)
{
  TRY(need_device(h));
  if (!thread_instr_per_s) return hpb_fail(HPB_ERR_INVALID, "fp64_issue_peak: null argument");
  TRY(sync_check(h, "fp64_issue_peak"));
  cudaDeviceProp p;
  HPB_CUDA(cudaGetDeviceProperties(&p, h->device));
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 2048;
  double* out = nullptr;
  HPB_CUDA(cudaMalloc((void**)&out, sizeof(double) * blocks * threads));
  cudaEvent_t a, b;
  HPB_CUDA(cudaEventCreate(&a)); HPB_CUDA(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(a, h->stream);
    k_fp64_peak<<<blocks, threads, 0, h->stream>>>(out, iters, 0.999999);
    cudaEventRecord(b, h->stream);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
  HPB_CUDA(cudaGetLastError());
  *thread_instr_per_s = (double)blocks * threads * iters * 64.0 / (best * 1e-3);
  return HPB_OK;
}

// ------------------------------------------------------------------------------------ RHS assembly (device)
// rhs = -hyp + par + source of TimeRHSFunctionExplicit.c:70-92, split at the viscous halo exchange.
// Production path: [Q-derivatives] | sweeps with the viscous flux and the gravity source inside.
// Exact path (use_fused = 0 or a configuration the fused kernels do not cover): generic kernels, par and
// source accumulated separately and combined in the reference's order, so that the result is bit-identical.
// Stage fusion (sweep_tma.cuh, RKF): when row s+1 of the explicit RK tableau has the single entry a_{s+1,s} -- every row of
// RK4, the second of SSPRK3 -- the last sweep of stage s also writes U_{s+1} = u + a dt k_s (TimeRK.c:131-141) with
// k_rk_combine's own two roundings, so the results do not change by a bit. Needs the sweeps to be the last contribution to
// the right-hand side: no sponge (added after the sweeps), viscous terms inside the sweeps or absent.
static bool stage_fusion_on(const hpb_solver* h)
{
  return h->stage_fusion && hpbk::stage_fusion_available(h) && !hpbk::has_sponge(h) && (fused_visc(h) || !viscous_on(h));
}
// where the sweeps of stage s (stage solution U) may put the next stage solution, and its coefficient; nullptr = not fusable
static double* fused_next_stage(hpb_solver* h, int s, const double* U, double* adt)
{
  const RKTableau& T = h->rk;
  if (s < 0 || T.glm || s + 1 >= T.ns || !stage_fusion_on(h)) return nullptr;
  for (int i = 0; i < s; i++) if (T.A[(s + 1) * T.ns + i] != 0.0) return nullptr;
  const double a = T.A[(s + 1) * T.ns + s];
  if (a == 0.0 || !h->d_U2) return nullptr;
  *adt = h->cfg.dt * a;
  return (U == h->d_U) ? h->d_U2 : h->d_U;
}

static int rhs_part_a(hpb_solver* h, const double* U, double* rhs, double* unext = nullptr, double adt = 0.0)
{
  if (fused_path(h)) {
    if (fused_visc(h)) {
      TRY(hpbk::qderiv_fused(h, U));
    } else {
      if (!hpbk::hyperbolic_fused(h, U, rhs, /*negate=*/true, /*with_source=*/true, rhs, nullptr, -1, unext, adt, h->d_u))
        return hpb_fail(HPB_ERR_CUDA, "right-hand side: the fused sweep refused the launch");
      if (viscous_on(h)) hpbk::parabolic_phase1(h, U);          // NavierStokes2D viscous terms: generic kernels
    }
    return HPB_OK;
  }
  TRY(ensure_generic(h));
  if (h->d_src) hpbk::set_zero(h, h->d_src, ncell(h));
  hpbk::hyperbolic_generic(h, U, rhs, /*negate=*/true, /*with_source=*/true, h->d_src);
  if (hpbk::has_sponge(h)) hpbk::sponge_source(h, U, h->d_src);      // SourceFunction.c:52-75: after the model's source
  if (viscous_on(h)) hpbk::parabolic_phase1(h, U);
  return HPB_OK;
}
static int rhs_part_b(hpb_solver* h, const double* U, double* rhs, double* unext = nullptr, double adt = 0.0)
{
  if (fused_path(h)) {
    if (fused_visc(h)) {
      if (!hpbk::hyperbolic_fused(h, U, rhs, true, true, rhs, h->d_qd4, -1, unext, adt, h->d_u))
        return hpb_fail(HPB_ERR_CUDA, "right-hand side: the fused sweep refused the launch");
    }
    else if (viscous_on(h)) hpbk::parabolic_phase2(h, U, rhs, /*accumulate=*/true);
    else if (h->cfg.model == HPB_MODEL_LINEAR_ADR) hpbk::parabolic_nc1(h, U, rhs, true);
    if (hpbk::has_sponge(h)) hpbk::sponge_source(h, U, rhs);        // production path: straight into the right-hand side
    return HPB_OK;
  }
  const double* par = nullptr;
  if (viscous_on(h)) { hpbk::parabolic_phase2(h, U, h->d_par, /*accumulate=*/false); par = h->d_par; }
  else if (h->cfg.model == HPB_MODEL_LINEAR_ADR) { hpbk::parabolic_nc1(h, U, h->d_par, false); par = h->d_par; }
  const bool src_on = h->d_src && (h->phys.has_grav || hpbk::has_sponge(h));
  if (par || src_on) hpbk::combine_rhs(h, rhs, par, src_on ? h->d_src : nullptr);
  return HPB_OK;
}
extern "C" int hpb_stage_overlap_supported(const hpb_solver* h);
static bool multi_rank(const hpb_solver* h)
{
  for (int k = 0; k < 6; k++) if (h->neighbor[k] >= 0) return true;
  return false;
}
#define SINGLE_RANK_ONLY(h, name) do { if (multi_rank(h)) return hpb_fail(HPB_ERR_INVALID, \
  name ": this rank has neighbours; use hpb_TimeStepsDistributed / hpb_RHSFunctionDistributed (include/hypar_b200.h)"); } while (0)

// ------------------------------------------------------------------------------------ HOST entry points
static int tmp(hpb_solver* h, int k) { return dalloc(&h->d_tmp[k], ncell(h)); }

extern "C" int hpb_ApplyBoundaryConditions(hpb_solver* h, double* u, double t)
{
  (void)t;
  TRY(need_device(h)); TRY(tmp(h, 0));
  TRY(upload(h, u, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  hpbk::apply_bc(h, h->d_tmp[0]);
  TRY(check_async(h, "ApplyBoundaryConditions"));
  return download(h, h->d_tmp[0], u, h->geo.npg, h->geo.nvars);
}

extern "C" int hpb_HyperbolicFunction(hpb_solver* h, double* hyp, const double* u, double t, int LimFlag)
{
  (void)t;
  TRY(need_device(h));
  if (!LimFlag && !h->cfg.no_limiting)
    return hpb_fail(HPB_ERR_INVALID, "HyperbolicFunction: LimFlag = 0 (frozen WENO weights) is used by implicit time "
                                     "integration only and is not implemented: the device path never stores weights");
  TRY(tmp(h, 0)); TRY(tmp(h, 1));
  TRY(upload(h, u, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  hpbk::set_zero(h, h->d_tmp[1], ncell(h));
  if (!fused_path(h)) TRY(ensure_generic(h));
  hpbk::hyperbolic(h, h->d_tmp[0], h->d_tmp[1], false, false, nullptr);
  stage_boundary_flux(h, h->d_tmp[0], CONS_SLOT_LAST);
  TRY(check_async(h, "HyperbolicFunction"));
  return download(h, h->d_tmp[1], hyp, h->geo.npg, h->geo.nvars);
}

extern "C" int hpb_ParabolicFunction(hpb_solver* h, double* par, const double* u, double t)
{
  (void)t;
  TRY(need_device(h));
  SINGLE_RANK_ONLY(h, "ParabolicFunction");
  TRY(tmp(h, 0)); TRY(tmp(h, 1));
  TRY(upload(h, u, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  hpbk::set_zero(h, h->d_tmp[1], ncell(h));
  TRY(ensure_generic(h));
  if (viscous_on(h)) { hpbk::parabolic_phase1(h, h->d_tmp[0]); hpbk::parabolic_phase2(h, h->d_tmp[0], h->d_tmp[1], true); }
  else if (h->cfg.model == HPB_MODEL_LINEAR_ADR) hpbk::parabolic_nc1(h, h->d_tmp[0], h->d_tmp[1], true);
  TRY(check_async(h, "ParabolicFunction"));
  return download(h, h->d_tmp[1], par, h->geo.npg, h->geo.nvars);
}

extern "C" int hpb_SourceFunction(hpb_solver* h, double* source, const double* u, double t)
{
  (void)t;
  TRY(need_device(h));
  TRY(tmp(h, 0)); TRY(tmp(h, 1)); TRY(tmp(h, 2));
  TRY(upload(h, u, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  hpbk::set_zero(h, h->d_tmp[1], ncell(h));
  // the source reuses the flux weights of the hyperbolic sweep of the same u (quirk Q5): the sweep is
  // re-evaluated here with the source enabled; its hyperbolic output goes to scratch
  TRY(ensure_generic(h));
  if (h->phys.has_grav) hpbk::hyperbolic_generic(h, h->d_tmp[0], h->d_tmp[2], false, true, h->d_tmp[1]);
  if (hpbk::has_sponge(h)) hpbk::sponge_source(h, h->d_tmp[0], h->d_tmp[1]);
  TRY(check_async(h, "SourceFunction"));
  return download(h, h->d_tmp[1], source, h->geo.npg, h->geo.nvars);
}

extern "C" int hpb_RHSFunction(hpb_solver* h, double* rhs, double* u, double t)
{
  (void)t;
  TRY(need_device(h));
  SINGLE_RANK_ONLY(h, "RHSFunction");
  TRY(upload(h, u, h->d_U, h->geo.npg, h->geo.nvars));
  hpbk::apply_bc(h, h->d_U);
  TRY(rhs_part_a(h, h->d_U, h->d_Udot[0]));
  TRY(rhs_part_b(h, h->d_U, h->d_Udot[0]));
  stage_boundary_flux(h, h->d_U, CONS_SLOT_LAST);
  TRY(check_async(h, "RHSFunction"));
  TRY(download(h, h->d_Udot[0], rhs, h->geo.npg, h->geo.nvars));
  return download(h, h->d_U, u, h->geo.npg, h->geo.nvars);
}

extern "C" int hpb_FFunction(hpb_solver* h, double* f, const double* u, int dir, double t)
{
  (void)t;
  TRY(need_device(h)); TRY(tmp(h, 0)); TRY(tmp(h, 1));
  if (dir < 0 || dir >= h->geo.ndims) return hpb_fail(HPB_ERR_INVALID, "FFunction: dir %d", dir);
  TRY(upload(h, u, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  hpbk::flux(h, h->d_tmp[0], h->d_tmp[1], dir);
  TRY(check_async(h, "FFunction"));
  return download(h, h->d_tmp[1], f, h->geo.npg, h->geo.nvars);
}

extern "C" int hpb_UFunction(hpb_solver* h, double* uC, const double* u, int dir, double t)
{
  (void)t; (void)dir;
  TRY(need_device(h)); TRY(tmp(h, 0)); TRY(tmp(h, 1));
  TRY(upload(h, u, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  hpbk::modified_solution(h, h->d_tmp[0], h->d_tmp[1]);
  TRY(check_async(h, "UFunction"));
  return download(h, h->d_tmp[1], uC, h->geo.npg, h->geo.nvars);
}

static long long woff(const hpb_solver* h, int dir)     // declared above ensure_pieces
{
  long long o = 0;
  for (int d = 0; d < dir; d++) o += 12 * nif(h, d) * h->geo.nvars;
  return o;
}

extern "C" int hpb_SetInterpLimiterVar(hpb_solver* h, const double* fC, const double* u, int dir)
{
  TRY(need_device(h)); TRY(tmp(h, 0)); TRY(tmp(h, 1));
  if (dir < 0 || dir >= h->geo.ndims) return hpb_fail(HPB_ERR_INVALID, "SetInterpLimiterVar: dir %d", dir);
  if (!hpb_scheme_has_weights(h->cfg.hyp_scheme))
    return HPB_OK;             // linear schemes: the reference leaves the pointer NULL (InitializeSolvers.c:194)
  TRY(dalloc(&h->d_w, woff(h, h->geo.ndims)));
  TRY(upload(h, fC, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  TRY(upload(h, u, h->d_tmp[1], h->geo.npg, h->geo.nvars));
  hpbk::weno_weights(h, h->d_tmp[0], h->d_tmp[1], dir, h->d_w + woff(h, dir));
  h->w_valid = true;
  return check_async(h, "SetInterpLimiterVar");
}

extern "C" int hpb_GetInterpWeights(hpb_solver* h, int dir, double* w)
{
  TRY(need_device(h));
  if (!h->d_w) return hpb_fail(HPB_ERR_INVALID, "GetInterpWeights: SetInterpLimiterVar has not been called");
  // device: [(3*blk+k)][v][q]  ->  host: [(3*blk+k)][q][v] (the reference's per-block AoS)
  const long long ni = nif(h, dir);
  for (int b = 0; b < 12; b++)
    TRY(download(h, h->d_w + woff(h, dir) + (long long)b * ni * h->geo.nvars, w + (long long)b * ni * h->geo.nvars, ni, h->geo.nvars));
  return HPB_OK;
}

extern "C" int hpb_InterpolateInterfacesHyp(hpb_solver* h, double* fI, const double* fC, const double* u,
                                            int upw, int dir, int uflag)
{
  TRY(need_device(h)); TRY(tmp(h, 0)); TRY(tmp(h, 1));
  if (dir < 0 || dir >= h->geo.ndims) return hpb_fail(HPB_ERR_INVALID, "InterpolateInterfacesHyp: dir %d", dir);
  TRY(dalloc(&h->d_w, woff(h, h->geo.ndims)));
  if (!h->w_valid) {
    // WENOInitialize.c:170-178: weights start at their optimal values (WENOFifthOrderInitializeWeights.c:33-140:
    // CRWENO5 (0.2,0.5,0.3) except on the two physical-boundary interfaces of every line)
    std::vector<double> w((size_t)woff(h, h->geo.ndims));
    for (int d = 0; d < h->geo.ndims; d++) {
      const long long nq = nif(h, d), n = nq * h->geo.nvars;
      const Geom& G = h->geo;
      const long long M0 = G.N[0] + (d == 0), M1 = G.N[1] + (d == 1);
      for (int b = 0; b < 4; b++) for (long long i = 0; i < n; i++) {
        const long long q = i % nq;
        const long long iI = (d == 0 ? q % M0 : d == 1 ? (q / M0) % M1 : q / (M0 * M1));
        const bool bnd = (iI == 0 && G.lo_phys[d]) || (iI == G.N[d] && G.hi_phys[d]);
        const bool cr = (h->cfg.hyp_scheme == HPB_SCHEME_CRWENO5) && !bnd;
        w[(size_t)(woff(h, d) + (3*b+0)*n + i)] = cr ? 0.2 : 0.1; w[(size_t)(woff(h, d) + (3*b+1)*n + i)] = cr ? 0.5 : 0.6;
        w[(size_t)(woff(h, d) + (3*b+2)*n + i)] = 0.3;
      }
    }
    HPB_CUDA(cudaMemcpy(h->d_w, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
    HPB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    h->w_valid = true;
  }
  TRY(dalloc(&h->d_iface[0], nif_max(h) * h->geo.nvars));
  TRY(ensure_pieces(h));
  TRY(upload(h, fC, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  TRY(upload(h, u, h->d_tmp[1], h->geo.npg, h->geo.nvars));
  hpbk::weno_interp(h, h->d_iface[0], h->d_tmp[0], h->d_tmp[1], h->d_w + woff(h, dir), upw, dir, uflag);
  TRY(check_async(h, "InterpolateInterfacesHyp"));
  return download(h, h->d_iface[0], fI, nif(h, dir), h->geo.nvars);
}

extern "C" int hpb_Upwind(hpb_solver* h, double* fI, const double* fL, const double* fR, const double* uL,
                          const double* uR, const double* u, int dir, double t)
{
  (void)t;
  TRY(need_device(h)); TRY(tmp(h, 0));
  if (dir < 0 || dir >= h->geo.ndims) return hpb_fail(HPB_ERR_INVALID, "Upwind: dir %d", dir);
  for (int k = 0; k < 5; k++) TRY(dalloc(&h->d_iface[k], nif_max(h) * h->geo.nvars));
  const long long ni = nif(h, dir);
  TRY(upload(h, fL, h->d_iface[1], ni, h->geo.nvars));
  TRY(upload(h, fR, h->d_iface[2], ni, h->geo.nvars));
  TRY(upload(h, uL, h->d_iface[3], ni, h->geo.nvars));
  TRY(upload(h, uR, h->d_iface[4], ni, h->geo.nvars));
  TRY(upload(h, u, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  hpbk::upwind(h, h->d_iface[0], h->d_iface[1], h->d_iface[2], h->d_iface[3], h->d_iface[4], h->d_tmp[0], dir);
  TRY(check_async(h, "Upwind"));
  return download(h, h->d_iface[0], fI, ni, h->geo.nvars);
}

extern "C" int hpb_FirstDerivativePar(hpb_solver* h, double* Df, const double* f, int dir, int bias)
{
  (void)bias;   // the fourth-order central operator ignores the bias, as in the reference
  TRY(need_device(h)); TRY(tmp(h, 0)); TRY(tmp(h, 1));
  if (dir < 0 || dir >= h->geo.ndims) return hpb_fail(HPB_ERR_INVALID, "FirstDerivativePar: dir %d", dir);
  TRY(upload(h, f, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  TRY(upload(h, Df, h->d_tmp[1], h->geo.npg, h->geo.nvars));      // entries outside the computed lines are kept
  hpbk::first_derivative(h, h->d_tmp[1], h->d_tmp[0], dir, h->geo.nvars);
  TRY(check_async(h, "FirstDerivativePar"));
  return download(h, h->d_tmp[1], Df, h->geo.npg, h->geo.nvars);
}

extern "C" int hpb_SecondDerivativePar(hpb_solver* h, double* D2f, const double* f, int dir)
{
  TRY(need_device(h)); TRY(tmp(h, 0)); TRY(tmp(h, 1));
  if (dir < 0 || dir >= h->geo.ndims) return hpb_fail(HPB_ERR_INVALID, "SecondDerivativePar: dir %d", dir);
  TRY(upload(h, f, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  TRY(upload(h, D2f, h->d_tmp[1], h->geo.npg, h->geo.nvars));
  hpbk::second_derivative(h, h->d_tmp[1], h->d_tmp[0], dir, h->geo.nvars, h->cfg.par_scheme == 4 ? 4 : 2);
  TRY(check_async(h, "SecondDerivativePar"));
  return download(h, h->d_tmp[1], D2f, h->geo.npg, h->geo.nvars);
}

extern "C" int hpb_ComputeCFL(hpb_solver* h, const double* u, double dt, double t, double* cfl)
{
  (void)t;
  TRY(need_device(h)); TRY(tmp(h, 0));
  TRY(upload(h, u, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  hpbk::cfl(h, h->d_tmp[0], dt, cfl);
  return check_async(h, "ComputeCFL");
}

// ------------------------------------------------------------------------------------ device-resident path
extern "C" int hpb_dev_set_solution(hpb_solver* h, const double* u_host)
{
  TRY(need_device(h));
  TRY(upload(h, u_host, h->d_u, h->geo.npg, h->geo.nvars));
  h->u_halo_valid = false;
  h->aux_valid = false;           // GLM-GEE: the auxiliary solution restarts with the new solution (TimeInitialize.c:156-169)
  return sync_check(h, "dev_set_solution");
}

extern "C" int hpb_dev_get_solution(hpb_solver* h, double* u_host)
{
  TRY(need_device(h));
  return download(h, h->d_u, u_host, h->geo.npg, h->geo.nvars);
}

extern "C" int hpb_dev_fill_solution_from_global(hpb_solver* h, const double* ug)
{
  // host-side gather of this rank's block out of the global AoS array (initial.inp layout), ghosts = 0
  TRY(need_device(h));
  const Geom& G = h->geo;
  const hpb_config& c = h->cfg;
  std::vector<double> loc((size_t)ncell(h), 0.0);
  const long long g0 = c.dim_global[0], g1 = (G.ndims > 1 ? c.dim_global[1] : 1);
  for (int k = 0; k < G.N[2]; k++) for (int j = 0; j < G.N[1]; j++) {
    const long long pg = (h->is_global[0]) + g0 * ((j + (G.ndims > 1 ? h->is_global[1] : 0)) + g1 * (k + (G.ndims > 2 ? h->is_global[2] : 0)));
    long long pl = G.g;
    if (G.ndims > 1) pl += (long long)G.P[0] * (j + G.g);
    if (G.ndims > 2) pl += (long long)G.P[0] * G.P[1] * (k + G.g);
    memcpy(&loc[(size_t)pl * G.nvars], ug + pg * G.nvars, (size_t)G.N[0] * G.nvars * sizeof(double));
  }
  return hpb_dev_set_solution(h, loc.data());
}

// stage solution U_s = u + dt sum_{i<s} a_si Udot_i (TimeRK.c:131-141). For s = 0 it equals u, whose ghosts
// TimePreStep has just filled: the device loop uses u itself instead of a copy (the boundary conditions
// re-applied by the RHS are idempotent), saving one pass over memory per step.
static double* stage_U(hpb_solver* h, int s)
{
  if (s == 0) { h->U_pre = nullptr; return h->d_u; }
  if (h->U_pre) { double* p = h->U_pre; h->U_pre = nullptr; return p; }     // formed by the last sweep of stage s-1
  hpbk::rk_stage(h, s);
  return h->d_U;
}

// TimeGLMGEE.c:45-155 on one rank. Every stage value is a combination of the solution, the auxiliary solution and the
// earlier stage right-hand sides (stage 0 included: C[0] need not be (1, 0)).
static int step_single_glm(hpb_solver* h)
{
  hpbk::apply_bc(h, h->d_u);                     // TimePreStep.c:50-76
  if (!h->aux_valid) hpbk::glm_aux_init(h);
  for (int j = 0; j < h->rk.ns; j++) {
    hpbk::glm_stage(h, j);                       // TimeGLMGEE.c:66-80
    hpbk::apply_bc(h, h->d_U);                   // TimeRHSFunctionExplicit.c:46
    TRY(rhs_part_a(h, h->d_U, h->d_Udot[j]));
    TRY(rhs_part_b(h, h->d_U, h->d_Udot[j]));
    stage_boundary_flux(h, h->d_U, j);           // :107-112
  }
  hpbk::glm_finish(h);                           // :116-151
  if (h->cfg.conservation_check) hpbk::step_boundary_integral(h, cons_slot(h, 0), cons_slot(h, CONS_SLOT_STEP));   // :131-140: dt B[0][i]
  h->t += h->cfg.dt;
  return check_async(h, "TimeStep (glm-gee)");
}

static int step_single(hpb_solver* h)
{
  if (h->rk.glm) return step_single_glm(h);
  // TimePreStep.c:50-76: boundary conditions on u (the step norm of TimePostStep is formed from the stage
  // right-hand sides afterwards: no copy of u is kept)
  hpbk::apply_bc(h, h->d_u);
  for (int s = 0; s < h->rk.ns; s++) {
    double* U = stage_U(h, s);                  // TimeRK.c:131-141
    if (s > 0) hpbk::apply_bc(h, U);            // TimeRHSFunctionExplicit.c:46 (stage 0 is u itself: just done)
    double adt = 0.0;
    double* unext = fused_next_stage(h, s, U, &adt);
    TRY(rhs_part_a(h, U, h->d_Udot[s], unext, adt));
    TRY(rhs_part_b(h, U, h->d_Udot[s], unext, adt));
    h->U_pre = unext;
    stage_boundary_flux(h, U, s);               // TimeRK.c:172-177 (BoundaryFlux[s] = StageBoundaryIntegral)
  }
  hpbk::rk_finish(h);                           // TimeRK.c:182-193
  if (h->cfg.conservation_check) hpbk::step_boundary_integral(h, cons_slot(h, 0), cons_slot(h, CONS_SLOT_STEP));
  h->t += h->cfg.dt;                            // TimePostStep.c:36
  return check_async(h, "TimeStep");
}

extern "C" int hpb_TimeStep(hpb_solver* h)
{
  TRY(need_device(h));
  SINGLE_RANK_ONLY(h, "TimeStep");
  return step_single(h);
}

extern "C" int hpb_TimeSteps(hpb_solver* h, int nsteps)
{
  TRY(need_device(h));
  SINGLE_RANK_ONLY(h, "TimeSteps");
  for (int i = 0; i < nsteps; i++) TRY(step_single(h));
  return sync_check(h, "TimeSteps");
}

extern "C" int hpb_TimeIntegrate(hpb_solver* h, double* u, int nsteps, double t0)
{
  TRY(need_device(h));
  SINGLE_RANK_ONLY(h, "TimeIntegrate");
  h->t = t0;
  TRY(upload(h, u, h->d_u, h->geo.npg, h->geo.nvars));
  h->aux_valid = false;
  for (int i = 0; i < nsteps; i++) TRY(step_single(h));
  return download(h, h->d_u, u, h->geo.npg, h->geo.nvars);
}

// ------------------------------------------------------------------------------------ pipelined host-array stepping
static int ensure_pipe(hpb_solver* h)
{
  if (h->s_h2d) return HPB_OK;
  TRY(dalloc(&h->d_pipe_in, ncell(h)));
  TRY(dalloc(&h->d_pipe_out, ncell(h)));
  HPB_CUDA(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
  for (int k = 0; k < 4; k++) HPB_CUDA(cudaEventCreateWithFlags(&h->ev_pipe[k], cudaEventDisableTiming));
  HPB_CUDA(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
  return HPB_OK;
}
enum { EV_IN_READY = 0, EV_IN_FREE = 1, EV_OUT_READY = 2, EV_OUT_FREE = 3 };

extern "C" int hpb_pipe_upload(hpb_solver* h, const double* u_in, double t0)
{
  TRY(need_device(h));
  if (!u_in) return hpb_fail(HPB_ERR_INVALID, "pipe_upload: null input");
  TRY(ensure_pipe(h));
  const size_t bytes = (size_t)ncell(h) * sizeof(double);
  // the staging array is free once the previous field has been transposed out of it
  HPB_CUDA(cudaStreamWaitEvent(h->s_h2d, h->ev_pipe[EV_IN_FREE], 0));
  HPB_CUDA(cudaMemcpyAsync(h->d_pipe_in, u_in, bytes, cudaMemcpyHostToDevice, h->s_h2d));
  HPB_CUDA(cudaEventRecord(h->ev_pipe[EV_IN_READY], h->s_h2d));
  // on the solver's stream (i.e. after the previous field's steps and its transposition to the outgoing array)
  HPB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_pipe[EV_IN_READY], 0));
  hpbk::aos_to_soa(h, h->d_pipe_in, h->d_u, h->geo.npg, h->geo.nvars);
  h->u_halo_valid = false;
  h->aux_valid = false;
  HPB_CUDA(cudaEventRecord(h->ev_pipe[EV_IN_FREE], h->stream));
  h->t = t0;
  return check_async(h, "pipe_upload");
}

extern "C" int hpb_pipe_download(hpb_solver* h, double* u_out)
{
  TRY(need_device(h));
  if (!u_out) return hpb_fail(HPB_ERR_INVALID, "pipe_download: null output");
  TRY(ensure_pipe(h));
  const size_t bytes = (size_t)ncell(h) * sizeof(double);
  HPB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_pipe[EV_OUT_FREE], 0));    // the previous field has left the array
  hpbk::soa_to_aos(h, h->d_u, h->d_pipe_out, h->geo.npg, h->geo.nvars);
  HPB_CUDA(cudaEventRecord(h->ev_pipe[EV_OUT_READY], h->stream));
  HPB_CUDA(cudaStreamWaitEvent(h->s_d2h, h->ev_pipe[EV_OUT_READY], 0));
  HPB_CUDA(cudaMemcpyAsync(u_out, h->d_pipe_out, bytes, cudaMemcpyDeviceToHost, h->s_d2h));
  HPB_CUDA(cudaEventRecord(h->ev_pipe[EV_OUT_FREE], h->s_d2h));
  return check_async(h, "pipe_download");
}

extern "C" int hpb_pipe_join(hpb_solver* h)
{
  TRY(need_device(h));
  if (!h->s_h2d) return HPB_OK;
  // work enqueued on the solver's stream from here on runs after every copy enqueued so far
  HPB_CUDA(cudaEventRecord(h->ev_pipe[EV_IN_READY], h->s_h2d));
  HPB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_pipe[EV_IN_READY], 0));
  HPB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_pipe[EV_OUT_FREE], 0));
  return HPB_OK;
}

extern "C" int hpb_pipe_wait(hpb_solver* h)
{
  TRY(need_device(h));
  if (h->s_h2d) {
    HPB_CUDA(cudaStreamSynchronize(h->s_h2d));
    TRY(sync_check(h, "pipe_wait"));
    HPB_CUDA(cudaStreamSynchronize(h->s_d2h));
    return HPB_OK;
  }
  return sync_check(h, "pipe_wait");
}

extern "C" int hpb_TimeIntegrateAsync(hpb_solver* h, const double* u_in, double* u_out, int nsteps, double t0)
{
  TRY(need_device(h));
  SINGLE_RANK_ONLY(h, "TimeIntegrateAsync");
  TRY(hpb_pipe_upload(h, u_in, t0));
  for (int i = 0; i < nsteps; i++) TRY(step_single(h));
  return hpb_pipe_download(h, u_out);
}

extern "C" int hpb_dev_ComputeCFL(hpb_solver* h, double* cfl_local_max)
{
  TRY(need_device(h));
  hpbk::cfl(h, h->d_u, h->cfg.dt, cfl_local_max);
  return check_async(h, "dev_ComputeCFL");
}

extern "C" int hpb_dev_StepNormSumSq(hpb_solver* h, double* sumsq_local)
{
  TRY(need_device(h));
  if (h->rk.glm) hpbk::sumsq_diff(h, h->d_u, h->d_U, sumsq_local);      // glm_finish left the previous solution in d_U
  else hpbk::step_norm_sumsq(h, sumsq_local);
  return check_async(h, "dev_StepNormSumSq");
}

// ------------------------------------------------------------------------------------ conservation / error diagnostics
extern "C" int hpb_dev_VolumeIntegral(hpb_solver* h, double* vol_local)
{
  TRY(need_device(h));
  if (!vol_local) return hpb_fail(HPB_ERR_INVALID, "dev_VolumeIntegral: null output");
  hpbk::volume_integral(h, h->d_u, vol_local);
  return check_async(h, "dev_VolumeIntegral");
}

static int cons_download(hpb_solver* h, int slot, double* out, const char* what)
{
  TRY(need_device(h));
  if (!h->cfg.conservation_check || !h->d_cons)
    return hpb_fail(HPB_ERR_INVALID, "%s: the boundary-flux bookkeeping is off (conservation_check = 0)", what);
  if (!out) return hpb_fail(HPB_ERR_INVALID, "%s: null output", what);
  HPB_CUDA(cudaMemcpyAsync(out, cons_slot(h, slot), nbf(h) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return sync_check(h, what);
}
extern "C" int hpb_dev_StageBoundaryIntegral(hpb_solver* h, int slot, double* sbi)
{
  if (h && (slot < -1 || slot >= h->rk.ns)) return hpb_fail(HPB_ERR_INVALID, "dev_StageBoundaryIntegral: slot %d", slot);
  return cons_download(h, slot < 0 ? CONS_SLOT_LAST : slot, sbi, "dev_StageBoundaryIntegral");
}
extern "C" int hpb_dev_StepBoundaryIntegral(hpb_solver* h, double* step_bi)
{
  return cons_download(h, CONS_SLOT_STEP, step_bi, "dev_StepBoundaryIntegral");
}

// BoundaryIntegral.c:36-48, this rank's part: sum_d (StepBI[2d] + StepBI[2d+1]) dS_d with dS_d the product of the
// other dimensions' spacings at the block's mid-point index (the reference's uniform-grid assumption). Host code.
extern "C" int hpb_BoundaryIntegral(const hpb_solver* h, const double* step_bi, double* bi_local)
{
  if (!h || !step_bi || !bi_local) return hpb_fail(HPB_ERR_INVALID, "BoundaryIntegral: null argument");
  const Geom& G = h->geo;
  for (int v = 0; v < G.nvars; v++) bi_local[v] = 0.0;
  for (int d = 0; d < G.ndims; d++) for (int v = 0; v < G.nvars; v++) {
    double dS = 1.0;
    for (int k = 0; k < G.ndims; k++) if (k != d) dS *= (1.0 / h->dxinv_h[G.xoff[k] + G.g + G.N[k] / 2]);
    bi_local[v] += step_bi[(2 * d + 0) * G.nvars + v] * dS;
    bi_local[v] += step_bi[(2 * d + 1) * G.nvars + v] * dS;
  }
  return HPB_OK;
}

// CalculateConservationError.c:24-36
extern "C" int hpb_CalculateConservationError(int nvars, const double* vol, const double* vol0, const double* tbi, double* err)
{
  if (!vol || !vol0 || !tbi || !err || nvars < 1) return hpb_fail(HPB_ERR_INVALID, "CalculateConservationError: bad argument");
  for (int v = 0; v < nvars; v++) {
    const double base = (fabs(vol0[v]) > 1.0) ? fabs(vol0[v]) : 1.0;
    const double e = (vol[v] + tbi[v] - vol0[v]) * (vol[v] + tbi[v] - vol0[v]);
    err[v] = sqrt(e) / base;
  }
  return HPB_OK;
}

extern "C" int hpb_dev_ErrorSums(hpb_solver* h, const double* uex_host, double* sums)
{
  TRY(need_device(h));
  if (!uex_host || !sums) return hpb_fail(HPB_ERR_INVALID, "dev_ErrorSums: null argument");
  TRY(tmp(h, 0));
  TRY(upload(h, uex_host, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  hpbk::diff_norm_sums(h, h->d_tmp[0], nullptr, sums);          // CalculateError.c:68-84 (solution norms)
  hpbk::diff_norm_sums(h, h->d_tmp[0], h->d_u, sums + 3);      // :87-104 (uex - u)
  return check_async(h, "dev_ErrorSums");
}

// ---- GLM-GEE: the auxiliary solution and the error estimate (TimeGetAuxSolutions.c, TimeError.c:43-127)
extern "C" double hpb_glmgee_gamma(const hpb_solver* h) { return h->rk.glm ? h->rk.gamma : 0.0; }
extern "C" int hpb_dev_get_aux_solution(hpb_solver* h, double* uaux_host)
{
  TRY(need_device(h));
  if (!h->rk.glm) return hpb_fail(HPB_ERR_INVALID, "dev_get_aux_solution: the time integrator keeps no auxiliary solution (glm-gee only)");
  if (!h->aux_valid) hpbk::glm_aux_init(h);
  return download(h, h->d_aux, uaux_host, h->geo.npg, h->geo.nvars);
}
extern "C" int hpb_dev_set_aux_solution(hpb_solver* h, const double* uaux_host)
{
  TRY(need_device(h));
  if (!h->rk.glm) return hpb_fail(HPB_ERR_INVALID, "dev_set_aux_solution: the time integrator keeps no auxiliary solution (glm-gee only)");
  TRY(upload(h, uaux_host, h->d_aux, h->geo.npg, h->geo.nvars));
  h->aux_valid = true;
  return sync_check(h, "dev_set_aux_solution");
}
extern "C" int hpb_dev_GLMGEEErrorSums(hpb_solver* h, const double* uex_host, double* sums)
{
  TRY(need_device(h));
  if (!h->rk.glm) return hpb_fail(HPB_ERR_INVALID, "dev_GLMGEEErrorSums: time_scheme is not glm-gee");
  if (!sums) return hpb_fail(HPB_ERR_INVALID, "dev_GLMGEEErrorSums: null argument");
  for (int k = 0; k < 3; k++) TRY(tmp(h, k));
  if (!h->aux_valid) hpbk::glm_aux_init(h);
  if (uex_host) TRY(upload(h, uex_host, h->d_tmp[0], h->geo.npg, h->geo.nvars));
  hpbk::glm_error_fields(h, uex_host ? h->d_tmp[0] : nullptr, h->d_tmp[1], h->d_tmp[2]);
  hpbk::diff_norm_sums(h, h->d_u, nullptr, sums);              // TimeError.c:55-70
  hpbk::diff_norm_sums(h, h->d_tmp[1], nullptr, sums + 3);    // :73-87
  hpbk::diff_norm_sums(h, h->d_tmp[2], nullptr, sums + 6);    // :89-106
  return check_async(h, "dev_GLMGEEErrorSums");
}

extern "C" int hpb_dev_RHS(hpb_solver* h, double t, double* rhs_host)
{
  (void)t;
  TRY(need_device(h));
  SINGLE_RANK_ONLY(h, "dev_RHS");
  hpbk::copy(h, h->d_U, h->d_u, ncell(h));
  hpbk::apply_bc(h, h->d_U);
  TRY(rhs_part_a(h, h->d_U, h->d_Udot[0]));
  TRY(rhs_part_b(h, h->d_U, h->d_Udot[0]));
  stage_boundary_flux(h, h->d_U, CONS_SLOT_LAST);
  TRY(check_async(h, "dev_RHS"));
  if (rhs_host) return download(h, h->d_Udot[0], rhs_host, h->geo.npg, h->geo.nvars);
  return sync_check(h, "dev_RHS");
}

// ------------------------------------------------------------------------------------ distributed step
// One time step of a domain-decomposed run, entirely inside the library: TimePreStep.c:50-76 (boundary conditions +
// halo of u), TimeRK.c:126-195 with TimeRHSFunctionExplicit.c:46-92 per stage (boundary conditions + halo of the stage
// solution, NavierStokes3DParabolicFunction.c:125-130: halo of QDerivX / QDerivY), and the step completion. The transport
// is comm.cu's (NCCL send/recv on a communication stream, or device-to-device copies between in-process ranks).
//
// The step is written once for a GROUP of solvers: one element with NCCL (one process per GPU), every rank of the
// decomposition with the in-process transport, which therefore advances in lock step.
//
// Overlapped schedule (h->overlap, the default), per stage s >= 1:
//     k_rk_faces      U_s on the face layers, straight into the send buffers
//     exchange of u   on the communication stream           ||  k_rk_combine: U_s on the whole block, physical BCs
//     unpack (one launch)
//     k_qderiv_*      (viscous) ; pack (one launch per slot)
//     exchange of the Q-derivatives, dimension 0, then dimensions 1.. as a second group
//     unpack dim 0 ; sweep x                                 ||  exchange of dimensions 1..
//     unpack dims 1.. ; sweep y ; sweep z
// and at the end of the step the same for u itself (faces of u + dt sum b_s k_s first, exchange under the full update):
// that exchange IS the TimePreStep exchange of the next step, so stage 0 (whose stage solution is u) starts at once.
// The exchange of the stage-0 solution the reference repeats (TimeRHSFunctionExplicit.c:60 right after TimePreStep.c:57
// on the same values) is not issued: 4 exchanges of u per RK4 step instead of 5. The serial schedule (h->overlap = 0)
// runs pack - exchange - unpack where the reference has its MPIExchangeBoundariesnD calls; results are bit-identical
// (tests/test_gpu_decomposed.py).
struct Grp { hpb_solver** hs; int n; };
#define EACH(h) for (int r_ = 0; r_ < G.n; r_++) if (hpb_solver* h = G.hs[r_])

static hpbc::RKCoef stage_coef(const hpb_solver* h, int stage)
{
  hpbc::RKCoef a; a.n = 0;
  for (int i = 0; i < stage; i++) {                       // the coefficients of hpbk::rk_stage
    const double c = h->cfg.dt * h->rk.A[stage * h->rk.ns + i];
    if (c == 0.0) continue;
    a.k[a.n] = h->d_Udot[i]; a.a[a.n] = c; a.n++;
  }
  return a;
}
static hpbc::RKCoef finish_coef(const hpb_solver* h)
{
  hpbc::RKCoef a; a.n = h->rk.ns;                         // hpbk::rk_finish
  for (int s = 0; s < h->rk.ns; s++) { a.k[s] = h->d_Udot[s]; a.a[s] = h->cfg.dt * h->rk.b[s]; }
  return a;
}

// serial exchange of a solution array: pack - exchange - wait - unpack
static int exchange_u_serial(Grp& G, bool stage_solution)
{
  TRY(hpbc::fill_begin(G.hs, G.n, hpbc::SLOT_U));
  EACH(h) { TRY(need_device(h)); hpbc::pack_faces(h, hpbc::SLOT_U, stage_solution ? h->U_cur : h->d_u); }
  TRY(hpbc::xchg_start(G.hs, G.n, hpbc::SLOT_U));
  TRY(hpbc::xchg_wait(G.hs, G.n, hpbc::SLOT_U));
  EACH(h) { TRY(need_device(h)); hpbc::unpack_faces(h, hpbc::SLOT_U, stage_solution ? h->U_cur : h->d_u); }
  return HPB_OK;
}

static int dist_prestep(Grp& G)
{
  bool valid = true;
  EACH(h) { TRY(need_device(h)); hpbk::apply_bc(h, h->d_u); valid = valid && h->u_halo_valid; }     // TimePreStep.c:50-56
  EACH(h) if (h->rk.glm && !h->aux_valid) hpbk::glm_aux_init(h);
  if (!valid) {
    TRY(exchange_u_serial(G, false));                                                                 // TimePreStep.c:57-76
    EACH(h) h->u_halo_valid = true;
  }
  return HPB_OK;
}

static int dist_stage(Grp& G, int s, bool fuse = true)
{
  const bool overlap = G.hs[0]->overlap != 0;
  // ---- stage solution with boundary conditions and halo (TimeRK.c:131-141, TimeRHSFunctionExplicit.c:46-60)
  if (G.hs[0]->rk.glm) {
    // GLM-GEE: every stage value combines the solution and the auxiliary solution (TimeGLMGEE.c:66-80); formed on the
    // whole block, then boundary conditions and the halo where the reference has them (serial exchange of the stage value)
    EACH(h) { TRY(need_device(h)); hpbk::glm_stage(h, s); h->U_cur = h->d_U; hpbk::apply_bc(h, h->U_cur); }
    TRY(exchange_u_serial(G, true));
  } else if (s == 0) {
    EACH(h) h->U_cur = h->d_u;                 // TimePreStep has just filled its ghosts
  } else if (overlap) {
    TRY(hpbc::fill_begin(G.hs, G.n, hpbc::SLOT_U));
    EACH(h) { TRY(need_device(h)); hpbc::rk_faces(h, h->d_u, stage_coef(h, s)); }
    TRY(hpbc::xchg_start(G.hs, G.n, hpbc::SLOT_U));
    EACH(h) { TRY(need_device(h)); h->U_cur = stage_U(h, s); hpbk::apply_bc(h, h->U_cur); }
    TRY(hpbc::xchg_wait(G.hs, G.n, hpbc::SLOT_U));
    EACH(h) { TRY(need_device(h)); hpbc::unpack_faces(h, hpbc::SLOT_U, h->U_cur); }
  } else {
    EACH(h) { TRY(need_device(h)); h->U_cur = stage_U(h, s); hpbk::apply_bc(h, h->U_cur); }
    TRY(exchange_u_serial(G, true));
  }
  // ---- right-hand side (its last sweep may form the next stage solution: fused_next_stage)
  const bool visc = viscous_on(G.hs[0]);
  std::vector<double*> unext(G.n, nullptr);
  std::vector<double> adt(G.n, 0.0);
  if (fuse) EACH(h) unext[r_] = fused_next_stage(h, s, h->U_cur, &adt[r_]);
  if (overlap && hpb_stage_overlap_supported(G.hs[0])) {
    const bool fv = fused_visc(G.hs[0]);
    const int nd = G.hs[0]->geo.ndims;
    if (fv) {
      EACH(h) { TRY(need_device(h)); TRY(hpbk::qderiv_fused(h, h->U_cur)); }
      for (int slot = hpbc::SLOT_Q0; slot <= hpbc::SLOT_Q12; slot++) {
        TRY(hpbc::fill_begin(G.hs, G.n, slot));
        EACH(h) { TRY(need_device(h)); hpbc::pack_faces(h, slot, nullptr); }
        TRY(hpbc::xchg_start(G.hs, G.n, slot));
      }
    }
    for (int d = 0; d < nd; d++) {
      if (fv && d < 2) {
        const int slot = d == 0 ? hpbc::SLOT_Q0 : hpbc::SLOT_Q12;
        TRY(hpbc::xchg_wait(G.hs, G.n, slot));
        EACH(h) { TRY(need_device(h)); hpbc::unpack_faces(h, slot, nullptr); }
      }
      EACH(h) {
        TRY(need_device(h));
        double* rhs = h->d_Udot[s];
        if (!hpbk::hyperbolic_fused(h, h->U_cur, rhs, true, true, rhs, fv ? h->d_qd4 : nullptr, d, unext[r_], adt[r_], h->d_u))
          return hpb_fail(HPB_ERR_CUDA, "distributed step: fused sweep of direction %d refused the launch", d);
        stage_boundary_flux(h, h->U_cur, s, d);
      }
    }
  } else {
    const hpb_solver* h0 = G.hs[0];
    bool split_compact = false;       // a compact scheme whose grid lines are split among ranks: the reconstructions couple the ranks
    if (hpb_scheme_is_compact(h0->cfg.hyp_scheme))
      for (int d = 0; d < h0->geo.ndims; d++) split_compact = split_compact || h0->cfg.iproc[d] > 1;
    if (split_compact) {
      std::vector<const double*> U(G.n);
      std::vector<double*> out(G.n), src(G.n);
      EACH(h) {
        TRY(need_device(h)); TRY(ensure_generic(h));
        if (h->d_src) hpbk::set_zero(h, h->d_src, ncell(h));
        U[r_] = h->U_cur; out[r_] = h->d_Udot[s]; src[r_] = h->d_src;
      }
      TRY(hpbk::hyperbolic_pieces_group(G.hs, G.n, U.data(), out.data(), /*negate=*/true, /*with_source=*/true, src.data()));
      EACH(h) {
        TRY(need_device(h));
        if (hpbk::has_sponge(h)) hpbk::sponge_source(h, h->U_cur, h->d_src);
        if (viscous_on(h)) hpbk::parabolic_phase1(h, h->U_cur);
      }
    } else {
      EACH(h) { TRY(need_device(h)); TRY(rhs_part_a(h, h->U_cur, h->d_Udot[s], unext[r_], adt[r_])); }
    }
    if (visc) {
      for (int slot = hpbc::SLOT_Q0; slot <= hpbc::SLOT_Q12; slot++) {
        TRY(hpbc::fill_begin(G.hs, G.n, slot));
        EACH(h) { TRY(need_device(h)); hpbc::pack_faces(h, slot, nullptr); }
        TRY(hpbc::xchg_start(G.hs, G.n, slot));
      }
      for (int slot = hpbc::SLOT_Q0; slot <= hpbc::SLOT_Q12; slot++) {
        TRY(hpbc::xchg_wait(G.hs, G.n, slot));
        EACH(h) { TRY(need_device(h)); hpbc::unpack_faces(h, slot, nullptr); }
      }
    }
    EACH(h) { TRY(need_device(h)); TRY(rhs_part_b(h, h->U_cur, h->d_Udot[s], unext[r_], adt[r_])); stage_boundary_flux(h, h->U_cur, s); }
  }
  EACH(h) h->U_pre = unext[r_];
  EACH(h) TRY(check_async(h, "distributed stage"));
  return HPB_OK;
}

static int dist_finish(Grp& G)
{
  if (G.hs[0]->rk.glm) {
    EACH(h) {
      TRY(need_device(h));
      hpbk::glm_finish(h);                                                                // TimeGLMGEE.c:116-151
      if (h->cfg.conservation_check) hpbk::step_boundary_integral(h, cons_slot(h, 0), cons_slot(h, CONS_SLOT_STEP));
      h->t += h->cfg.dt;
      h->u_halo_valid = false;
    }
    EACH(h) TRY(check_async(h, "distributed step (glm-gee)"));
    return HPB_OK;
  }
  const bool overlap = G.hs[0]->overlap != 0;
  if (overlap) {
    // u^{n+1} on the face layers first (reads the old u), its exchange under the full update
    TRY(hpbc::fill_begin(G.hs, G.n, hpbc::SLOT_U));
    EACH(h) { TRY(need_device(h)); hpbc::rk_faces(h, h->d_u, finish_coef(h)); }
    TRY(hpbc::xchg_start(G.hs, G.n, hpbc::SLOT_U));
  }
  EACH(h) {
    TRY(need_device(h));
    hpbk::rk_finish(h);                                                                   // TimeRK.c:182-193
    if (h->cfg.conservation_check) hpbk::step_boundary_integral(h, cons_slot(h, 0), cons_slot(h, CONS_SLOT_STEP));
    h->t += h->cfg.dt;                                                                    // TimePostStep.c:36
    h->u_halo_valid = false;
  }
  if (overlap) {
    TRY(hpbc::xchg_wait(G.hs, G.n, hpbc::SLOT_U));
    EACH(h) { TRY(need_device(h)); hpbc::unpack_faces(h, hpbc::SLOT_U, h->d_u); h->u_halo_valid = true; }
  }
  EACH(h) TRY(check_async(h, "distributed step"));
  return HPB_OK;
}

static int dist_check(Grp& G, const char* what, int want_kind)
{
  if (!G.hs || G.n < 1) return hpb_fail(HPB_ERR_INVALID, "%s: no solver", what);
  EACH(h) {
    TRY(need_device(h));
    if (!hpbc::comm_ready(h)) {
      if (multi_rank(h)) return hpb_fail(HPB_ERR_INVALID, "%s: this rank has neighbours but no transport "
                                         "(hpb_comm_init_nccl / hpb_comm_init_local)", what);
    } else if (hpb_comm_kind(h) != want_kind)
      return hpb_fail(HPB_ERR_INVALID, "%s: the solver's transport is of the other kind (NCCL: hpb_*Distributed, in-process: hpb_*Local)", what);
    if (h->overlap != G.hs[0]->overlap) return hpb_fail(HPB_ERR_INVALID, "%s: the ranks disagree on the schedule", what);
  }
  return HPB_OK;
}

static int dist_steps(Grp& G, int nsteps)
{
  for (int i = 0; i < nsteps; i++) {
    TRY(dist_prestep(G));
    for (int s = 0; s < G.hs[0]->rk.ns; s++) TRY(dist_stage(G, s));
    TRY(dist_finish(G));
  }
  return HPB_OK;
}
static int dist_rhs(Grp& G)
{
  TRY(dist_prestep(G));
  return dist_stage(G, 0, /*fuse=*/false);
}

// NCCL transport: this rank's part; only enqueues (no host synchronisation: hpb_synchronize when the result is needed)
extern "C" int hpb_TimeStepsDistributed(hpb_solver* h, int nsteps)
{
  Grp G{ &h, 1 };
  TRY(dist_check(G, "TimeStepsDistributed", 1));
  return dist_steps(G, nsteps);
}
extern "C" int hpb_TimeStepDistributed(hpb_solver* h) { return hpb_TimeStepsDistributed(h, 1); }
extern "C" int hpb_RHSFunctionDistributed(hpb_solver* h)
{
  Grp G{ &h, 1 };
  TRY(dist_check(G, "RHSFunctionDistributed", 1));
  return dist_rhs(G);
}
// in-process transport: all ranks
extern "C" int hpb_TimeStepsLocal(hpb_solver** hs, int nranks, int nsteps)
{
  Grp G{ hs, nranks };
  TRY(dist_check(G, "TimeStepsLocal", 2));
  return dist_steps(G, nsteps);
}
extern "C" int hpb_RHSFunctionLocal(hpb_solver** hs, int nranks)
{
  Grp G{ hs, nranks };
  TRY(dist_check(G, "RHSFunctionLocal", 2));
  return dist_rhs(G);
}

// MPIExchangeBoundariesnD on the device solution (ghost faces of u <- the neighbours' interior layers), blocking
extern "C" int hpb_ExchangeBoundariesnD(hpb_solver* h)
{
  Grp G{ &h, 1 };
  TRY(dist_check(G, "ExchangeBoundariesnD", 1));
  TRY(exchange_u_serial(G, false));
  h->u_halo_valid = true;
  return sync_check(h, "ExchangeBoundariesnD");
}

extern "C" int hpb_ExchangeBoundariesLocal(hpb_solver** hs, int nranks)
{
  Grp G{ hs, nranks };
  TRY(dist_check(G, "ExchangeBoundariesLocal", 2));
  TRY(exchange_u_serial(G, false));
  EACH(h) { h->u_halo_valid = true; TRY(sync_check(h, "ExchangeBoundariesLocal")); }
  return HPB_OK;
}

extern "C" int hpb_set_stage_fusion(hpb_solver* h, int on)
{
  if (!h) return hpb_fail(HPB_ERR_INVALID, "set_stage_fusion: null solver");
  h->stage_fusion = on ? 1 : 0;
  h->U_pre = nullptr;
  if (on && h->device_ready && stage_fusion_on(h) && !h->rk.glm && h->rk.ns > 1) TRY(dalloc(&h->d_U2, ncell(h)));
  return HPB_OK;
}
extern "C" int hpb_stage_fusion_active(const hpb_solver* h)
{
  return (stage_fusion_on(h) && h->d_U2 && !h->rk.glm) ? 1 : 0;
}

extern "C" int hpb_set_overlap(hpb_solver* h, int on)
{
  if (!h) return hpb_fail(HPB_ERR_INVALID, "set_overlap: null solver");
  h->overlap = on ? 1 : 0;
  return HPB_OK;
}

// 1 when the production path of this configuration is driven sweep by sweep in the overlapped schedule (fused sweeps
// with everything but the Q-derivatives inside them)
extern "C" int hpb_stage_overlap_supported(const hpb_solver* h)
{
  if (!fused_path(h)) return 0;
  if (hpbk::has_sponge(h)) return 0;          // the sponge source is added after the last sweep (rhs_part_b)
  if (viscous_on(h) && !fused_visc(h)) return 0;
  if (h->cfg.model == HPB_MODEL_LINEAR_ADR) return 0;
  return 1;
}

extern "C" int hpb_dev_get_stage_rhs(hpb_solver* h, int stage, double* rhs_host)
{
  TRY(need_device(h));
  if (stage < 0 || stage >= h->rk.ns || !rhs_host) return hpb_fail(HPB_ERR_INVALID, "dev_get_stage_rhs: stage %d", stage);
  return download(h, h->d_Udot[stage], rhs_host, h->geo.npg, h->geo.nvars);
}

// comm.cu -- the halo exchange of MPIExchangeBoundariesnD (reference src/MPIFunctions/MPIExchangeBoundariesnD.c:42-173;
// its CUDA twin MPIExchangeBoundariesnD_GPU.cu:435-558) inside the library: face layers -> persistent device send
// buffers -> ncclSend / ncclRecv over NVLink on a communication stream -> persistent receive buffers -> ghost layers.
//
// What the reference does per call: for every dimension, copy `ghosts` interior layers next to each face that has a
// neighbour into a send buffer, post MPI_Irecv / MPI_Isend (tags 1630 / 1631), MPI_Waitall, copy the received layers into
// the ghost points. Faces only -- edge and corner ghosts are never touched (:60-76). The GPU twin does the same with two
// kernels per face and a blocking MPI exchange of host-staged buffers, fully synchronous (:491-525).
//
// Here:
//   * ONE pack launch and ONE unpack launch per exchange cover all faces (k_faces); the stage vector U_s and the step
//     completion are evaluated ON the face layers straight into the send buffers by k_rk_faces, so the exchange of u runs
//     under the full-array update that follows (capi.cu: dist_step);
//   * transport = NCCL point-to-point, grouped (ncclGroupStart ... ncclGroupEnd), on a high-priority communication stream
//     ordered against the compute stream by CUDA events only -- no host synchronisation anywhere in the step. libnccl is
//     loaded at run time (dlopen "libnccl.so.2": the copy PyTorch has already loaded when the caller is bench.py, the
//     system's for a C caller such as HyPar), so the library has no link-time dependency on it;
//   * an in-process transport (all ranks of the decomposition in one process: device-to-device copies between the ranks'
//     buffers) drives the same step in lock step for the single-GPU tests and for a caller that owns several GPUs;
//   * the Q-derivative exchange of the fused viscous path (NavierStokes3DParabolicFunction.c:125-130 exchanges QDerivX and
//     QDerivY, 5 components x 3 layers each) carries only what the sweeps read: per face the normal derivative's 4
//     components (u, v, w, T) + the 2 transverse ones the viscous flux of that direction uses, 2 layers (the fourth-order
//     derivative of the viscous flux reaches 2 cells) -- 134 MB instead of 378 MB per stage and rank at 512^3.
//
// Message matching. NCCL point-to-point has no tags: between one pair of ranks messages match in issue order. Per
// dimension every rank issues send(low face), send(high face), recv(high ghosts), recv(low ghosts): with iproc = 2 and
// periodic boundaries (both neighbours are the same peer) the peer's low-face send is the first message it sends us and
// lands in our high ghosts -- the role of the reference's two tags.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include "hpb_internal.h"

// ------------------------------------------------------------------------------------------ NCCL, loaded at run time
// the part of nccl.h (NCCL 2.x public API) this file uses
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                       // ncclSuccess = 0
enum { NCCL_FLOAT64 = 8 };                      // ncclDataType_t: ncclFloat64 / ncclDouble
enum { NCCL_SUM = 0, NCCL_MAX = 2 };            // ncclRedOp_t

namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  // optional (NCCL >= 2.19): buffers NCCL can map into the peers for zero-copy point-to-point
  ncclResult_t (*MemAlloc)(void**, size_t) = nullptr;
  ncclResult_t (*MemFree)(void*) = nullptr;
  ncclResult_t (*CommRegister)(ncclComm_t, void*, size_t, void**) = nullptr;
  ncclResult_t (*CommDeregister)(ncclComm_t, void*) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

int load_nccl()
{
  std::lock_guard<std::mutex> lk(g_nccl_mutex);
  if (g_nccl.lib) return HPB_OK;
  const char* names[] = { getenv("HPB_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
  void* lib = nullptr;
  for (const char* nm : names) if (nm && nm[0] && (lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!lib) return hpb_fail(HPB_ERR_INVALID, "cannot load libnccl.so.2 (%s): multi-GPU runs need NCCL", dlerror());
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(lib, name); \
  if (!g_nccl.field) { dlclose(lib); return hpb_fail(HPB_ERR_INVALID, "libnccl lacks %s", name); }
  SYM(GetVersion, "ncclGetVersion") SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
  SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  *(void**)(&g_nccl.MemAlloc) = dlsym(lib, "ncclMemAlloc");
  *(void**)(&g_nccl.MemFree) = dlsym(lib, "ncclMemFree");
  *(void**)(&g_nccl.CommRegister) = dlsym(lib, "ncclCommRegister");
  *(void**)(&g_nccl.CommDeregister) = dlsym(lib, "ncclCommDeregister");
  g_nccl.lib = lib;
  return HPB_OK;
}
#define HPB_NCCL(call) do { ncclResult_t r_ = (call); if (r_ != 0) \
  return hpb_fail(HPB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); } while (0)
} // namespace

// the transport of one solver
struct HpbComm {
  int kind = 0;                         // 1 = NCCL, 2 = in-process group
  int nranks = 0;
  ncclComm_t nccl = nullptr;
  std::vector<void*> reg;               // ncclCommRegister handles
  std::vector<hpb_solver*> group;       // in-process: every rank, indexed by rank
};

// ------------------------------------------------------------------------------------------ face kernels
namespace {

constexpr int MAXC = 10;
struct FaceDesc {
  double* buf;                          // send (pack) or receive (unpack) buffer of this face
  long long start;                      // first work item of this face in the launch
  long long npts;                       // points of the face box
  int d, off, nl;                       // dimension, first layer along d (interior-relative index), number of layers
  int ncomp;
  long long comp[MAXC];                 // offset (doubles) of each component's array from the base pointer
};
struct FaceSet { Geom G; int nf; long long total; FaceDesc f[6]; };

__device__ __forceinline__ bool face_point(const FaceSet& S, long long i, int& fi, long long& p2, long long& p1)
{
  if (i >= S.total) return false;
  fi = 0;
#pragma unroll
  for (int k = 1; k < 6; k++) if (k < S.nf && i >= S.f[k].start) fi = k;
  const FaceDesc& F = S.f[fi];
  const Geom& G = S.G;
  p2 = i - F.start;
  int b[3] = { G.N[0], G.N[1], G.N[2] };
  b[F.d] = F.nl;
  int s[3] = { (int)(p2 % b[0]), (int)((p2 / b[0]) % b[1]), (int)(p2 / ((long long)b[0] * b[1])) };
  s[F.d] += F.off;
  long long p = s[0] + G.g;
  if (G.ndims > 1) p += (long long)G.P[0] * (s[1] + G.g);
  if (G.ndims > 2) p += (long long)G.P[0] * G.P[1] * (s[2] + G.g);
  p1 = p;
  return true;
}

// buffer layout per face: component-major, then the face box with dimension 0 fastest (as k_face_copy of round 1 and
// the reference's buffers per component, MPIExchangeBoundariesnD.c:78-90)
__global__ void __launch_bounds__(256) k_faces(const FaceSet S, double* __restrict__ a, int to_buf)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int fi; long long p2, p1;
  if (!face_point(S, i, fi, p2, p1)) return;
  const FaceDesc& F = S.f[fi];
  if (to_buf) { for (int c = 0; c < F.ncomp; c++) F.buf[c * F.npts + p2] = a[F.comp[c] + p1]; }
  else        { for (int c = 0; c < F.ncomp; c++) a[F.comp[c] + p1] = F.buf[c * F.npts + p2]; }
}

// send buffers <- u + sum_s a_s k_s on the face layers: the arithmetic of k_rk_combine (kernels.cu), operation by
// operation, so that a face value is bit-identical to the one the full-array update writes afterwards
__global__ void __launch_bounds__(256) k_rk_faces(const FaceSet S, const double* __restrict__ u, const hpbc::RKCoef rk)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int fi; long long p2, p1;
  if (!face_point(S, i, fi, p2, p1)) return;
  const FaceDesc& F = S.f[fi];
  for (int c = 0; c < F.ncomp; c++) {
    const long long q = F.comp[c] + p1;
    double t = u[q];
    for (int s = 0; s < rk.n; s++) t = __dadd_rn(t, __dmul_rn(rk.a[s], rk.k[s][q]));
    F.buf[c * F.npts + p2] = t;
  }
}

inline bool viscous_on(const hpb_solver* h)
{
  return (h->cfg.model == HPB_MODEL_NS3D || h->cfg.model == HPB_MODEL_NS2D) && h->phys.Re > 0;
}
inline bool fused_visc(const hpb_solver* h)
{
  return hpbk::fused_available(h) && h->cfg.model == HPB_MODEL_NS3D && viscous_on(h);
}

// Q-derivative scalars of the fused viscous path that cross the faces of dimension d: indices dir*4 + comp into d_qd4
// (viscous_fused.cu; comp = u, v, w, T). QDerivZ is never exchanged (quirk Q1), so only x- and y-derivatives appear:
//   d = 0: ux vx wx Tx (normal) + uy vy ;  d = 1: uy vy wy Ty (normal) + ux vx ;  d = 2: ux wx + vy wy
// = the entries < 8 of the tables the sweeps read (sweep_fused.cu: hyperbolic_fused, tab[d]).
const int QD_NCOMP[3] = { 6, 6, 4 };
const int QD_COMP[3][6] = { { 0, 1, 2, 3, 4, 5 }, { 4, 5, 6, 7, 0, 1 }, { 0, 2, 5, 6, -1, -1 } };
constexpr int QD_LAYERS = 2;

// one message of an exchange: face k = 2*d + side, buffers of `field`
struct Msg { int k, field; size_t count; };

// the faces of `slot` this rank exchanges, in issue order (per dimension: low, high)
struct SlotPlan { int nmsg; Msg m[12]; };

long long face_cells(const hpb_solver* h, int d)
{
  const Geom& G = h->geo;
  long long n = 1;
  for (int k = 0; k < G.ndims; k++) if (k != d) n *= G.N[k];
  return n;
}

SlotPlan slot_plan(const hpb_solver* h, int slot)
{
  SlotPlan P; P.nmsg = 0;
  const Geom& G = h->geo;
  const int mask = hpbc::slot_dimmask(h, slot);
  const bool qd = hpbc::slot_bufset(slot) == hpbc::BUF_QD;
  if (qd && !viscous_on(h)) return P;
  for (int d = 0; d < G.ndims; d++) {
    if (!((mask >> d) & 1)) continue;
    for (int side = 0; side < 2; side++) {
      const int k = 2 * d + side;
      if (h->neighbor[k] < 0) continue;
      if (!qd) P.m[P.nmsg++] = Msg{ k, HPB_FIELD_U, (size_t)(face_cells(h, d) * G.g * G.nvars) };
      else if (fused_visc(h)) P.m[P.nmsg++] = Msg{ k, HPB_FIELD_QDERIVX, (size_t)(face_cells(h, d) * QD_LAYERS * QD_NCOMP[d]) };
      else {
        P.m[P.nmsg++] = Msg{ k, HPB_FIELD_QDERIVX, (size_t)(face_cells(h, d) * G.g * G.nvars) };
        P.m[P.nmsg++] = Msg{ k, HPB_FIELD_QDERIVY, (size_t)(face_cells(h, d) * G.g * G.nvars) };
      }
    }
  }
  return P;
}

// face descriptors of a pack (to_buf) or unpack launch of `slot`; for the exact viscous path (two 5-component fields)
// `field` selects QDerivX / QDerivY
void build_faceset(const hpb_solver* h, int slot, int field, bool to_buf, FaceSet* S)
{
  const Geom& G = h->geo;
  S->G = G; S->nf = 0; S->total = 0;
  const int mask = hpbc::slot_dimmask(h, slot);
  const bool qd = hpbc::slot_bufset(slot) == hpbc::BUF_QD;
  const bool fq = qd && fused_visc(h);
  for (int d = 0; d < G.ndims; d++) {
    if (!((mask >> d) & 1)) continue;
    for (int side = 0; side < 2; side++) {
      const int k = 2 * d + side;
      if (h->neighbor[k] < 0) continue;
      FaceDesc& F = S->f[S->nf++];
      F.d = d;
      F.nl = fq ? QD_LAYERS : G.g;
      // pack: the nl interior layers next to the face; unpack: the nl ghost layers next to it
      if (to_buf) F.off = side ? G.N[d] - F.nl : 0;
      else        F.off = side ? G.N[d] : -F.nl;
      F.npts = face_cells(h, d) * F.nl;
      F.start = S->total;
      S->total += F.npts;
      F.buf = to_buf ? h->d_send[field][k] : h->d_recv[field][k];
      if (fq) {
        F.ncomp = QD_NCOMP[d];
        for (int c = 0; c < F.ncomp; c++) F.comp[c] = (long long)QD_COMP[d][c] * G.npg;
      } else {
        F.ncomp = G.nvars;
        for (int c = 0; c < F.ncomp; c++) F.comp[c] = (long long)c * G.npg;
      }
    }
  }
}

void launch_faces(hpb_solver* h, const FaceSet& S, double* a, int to_buf)
{
  if (S.total <= 0) return;
  k_faces<<<(unsigned)((S.total + 255) / 256), 256, 0, h->stream>>>(S, a, to_buf);
  h->launches++;
}

} // namespace

// ------------------------------------------------------------------------------------------ primitives
namespace hpbc {

int slot_bufset(int slot) { return slot == SLOT_U ? BUF_U : BUF_QD; }
int slot_dimmask(const hpb_solver* h, int slot)
{
  const int all = (1 << h->geo.ndims) - 1;
  if (slot == SLOT_U) return all;
  if (slot == SLOT_Q0) return 1;
  return all & ~1;
}
bool comm_ready(const hpb_solver* h) { return h->comm != nullptr; }

static double* qd_array(hpb_solver* h, int field)
{
  return fused_visc(h) ? h->d_qd4 : h->d_QD[field - 1];
}

void pack_faces(hpb_solver* h, int slot, const double* a)
{
  ProfScope ps(h, HPB_PROF_HALO);
  FaceSet S;
  if (slot_bufset(slot) == BUF_U) { build_faceset(h, slot, HPB_FIELD_U, true, &S); launch_faces(h, S, (double*)a, 1); return; }
  if (!viscous_on(h)) return;
  build_faceset(h, slot, HPB_FIELD_QDERIVX, true, &S); launch_faces(h, S, qd_array(h, HPB_FIELD_QDERIVX), 1);
  if (!fused_visc(h)) { build_faceset(h, slot, HPB_FIELD_QDERIVY, true, &S); launch_faces(h, S, qd_array(h, HPB_FIELD_QDERIVY), 1); }
}

void unpack_faces(hpb_solver* h, int slot, double* a)
{
  ProfScope ps(h, HPB_PROF_HALO);
  FaceSet S;
  if (slot_bufset(slot) == BUF_U) { build_faceset(h, slot, HPB_FIELD_U, false, &S); launch_faces(h, S, a, 0); return; }
  if (!viscous_on(h)) return;
  build_faceset(h, slot, HPB_FIELD_QDERIVX, false, &S); launch_faces(h, S, qd_array(h, HPB_FIELD_QDERIVX), 0);
  if (!fused_visc(h)) { build_faceset(h, slot, HPB_FIELD_QDERIVY, false, &S); launch_faces(h, S, qd_array(h, HPB_FIELD_QDERIVY), 0); }
}

void rk_faces(hpb_solver* h, const double* u, const RKCoef& c)
{
  ProfScope ps(h, HPB_PROF_HALO);
  FaceSet S;
  build_faceset(h, SLOT_U, HPB_FIELD_U, true, &S);
  if (S.total <= 0) return;
  k_rk_faces<<<(unsigned)((S.total + 255) / 256), 256, 0, h->stream>>>(S, u, c);
  h->launches++;
}

static int ensure_comm_stream(hpb_solver* h)
{
  if (h->s_comm) return HPB_OK;
  int lo = 0, hi = 0;
  HPB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));        // hi = numerically lowest = highest priority
  HPB_CUDA(cudaStreamCreateWithPriority(&h->s_comm, cudaStreamNonBlocking, hi));
  for (int s = 0; s < 3; s++) for (int k = 0; k < 2; k++)
    HPB_CUDA(cudaEventCreateWithFlags(&h->ev_x[s][k], cudaEventDisableTiming));
  return HPB_OK;
}

int fill_begin(hpb_solver** hs, int n, int slot)
{
  // NCCL: the previous sends out of these buffers completed before this rank's compute stream passed xchg_wait.
  // In-process: the NEIGHBOURS copy out of this rank's send buffers on THEIR communication streams.
  for (int r = 0; r < n; r++) {
    hpb_solver* h = hs[r];
    if (!h->comm || h->comm->kind != 2) continue;
    HPB_CUDA(cudaSetDevice(h->device));
    const SlotPlan P = slot_plan(h, slot);
    for (int i = 0; i < P.nmsg; i++) {
      hpb_solver* peer = h->comm->group[h->neighbor[P.m[i].k]];
      HPB_CUDA(cudaStreamWaitEvent(h->stream, peer->ev_x[slot][1], 0));
    }
  }
  return HPB_OK;
}

int xchg_start(hpb_solver** hs, int n, int slot)
{
  // pass A: every rank marks its send buffers as filled
  for (int r = 0; r < n; r++) {
    hpb_solver* h = hs[r];
    if (!h->comm) return hpb_fail(HPB_ERR_INVALID, "halo exchange: no transport (call hpb_comm_init_nccl or hpb_comm_init_local)");
    HPB_CUDA(cudaSetDevice(h->device));
    HPB_CUDA(cudaEventRecord(h->ev_x[slot][0], h->stream));
  }
  // pass B: transfers on the communication streams
  for (int r = 0; r < n; r++) {
    hpb_solver* h = hs[r];
    HPB_CUDA(cudaSetDevice(h->device));
    const SlotPlan P = slot_plan(h, slot);
    HPB_CUDA(cudaStreamWaitEvent(h->s_comm, h->ev_x[slot][0], 0));
    if (h->comm->kind == 1 && P.nmsg > 0) {
      HPB_NCCL(g_nccl.GroupStart());
      // per dimension: send(low), send(high), recv(high), recv(low); the plan lists low then high per dimension
      int i = 0;
      while (i < P.nmsg) {
        const int d = P.m[i].k / 2;
        int j = i;
        while (j < P.nmsg && P.m[j].k / 2 == d) j++;
        for (int q = i; q < j; q++) {
          const Msg& m = P.m[q];
          HPB_NCCL(g_nccl.Send(h->d_send[m.field][m.k], m.count, NCCL_FLOAT64, h->neighbor[m.k], h->comm->nccl, h->s_comm));
          h->xchg_count++; h->xchg_bytes += (long long)m.count * 8;
        }
        for (int side = 1; side >= 0; side--) for (int q = i; q < j; q++) {
          const Msg& m = P.m[q];
          if ((m.k & 1) != side) continue;
          HPB_NCCL(g_nccl.Recv(h->d_recv[m.field][m.k], m.count, NCCL_FLOAT64, h->neighbor[m.k], h->comm->nccl, h->s_comm));
        }
        i = j;
      }
      HPB_NCCL(g_nccl.GroupEnd());
    } else if (h->comm->kind == 2) {
      // pull: my ghosts across face k come from the peer's send buffer of the opposite face
      for (int i = 0; i < P.nmsg; i++) {
        const Msg& m = P.m[i];
        hpb_solver* peer = h->comm->group[h->neighbor[m.k]];
        HPB_CUDA(cudaStreamWaitEvent(h->s_comm, peer->ev_x[slot][0], 0));
        HPB_CUDA(cudaMemcpyAsync(h->d_recv[m.field][m.k], peer->d_send[m.field][m.k ^ 1], m.count * sizeof(double),
                                 cudaMemcpyDeviceToDevice, h->s_comm));
        h->xchg_count++; h->xchg_bytes += (long long)m.count * 8;
      }
    }
    HPB_CUDA(cudaEventRecord(h->ev_x[slot][1], h->s_comm));
  }
  return HPB_OK;
}

int xchg_wait(hpb_solver** hs, int n, int slot)
{
  for (int r = 0; r < n; r++) {
    hpb_solver* h = hs[r];
    HPB_CUDA(cudaSetDevice(h->device));
    HPB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_x[slot][1], 0));
  }
  return HPB_OK;
}

// ---- small messages along one dimension's line of ranks (compact schemes across ranks). They run on the COMPUTE
// streams: the solve they serve is a chain of tiny dependent steps. In-process transport: all streams are
// synchronised around the copies (this is the reference-exact path: simplicity over speed).
int line_rank(const hpb_solver* h, int dir, int k)
{
  int ip[3] = { h->ip[0], h->ip[1], h->ip[2] };
  ip[dir] = k;
  return hpb_rank1d(h->geo.ndims, h->cfg.iproc, ip);
}

static int local_sync_all(hpb_solver** hs, int n)
{
  for (int r = 0; r < n; r++) {
    HPB_CUDA(cudaSetDevice(hs[r]->device));
    HPB_CUDA(cudaStreamSynchronize(hs[r]->stream));
  }
  return HPB_OK;
}

int line_swap(hpb_solver** hs, int n, int dir, double* const* send_lo, double* const* send_hi, double* const* recv_lo,
              double* const* recv_hi, const long long* counts, const int* active)
{
  if (n < 1 || !hs[0]->comm) return hpb_fail(HPB_ERR_INVALID, "line exchange: no transport");
  const int kind = hs[0]->comm->kind;
  if (kind == 2) { int rc = local_sync_all(hs, n); if (rc) return rc; }
  for (int r = 0; r < n; r++) {
    hpb_solver* h = hs[r];
    if (active && !active[r]) continue;
    HPB_CUDA(cudaSetDevice(h->device));
    const int ip = h->ip[dir], np = h->cfg.iproc[dir];
    const int lo = ip > 0 ? line_rank(h, dir, ip - 1) : -1, hi = ip < np - 1 ? line_rank(h, dir, ip + 1) : -1;
    const size_t count = (size_t)counts[r];          // the members of one line share their transverse extents
    if (kind == 1) {
      HPB_NCCL(g_nccl.GroupStart());
      if (lo >= 0 && send_lo && send_lo[r]) HPB_NCCL(g_nccl.Send(send_lo[r], count, NCCL_FLOAT64, lo, h->comm->nccl, h->stream));
      if (hi >= 0 && send_hi && send_hi[r]) HPB_NCCL(g_nccl.Send(send_hi[r], count, NCCL_FLOAT64, hi, h->comm->nccl, h->stream));
      if (lo >= 0 && recv_lo && recv_lo[r]) HPB_NCCL(g_nccl.Recv(recv_lo[r], count, NCCL_FLOAT64, lo, h->comm->nccl, h->stream));
      if (hi >= 0 && recv_hi && recv_hi[r]) HPB_NCCL(g_nccl.Recv(recv_hi[r], count, NCCL_FLOAT64, hi, h->comm->nccl, h->stream));
      HPB_NCCL(g_nccl.GroupEnd());
    } else {
      // pull: my recv_lo <- the low neighbour's send_hi, my recv_hi <- the high neighbour's send_lo
      if (lo >= 0 && recv_lo && recv_lo[r] && send_hi && send_hi[lo])
        HPB_CUDA(cudaMemcpyAsync(recv_lo[r], send_hi[lo], count * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      if (hi >= 0 && recv_hi && recv_hi[r] && send_lo && send_lo[hi])
        HPB_CUDA(cudaMemcpyAsync(recv_hi[r], send_lo[hi], count * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    h->xchg_count += (lo >= 0 && send_lo ? 1 : 0) + (hi >= 0 && send_hi ? 1 : 0);
  }
  if (kind == 2) { int rc = local_sync_all(hs, n); if (rc) return rc; }
  return HPB_OK;
}

int line_shift(hpb_solver** hs, int n, int dir, int step, double* const* send, double* const* recv, const long long* counts, const int* active)
{
  // step +1: to the high neighbour (received from the low one); step -1: to the low neighbour
  if (step > 0) return line_swap(hs, n, dir, nullptr, send, recv, nullptr, counts, active);
  return line_swap(hs, n, dir, send, nullptr, nullptr, recv, counts, active);
}

int line_gather(hpb_solver** hs, int n, int dir, double* const* d_val, double* const* d_scratch, double (*h_out)[64], const int* active)
{
  if (n < 1 || !hs[0]->comm) return hpb_fail(HPB_ERR_INVALID, "line gather: no transport");
  const int kind = hs[0]->comm->kind;
  if (kind == 2) {
    int rc = local_sync_all(hs, n); if (rc) return rc;
    for (int r = 0; r < n; r++) {
      if (active && !active[r]) continue;
      hpb_solver* h = hs[r];
      for (int k = 0; k < h->cfg.iproc[dir]; k++) {
        hpb_solver* m = hs[line_rank(h, dir, k)];
        HPB_CUDA(cudaSetDevice(m->device));
        HPB_CUDA(cudaMemcpy(&h_out[r][k], d_val[line_rank(h, dir, k)], sizeof(double), cudaMemcpyDeviceToHost));
      }
    }
    return HPB_OK;
  }
  hpb_solver* h = hs[0];
  if (active && !active[0]) return HPB_OK;
  HPB_CUDA(cudaSetDevice(h->device));
  const int np = h->cfg.iproc[dir], me = h->ip[dir];
  HPB_NCCL(g_nccl.GroupStart());
  for (int k = 0; k < np; k++) if (k != me) {
    const int peer = line_rank(h, dir, k);
    HPB_NCCL(g_nccl.Send(d_val[0], 1, NCCL_FLOAT64, peer, h->comm->nccl, h->stream));
    HPB_NCCL(g_nccl.Recv(d_scratch[0] + k, 1, NCCL_FLOAT64, peer, h->comm->nccl, h->stream));
  }
  HPB_NCCL(g_nccl.GroupEnd());
  HPB_CUDA(cudaMemcpyAsync(d_scratch[0] + me, d_val[0], sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  HPB_CUDA(cudaMemcpyAsync(h->h_mr, d_scratch[0], np * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  HPB_CUDA(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < np; k++) h_out[0][k] = h->h_mr[k];
  return HPB_OK;
}

int comm_free(hpb_solver* h)
{
  if (!h) return HPB_OK;
  if (h->stream || h->s_comm) cudaSetDevice(h->device);
  if (h->s_comm) cudaStreamSynchronize(h->s_comm);
  if (h->comm) {
    if (h->comm->kind == 1 && h->comm->nccl) {
      if (g_nccl.CommDeregister) for (void* r : h->comm->reg) g_nccl.CommDeregister(h->comm->nccl, r);
      g_nccl.CommDestroy(h->comm->nccl);
    }
    delete h->comm;
    h->comm = nullptr;
  }
  for (int s = 0; s < 3; s++) for (int k = 0; k < 2; k++) if (h->ev_x[s][k]) { cudaEventDestroy(h->ev_x[s][k]); h->ev_x[s][k] = nullptr; }
  if (h->s_comm) { cudaStreamDestroy(h->s_comm); h->s_comm = nullptr; }
  if (h->halo_nccl_mem && g_nccl.MemFree) {
    for (int f = 0; f < 3; f++) for (int k = 0; k < 6; k++) {
      if (h->d_send[f][k]) { g_nccl.MemFree(h->d_send[f][k]); h->d_send[f][k] = nullptr; }
      if (h->d_recv[f][k]) { g_nccl.MemFree(h->d_recv[f][k]); h->d_recv[f][k] = nullptr; }
    }
    h->halo_nccl_mem = false;
  }
  return HPB_OK;
}

} // namespace hpbc

// ------------------------------------------------------------------------------------------ C ABI: transport set-up
extern "C" int hpb_comm_get_unique_id(void* id128)
{
  if (!id128) return hpb_fail(HPB_ERR_INVALID, "comm_get_unique_id: null argument");
  int rc = load_nccl(); if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == HPB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  HPB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return HPB_OK;
}

extern "C" int hpb_comm_nccl_version(void)
{
  if (load_nccl()) return 0;
  int v = 0;
  if (g_nccl.GetVersion(&v) != 0) return 0;
  return v;
}

extern "C" int hpb_comm_init_nccl(hpb_solver* h, const void* id128, int nranks)
{
  if (!h || !id128) return hpb_fail(HPB_ERR_INVALID, "comm_init_nccl: null argument");
  if (!h->device_ready) return hpb_fail(HPB_ERR_NO_DEVICE, "no CUDA device: hypar_b200 has no CPU path");
  int np = 1;
  for (int d = 0; d < h->geo.ndims; d++) np *= h->cfg.iproc[d];
  if (nranks != np) return hpb_fail(HPB_ERR_INVALID, "comm_init_nccl: %d ranks, but iproc describes %d blocks", nranks, np);
  if (h->comm) return hpb_fail(HPB_ERR_INVALID, "comm_init_nccl: this solver already has a transport");
  int rc = load_nccl(); if (rc) return rc;
  HPB_CUDA(cudaSetDevice(h->device));
  rc = hpbc::ensure_comm_stream(h); if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  HpbComm* c = new HpbComm();
  c->kind = 1; c->nranks = nranks;
  ncclResult_t r = g_nccl.CommInitRank(&c->nccl, nranks, id, h->cfg.rank);
  if (r != 0) { delete c; return hpb_fail(HPB_ERR_CUDA, "ncclCommInitRank(rank %d of %d): %s", h->cfg.rank, nranks, g_nccl.GetErrorString(r)); }
  h->comm = c;
  // halo buffers NCCL can map into the peers (zero-copy point-to-point over NVLink): ncclMemAlloc + ncclCommRegister
  // where the library offers them; otherwise the cudaMalloc'ed buffers of hpb_create stay (NCCL then stages through
  // its own FIFO buffers). HPB_NCCL_REGISTER=0 turns this off.
  const char* reg = getenv("HPB_NCCL_REGISTER");
  if (!(reg && reg[0] == '0') && g_nccl.MemAlloc && g_nccl.MemFree && g_nccl.CommRegister) {
    bool ok = true;
    void* fresh[3][6][2] = {};
    for (int f = 0; f < 3 && ok; f++) for (int k = 0; k < 6 && ok; k++) {
      if (!h->d_send[f][k]) continue;
      for (int w = 0; w < 2 && ok; w++) ok = (g_nccl.MemAlloc(&fresh[f][k][w], h->face_bytes[k]) == 0);
    }
    if (ok) {
      for (int f = 0; f < 3; f++) for (int k = 0; k < 6; k++) {
        if (!h->d_send[f][k]) continue;
        cudaFree(h->d_send[f][k]); cudaFree(h->d_recv[f][k]);
        h->d_send[f][k] = (double*)fresh[f][k][0]; h->d_recv[f][k] = (double*)fresh[f][k][1];
        cudaMemset(h->d_send[f][k], 0, h->face_bytes[k]); cudaMemset(h->d_recv[f][k], 0, h->face_bytes[k]);
        for (int w = 0; w < 2; w++) {
          void* handle = nullptr;
          if (g_nccl.CommRegister(c->nccl, fresh[f][k][w], h->face_bytes[k], &handle) == 0 && handle) c->reg.push_back(handle);
        }
      }
      cudaStreamSynchronize(cudaStreamLegacy);
      h->halo_nccl_mem = true;
    } else {
      for (int f = 0; f < 3; f++) for (int k = 0; k < 6; k++) for (int w = 0; w < 2; w++) if (fresh[f][k][w]) g_nccl.MemFree(fresh[f][k][w]);
    }
  }
  h->u_halo_valid = false;
  return HPB_OK;
}

extern "C" int hpb_comm_init_local(hpb_solver** hs, int nranks)
{
  if (!hs || nranks < 1) return hpb_fail(HPB_ERR_INVALID, "comm_init_local: bad argument");
  int np = 1;
  for (int d = 0; d < hs[0]->geo.ndims; d++) np *= hs[0]->cfg.iproc[d];
  if (nranks != np) return hpb_fail(HPB_ERR_INVALID, "comm_init_local: %d solvers, but iproc describes %d blocks", nranks, np);
  for (int r = 0; r < nranks; r++) {
    if (!hs[r] || !hs[r]->device_ready) return hpb_fail(HPB_ERR_NO_DEVICE, "comm_init_local: solver %d has no CUDA device", r);
    if (hs[r]->cfg.rank != r) return hpb_fail(HPB_ERR_INVALID, "comm_init_local: solver %d was created as rank %d", r, hs[r]->cfg.rank);
    if (hs[r]->comm) return hpb_fail(HPB_ERR_INVALID, "comm_init_local: solver %d already has a transport", r);
  }
  for (int r = 0; r < nranks; r++) {
    HPB_CUDA(cudaSetDevice(hs[r]->device));
    int rc = hpbc::ensure_comm_stream(hs[r]); if (rc) return rc;
    HpbComm* c = new HpbComm();
    c->kind = 2; c->nranks = nranks;
    c->group.assign(hs, hs + nranks);
    hs[r]->comm = c;
    hs[r]->u_halo_valid = false;
    // peers on other devices of this process: enable direct access once (ignored when already on / not possible:
    // cudaMemcpyAsync device-to-device then stages through the host)
    for (int q = 0; q < nranks; q++) if (hs[q]->device != hs[r]->device) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, hs[r]->device, hs[q]->device) == cudaSuccess && can) cudaDeviceEnablePeerAccess(hs[q]->device, 0);
      cudaGetLastError();
    }
  }
  return HPB_OK;
}

extern "C" int hpb_comm_finalize(hpb_solver* h) { return hpbc::comm_free(h); }

extern "C" int hpb_comm_kind(const hpb_solver* h) { return (h && h->comm) ? h->comm->kind : 0; }

// MPIMax_double / MPISum_double of the reference (src/MPIFunctions/MPIMax.c, MPISum.c) for the scalars of
// TimePreStep.c:81-107 and TimePostStep.c:44-63, over the solver's transport. op: 0 = sum, 1 = max. In place.
extern "C" int hpb_comm_allreduce(hpb_solver* h, double* v, int n, int op)
{
  if (!h || !v || n < 1 || n > 8) return hpb_fail(HPB_ERR_INVALID, "comm_allreduce: bad argument (1..8 values)");
  if (!h->comm) return HPB_OK;                       // single rank
  if (h->comm->kind == 2) return hpb_fail(HPB_ERR_INVALID, "comm_allreduce: in-process ranks reduce on the host");
  HPB_CUDA(cudaSetDevice(h->device));
  HPB_CUDA(cudaMemcpyAsync(h->d_red, v, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  HPB_NCCL(g_nccl.AllReduce(h->d_red, h->d_red, (size_t)n, NCCL_FLOAT64, op ? NCCL_MAX : NCCL_SUM, h->comm->nccl, h->stream));
  HPB_CUDA(cudaMemcpyAsync(v, h->d_red, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  HPB_CUDA(cudaStreamSynchronize(h->stream));
  return HPB_OK;
}

// the ordered point-to-point operations of one exchange of `slot` on this rank (host logic; no device needed):
// ops[3*i] = 0 send / 1 recv, ops[3*i+1] = face 2*d + side, ops[3*i+2] = peer rank; counts[i] = doubles.
extern "C" int hpb_exchange_plan(const hpb_solver* h, int slot, int* ops, long long* counts, int* nops)
{
  if (!h || !ops || !nops || slot < 0 || slot > 2) return hpb_fail(HPB_ERR_INVALID, "exchange_plan: bad argument");
  const SlotPlan P = slot_plan(h, slot);
  int n = 0, i = 0;
  while (i < P.nmsg) {
    const int d = P.m[i].k / 2;
    int j = i;
    while (j < P.nmsg && P.m[j].k / 2 == d) j++;
    for (int q = i; q < j; q++) { ops[3*n] = 0; ops[3*n+1] = P.m[q].k; ops[3*n+2] = h->neighbor[P.m[q].k]; if (counts) counts[n] = (long long)P.m[q].count; n++; }
    for (int side = 1; side >= 0; side--) for (int q = i; q < j; q++) if ((P.m[q].k & 1) == side) {
      ops[3*n] = 1; ops[3*n+1] = P.m[q].k; ops[3*n+2] = h->neighbor[P.m[q].k]; if (counts) counts[n] = (long long)P.m[q].count; n++;
    }
    i = j;
  }
  *nops = n;
  return HPB_OK;
}

extern "C" int hpb_comm_stats(const hpb_solver* h, long long* messages, long long* bytes)
{
  if (!h) return hpb_fail(HPB_ERR_INVALID, "comm_stats: null solver");
  if (messages) *messages = h->xchg_count;
  if (bytes) *bytes = h->xchg_bytes;
  return HPB_OK;
}

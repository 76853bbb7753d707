/* mpishim_mp.c -- multi-process MPI shim over shared memory (same mpi.h as the single-rank shim).
 * TEST INFRASTRUCTURE ONLY: lets the UNMODIFIED reference (and HyPar's main with the library attached) run with several
 * ranks on one box that has no MPI implementation, so that decomposed runs are pinned against the real reference, not
 * against an emulation of its exchange.
 *
 *   HPB_MPI_NP=4 ./hypar_ref_mp rhs        MPI_Init forks NP-1 children (rank 0 = the parent); no launcher needed
 *
 * Semantics provided (what the reference uses, src/MPIFunctions/ and callers): eager point-to-point with (source, tag,
 * communicator) matching in FIFO order, Wait / Waitall, blocking Send / Recv, Bcast / Allreduce (sum, max, min; int,
 * double) / Allgather / Gatherv / Scatterv / Barrier built on point-to-point, Comm_dup / Comm_split / Comm_create,
 * MPI-IO reads on stdio. Reductions are evaluated on the communicator's rank 0 in rank order.
 * Transport: one anonymous shared mapping created before the fork: a bump allocator with per-size-class free lists, and
 * one inbox (linked list under a spin lock) per rank.
 */
#define _GNU_SOURCE
#include "mpi.h"
#include <sched.h>
#include <signal.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#define MAXP      64
#define NCLASS    48
#define HEAP_SIZE ((size_t)48 << 30)           /* virtual; MAP_NORESERVE */

typedef struct {
  size_t next;                                 /* offset of the next message in the inbox, 0 = none */
  int src, tag;                                /* src = WORLD rank of the sender */
  long long ctx;
  size_t bytes, cls_bytes;
} msg_t;

typedef struct { atomic_flag lock; size_t head, tail; } inbox_t;

typedef struct {
  int nproc;
  atomic_int abort_flag;
  atomic_flag heap_lock;
  size_t heap_top;
  size_t free_head[NCLASS];
  inbox_t inbox[MAXP];
} shared_t;

static shared_t* S = NULL;
static char* H = NULL;                          /* heap base (offsets are relative to it) */
static int g_rank = 0, g_np = 1;
static pid_t g_child[MAXP];

static void die(const char* m) { fprintf(stderr, "mpishim_mp[rank %d]: %s\n", g_rank, m); if (S) atomic_store(&S->abort_flag, 1); _exit(70); }
static void lock(atomic_flag* f) { while (atomic_flag_test_and_set_explicit(f, memory_order_acquire)) sched_yield(); }
static void unlock(atomic_flag* f) { atomic_flag_clear_explicit(f, memory_order_release); }
static void check_abort(void) { if (atomic_load(&S->abort_flag)) _exit(71); }

static size_t tsize(MPI_Datatype t)
{
  switch (t) { case MPI_CHAR: case MPI_BYTE: return 1; case MPI_INT: return sizeof(int); case MPI_DOUBLE: return sizeof(double); }
  die("unknown datatype"); return 0;
}

/* ---- shared heap ---- */
static int size_class(size_t n, size_t* rounded)
{
  size_t c = 64; int k = 0;
  while (c < n) { c <<= 1; k++; }
  if (k >= NCLASS) die("message too large");
  *rounded = c;
  return k;
}
static size_t heap_alloc(size_t payload, size_t* cls_bytes)
{
  size_t need = sizeof(msg_t) + payload, rounded;
  const int k = size_class(need, &rounded);
  lock(&S->heap_lock);
  size_t off = S->free_head[k];
  if (off) S->free_head[k] = ((msg_t*)(H + off))->next;
  else {
    off = S->heap_top;
    S->heap_top += rounded;
    if (S->heap_top > HEAP_SIZE) { unlock(&S->heap_lock); die("shared heap exhausted"); }
  }
  unlock(&S->heap_lock);
  *cls_bytes = rounded;
  return off;
}
static void heap_free(size_t off)
{
  msg_t* m = (msg_t*)(H + off);
  size_t rounded;
  const int k = size_class(m->cls_bytes, &rounded);
  lock(&S->heap_lock);
  m->next = S->free_head[k];
  S->free_head[k] = off;
  unlock(&S->heap_lock);
}

/* ---- communicators (per process; creation calls are collective and ordered, so the tables agree) ---- */
typedef struct { int used, size, rank, nchild; long long ctx; int world[MAXP]; } comm_t;
#define MAXCOMM 256
static comm_t C[MAXCOMM];
typedef struct { int used, n; int world[MAXP]; } group_t;
static group_t G[MAXCOMM];

static comm_t* comm_of(MPI_Comm c) { if (c < 0 || c >= MAXCOMM || !C[c].used) die("invalid communicator"); return &C[c]; }
static MPI_Comm new_comm(void) { for (int i = 1; i < MAXCOMM; i++) if (!C[i].used) { memset(&C[i], 0, sizeof(comm_t)); C[i].used = 1; return i; } die("too many communicators"); return -1; }
static long long child_ctx(comm_t* p, int color) { p->nchild++; return p->ctx * 1000003LL + (long long)p->nchild * 7919LL + (long long)(color + 1) * 104729LL; }

/* ---- point-to-point ---- */
static void push_msg(int dst_world, const void* buf, size_t bytes, int tag, long long ctx)
{
  size_t cls;
  const size_t off = heap_alloc(bytes, &cls);
  msg_t* m = (msg_t*)(H + off);
  m->next = 0; m->src = g_rank; m->tag = tag; m->ctx = ctx; m->bytes = bytes; m->cls_bytes = cls;
  if (bytes) memcpy((char*)(m + 1), buf, bytes);
  inbox_t* ib = &S->inbox[dst_world];
  lock(&ib->lock);
  if (ib->tail) ((msg_t*)(H + ib->tail))->next = off; else ib->head = off;
  ib->tail = off;
  unlock(&ib->lock);
}
/* first message from (src_world, tag, ctx) in my inbox; 1 if found and copied out */
static int try_pop(int src_world, int tag, long long ctx, void* buf, size_t cap)
{
  inbox_t* ib = &S->inbox[g_rank];
  lock(&ib->lock);
  size_t prev = 0, off = ib->head;
  while (off) {
    msg_t* m = (msg_t*)(H + off);
    if (m->src == src_world && m->tag == tag && m->ctx == ctx) break;
    prev = off; off = m->next;
  }
  if (!off) { unlock(&ib->lock); return 0; }
  msg_t* m = (msg_t*)(H + off);
  if (prev) ((msg_t*)(H + prev))->next = m->next; else ib->head = m->next;
  if (ib->tail == off) ib->tail = prev;
  unlock(&ib->lock);
  if (m->bytes > cap) die("message truncated");
  if (m->bytes) memcpy(buf, (char*)(m + 1), m->bytes);
  heap_free(off);
  return 1;
}
static void recv_blocking(int src_world, int tag, long long ctx, void* buf, size_t cap)
{
  while (!try_pop(src_world, tag, ctx, buf, cap)) { check_abort(); sched_yield(); }
}

typedef struct { int active; void* buf; size_t cap; int src_world, tag; long long ctx; } req_t;
#define MAXREQ 4096
static req_t R[MAXREQ];

int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request* req)
{
  comm_t* cm = comm_of(c);
  if (dest < 0 || dest >= cm->size) die("Isend: bad destination");
  push_msg(cm->world[dest], buf, (size_t)count * tsize(t), tag, cm->ctx);
  if (req) *req = MPI_REQUEST_NULL;             /* eager: complete */
  return MPI_SUCCESS;
}
int MPI_Irecv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* req)
{
  comm_t* cm = comm_of(c);
  if (src < 0 || src >= cm->size) die("Irecv: bad source");
  int i;
  for (i = 0; i < MAXREQ && R[i].active; i++) ;
  if (i == MAXREQ) die("too many pending receives");
  R[i].active = 1; R[i].buf = buf; R[i].cap = (size_t)count * tsize(t); R[i].src_world = cm->world[src]; R[i].tag = tag; R[i].ctx = cm->ctx;
  *req = i;
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request* r, MPI_Status* s)
{
  (void)s;
  if (!r || *r == MPI_REQUEST_NULL) return MPI_SUCCESS;
  req_t* q = &R[*r];
  if (!q->active) die("Wait on an inactive request");
  recv_blocking(q->src_world, q->tag, q->ctx, q->buf, q->cap);
  q->active = 0; *r = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}
int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s) { (void)s; for (int i = 0; i < n; i++) MPI_Wait(&r[i], NULL); return MPI_SUCCESS; }
int MPI_Send(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c) { return MPI_Isend(buf, count, t, dest, tag, c, NULL); }
int MPI_Recv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st)
{
  comm_t* cm = comm_of(c);
  if (src < 0 || src >= cm->size) die("Recv: bad source");
  recv_blocking(cm->world[src], tag, cm->ctx, buf, (size_t)count * tsize(t));
  if (st) { st->MPI_SOURCE = src; st->MPI_TAG = tag; st->MPI_ERROR = 0; }
  return MPI_SUCCESS;
}

/* ---- collectives on point-to-point (internal tags are negative) ---- */
enum { TAG_BARRIER = -101, TAG_BCAST = -102, TAG_REDUCE = -103, TAG_GATHER = -104, TAG_SCATTER = -105 };

int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c)
{
  comm_t* cm = comm_of(c);
  const size_t bytes = (size_t)n * tsize(t);
  if (cm->rank == root) { for (int r = 0; r < cm->size; r++) if (r != root) push_msg(cm->world[r], b, bytes, TAG_BCAST, cm->ctx); }
  else recv_blocking(cm->world[root], TAG_BCAST, cm->ctx, b, bytes);
  return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm c)
{
  comm_t* cm = comm_of(c);
  char z = 0;
  if (cm->rank == 0) { for (int r = 1; r < cm->size; r++) recv_blocking(cm->world[r], TAG_BARRIER, cm->ctx, &z, 1); }
  else push_msg(cm->world[0], &z, 1, TAG_BARRIER, cm->ctx);
  return MPI_Bcast(&z, 1, MPI_CHAR, 0, c);
}
static void reduce_into(void* acc, const void* x, int n, MPI_Datatype t, MPI_Op op)
{
  if (t == MPI_DOUBLE) {
    double* a = (double*)acc; const double* b = (const double*)x;
    for (int i = 0; i < n; i++) a[i] = (op == MPI_SUM) ? a[i] + b[i] : (op == MPI_MAX) ? (b[i] > a[i] ? b[i] : a[i]) : (b[i] < a[i] ? b[i] : a[i]);
  } else if (t == MPI_INT) {
    int* a = (int*)acc; const int* b = (const int*)x;
    for (int i = 0; i < n; i++) a[i] = (op == MPI_SUM) ? a[i] + b[i] : (op == MPI_MAX) ? (b[i] > a[i] ? b[i] : a[i]) : (b[i] < a[i] ? b[i] : a[i]);
  } else die("Allreduce: datatype");
}
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
  comm_t* cm = comm_of(c);
  const size_t bytes = (size_t)n * tsize(t);
  if (s != MPI_IN_PLACE && s != r) memmove(r, s, bytes);
  if (cm->rank == 0) {
    void* tmp = malloc(bytes ? bytes : 1);
    for (int q = 1; q < cm->size; q++) { recv_blocking(cm->world[q], TAG_REDUCE, cm->ctx, tmp, bytes); reduce_into(r, tmp, n, t, op); }
    free(tmp);
  } else push_msg(cm->world[0], r, bytes, TAG_REDUCE, cm->ctx);
  return MPI_Bcast(r, n, t, 0, c);
}
int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c)
{
  comm_t* cm = comm_of(c);
  const size_t blk = (size_t)rn * tsize(rt);
  if (s != MPI_IN_PLACE) memmove((char*)r + blk * cm->rank, s, (size_t)sn * tsize(st));
  if (cm->rank == 0) { for (int q = 1; q < cm->size; q++) recv_blocking(cm->world[q], TAG_GATHER, cm->ctx, (char*)r + blk * q, blk); }
  else push_msg(cm->world[0], (char*)r + blk * cm->rank, blk, TAG_GATHER, cm->ctx);
  return MPI_Bcast(r, rn * cm->size, rt, 0, c);
}
int MPI_Gatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rc, const int* displ, MPI_Datatype rt, int root, MPI_Comm c)
{
  comm_t* cm = comm_of(c);
  if (cm->rank == root) {
    for (int q = 0; q < cm->size; q++) {
      char* dst = (char*)r + (size_t)displ[q] * tsize(rt);
      if (q == root) { if (s != MPI_IN_PLACE) memmove(dst, s, (size_t)sn * tsize(st)); }
      else recv_blocking(cm->world[q], TAG_GATHER, cm->ctx, dst, (size_t)rc[q] * tsize(rt));
    }
  } else push_msg(cm->world[root], s, (size_t)sn * tsize(st), TAG_GATHER, cm->ctx);
  return MPI_SUCCESS;
}
int MPI_Scatterv(const void* s, const int* sc, const int* displ, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c)
{
  comm_t* cm = comm_of(c);
  if (cm->rank == root) {
    for (int q = 0; q < cm->size; q++) {
      const char* src = (const char*)s + (size_t)displ[q] * tsize(st);
      if (q == root) { if (r != MPI_IN_PLACE) memmove(r, src, (size_t)rn * tsize(rt)); }
      else push_msg(cm->world[q], src, (size_t)sc[q] * tsize(st), TAG_SCATTER, cm->ctx);
    }
  } else recv_blocking(cm->world[root], TAG_SCATTER, cm->ctx, r, (size_t)rn * tsize(rt));
  return MPI_SUCCESS;
}

/* ---- environment ---- */
static void on_sigchld(int sig)
{
  (void)sig;
  int st;
  pid_t p;
  while ((p = waitpid(-1, &st, WNOHANG)) > 0)
    if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0)) { if (S) atomic_store(&S->abort_flag, 1); }
}

int MPI_Init(int* a, char*** b)
{
  (void)a; (void)b;
  const char* e = getenv("HPB_MPI_NP");
  g_np = e ? atoi(e) : 1;
  if (g_np < 1 || g_np > MAXP) { fprintf(stderr, "mpishim_mp: HPB_MPI_NP must be 1..%d\n", MAXP); exit(2); }
  void* m = mmap(NULL, sizeof(shared_t) + HEAP_SIZE, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (m == MAP_FAILED) { perror("mpishim_mp: mmap"); exit(2); }
  S = (shared_t*)m;
  H = (char*)m + sizeof(shared_t);
  memset(S, 0, sizeof(shared_t));
  S->nproc = g_np;
  S->heap_top = 64;                              /* offset 0 means "none" */
  fflush(stdout); fflush(stderr);
  g_rank = 0;
  if (g_np > 1) {
    struct sigaction sa; memset(&sa, 0, sizeof(sa)); sa.sa_handler = on_sigchld; sa.sa_flags = SA_RESTART | SA_NOCLDSTOP;
    sigaction(SIGCHLD, &sa, NULL);
    for (int r = 1; r < g_np; r++) {
      pid_t p = fork();
      if (p < 0) { perror("mpishim_mp: fork"); exit(2); }
      if (p == 0) { g_rank = r; signal(SIGCHLD, SIG_DFL); break; }
      g_child[r] = p;
    }
  }
  memset(C, 0, sizeof(C));
  C[0].used = 1; C[0].size = g_np; C[0].rank = g_rank; C[0].ctx = 1;
  for (int r = 0; r < g_np; r++) C[0].world[r] = r;
  return MPI_SUCCESS;
}
int MPI_Finalize(void)
{
  MPI_Barrier(MPI_COMM_WORLD);
  fflush(stdout); fflush(stderr);
  if (g_rank != 0) _exit(0);
  signal(SIGCHLD, SIG_DFL);
  int bad = atomic_load(&S->abort_flag);
  for (int r = 1; r < g_np; r++) { int st; if (waitpid(g_child[r], &st, 0) > 0 && !(WIFEXITED(st) && WEXITSTATUS(st) == 0)) bad = 1; }
  if (bad) { fprintf(stderr, "mpishim_mp: a rank failed\n"); exit(72); }
  return MPI_SUCCESS;
}
int MPI_Comm_rank(MPI_Comm c, int* r) { *r = comm_of(c)->rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int* n) { *n = comm_of(c)->size; return MPI_SUCCESS; }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm* o)
{
  comm_t* p = comm_of(c);
  const long long ctx = child_ctx(p, 0);
  const MPI_Comm n = new_comm();
  C[n] = *p; C[n].nchild = 0; C[n].ctx = ctx;
  *o = n;
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm* c) { if (c && *c > 0 && *c < MAXCOMM) C[*c].used = 0; return MPI_SUCCESS; }
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* o)
{
  comm_t* p = comm_of(c);
  int mine[2] = { color, key }, all[2 * MAXP];
  MPI_Allgather(mine, 2, MPI_INT, all, 2, MPI_INT, c);
  const long long ctx = child_ctx(p, color);
  const MPI_Comm n = new_comm();
  comm_t* q = &C[n];
  q->ctx = ctx; q->size = 0;
  /* members with my colour, ordered by (key, rank in the parent) */
  int idx[MAXP], m = 0;
  for (int r = 0; r < p->size; r++) if (all[2*r] == color) idx[m++] = r;
  for (int i = 1; i < m; i++) { int v = idx[i], j = i - 1; while (j >= 0 && all[2*idx[j]+1] > all[2*v+1]) { idx[j+1] = idx[j]; j--; } idx[j+1] = v; }
  for (int i = 0; i < m; i++) { q->world[i] = p->world[idx[i]]; if (idx[i] == p->rank) q->rank = i; }
  q->size = m;
  *o = n;
  return MPI_SUCCESS;
}
int MPI_Comm_group(MPI_Comm c, MPI_Group* g)
{
  comm_t* p = comm_of(c);
  for (int i = 0; i < MAXCOMM; i++) if (!G[i].used) { G[i].used = 1; G[i].n = p->size; memcpy(G[i].world, p->world, sizeof(p->world)); *g = i; return MPI_SUCCESS; }
  die("too many groups"); return 1;
}
int MPI_Group_incl(MPI_Group g, int n, const int* r, MPI_Group* o)
{
  for (int i = 0; i < MAXCOMM; i++) if (!G[i].used) {
    G[i].used = 1; G[i].n = n;
    for (int k = 0; k < n; k++) G[i].world[k] = G[g].world[r[k]];
    *o = i; return MPI_SUCCESS;
  }
  die("too many groups"); return 1;
}
int MPI_Group_free(MPI_Group* g) { if (g && *g >= 0 && *g < MAXCOMM) G[*g].used = 0; return MPI_SUCCESS; }
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm* o)
{
  comm_t* p = comm_of(c);
  const long long ctx = child_ctx(p, 0);
  int me = -1;
  for (int k = 0; k < G[g].n; k++) if (G[g].world[k] == g_rank) me = k;
  if (me < 0) { *o = -1; return MPI_SUCCESS; }    /* MPI_COMM_NULL */
  const MPI_Comm n = new_comm();
  C[n].ctx = ctx; C[n].size = G[g].n; C[n].rank = me;
  memcpy(C[n].world, G[g].world, sizeof(G[g].world));
  *o = n;
  return MPI_SUCCESS;
}

/* ---- MPI-IO on top of stdio (ReadArray.c mpi-io input mode only) ---- */
struct hpb_mpishim_file { FILE* f; };
int MPI_File_open(MPI_Comm c, const char* name, int mode, MPI_Info info, MPI_File* fh)
{
  (void)c; (void)mode; (void)info;
  FILE* f = fopen(name, "rb");
  if (!f) return 1;
  *fh = (MPI_File) malloc(sizeof(struct hpb_mpishim_file));
  (*fh)->f = f;
  return MPI_SUCCESS;
}
int MPI_File_seek(MPI_File fh, MPI_Offset off, int whence) { (void)whence; return fseek(fh->f, (long)off, SEEK_SET); }
int MPI_File_read(MPI_File fh, void* buf, int n, MPI_Datatype t, MPI_Status* s)
{ (void)s; size_t got = fread(buf, tsize(t), (size_t)n, fh->f); return got == (size_t)n ? MPI_SUCCESS : 1; }
int MPI_File_close(MPI_File* fh) { fclose((*fh)->f); free(*fh); *fh = NULL; return MPI_SUCCESS; }

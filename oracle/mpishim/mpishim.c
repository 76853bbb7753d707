/* mpishim.c -- single-rank MPI shim (see mpi.h). TEST INFRASTRUCTURE ONLY. */
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static size_t tsize(MPI_Datatype t)
{
  switch (t) {
    case MPI_CHAR: case MPI_BYTE: return 1;
    case MPI_INT:    return sizeof(int);
    case MPI_DOUBLE: return sizeof(double);
  }
  fprintf(stderr, "mpishim: unknown datatype %d\n", t); abort();
}

/* ---- point-to-point: messages to self, matched by tag in FIFO order ---- */
typedef struct msg { int tag; size_t bytes; void* data; struct msg* next; } msg_t;     /* unexpected sends */
typedef struct rcv { int tag; size_t bytes; void* buf; int done; struct rcv* next; } rcv_t; /* posted recvs */
static msg_t *msg_head = NULL, *msg_tail = NULL;
static rcv_t *rcv_head = NULL, *rcv_tail = NULL;

static void check_self(int peer)
{
  if (peer != 0) { fprintf(stderr, "mpishim: single-rank shim got peer %d\n", peer); abort(); }
}

int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request* req)
{
  (void)c; check_self(dest);
  size_t bytes = (size_t)count * tsize(t);
  rcv_t *r = rcv_head, *prev = NULL;
  while (r && (r->tag != tag || r->done)) { prev = r; r = r->next; }
  if (r) {
    if (bytes > r->bytes) { fprintf(stderr, "mpishim: message truncated\n"); abort(); }
    memcpy(r->buf, buf, bytes);
    r->done = 1;
    /* unlink */
    if (prev) prev->next = r->next; else rcv_head = r->next;
    if (rcv_tail == r) rcv_tail = prev;
    free(r);
  } else {
    msg_t* m = (msg_t*) malloc(sizeof(msg_t));
    m->tag = tag; m->bytes = bytes; m->data = malloc(bytes ? bytes : 1); m->next = NULL;
    memcpy(m->data, buf, bytes);
    if (msg_tail) msg_tail->next = m; else msg_head = m;
    msg_tail = m;
  }
  if (req) *req = 0;
  return MPI_SUCCESS;
}

int MPI_Irecv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* req)
{
  (void)c; check_self(src);
  size_t bytes = (size_t)count * tsize(t);
  msg_t *m = msg_head, *prev = NULL;
  while (m && m->tag != tag) { prev = m; m = m->next; }
  if (m) {
    if (m->bytes > bytes) { fprintf(stderr, "mpishim: message truncated\n"); abort(); }
    memcpy(buf, m->data, m->bytes);
    if (prev) prev->next = m->next; else msg_head = m->next;
    if (msg_tail == m) msg_tail = prev;
    free(m->data); free(m);
  } else {
    rcv_t* r = (rcv_t*) malloc(sizeof(rcv_t));
    r->tag = tag; r->bytes = bytes; r->buf = buf; r->done = 0; r->next = NULL;
    if (rcv_tail) rcv_tail->next = r; else rcv_head = r;
    rcv_tail = r;
  }
  if (req) *req = 0;
  return MPI_SUCCESS;
}

int MPI_Send(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c)
{ MPI_Request r; return MPI_Isend(buf, count, t, dest, tag, c, &r); }

int MPI_Recv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st)
{
  (void)st; (void)c; check_self(src);
  msg_t* m = msg_head;
  while (m && m->tag != tag) m = m->next;
  if (!m) { fprintf(stderr, "mpishim: blocking MPI_Recv(tag %d) with no matching send: deadlock\n", tag); abort(); }
  MPI_Request r; return MPI_Irecv(buf, count, t, src, tag, c, &r);
}

int MPI_Wait(MPI_Request* r, MPI_Status* s) { (void)s; if (r) *r = MPI_REQUEST_NULL; return MPI_SUCCESS; }
int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s)
{
  (void)s;
  /* every posted receive must have been matched by now (all sends are to self and eager) */
  if (rcv_head) { fprintf(stderr, "mpishim: MPI_Waitall with an unmatched receive (tag %d)\n", rcv_head->tag); abort(); }
  for (int i = 0; i < n; i++) r[i] = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}

/* ---- environment / communicators ---- */
int MPI_Init(int* a, char*** b) { (void)a; (void)b; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int* n) { (void)c; *n = 1; return MPI_SUCCESS; }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm* o) { *o = c; return MPI_SUCCESS; }
int MPI_Comm_free(MPI_Comm* c) { (void)c; return MPI_SUCCESS; }
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* o) { (void)color; (void)key; *o = c; return MPI_SUCCESS; }
int MPI_Comm_group(MPI_Comm c, MPI_Group* g) { (void)c; *g = 0; return MPI_SUCCESS; }
int MPI_Group_incl(MPI_Group g, int n, const int* r, MPI_Group* o) { (void)g; (void)n; (void)r; *o = 0; return MPI_SUCCESS; }
int MPI_Group_free(MPI_Group* g) { (void)g; return MPI_SUCCESS; }
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm* o) { (void)g; *o = c; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }

/* ---- collectives over one rank are copies ---- */
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)root; (void)c; return MPI_SUCCESS; }

int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
  (void)op; (void)c;
  if (s != MPI_IN_PLACE && s != r) memmove(r, s, (size_t)n * tsize(t));
  return MPI_SUCCESS;
}

int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c)
{
  (void)rn; (void)rt; (void)c;
  if (s != MPI_IN_PLACE && s != r) memmove(r, s, (size_t)sn * tsize(st));
  return MPI_SUCCESS;
}

int MPI_Gatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rc, const int* displ,
                MPI_Datatype rt, int root, MPI_Comm c)
{
  (void)rc; (void)root; (void)c;
  if (s != MPI_IN_PLACE) memmove((char*)r + (size_t)displ[0]*tsize(rt), s, (size_t)sn * tsize(st));
  return MPI_SUCCESS;
}

int MPI_Scatterv(const void* s, const int* sc, const int* displ, MPI_Datatype st, void* r, int rn,
                 MPI_Datatype rt, int root, MPI_Comm c)
{
  (void)sc; (void)root; (void)c;
  if (r != MPI_IN_PLACE) memmove(r, (const char*)s + (size_t)displ[0]*tsize(st), (size_t)rn * tsize(rt));
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

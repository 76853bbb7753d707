/* mpi.h -- single-rank MPI shim. TEST INFRASTRUCTURE ONLY.
 *
 * This container (and the GPU box) has no MPI implementation. The reference's serial build
 * (-Dserial) compiles MPIExchangeBoundariesnD to an empty function
 * (src/MPIFunctions/MPIExchangeBoundariesnD.c:51,171), so it is NOT semantically identical to
 * the published MPI build for periodic viscous cases: the MPI
 * build with one rank self-sends and thereby makes QDerivX/QDerivY periodic.
 * This shim lets the unmodified reference be compiled WITHOUT -Dserial for exactly one rank:
 * a self Isend is matched to the posted self Irecv by tag (FIFO per tag), collectives are
 * copies. Only the symbols the reference uses are provided.
 */
#ifndef HPB_MPISHIM_H
#define HPB_MPISHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Group;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef struct hpb_mpishim_file* MPI_File;

#define MPI_COMM_WORLD     0
#define MPI_SUCCESS        0
#define MPI_REQUEST_NULL  (-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE      ((void*)1)
#define MPI_INFO_NULL      0
#define MPI_MODE_RDONLY    1
#define MPI_SEEK_SET       0

#define MPI_CHAR    1
#define MPI_BYTE    2
#define MPI_INT     3
#define MPI_DOUBLE  4

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

int MPI_Init(int*, char***);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Comm_size(MPI_Comm, int*);
int MPI_Comm_dup(MPI_Comm, MPI_Comm*);
int MPI_Comm_free(MPI_Comm*);
int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm*);
int MPI_Comm_group(MPI_Comm, MPI_Group*);
int MPI_Group_incl(MPI_Group, int, const int*, MPI_Group*);
int MPI_Group_free(MPI_Group*);
int MPI_Comm_create(MPI_Comm, MPI_Group, MPI_Comm*);
int MPI_Barrier(MPI_Comm);
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Gatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, int, MPI_Comm);
int MPI_Scatterv(const void*, const int*, const int*, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*);
int MPI_Wait(MPI_Request*, MPI_Status*);
int MPI_Waitall(int, MPI_Request*, MPI_Status*);
int MPI_File_open(MPI_Comm, const char*, int, MPI_Info, MPI_File*);
int MPI_File_seek(MPI_File, MPI_Offset, int);
int MPI_File_read(MPI_File, void*, int, MPI_Datatype, MPI_Status*);
int MPI_File_close(MPI_File*);

#ifdef __cplusplus
}
#endif
#endif

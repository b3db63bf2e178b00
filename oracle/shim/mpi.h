/* Single-rank stand-in for <mpi.h> (the reference needs MVAPICH2, which is not
 * in the image).  Every collective degenerates to the identity on one rank.
 * Written for this repo; test infrastructure only. */
#ifndef GMM_SHIM_MPI_H
#define GMM_SHIM_MPI_H
#include <string.h>
#include <chrono>
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_THREAD_MULTIPLE 3
#define MPI_THREAD_FUNNELED 1
#define MPI_COMM_WORLD 0
#define MPI_INT 1
#define MPI_FLOAT 2
#define MPI_SUM 1
#define MPI_IN_PLACE ((void*)1)
typedef int MPI_Comm; typedef int MPI_Datatype; typedef int MPI_Op;
typedef struct { int dummy; } MPI_Status;
static inline int MPI_Init_thread(int*, char***, int req, int* prov) { *prov = req; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int* n) { *n = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return 0; }
static inline int MPI_Get_processor_name(char* name, int* len) { strcpy(name, "localhost"); *len = 9; return 0; }
static inline double MPI_Wtime() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm) { return 0; }
static inline int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm) { return 0; }
static inline int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*) { return 0; }
static inline int MPI_Finalize() { return 0; }
#endif

// In-process mock of the MPI subset schwarz-lib uses, for
// running the reference's OWN sources as an oracle: a rank is a thread, a
// message is a memcpy through a mailbox, a window is a table of base pointers.
// TEST INFRASTRUCTURE (oracle/_ref); never part of the product.
#ifndef MPI_H_SHIM
#define MPI_H_SHIM
#include <complex>
#include <cstddef>
#include <cstdint>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef int MPI_Win;
typedef int MPI_Aint;
struct MPI_Status { int source, tag; };
struct MPI_Request_impl;
typedef MPI_Request_impl *MPI_Request;

enum { MPI_COMM_WORLD = 0 };
// datatype = (kind << 8) | size ; kind 1 signed int, 2 unsigned int, 3 floating, 4 complex
#define MOCK_DT(kind, size) (((kind) << 8) | (size))
enum { MPI_CHAR = MOCK_DT(1, 1), MPI_UNSIGNED_CHAR = MOCK_DT(2, 1), MPI_UNSIGNED = MOCK_DT(2, 4),
       MPI_INT = MOCK_DT(1, 4), MPI_UNSIGNED_LONG = MOCK_DT(2, 8), MPI_UNSIGNED_SHORT = MOCK_DT(2, 2),
       MPI_LONG = MOCK_DT(1, 8), MPI_FLOAT = MOCK_DT(3, 4), MPI_DOUBLE = MOCK_DT(3, 8),
       MPI_LONG_DOUBLE = MOCK_DT(3, 16), MPI_COMPLEX = MOCK_DT(4, 8), MPI_DOUBLE_COMPLEX = MOCK_DT(4, 16) };
enum { MPI_SUM = 1, MPI_MIN = 2, MPI_MAX = 3 };
enum { MPI_INFO_NULL = 0, MPI_LOCK_SHARED = 1, MPI_COMM_TYPE_SHARED = 1, MPI_THREAD_MULTIPLE = 3,
       MPI_SUCCESS = 0 };

int MPI_Init(int *, char ***);
int MPI_Init_thread(int *, char ***, int, int *);
int MPI_Finalize();
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Comm_split_type(MPI_Comm, int, int, MPI_Info, MPI_Comm *);
int MPI_Barrier(MPI_Comm);
double MPI_Wtime();
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm);
int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm, MPI_Request *);
int MPI_Irecv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Request *);
int MPI_Wait(MPI_Request *, MPI_Status *);
int MPI_Alltoall(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, MPI_Comm);
int MPI_Allgather(const void *s, int sc, MPI_Datatype st, void *r, int rc, MPI_Datatype rt, MPI_Comm);
int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype t, MPI_Op op, MPI_Comm);
int MPI_Win_create(void *base, MPI_Aint size, int disp_unit, MPI_Info, MPI_Comm, MPI_Win *);
int MPI_Win_lock_all(int, MPI_Win);
int MPI_Win_unlock_all(MPI_Win);
int MPI_Win_lock(int, int, int, MPI_Win);
int MPI_Win_unlock(int, MPI_Win);
int MPI_Win_flush(int, MPI_Win);
int MPI_Win_flush_local(int, MPI_Win);
int MPI_Win_free(MPI_Win *);
int MPI_Put(const void *o, int oc, MPI_Datatype ot, int target, MPI_Aint disp, int tc, MPI_Datatype tt, MPI_Win);
int MPI_Get(void *o, int oc, MPI_Datatype ot, int target, MPI_Aint disp, int tc, MPI_Datatype tt, MPI_Win);
int MPI_Accumulate(const void *o, int oc, MPI_Datatype ot, int target, MPI_Aint disp, int tc,
                   MPI_Datatype tt, MPI_Op op, MPI_Win);

// launcher of the mock (ref_driver.cpp)
namespace mockmpi {
void run(int nranks, void (*fn)(int rank, void *arg), void *arg);
}
#endif

/* TEST INFRASTRUCTURE — single-rank stand-in for <mpi.h> so the UNMODIFIED reference pRIblast can be built
 * in a container without MPI.  Covers exactly the symbols the reference uses:
 * rank 0 of 1, RMA window = the caller's local int, gathers/scatters = memcpy, Send/Recv unreachable. */
#ifndef PRIB_MPI_SHIM_H
#define PRIB_MPI_SHIM_H
#include <cstdlib>
#include <cstring>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long MPI_Aint;
typedef struct { int unused; } MPI_Status;
typedef struct MPI_Win_s { void *base; } *MPI_Win;

#define MPI_COMM_WORLD 0
#define MPI_INT 4
#define MPI_UNSIGNED_CHAR 1
#define MPI_INFO_NULL 0
#define MPI_SUM 1
#define MPI_REPLACE 2
#define MPI_LOCK_SHARED 1
#define MPI_LOCK_EXCLUSIVE 2
#define MPI_ANY_SOURCE (-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)

static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize() { return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Scatterv(const void *sb, const int *cnt, const int *displ, MPI_Datatype t, void *rb, int rc,
                               MPI_Datatype, int, MPI_Comm) {
  std::memcpy(rb, (const char *)sb + (size_t)displ[0] * t, (size_t)(rc < cnt[0] ? rc : cnt[0]) * t);
  return 0;
}
static inline int MPI_Gather(const void *sb, int sc, MPI_Datatype t, void *rb, int, MPI_Datatype, int, MPI_Comm) {
  std::memcpy(rb, sb, (size_t)sc * t);
  return 0;
}
static inline int MPI_Gatherv(const void *sb, int sc, MPI_Datatype t, void *rb, const int *, const int *displ,
                              MPI_Datatype, int, MPI_Comm) {
  std::memcpy((char *)rb + (size_t)displ[0] * t, sb, (size_t)sc * t);
  return 0;
}
static inline int MPI_Alloc_mem(MPI_Aint size, MPI_Info, void *baseptr) {
  *(void **)baseptr = std::malloc((size_t)size);
  return 0;
}
static inline int MPI_Free_mem(void *p) { std::free(p); return 0; }
static inline int MPI_Win_create(void *base, MPI_Aint, int, MPI_Info, MPI_Comm, MPI_Win *win) {
  *win = (MPI_Win)std::malloc(sizeof(struct MPI_Win_s));
  (*win)->base = base;
  return 0;
}
static inline int MPI_Win_free(MPI_Win *win) { std::free(*win); *win = 0; return 0; }
static inline int MPI_Win_lock(int, int, int, MPI_Win) { return 0; }
static inline int MPI_Win_unlock(int, MPI_Win) { return 0; }
static inline int MPI_Fetch_and_op(const void *origin, void *result, MPI_Datatype, int, MPI_Aint disp, MPI_Op op,
                                   MPI_Win win) {
  int *target = (int *)win->base + disp;
  *(int *)result = *target;
  if (op == MPI_SUM) *target += *(const int *)origin;
  else *target = *(const int *)origin;
  return 0;
}
static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return 0; }
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return 0; }
#endif

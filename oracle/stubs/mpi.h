#ifndef STUB_MPI_H
#define STUB_MPI_H
typedef int MPI_Comm; typedef int MPI_Datatype; typedef int MPI_Op; typedef int MPI_Win; typedef int MPI_Info; typedef int MPI_Group; typedef long MPI_Aint;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 1
#define MPI_FLOAT 2
#define MPI_INT 3
#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_IN_PLACE ((void*)1)
#define MPI_PROC_NULL (-2)
#ifdef __cplusplus
extern "C" {
#endif
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Barrier(MPI_Comm);
#ifdef __cplusplus
}
#endif
#endif

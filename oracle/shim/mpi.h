/* Single-rank MPI shim for building the BOUT++ reference sources as a CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md). Written from the MPI standard's
 * public function signatures; provides exactly the subset BOUT++'s hot path calls
 * (reference: include/bout/mpi_wrapper.hxx:37-154, include/bout/cyclic_reduction.hxx:215-440,
 * src/mesh/impls/bout/boutmesh.cxx:668-1008).  World size is always 1, rank 0.
 * Messages to self are matched by (communicator, tag) in FIFO order.
 */
#ifndef BOUT_B200_ORACLE_MPI_SHIM_H
#define BOUT_B200_ORACLE_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Group;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct MPI_Status {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_UNDEFINED (-32766)
#define MPI_ANY_TAG (-1)
#define MPI_ANY_SOURCE (-2)

#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2

#define MPI_GROUP_NULL (-1)
#define MPI_GROUP_EMPTY 0

#define MPI_REQUEST_NULL (-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

/* datatypes: the value encodes nothing; sizes are looked up in the shim */
#define MPI_DATATYPE_NULL 0
#define MPI_CHAR 1
#define MPI_BYTE 2
#define MPI_INT 3
#define MPI_DOUBLE 4
#define MPI_DOUBLE_COMPLEX 5
#define MPI_C_BOOL 6
#define MPI_UNSIGNED 7
#define MPI_LONG 8
#define MPI_FLOAT 9
#define MPI_C_DOUBLE_COMPLEX 5

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_LAND 4
#define MPI_LOR 5
#define MPI_BOR 6
#define MPI_BAND 7
#define MPI_PROD 8

int MPI_Init(int* argc, char*** argv);
int MPI_Initialized(int* flag);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);

int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm* newcomm);
int MPI_Comm_free(MPI_Comm* comm);
int MPI_Comm_group(MPI_Comm comm, MPI_Group* group);
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm* newcomm);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm* newcomm);

int MPI_Group_range_incl(MPI_Group group, int n, int ranges[][3], MPI_Group* newgroup);
int MPI_Group_union(MPI_Group g1, MPI_Group g2, MPI_Group* newgroup);
int MPI_Group_free(MPI_Group* group);

int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype,
                    MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* datatype);
int MPI_Type_free(MPI_Datatype* datatype);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void* buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype,
                  MPI_Op op, MPI_Comm comm);
int MPI_Scan(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype,
             MPI_Op op, MPI_Comm comm);
int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf,
                const int* recvcounts, const int* displs, MPI_Datatype recvtype, int root,
                MPI_Comm comm);
int MPI_Scatterv(const void* sendbuf, const int* sendcounts, const int* displs,
                 MPI_Datatype sendtype, void* recvbuf, int recvcount,
                 MPI_Datatype recvtype, int root, MPI_Comm comm);

int MPI_Send(const void* buf, int count, MPI_Datatype datatype, int dest, int tag,
             MPI_Comm comm);
int MPI_Recv(void* buf, int count, MPI_Datatype datatype, int source, int tag,
             MPI_Comm comm, MPI_Status* status);
int MPI_Isend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag,
              MPI_Comm comm, MPI_Request* request);
int MPI_Irecv(void* buf, int count, MPI_Datatype datatype, int source, int tag,
              MPI_Comm comm, MPI_Request* request);
int MPI_Sendrecv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, int dest,
                 int sendtag, void* recvbuf, int recvcount, MPI_Datatype recvtype,
                 int source, int recvtag, MPI_Comm comm, MPI_Status* status);
int MPI_Wait(MPI_Request* request, MPI_Status* status);
int MPI_Waitall(int count, MPI_Request array_of_requests[], MPI_Status array_of_statuses[]);
int MPI_Waitany(int count, MPI_Request array_of_requests[], int* indx, MPI_Status* status);

#ifdef __cplusplus
}
#endif
#endif

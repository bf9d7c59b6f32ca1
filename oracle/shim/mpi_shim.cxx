// Single-rank MPI shim implementation (test infrastructure; see mpi.h).
#include "mpi.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <vector>

namespace {
struct Message {
  MPI_Comm comm;
  int tag;
  std::vector<char> bytes;
};
struct Request {
  bool active = false;
  bool is_recv = false;
  void* buf = nullptr;
  size_t nbytes = 0;
  MPI_Comm comm = 0;
  int tag = 0;
};
struct VectorType {
  int count, blocklength, stride;
  MPI_Datatype base;
};

std::deque<Message> mailbox;
std::vector<Request> requests;
std::vector<VectorType> vector_types; // handle = 1000 + index
int next_comm = 16;
bool initialised = false;

size_t elemSize(MPI_Datatype t) {
  switch (t) {
  case MPI_CHAR:
  case MPI_BYTE:
    return 1;
  case MPI_C_BOOL:
    return sizeof(bool);
  case MPI_INT:
  case MPI_UNSIGNED:
    return 4;
  case MPI_FLOAT:
    return 4;
  case MPI_LONG:
    return sizeof(long);
  case MPI_DOUBLE:
    return 8;
  case MPI_DOUBLE_COMPLEX:
    return 16;
  default:
    break;
  }
  if (t >= 1000 && static_cast<size_t>(t - 1000) < vector_types.size()) {
    const auto& v = vector_types[t - 1000];
    return static_cast<size_t>(v.count) * v.blocklength * elemSize(v.base);
  }
  std::fprintf(stderr, "mpi_shim: unknown datatype %d\n", t);
  std::abort();
}

// Copy `count` items of type t between a (possibly strided) user buffer and packed bytes.
void packBytes(const void* buf, int count, MPI_Datatype t, std::vector<char>& out) {
  if (t >= 1000) {
    const auto& v = vector_types[t - 1000];
    const size_t es = elemSize(v.base);
    const char* p = static_cast<const char*>(buf);
    for (int c = 0; c < count; ++c) {
      for (int i = 0; i < v.count; ++i) {
        const char* blk = p + static_cast<size_t>(i) * v.stride * es;
        out.insert(out.end(), blk, blk + static_cast<size_t>(v.blocklength) * es);
      }
      // extent of a vector type: ((count-1)*stride + blocklength) elements
      p += (static_cast<size_t>(v.count - 1) * v.stride + v.blocklength) * es;
    }
  } else {
    const char* p = static_cast<const char*>(buf);
    out.insert(out.end(), p, p + static_cast<size_t>(count) * elemSize(t));
  }
}
void unpackBytes(void* buf, int count, MPI_Datatype t, const std::vector<char>& in) {
  if (t >= 1000) {
    const auto& v = vector_types[t - 1000];
    const size_t es = elemSize(v.base);
    char* p = static_cast<char*>(buf);
    size_t off = 0;
    for (int c = 0; c < count; ++c) {
      for (int i = 0; i < v.count; ++i) {
        const size_t n = static_cast<size_t>(v.blocklength) * es;
        if (off + n > in.size()) {
          return;
        }
        std::memcpy(p + static_cast<size_t>(i) * v.stride * es, in.data() + off, n);
        off += n;
      }
      p += (static_cast<size_t>(v.count - 1) * v.stride + v.blocklength) * es;
    }
  } else {
    size_t n = static_cast<size_t>(count) * elemSize(t);
    if (n > in.size()) {
      n = in.size();
    }
    std::memcpy(buf, in.data(), n);
  }
}

struct RecvDesc {
  void* buf;
  int count;
  MPI_Datatype type;
};
std::vector<RecvDesc> recv_desc; // parallel to requests

bool tryComplete(int r, MPI_Status* status) {
  Request& q = requests[r];
  if (!q.active) {
    return true;
  }
  if (!q.is_recv) {
    q.active = false;
    return true;
  }
  for (auto it = mailbox.begin(); it != mailbox.end(); ++it) {
    if (it->comm == q.comm && (q.tag == MPI_ANY_TAG || it->tag == q.tag)) {
      unpackBytes(recv_desc[r].buf, recv_desc[r].count, recv_desc[r].type, it->bytes);
      if (status != nullptr) {
        status->MPI_SOURCE = 0;
        status->MPI_TAG = it->tag;
        status->MPI_ERROR = MPI_SUCCESS;
      }
      mailbox.erase(it);
      q.active = false;
      return true;
    }
  }
  return false;
}

int newRequest() {
  for (size_t i = 0; i < requests.size(); ++i) {
    if (!requests[i].active) {
      return static_cast<int>(i);
    }
  }
  requests.emplace_back();
  recv_desc.emplace_back();
  return static_cast<int>(requests.size() - 1);
}

[[noreturn]] void deadlock(const char* where) {
  std::fprintf(stderr, "mpi_shim: %s would block forever on a single rank\n", where);
  std::abort();
}
} // namespace

extern "C" {

int MPI_Init(int*, char***) {
  initialised = true;
  return MPI_SUCCESS;
}
int MPI_Initialized(int* flag) {
  *flag = initialised ? 1 : 0;
  return MPI_SUCCESS;
}
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int errorcode) { std::exit(errorcode); }
double MPI_Wtime(void) {
  using clock = std::chrono::steady_clock;
  static const auto t0 = clock::now();
  return std::chrono::duration<double>(clock::now() - t0).count();
}

int MPI_Comm_size(MPI_Comm, int* size) {
  *size = 1;
  return MPI_SUCCESS;
}
int MPI_Comm_rank(MPI_Comm, int* rank) {
  *rank = 0;
  return MPI_SUCCESS;
}
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm* newcomm) {
  *newcomm = (comm == MPI_COMM_NULL) ? MPI_COMM_NULL : next_comm++;
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm* comm) {
  *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}
int MPI_Comm_group(MPI_Comm, MPI_Group* group) {
  *group = 1; // the group {0}
  return MPI_SUCCESS;
}
int MPI_Comm_create(MPI_Comm, MPI_Group group, MPI_Comm* newcomm) {
  *newcomm = (group == 1) ? next_comm++ : MPI_COMM_NULL;
  return MPI_SUCCESS;
}
int MPI_Comm_split(MPI_Comm, int color, int, MPI_Comm* newcomm) {
  *newcomm = (color == MPI_UNDEFINED) ? MPI_COMM_NULL : next_comm++;
  return MPI_SUCCESS;
}

int MPI_Group_range_incl(MPI_Group group, int n, int ranges[][3], MPI_Group* newgroup) {
  bool has0 = false;
  if (group == 1) {
    for (int i = 0; i < n; ++i) {
      const int first = ranges[i][0], last = ranges[i][1], stride = ranges[i][2];
      if (stride > 0) {
        for (int r = first; r <= last; r += stride) {
          has0 = has0 || (r == 0);
        }
      } else if (stride < 0) {
        for (int r = first; r >= last; r += stride) {
          has0 = has0 || (r == 0);
        }
      }
    }
  }
  *newgroup = has0 ? 1 : MPI_GROUP_EMPTY;
  return MPI_SUCCESS;
}
int MPI_Group_union(MPI_Group g1, MPI_Group g2, MPI_Group* newgroup) {
  *newgroup = (g1 == 1 || g2 == 1) ? 1 : MPI_GROUP_EMPTY;
  return MPI_SUCCESS;
}
int MPI_Group_free(MPI_Group* group) {
  *group = MPI_GROUP_NULL;
  return MPI_SUCCESS;
}

int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype,
                    MPI_Datatype* newtype) {
  vector_types.push_back({count, blocklength, stride, oldtype});
  *newtype = 1000 + static_cast<int>(vector_types.size()) - 1;
  return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype* datatype) {
  *datatype = MPI_DATATYPE_NULL;
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype,
                  MPI_Op, MPI_Comm) {
  if (sendbuf != recvbuf) {
    std::memmove(recvbuf, sendbuf, static_cast<size_t>(count) * elemSize(datatype));
  }
  return MPI_SUCCESS;
}
int MPI_Scan(const void* sendbuf, void* recvbuf, int count, MPI_Datatype datatype,
             MPI_Op op, MPI_Comm comm) {
  return MPI_Allreduce(sendbuf, recvbuf, count, datatype, op, comm);
}
int MPI_Gatherv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, void* recvbuf,
                const int*, const int* displs, MPI_Datatype recvtype, int, MPI_Comm) {
  char* dst = static_cast<char*>(recvbuf)
              + static_cast<size_t>(displs != nullptr ? displs[0] : 0) * elemSize(recvtype);
  std::memmove(dst, sendbuf, static_cast<size_t>(sendcount) * elemSize(sendtype));
  return MPI_SUCCESS;
}
int MPI_Scatterv(const void* sendbuf, const int*, const int* displs, MPI_Datatype sendtype,
                 void* recvbuf, int recvcount, MPI_Datatype recvtype, int, MPI_Comm) {
  const char* src = static_cast<const char*>(sendbuf)
                    + static_cast<size_t>(displs != nullptr ? displs[0] : 0) * elemSize(sendtype);
  std::memmove(recvbuf, src, static_cast<size_t>(recvcount) * elemSize(recvtype));
  return MPI_SUCCESS;
}

int MPI_Send(const void* buf, int count, MPI_Datatype datatype, int, int tag, MPI_Comm comm) {
  Message m;
  m.comm = comm;
  m.tag = tag;
  packBytes(buf, count, datatype, m.bytes);
  mailbox.push_back(std::move(m));
  return MPI_SUCCESS;
}
int MPI_Isend(const void* buf, int count, MPI_Datatype datatype, int dest, int tag,
              MPI_Comm comm, MPI_Request* request) {
  MPI_Send(buf, count, datatype, dest, tag, comm);
  const int r = newRequest();
  requests[r].active = true;
  requests[r].is_recv = false;
  *request = r;
  return MPI_SUCCESS;
}
int MPI_Irecv(void* buf, int count, MPI_Datatype datatype, int, int tag, MPI_Comm comm,
              MPI_Request* request) {
  const int r = newRequest();
  requests[r].active = true;
  requests[r].is_recv = true;
  requests[r].comm = comm;
  requests[r].tag = tag;
  recv_desc[r] = {buf, count, datatype};
  *request = r;
  return MPI_SUCCESS;
}
int MPI_Recv(void* buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm,
             MPI_Status* status) {
  MPI_Request r;
  MPI_Irecv(buf, count, datatype, source, tag, comm, &r);
  return MPI_Wait(&r, status);
}
int MPI_Sendrecv(const void* sendbuf, int sendcount, MPI_Datatype sendtype, int dest,
                 int sendtag, void* recvbuf, int recvcount, MPI_Datatype recvtype, int source,
                 int recvtag, MPI_Comm comm, MPI_Status* status) {
  MPI_Send(sendbuf, sendcount, sendtype, dest, sendtag, comm);
  return MPI_Recv(recvbuf, recvcount, recvtype, source, recvtag, comm, status);
}
int MPI_Wait(MPI_Request* request, MPI_Status* status) {
  if (*request == MPI_REQUEST_NULL) {
    return MPI_SUCCESS;
  }
  if (!tryComplete(*request, status)) {
    deadlock("MPI_Wait");
  }
  *request = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}
int MPI_Waitall(int count, MPI_Request array_of_requests[], MPI_Status array_of_statuses[]) {
  for (int i = 0; i < count; ++i) {
    MPI_Wait(&array_of_requests[i],
             array_of_statuses != nullptr ? &array_of_statuses[i] : nullptr);
  }
  return MPI_SUCCESS;
}
int MPI_Waitany(int count, MPI_Request array_of_requests[], int* indx, MPI_Status* status) {
  bool any_active = false;
  for (int i = 0; i < count; ++i) {
    if (array_of_requests[i] == MPI_REQUEST_NULL) {
      continue;
    }
    any_active = true;
    if (tryComplete(array_of_requests[i], status)) {
      array_of_requests[i] = MPI_REQUEST_NULL;
      *indx = i;
      return MPI_SUCCESS;
    }
  }
  if (any_active) {
    deadlock("MPI_Waitany");
  }
  *indx = MPI_UNDEFINED;
  return MPI_SUCCESS;
}
} // extern "C"

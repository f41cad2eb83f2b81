// NCCL is bound at run time (dlopen) so that libdevqa.so loads on hosts without NCCL and shares
// the copy a host application (e.g. PyTorch) has already loaded instead of bringing a second one.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace devqa {

struct ncclUniqueId_ { char internal[128]; };   // layout of ncclUniqueId (nccl.h)

struct NcclApi {
  int (*GetVersion)(int*);
  int (*GetUniqueId)(ncclUniqueId_*);
  int (*CommInitRank)(void** comm, int nranks, ncclUniqueId_ id, int rank);
  int (*CommDestroy)(void* comm);
  int (*AllGather)(const void* send, void* recv, size_t count, int dtype, void* comm, cudaStream_t stream);
  const char* (*GetErrorString)(int);
};

// nullptr (and devqa_last_error set) when NCCL cannot be found
const NcclApi* nccl_api();

}  // namespace devqa

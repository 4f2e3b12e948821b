// Host-side helpers shared by the C-ABI translation units: thread-local error string and a lazily dlopen()ed NCCL
// (so the library loads on hosts without libnccl and uses the copy already mapped by the process, e.g. PyTorch's).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>

namespace dmv {
int set_error(int code, const char* fmt, ...);
int nccl_unique_id(void* id128);
int nccl_init(void** comm, int nranks, int rank, const void* id128);
int nccl_allreduce_double(void* comm, double* buf, size_t count, cudaStream_t s);
void nccl_destroy(void* comm);
}  // namespace dmv

// liger_b200 -- NCCL loaded at run time (dlopen), so single-GPU use has no NCCL dependency.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace lgr {
struct NcclComm;  // opaque
void nccl_unique_id(void* out128);
NcclComm* nccl_init(const void* id128, int rank, int world);
void nccl_allreduce_sum_f32(NcclComm* c, float* buf, size_t n, cudaStream_t st);
void nccl_allreduce_sum_f64(NcclComm* c, double* buf, size_t n, cudaStream_t st);
void nccl_group_start();
void nccl_group_end();
void nccl_destroy(NcclComm* c);
}  // namespace lgr

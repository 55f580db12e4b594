// Kernel-side view of a2f_gemm_args shared by the SIMT and tcgen05 GEMM back ends.
#pragma once
#include "a2f_common.cuh"

namespace a2f {

struct GemmParams {
    int M, N, K;
    const void* A;
    long long a_row_stride, a_batch_stride;
    int rows_per_batch;
    const void* W;
    long long ldw;
    const float* bias;
    int act;
    const void* resid;
    int resid_bf16;
    long long ldr;
    const float* tmpl;
    int rows_per_tmpl;
    void* C;
    long long ldc;
    long long c_batch_stride;   // elements between the output blocks of consecutive batches (rows_per_batch*ldc when dense)
};

int gemm_simt(const GemmParams& p, int a_bf16, int c_bf16, cudaStream_t s);
int posconv_simt(const GemmParams& p, int a_bf16, int c_bf16, cudaStream_t s);
// tcgen05 back end (bf16 operands).  mode 0: plain / strided-row implicit GEMM, mode 2: positional conv.
int gemm_tc(const GemmParams& p, int c_bf16, int mode, cudaStream_t s);

}  // namespace a2f

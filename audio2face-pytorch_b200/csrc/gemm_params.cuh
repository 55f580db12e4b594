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
    // backward-pass extensions (a2f.h): explicit K segments with row / column offsets, zero rows outside [0,a_rows),
    // batch stride of resid, resid_mode 1 = multiply by act'(resid)
    int a_rows = 0;               // 0 = rows_per_batch (normalised by normalize_gemm)
    int n_seg = 0;
    int seg_row_off[4] = {0, 0, 0, 0};
    int seg_col_off[4] = {0, 0, 0, 0};
    long long r_batch_stride = 0; // 0 = rows_per_batch * ldr
    int resid_mode = 0;
    // SIMT back end only: K range split over gridDim.z (fp32 atomicAdd into a zeroed C), set by gemm_simt itself
    int split_k = 1;
    int k_per_split = 0;
    // tcgen05 CTA-pair kernel: second output C2 = act(z) while C keeps the pre-activation z (a2f.h a2f_gemm_args::C2)
    void* C2 = nullptr;
    long long ldc2 = 0;
    // tcgen05 vertex head with the reconstruction / velocity loss fused into its epilogue (a2f_vertex_head_loss): gt [M,N]
    // fp32; the epilogue accumulates sum (y-g)^2 and sum ((y1-y0)-(g1-g0))^2 over row pairs (2k,2k+1) into per-warp fp64
    // partials, writes dL/dy as bf16 [M, ld_dy] and stores y itself only when C != NULL
    const float* loss_gt = nullptr;
    void* loss_dy = nullptr;
    long long ld_dy = 0;
    float c_rec = 0.f, c_vel = 0.f;
    double* loss_partial = nullptr;   // [grid][8 epilogue warps][2]
    // tcgen05 vertex head that follows the decoder rollout (a2f_vertex_head_stream).  A is "frame-major": row
    // r = frame * perm_rows + utterance; a batch (rows_per_batch = 128 rows) is a group of 128 / perm_rows frames.  The
    // wide scalar epilogue sends row r of batch b to C + b * c_batch_stride + (r / perm_rows) * perm_stride +
    // (r % perm_rows) * ldc and adds template row r % perm_rows; rows at or beyond live_rows are not stored.  The producer
    // of a tile of batch b waits until wait_counters[min(wait_n, (b + 1) * wait_per_batch) - 1] >= wait_target (acquire).
    int perm_rows = 0;
    long long perm_stride = 0;
    long long live_rows = 0;
    const unsigned* wait_counters = nullptr;
    int wait_target = 0, wait_per_batch = 0, wait_n = 0;
    int max_ctas = 0;                 // 0 = one CTA per SM; else the grid is capped (SMs left to the kernel that is waited for)
};

inline void normalize_gemm(GemmParams& p) {
    if (p.a_rows <= 0) p.a_rows = p.rows_per_batch;
    if (p.r_batch_stride <= 0) p.r_batch_stride = (long long)p.rows_per_batch * p.ldr;
}

struct WgradParams {
    int M, N, K;
    const void* dY;
    long long dy_row_stride, dy_batch_stride;
    const void* X;
    long long x_row_stride, x_batch_stride;
    int rows_per_batch, x_rows;
    int n_seg;
    int x_row_off[4];
    int x_col_off[4];
    float* dW;
    long long ldw;
    int x_row_step;
    A2F_HD int row_off(int s) const { return x_row_step != 0 ? x_row_off[0] + s * x_row_step : x_row_off[s & 3]; }
    A2F_HD int col_off(int s) const { return x_row_step != 0 ? x_col_off[0] : x_col_off[s & 3]; }
};

int gemm_simt(const GemmParams& p, int a_bf16, int c_bf16, cudaStream_t s);
int posconv_simt(const GemmParams& p, int a_bf16, int c_bf16, cudaStream_t s);
// tcgen05 back end (bf16 operands).  mode 0: plain / strided-row implicit GEMM, mode 2: positional conv.
int gemm_tc(const GemmParams& p, int c_bf16, int mode, cudaStream_t s);
// dedicated positional-conv kernel (posconv_tc.cu): bf16 channels-last [B,T,768] in / out, chunked weight layout
int posconv_tc(const void* A, const void* W, const float* bias, const void* resid, int resid_mode, void* C, int B, int T,
               int row_off, int act, cudaStream_t s);
// element index of (group, n, tap, c) in the chunked layout [16][128 taps][6 chunks][48 n][8]
A2F_HD long long posconv_chunked_index(int grp, int n, int tap, int c) {
    return ((((long long)grp * 128 + tap) * 6 + (c >> 3)) * 48 + n) * 8 + (c & 7);
}
int wgrad_simt(const WgradParams& p, int bf16_in, cudaStream_t s);
int wgrad_tc(const WgradParams& p, cudaStream_t s);

// derivative of the forward activation at pre-activation z (resid_mode A2F_RESID_DACT)
A2F_D float act_grad(float z, int act) {
    switch (act) {
        case A2F_ACT_RELU: return z > 0.f ? 1.f : 0.f;
        case A2F_ACT_GELU: {
            const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
            return cdf + z * 0.39894228040143267794f * __expf(-0.5f * z * z);
        }
        case A2F_ACT_TANH: {
            const float t = tanhf(z);
            return 1.f - t * t;
        }
        default: return 1.f;
    }
}

// GELU'(z) = Phi(z) + z phi(z) for the bf16 tensor-core path: Phi through the tanh form on MUFU.TANH (|error| <= 5e-4,
// the same approximation the forward epilogues use), phi through MUFU.EX2 -- ~12 instructions instead of erff + expf
// (~45), in an epilogue that is instruction-issue bound (the data-gradient GEMM of FFN2 evaluates it 7.4 M times).
A2F_D float gelu_grad_fast(float z) {
    const float z2 = z * z;
    const float u = z * fmaf(0.0356774081f, z2, 0.7978845608f);
    float t, e;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.72134752044f * z2));     // exp(-z^2/2)
    return fmaf(z * 0.39894228040f, e, fmaf(0.5f, t, 0.5f));
}

}  // namespace a2f

// True-fp32 SIMT GEMM with the fused epilogue of a2f_gemm (bias, activation, residual, template add).
// This is the 1e-5 parity path (TF32/bf16 tensor cores cannot meet 1e-5); the tensor-core
// path lives in gemm_tc.cu.  64x64x16 tiles, 256 threads, 4x4 register micro-tiles, fp32 FMA accumulation in
// ascending-k order.
#include "a2f_common.cuh"
#include "gemm_params.cuh"

namespace a2f {

constexpr int SBM = 64, SBN = 64, SBK = 16;

// MODE 0: A rows addressed by (batch, row) strides, K contiguous.
// MODE 2: positional conv: k = tap*48 + c -> h[b, t + tap - 64, g*48 + c], zero outside [0,T).
template <typename TA, typename TC, int MODE>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmParams p) {
    __shared__ float As[SBK][SBM + 4];
    __shared__ float Bs[SBK][SBN + 4];

    const TA* __restrict__ A = static_cast<const TA*>(p.A);
    const TA* __restrict__ W = static_cast<const TA*>(p.W);
    const float* __restrict__ bias = p.bias;
    const int g = (MODE == 2) ? blockIdx.z : 0;
    int col_off = 0;
    if (MODE == 2) {
        W += (long long)g * p.N * p.ldw;
        col_off = g * p.N;
    }

    const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int lrow = tid / 4, lk = (tid % 4) * 4;

    // per-thread A row base (row index fixed across the k loop)
    const int am = m0 + lrow;
    const bool a_row_ok = am < p.M;
    long long a_base = 0;
    int a_t = 0;
    if (a_row_ok) {
        int b = am / p.rows_per_batch, r = am % p.rows_per_batch;
        if (MODE == 0) a_base = (long long)b * p.a_batch_stride + (long long)r * p.a_row_stride;
        else {
            a_base = (long long)b * p.rows_per_batch * p.a_row_stride;
            a_t = r;
        }
    }
    const int wn = n0 + lrow;
    const bool w_row_ok = wn < p.N;
    const long long w_base = (long long)wn * p.ldw;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += SBK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + lk + j;
            float av = 0.f, wv = 0.f;
            if (k < p.K) {
                if (a_row_ok) {
                    if (MODE == 0) av = ld_as_float(A + a_base + k);
                    else {
                        int tap = k / 48, c = k - tap * 48;
                        int t = a_t + tap - 64;
                        if (t >= 0 && t < p.rows_per_batch)
                            av = ld_as_float(A + a_base + (long long)t * p.a_row_stride + g * 48 + c);
                    }
                }
                if (w_row_ok) wv = ld_as_float(W + w_base + k);
            }
            As[lk + j][lrow] = av;
            Bs[lk + j][lrow] = wv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SBK; ++kk) {
            float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    TC* __restrict__ C = static_cast<TC*>(p.C);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            int nc = n + col_off;
            float v = acc[i][j];
            if (bias) v += bias[nc];
            v = apply_act_rt(v, p.act);
            if (p.resid) {
                long long ri = (long long)m * p.ldr + nc;
                v += p.resid_bf16 ? __bfloat162float(static_cast<const bf16*>(p.resid)[ri])
                                  : static_cast<const float*>(p.resid)[ri];
            }
            if (p.tmpl) v += p.tmpl[(long long)(m / p.rows_per_tmpl) * p.N + n];
            const long long crow = (long long)(m / p.rows_per_batch) * p.c_batch_stride + (long long)(m % p.rows_per_batch) * p.ldc;
            st_from_float(C + crow + nc, v);
        }
    }
}

template <int MODE> static int launch_simt(const GemmParams& p, int a_bf16, int c_bf16, int groups, cudaStream_t s) {
    dim3 grid((p.N + SBN - 1) / SBN, (p.M + SBM - 1) / SBM, groups);
    if (!a_bf16 && !c_bf16) gemm_simt_kernel<float, float, MODE><<<grid, 256, 0, s>>>(p);
    else if (!a_bf16 && c_bf16) gemm_simt_kernel<float, bf16, MODE><<<grid, 256, 0, s>>>(p);
    else if (a_bf16 && !c_bf16) gemm_simt_kernel<bf16, float, MODE><<<grid, 256, 0, s>>>(p);
    else gemm_simt_kernel<bf16, bf16, MODE><<<grid, 256, 0, s>>>(p);
    A2F_CHECK_LAUNCH("gemm_simt_kernel");
    count_launch();
    return A2F_OK;
}

int gemm_simt(const GemmParams& p, int a_bf16, int c_bf16, cudaStream_t s) {
    if (p.M <= 0 || p.N <= 0) return A2F_OK;
    return launch_simt<0>(p, a_bf16, c_bf16, 1, s);
}
int posconv_simt(const GemmParams& p, int a_bf16, int c_bf16, cudaStream_t s) {
    return launch_simt<2>(p, a_bf16, c_bf16, 16, s);
}

}  // namespace a2f

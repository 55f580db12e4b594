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
    pdl_sync();   // PDL: wait for the previous kernel's results, let the next kernel's prologue start
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
    long long a_base = 0, a_bbase = 0;
    int a_t = 0;
    bool a_in_range = a_row_ok;
    const int kseg = p.n_seg > 0 ? p.K / p.n_seg : p.K;
    if (a_row_ok) {
        int b = am / p.rows_per_batch, r = am % p.rows_per_batch;
        a_t = r;
        if (MODE == 0) {
            a_bbase = (long long)b * p.a_batch_stride;
            a_base = a_bbase + (long long)r * p.a_row_stride;
            a_in_range = r < p.a_rows;
        } else {
            a_base = (long long)b * p.rows_per_batch * p.a_row_stride;
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

    // operands of one k-block for this thread (4 consecutive k of one A row and one W row)
    const int k_lim = (MODE == 0 && p.split_k > 1) ? min(p.K, ((int)blockIdx.z + 1) * p.k_per_split) : p.K;
    auto fetch = [&](int k0, float* av, float* wv) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + lk + j;
            av[j] = 0.f;
            wv[j] = 0.f;
            if (k < k_lim) {
                if (a_row_ok) {
                    if (MODE == 0) {
                        if (p.n_seg > 0) {
                            const int sg = k / kseg, kk = k - sg * kseg;
                            const int row = a_t + p.seg_row_off[sg];
                            if (row >= 0 && row < p.a_rows)
                                av[j] = ld_as_float(A + a_bbase + (long long)row * p.a_row_stride + p.seg_col_off[sg] + kk);
                        } else if (a_in_range) {
                            av[j] = ld_as_float(A + a_base + k);
                        }
                    } else {
                        int tap = k / 48, c = k - tap * 48;
                        int t = a_t + tap - 64 + p.seg_row_off[0];
                        if (t >= 0 && t < p.rows_per_batch)
                            av[j] = ld_as_float(A + a_base + (long long)t * p.a_row_stride + g * 48 + c);
                    }
                }
                if (w_row_ok) wv[j] = ld_as_float(W + w_base + k);
            }
        }
    };

    // software pipeline: the global loads of k-block i+1 are in flight while k-block i is multiplied (small grids run
    // one CTA per SM, where nothing else hides the load latency -- Audio2Mesh at 64 windows, the 64-wide decoder GEMMs)
    // split-K (MODE 0 only): this CTA's slice of the contraction
    const int k_begin = (MODE == 0 && p.split_k > 1) ? (int)blockIdx.z * p.k_per_split : 0;
    const int k_end = (MODE == 0 && p.split_k > 1) ? min(p.K, k_begin + p.k_per_split) : p.K;
    float av[4], wv[4];
    fetch(k_begin, av, wv);
    for (int k0 = k_begin; k0 < k_end; k0 += SBK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[lk + j][lrow] = av[j];
            Bs[lk + j][lrow] = wv[j];
        }
        __syncthreads();
        if (k0 + SBK < k_end) fetch(k0 + SBK, av, wv);
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
            if (MODE == 0 && p.split_k > 1) {
                // partial sum of one K slice: accumulate into the zeroed output (bias once, from slice 0)
                if (bias && blockIdx.z == 0) v += bias[nc];
                const long long crow_s = (long long)(m / p.rows_per_batch) * p.c_batch_stride + (long long)(m % p.rows_per_batch) * p.ldc;
                atomicAdd(reinterpret_cast<float*>(C) + crow_s + nc, v);
                continue;
            }
            if (p.resid_mode == A2F_RESID_DACT) {
                const long long ri = (long long)(m / p.rows_per_batch) * p.r_batch_stride +
                                     (long long)(m % p.rows_per_batch) * p.ldr + nc;
                const float z = p.resid_bf16 ? __bfloat162float(static_cast<const bf16*>(p.resid)[ri])
                                             : static_cast<const float*>(p.resid)[ri];
                v *= act_grad(z, p.act);
            } else {
                if (bias) v += bias[nc];
                v = apply_act_rt(v, p.act);
                if (p.resid) {
                    const long long ri = (long long)(m / p.rows_per_batch) * p.r_batch_stride +
                                         (long long)(m % p.rows_per_batch) * p.ldr + nc;
                    v += p.resid_bf16 ? __bfloat162float(static_cast<const bf16*>(p.resid)[ri])
                                      : static_cast<const float*>(p.resid)[ri];
                }
            }
            if (p.tmpl) v += p.tmpl[(long long)(m / p.rows_per_tmpl) * p.N + n];
            const long long crow = (long long)(m / p.rows_per_batch) * p.c_batch_stride + (long long)(m % p.rows_per_batch) * p.ldc;
            st_from_float(C + crow + nc, v);
        }
    }
}

template <int MODE> static int launch_simt(const GemmParams& p, int a_bf16, int c_bf16, int groups, cudaStream_t s) {
    dim3 grid((p.N + SBN - 1) / SBN, (p.M + SBM - 1) / SBM, groups);
    if (!a_bf16 && !c_bf16) A2F_CHECK_CUDA(launch_pdl((gemm_simt_kernel<float, float, MODE>), dim3(grid), dim3(256), 0, s, p));
    else if (!a_bf16 && c_bf16) A2F_CHECK_CUDA(launch_pdl((gemm_simt_kernel<float, bf16, MODE>), dim3(grid), dim3(256), 0, s, p));
    else if (a_bf16 && !c_bf16) A2F_CHECK_CUDA(launch_pdl((gemm_simt_kernel<bf16, float, MODE>), dim3(grid), dim3(256), 0, s, p));
    else A2F_CHECK_CUDA(launch_pdl((gemm_simt_kernel<bf16, bf16, MODE>), dim3(grid), dim3(256), 0, s, p));
    A2F_CHECK_LAUNCH("gemm_simt_kernel");
    count_launch();
    return A2F_OK;
}

int gemm_simt(const GemmParams& p_in, int a_bf16, int c_bf16, cudaStream_t s) {
    if (p_in.M <= 0 || p_in.N <= 0) return A2F_OK;
    GemmParams p = p_in;
    normalize_gemm(p);
    // split-K: a long contraction behind a handful of output tiles (the vertex-head data gradient of the conv models is
    // M=128, N=50, K=15069 -- two CTAs walking 942 k-blocks, 0.77 ms) is spread over gridDim.z; plain linear epilogues only
    const long long tiles = (long long)((p.N + SBN - 1) / SBN) * ((p.M + SBM - 1) / SBM);
    if (!c_bf16 && p.K >= 2048 && tiles * 4 <= sm_count() && p.act == A2F_ACT_NONE && p.resid == nullptr && p.tmpl == nullptr &&
        p.n_seg == 0 && p.rows_per_batch == p.M) {
        int splits = (int)(2LL * sm_count() / tiles);
        const int max_splits = p.K / 256;
        if (splits > max_splits) splits = max_splits;
        if (splits > 1) {
            p.k_per_split = ((p.K + splits - 1) / splits + SBK - 1) / SBK * SBK;
            p.split_k = (p.K + p.k_per_split - 1) / p.k_per_split;
            A2F_CHECK_CUDA(cudaMemset2DAsync(p.C, (size_t)p.ldc * sizeof(float), 0, (size_t)p.N * sizeof(float), (size_t)p.M, s));
            return launch_simt<0>(p, a_bf16, c_bf16, p.split_k, s);
        }
    }
    return launch_simt<0>(p, a_bf16, c_bf16, 1, s);
}
int posconv_simt(const GemmParams& p_in, int a_bf16, int c_bf16, cudaStream_t s) {
    GemmParams p = p_in;
    normalize_gemm(p);
    return launch_simt<2>(p, a_bf16, c_bf16, 16, s);
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[n, s*K + k] += sum_m dY[m, n] * X[(b, r + roff[s]), coff[s] + k].  64x64 output tiles, the M reduction split over
// blockIdx.z, fp32 atomics into dW (the fp32 parity path; ordering noise ~1e-7 relative).
// Reduction rows per CTA are chosen by the launcher so that small outputs (the decoder's 64x64 matrices over a few
// thousand rows) still fill the machine: a fixed 512-row slice left them on 5 CTAs (95 us for 20 MFLOP).
template <typename TI>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradParams p, int WG_ROWS) {
    __shared__ float As[SBK][SBM + 4];   // [m][n]
    __shared__ float Bs[SBK][SBN + 4];   // [m][kcol]
    const TI* __restrict__ dY = static_cast<const TI*>(p.dY);
    const TI* __restrict__ X = static_cast<const TI*>(p.X);
    const int n0 = blockIdx.y * SBM, c0 = blockIdx.x * SBN;
    const int ktot = p.K * p.n_seg;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int lcol = tid % 64, lm = tid / 64;      // 4 m-rows per pass, 64 columns
    // column of this thread's X loads: segment / offset are fixed across the reduction
    const int xc = c0 + lcol;
    const bool xc_ok = xc < ktot;
    const int sg = xc_ok ? xc / p.K : 0;
    const int xk = xc_ok ? xc - sg * p.K : 0;
    const int roff = p.row_off(sg), coff = p.col_off(sg);
    const int yn = n0 + lcol;
    const bool yn_ok = yn < p.N;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const long long m_begin = (long long)blockIdx.z * WG_ROWS;
    const long long m_end = m_begin + WG_ROWS < p.M ? m_begin + WG_ROWS : p.M;
    for (long long mb = m_begin; mb < m_end; mb += SBK) {
#pragma unroll
        for (int q = 0; q < SBK / 4; ++q) {
            const int kk = q * 4 + lm;
            const long long m = mb + kk;
            float yv = 0.f, xv = 0.f;
            if (m < m_end) {
                const int b = (int)(m / p.rows_per_batch), r = (int)(m % p.rows_per_batch);
                if (yn_ok) yv = ld_as_float(dY + (long long)b * p.dy_batch_stride + (long long)r * p.dy_row_stride + yn);
                const int xr = r + roff;
                if (xc_ok && xr >= 0 && xr < p.x_rows)
                    xv = ld_as_float(X + (long long)b * p.x_batch_stride + (long long)xr * p.x_row_stride + coff + xk);
            }
            As[kk][lcol] = yv;
            Bs[kk][lcol] = xv;
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
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= p.N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + tx * 4 + j;
            if (c < ktot) atomicAdd(p.dW + (long long)n * p.ldw + c, acc[i][j]);
        }
    }
}

int wgrad_simt(const WgradParams& p, int bf16_in, cudaStream_t s) {
    const int ktot = p.K * p.n_seg;
    const int gx = (ktot + SBN - 1) / SBN, gy = (p.N + SBM - 1) / SBM;
    const int want_z = (2 * sm_count() + gx * gy - 1) / (gx * gy);            // ~2 CTAs per SM overall
    long long rows = (p.M + want_z - 1) / want_z;
    rows = (rows + SBK - 1) / SBK * SBK;
    if (rows < 2 * SBK) rows = 2 * SBK;
    if (rows > 512) rows = 512;
    dim3 grid(gx, gy, (unsigned)((p.M + rows - 1) / rows));
    if (bf16_in) wgrad_simt_kernel<bf16><<<grid, 256, 0, s>>>(p, (int)rows);
    else wgrad_simt_kernel<float><<<grid, 256, 0, s>>>(p, (int)rows);
    A2F_CHECK_LAUNCH("wgrad_simt_kernel");
    count_launch();
    return A2F_OK;
}

}  // namespace a2f

// VOCA trunk (ref:src/model/voca.py:19-46) fused into one kernel: one-hot tiling, 4 x (Conv2d(3x1,s2,p1)+ReLU),
// concat(one_hot[:8]), Linear 72->72, Linear 72->128, tanh, Linear 128->50.  The last Linear 50->15069 + template add
// is the shared vertex-head GEMM (a2f_gemm with tmpl), which is where VOCA's HBM bytes are.
//
// Each CTA keeps WPB windows in shared memory and walks the layers; a thread owns one output unit and reuses every
// weight it loads for all WPB windows.  Accumulation order per output: bias first, then (ci, kh) ascending.
#include "a2f_common.cuh"

namespace a2f {

constexpr int WPB = 8;   // windows per CTA iteration

struct VocaW {
    const float* cw[4];
    const float* cb[4];
    const float* fw[3];
    const float* fb[3];
};

template <int CIN, int HIN, int COUT>
__device__ __forceinline__ void voca_conv(const float* __restrict__ in /*[WPB][CIN][HIN]*/,
                                          float* __restrict__ out /*[WPB][out_stride]*/, int out_stride,
                                          const float* __restrict__ w, const float* __restrict__ b) {
    constexpr int HOUT = HIN / 2;
    for (int o = threadIdx.x; o < COUT * HOUT; o += blockDim.x) {
        const int co = o / HOUT, ho = o % HOUT;
        float acc[WPB];
        const float bv = __ldg(b + co);
#pragma unroll
        for (int i = 0; i < WPB; ++i) acc[i] = bv;
        for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hi = 2 * ho + kh - 1;
                if (hi < 0 || hi >= HIN) continue;
                const float wv = __ldg(w + (co * CIN + ci) * 3 + kh);
#pragma unroll
                for (int i = 0; i < WPB; ++i) acc[i] = fmaf(wv, in[(i * CIN + ci) * HIN + hi], acc[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < WPB; ++i) out[i * out_stride + co * HOUT + ho] = relu(acc[i]);
    }
}

template <int K, int N, int ACT>
__device__ __forceinline__ void voca_fc(const float* __restrict__ in, int in_stride, float* __restrict__ out,
                                        int out_stride, const float* __restrict__ w, const float* __restrict__ b) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float acc[WPB];
        const float bv = __ldg(b + n);
#pragma unroll
        for (int i = 0; i < WPB; ++i) acc[i] = bv;
        for (int k = 0; k < K; ++k) {
            const float wv = __ldg(w + n * K + k);
#pragma unroll
            for (int i = 0; i < WPB; ++i) acc[i] = fmaf(wv, in[i * in_stride + k], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < WPB; ++i) out[i * out_stride + n] = apply_act<ACT>(acc[i]);
    }
}

template <typename TZ>
__global__ void __launch_bounds__(256) voca_trunk_kernel(VocaW w, const float* __restrict__ x,
                                                         const float* __restrict__ one_hot, int n_onehot,
                                                         TZ* __restrict__ z, int ldz, int B) {
    __shared__ float s_in[WPB * 37 * 16];
    __shared__ float s_a1[WPB * 32 * 8];
    __shared__ float s_a2[WPB * 32 * 4];
    __shared__ float s_a3[WPB * 64 * 2];
    __shared__ float s_a4[WPB * 72];
    __shared__ float s_f1[WPB * 72];
    __shared__ float s_f2[WPB * 128];
    __shared__ float s_z[WPB * 50];

    for (int w0 = blockIdx.x * WPB; w0 < B; w0 += gridDim.x * WPB) {
        // input assembly: channels 0..28 = features, 29..36 = tiled one-hot: emb[r][c] = oh8[(16 r + c) % 8]
        for (int i = threadIdx.x; i < WPB * 37 * 16; i += blockDim.x) {
            const int wi = i / (37 * 16), rem = i % (37 * 16), ch = rem / 16, h = rem % 16;
            const int bw = w0 + wi;
            float v = 0.f;
            if (bw < B) {
                if (ch < 29) v = x[((long long)bw * 29 + ch) * 16 + h];
                else v = one_hot[(long long)bw * n_onehot + ((16 * (ch - 29) + h) % 8)];
            }
            s_in[i] = v;
        }
        for (int i = threadIdx.x; i < WPB * 8; i += blockDim.x) {
            const int wi = i / 8, j = i % 8, bw = w0 + wi;
            s_a4[wi * 72 + 64 + j] = (bw < B) ? one_hot[(long long)bw * n_onehot + j] : 0.f;
        }
        __syncthreads();
        voca_conv<37, 16, 32>(s_in, s_a1, 32 * 8, w.cw[0], w.cb[0]);
        __syncthreads();
        voca_conv<32, 8, 32>(s_a1, s_a2, 32 * 4, w.cw[1], w.cb[1]);
        __syncthreads();
        voca_conv<32, 4, 64>(s_a2, s_a3, 64 * 2, w.cw[2], w.cb[2]);
        __syncthreads();
        voca_conv<64, 2, 64>(s_a3, s_a4, 72, w.cw[3], w.cb[3]);   // -> first 64 of the 72-vector
        __syncthreads();
        voca_fc<72, 72, A2F_ACT_NONE>(s_a4, 72, s_f1, 72, w.fw[0], w.fb[0]);
        __syncthreads();
        voca_fc<72, 128, A2F_ACT_TANH>(s_f1, 72, s_f2, 128, w.fw[1], w.fb[1]);
        __syncthreads();
        voca_fc<128, 50, A2F_ACT_NONE>(s_f2, 128, s_z, 50, w.fw[2], w.fb[2]);
        __syncthreads();
        for (int i = threadIdx.x; i < WPB * ldz; i += blockDim.x) {
            const int wi = i / ldz, j = i % ldz, bw = w0 + wi;
            if (bw < B) st_from_float(z + (long long)bw * ldz + j, j < 50 ? s_z[wi * 50 + j] : 0.f);
        }
        __syncthreads();
    }
}

}  // namespace a2f

using namespace a2f;

extern "C" int a2f_voca_trunk(const a2f_voca_weights* w, const float* x, const float* one_hot, int n_onehot, void* z,
                              int z_dtype, int ldz, int B, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(w && x && one_hot && z, "a2f_voca_trunk: NULL argument");
    A2F_REQUIRE(n_onehot >= 8 && ldz >= 50, "a2f_voca_trunk: need n_onehot >= 8 and ldz >= 50");
    if (B <= 0) return A2F_OK;
    VocaW vw;
    for (int i = 0; i < 4; ++i) {
        A2F_REQUIRE(w->conv_w[i] && w->conv_b[i], "a2f_voca_trunk: NULL conv weight");
        vw.cw[i] = w->conv_w[i];
        vw.cb[i] = w->conv_b[i];
    }
    for (int i = 0; i < 3; ++i) {
        A2F_REQUIRE(w->fc_w[i] && w->fc_b[i], "a2f_voca_trunk: NULL fc weight");
        vw.fw[i] = w->fc_w[i];
        vw.fb[i] = w->fc_b[i];
    }
    int grid = (B + WPB - 1) / WPB;
    const int cap = 4 * sm_count();
    if (grid > cap) grid = cap;
    if (z_dtype == A2F_BF16)
        voca_trunk_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(vw, x, one_hot, n_onehot, static_cast<bf16*>(z), ldz, B);
    else
        voca_trunk_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(vw, x, one_hot, n_onehot, static_cast<float*>(z), ldz, B);
    A2F_CHECK_LAUNCH("voca_trunk_kernel");
    count_launch();
    return A2F_OK;
}

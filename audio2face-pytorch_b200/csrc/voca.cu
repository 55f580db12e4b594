// VOCA trunk (ref:src/model/voca.py:19-46) fused into one kernel: one-hot tiling, 4 x (Conv2d(3x1,s2,p1)+ReLU),
// concat(one_hot[:8]), Linear 72->72, Linear 72->128, tanh, Linear 128->50.  The last Linear 50->15069 + template add
// is the shared vertex-head GEMM (a2f_gemm with tmpl), which is where VOCA's HBM bytes are.
//
// Each CTA keeps WPB windows in shared memory and walks the layers.  Activations are stored unit-major with the WPB
// windows of a unit contiguous ([unit][WPB]), so a thread that owns one output unit reads the inputs of all its windows
// as broadcast 16-byte loads; every layer's weights are transposed into shared memory first ([k][n], pitch n+1), so the
// weight reads of a warp are consecutive.  Layers with fewer outputs than threads split the windows over 2 or 4 threads
// per output.  Accumulation order per output: bias first, then (ci, kh) ascending (same as the oracle's loops).
#include "a2f_common.cuh"

namespace a2f {

constexpr int WPB = 16;              // windows per CTA iteration
constexpr int VT = 256;              // threads per CTA
constexpr int VOCA_W_FLOATS = 192 * 65;          // largest transposed weight block (conv 64->64: K = 192, N = 64)
constexpr int VOCA_A_FLOATS = 37 * 16 * WPB;     // buffer A: input (592 units), later a2 / a4 / f2
constexpr int VOCA_B_FLOATS = 32 * 8 * WPB;      // buffer B: a1 (256 units), later a3 / f1 / z
constexpr size_t VOCA_SMEM = (size_t)(VOCA_W_FLOATS + VOCA_A_FLOATS + VOCA_B_FLOATS) * sizeof(float);

struct VocaW {
    const float* cw[4];
    const float* cb[4];
    const float* fw[3];
    const float* fb[3];
};

// w [N][K] (row-major, K = ci*3+kh for the convs) -> sw [K][N+1]
template <int K, int N>
__device__ __forceinline__ void voca_stage_w(const float* __restrict__ w, float* __restrict__ sw) {
    constexpr int UB = 8;                      // loads in flight per thread (the staging is L2-latency bound)
#pragma unroll 1
    for (int e0 = threadIdx.x; e0 < N * K; e0 += VT * UB) {
        float v[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) v[u] = (e0 + u * VT < N * K) ? __ldg(w + e0 + u * VT) : 0.f;
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int e = e0 + u * VT;
            if (e < N * K) {
                const int n = e / K, k = e - n * K;
                sw[k * (N + 1) + n] = v[u];
            }
        }
    }
}

// one output unit x WN windows: acc[i] = bias, then += w[k] * in[unit(k)][win0 + i] over the live taps
template <int CIN, int HIN, int COUT, int SPLIT>
__device__ __forceinline__ void voca_conv(const float* __restrict__ in /*[CIN*HIN][WPB]*/,
                                          float* __restrict__ out /*[out_units][WPB]*/,
                                          const float* __restrict__ sw /*[CIN*3][COUT+1]*/, const float* __restrict__ b) {
    constexpr int HOUT = HIN / 2, O = COUT * HOUT, WN = WPB / SPLIT;
    static_assert(O * SPLIT <= VT && WN % 4 == 0, "voca_conv: mapping");
    const int t = threadIdx.x;
    if (t >= O * SPLIT) return;
    const int part = t / O, o = t - part * O;
    const int ho = o / COUT, co = o - ho * COUT;        // consecutive threads: consecutive output channels
    // packed fp32 FMAs (FFMA2: two windows per instruction, same per-accumulator operation order -> identical bits): the
    // kernel is instruction-issue bound (16 FMAs + 5 shared loads per tap and thread), so halving the FMA slots is worth 1.5x
    float2 acc[WN / 2];
    const float bv = __ldg(b + co);
#pragma unroll
    for (int i = 0; i < WN / 2; ++i) acc[i] = make_float2(bv, bv);
    const float* inp = in + part * WN;
#pragma unroll 2
    for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int hi = 2 * ho + kh - 1;
            if (hi < 0 || hi >= HIN) continue;
            const float wv = sw[(ci * 3 + kh) * (COUT + 1) + co];
            const float2 w2 = make_float2(wv, wv);
            const float4* ip = reinterpret_cast<const float4*>(inp + (ci * HIN + hi) * WPB);
#pragma unroll
            for (int i = 0; i < WN / 4; ++i) {
                const float4 v = ip[i];
                acc[2 * i + 0] = ffma2(w2, make_float2(v.x, v.y), acc[2 * i + 0]);
                acc[2 * i + 1] = ffma2(w2, make_float2(v.z, v.w), acc[2 * i + 1]);
            }
        }
    }
    float4* op = reinterpret_cast<float4*>(out + (co * HOUT + ho) * WPB + part * WN);
#pragma unroll
    for (int i = 0; i < WN / 4; ++i)
        op[i] = make_float4(relu(acc[2 * i].x), relu(acc[2 * i].y), relu(acc[2 * i + 1].x), relu(acc[2 * i + 1].y));
}

template <int K, int N, int ACT, int SPLIT>
__device__ __forceinline__ void voca_fc(const float* __restrict__ in /*[K][WPB]*/, float* __restrict__ out /*[N][WPB]*/,
                                        const float* __restrict__ sw /*[K][N+1]*/, const float* __restrict__ b) {
    constexpr int WN = WPB / SPLIT;
    static_assert(N * SPLIT <= VT && WN % 4 == 0, "voca_fc: mapping");
    const int t = threadIdx.x;
    if (t >= N * SPLIT) return;
    const int part = t / N, n = t - part * N;
    float2 acc[WN / 2];
    const float bv = __ldg(b + n);
#pragma unroll
    for (int i = 0; i < WN / 2; ++i) acc[i] = make_float2(bv, bv);
    const float* inp = in + part * WN;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float wv = sw[k * (N + 1) + n];
        const float2 w2 = make_float2(wv, wv);
        const float4* ip = reinterpret_cast<const float4*>(inp + k * WPB);
#pragma unroll
        for (int i = 0; i < WN / 4; ++i) {
            const float4 v = ip[i];
            acc[2 * i + 0] = ffma2(w2, make_float2(v.x, v.y), acc[2 * i + 0]);
            acc[2 * i + 1] = ffma2(w2, make_float2(v.z, v.w), acc[2 * i + 1]);
        }
    }
    float4* op = reinterpret_cast<float4*>(out + n * WPB + part * WN);
#pragma unroll
    for (int i = 0; i < WN / 4; ++i)
        op[i] = make_float4(apply_act<ACT>(acc[2 * i].x), apply_act<ACT>(acc[2 * i].y), apply_act<ACT>(acc[2 * i + 1].x),
                            apply_act<ACT>(acc[2 * i + 1].y));
}

template <typename TZ>
__global__ void __launch_bounds__(VT, 2) voca_trunk_kernel(VocaW w, const float* __restrict__ x,
                                                           const float* __restrict__ one_hot, int n_onehot,
                                                           TZ* __restrict__ z, int ldz, int B) {
    extern __shared__ __align__(16) float vsm[];
    float* sw = vsm;
    float* sA = vsm + VOCA_W_FLOATS;
    float* sB = sA + VOCA_A_FLOATS;

    for (int w0 = blockIdx.x * WPB; w0 < B; w0 += gridDim.x * WPB) {
        // input assembly: channels 0..28 = features, 29..36 = tiled one-hot: emb[r][c] = oh8[(16 r + c) % 8]
        for (int i = threadIdx.x; i < WPB * 29 * 16; i += VT) {
            const int wi = i / (29 * 16), u = i - wi * (29 * 16), bw = w0 + wi;      // coalesced over a window's features
            sA[u * WPB + wi] = (bw < B) ? x[(long long)bw * (29 * 16) + u] : 0.f;
        }
        for (int i = threadIdx.x; i < WPB * 8 * 16; i += VT) {
            const int wi = i % WPB, u = i / WPB, ch = u / 16, h = u % 16, bw = w0 + wi;
            sA[(29 * 16 + u) * WPB + wi] = (bw < B) ? one_hot[(long long)bw * n_onehot + ((16 * ch + h) % 8)] : 0.f;
        }
        voca_stage_w<37 * 3, 32>(w.cw[0], sw);
        __syncthreads();
        voca_conv<37, 16, 32, 1>(sA, sB, sw, w.cb[0]);            // in (A) -> a1 (B)
        __syncthreads();
        voca_stage_w<32 * 3, 32>(w.cw[1], sw);
        __syncthreads();
        voca_conv<32, 8, 32, 2>(sB, sA, sw, w.cb[1]);             // a1 (B) -> a2 (A)
        __syncthreads();
        voca_stage_w<32 * 3, 64>(w.cw[2], sw);
        __syncthreads();
        voca_conv<32, 4, 64, 2>(sA, sB, sw, w.cb[2]);             // a2 (A) -> a3 (B)
        __syncthreads();
        voca_stage_w<64 * 3, 64>(w.cw[3], sw);
        for (int i = threadIdx.x; i < WPB * 8; i += VT) {         // one_hot[:8] behind the 64 conv features
            const int wi = i % WPB, j = i / WPB, bw = w0 + wi;
            sA[(64 + j) * WPB + wi] = (bw < B) ? one_hot[(long long)bw * n_onehot + j] : 0.f;
        }
        __syncthreads();
        voca_conv<64, 2, 64, 4>(sB, sA, sw, w.cb[3]);             // a3 (B) -> a4[0..63] (A)
        __syncthreads();
        voca_stage_w<72, 72>(w.fw[0], sw);
        __syncthreads();
        voca_fc<72, 72, A2F_ACT_NONE, 2>(sA, sB, sw, w.fb[0]);    // a4 (A) -> f1 (B)
        __syncthreads();
        voca_stage_w<72, 128>(w.fw[1], sw);
        __syncthreads();
        voca_fc<72, 128, A2F_ACT_TANH, 2>(sB, sA, sw, w.fb[1]);   // f1 (B) -> f2 (A)
        __syncthreads();
        voca_stage_w<128, 50>(w.fw[2], sw);
        __syncthreads();
        voca_fc<128, 50, A2F_ACT_NONE, 4>(sA, sB, sw, w.fb[2]);   // f2 (A) -> z (B)
        __syncthreads();
        for (int i = threadIdx.x; i < WPB * ldz; i += VT) {
            const int wi = i / ldz, j = i - wi * ldz, bw = w0 + wi;
            if (bw < B) st_from_float(z + (long long)bw * ldz + j, j < 50 ? sB[j * WPB + wi] : 0.f);
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
    const int cap = 2 * sm_count();
    if (grid > cap) grid = cap;
    static bool attr_done = false;
    if (!attr_done) {
        A2F_CHECK_CUDA(cudaFuncSetAttribute(voca_trunk_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VOCA_SMEM));
        A2F_CHECK_CUDA(cudaFuncSetAttribute(voca_trunk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VOCA_SMEM));
        attr_done = true;
    }
    if (z_dtype == A2F_BF16)
        voca_trunk_kernel<bf16><<<grid, VT, VOCA_SMEM, as_stream(stream)>>>(vw, x, one_hot, n_onehot, static_cast<bf16*>(z), ldz, B);
    else
        voca_trunk_kernel<float><<<grid, VT, VOCA_SMEM, as_stream(stream)>>>(vw, x, one_hot, n_onehot, static_cast<float*>(z), ldz, B);
    A2F_CHECK_LAUNCH("voca_trunk_kernel");
    count_launch();
    return A2F_OK;
}

// Song2Face (ref:src/model/song2face.py:5-72; registry entry "song2face" of ref:src/model/lightning_model.py:50-58):
// conv stack -> two LSTMs over the CHANNEL axis (the reference feeds [B, C=256, H=64] to a batch_first LSTM: 256 steps of
// 64 features) -> bilinear resize of the hidden axis 256 -> 32 -> four convs along it -> the Audio2Mesh output MLP + vertex
// head.  The convolutions, the LSTM input projections, the MLP and the head reuse the kernels of the other models
// (a2f_im2col1d + a2f_gemm, a2f_a2m_mlp); this file holds what is specific to the model:
//   a2f_transpose_batched   [B, R, C] -> [B, C, R] fp32 (conv output [B, H, C] -> LSTM input [B, steps = C, features = H])
//   a2f_lstm_recurrence     h_t, c_t of a batch_first single-layer LSTM given the input projections of all steps
//   a2f_song2face_resize    [B, steps, hidden] -> [B, 32, steps]: the bilinear 256 -> 32 resize of the hidden axis
//                           (F.interpolate(size=(32,1)): source index 8 i + 3.5, the mean of two neighbours) written in
//                           the channels-last layout the regression convs read
#include "a2f_common.cuh"

namespace a2f {

__global__ void __launch_bounds__(256) transpose_batched_kernel(const float* __restrict__ x, float* __restrict__ y, int R, int C) {
    __shared__ float tile[32][33];
    pdl_sync();
    const long long b = blockIdx.z;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* xb = x + b * R * C;
    float* yb = y + b * R * C;
    for (int j = ty; j < 32; j += 8)
        if (r0 + j < R && c0 + tx < C) tile[j][tx] = xb[(long long)(r0 + j) * C + c0 + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
        if (c0 + j < C && r0 + tx < R) yb[(long long)(c0 + j) * R + r0 + tx] = tile[tx][j];
}

// One CTA per LSTM_BT batch elements (2 for small batches: more CTAs in flight; 4 from 128 windows on: less L2 traffic --
// measured at 64 windows: BT 1 / 2 / 4 = 15.1 / 6.8 / 9.3 ms per forward); thread (u, part) owns hidden unit u and 1/LSTM_KP of the contraction (HID = 256,
// blockDim = HID * LSTM_KP).  Part 0 keeps c[b][u] in registers for the whole sequence.  Per step the four gates of unit u
// are xp[b, t, g*HID + u] (input projection with both biases, precomputed for all steps by one GEMM)
// + sum_k W_hh[g*HID + u, k] h_{t-1}[b, k]; W_hh is read TRANSPOSED ([k][4*HID], coalesced across the threads) from L2
// every step and shared by the LSTM_BT batch elements of the CTA.  The step is a chain of dependent L2 round trips, so
// the k loop is unrolled 8-fold (32 independent loads in flight per thread) and split over LSTM_KP thread groups whose
// partial sums meet in shared memory: 83 -> ~10 us per step at 64 windows.  (Next step: W_hh slices resident in the
// shared memory of an 8-CTA cluster, h exchanged through DSMEM.)  PyTorch gate order i, f, g, o; fp32 throughout.
constexpr int LSTM_KP = 4;
template <int LSTM_BT>
__global__ void __launch_bounds__(256 * LSTM_KP) lstm_recurrence_kernel(const float* __restrict__ xp, const float* __restrict__ whh_t,
                                                                       float* __restrict__ hout, int B, int T, int HID) {
    extern __shared__ float lstm_sm[];               // [2][LSTM_BT][HID] hidden states, then [LSTM_KP-1][4][LSTM_BT][HID] partials
    float* part_sm = lstm_sm + 2 * LSTM_BT * HID;
    pdl_sync();
    const int u = threadIdx.x % HID, part = threadIdx.x / HID;
    const int b0 = blockIdx.x * LSTM_BT;
    float c[LSTM_BT];
#pragma unroll
    for (int i = 0; i < LSTM_BT; ++i) {
        c[i] = 0.f;
        if (part == 0) lstm_sm[i * HID + u] = 0.f;
    }
    __syncthreads();
    const int G = 4 * HID;
    const int kper = HID / LSTM_KP, k0 = part * kper;
    for (int t = 0; t < T; ++t) {
        const float* hp = lstm_sm + (t & 1) * LSTM_BT * HID;
        float* hn = lstm_sm + ((t + 1) & 1) * LSTM_BT * HID;
        float acc[4][LSTM_BT];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int i = 0; i < LSTM_BT; ++i) {
                const int b = b0 + i;
                acc[g][i] = (part == 0 && b < B) ? __ldg(xp + ((long long)b * T + t) * G + g * HID + u) : 0.f;
            }
#pragma unroll 1
        for (int kk = 0; kk < kper; kk += 8) {
            float w[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float* wr = whh_t + (long long)(k0 + kk + j) * G + u;
                w[j][0] = __ldg(wr); w[j][1] = __ldg(wr + HID); w[j][2] = __ldg(wr + 2 * HID); w[j][3] = __ldg(wr + 3 * HID);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < LSTM_BT; ++i) {
                    const float h = hp[i * HID + k0 + kk + j];
                    acc[0][i] = fmaf(w[j][0], h, acc[0][i]);
                    acc[1][i] = fmaf(w[j][1], h, acc[1][i]);
                    acc[2][i] = fmaf(w[j][2], h, acc[2][i]);
                    acc[3][i] = fmaf(w[j][3], h, acc[3][i]);
                }
        }
        if (part > 0) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int i = 0; i < LSTM_BT; ++i) part_sm[(((part - 1) * 4 + g) * LSTM_BT + i) * HID + u] = acc[g][i];
        }
        __syncthreads();
        if (part == 0) {
#pragma unroll
            for (int i = 0; i < LSTM_BT; ++i) {
                float gsum[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float v = acc[g][i];
#pragma unroll
                    for (int pp = 0; pp < LSTM_KP - 1; ++pp) v += part_sm[((pp * 4 + g) * LSTM_BT + i) * HID + u];
                    gsum[g] = v;
                }
                const float ig = 1.f / (1.f + expf(-gsum[0])), fg = 1.f / (1.f + expf(-gsum[1]));
                const float gg = tanhf(gsum[2]), og = 1.f / (1.f + expf(-gsum[3]));
                c[i] = fmaf(fg, c[i], ig * gg);
                const float h = og * tanhf(c[i]);
                hn[i * HID + u] = h;
                const int b = b0 + i;
                if (b < B) hout[((long long)b * T + t) * HID + u] = h;
            }
        }
        __syncthreads();
    }
}

// out[b, i, t] = 0.5 * (h[b, t, 8 i + 3] + h[b, t, 8 i + 4]) generalised: bilinear source index of F.interpolate
// (align_corners = False) along the hidden axis, output channels-last [B, out_h, steps]
__global__ void __launch_bounds__(256) song2face_resize_kernel(const float* __restrict__ h, int B, int T, int HID, int out_h,
                                                               float* __restrict__ out) {
    pdl_sync();
    const long long total = (long long)B * out_h * T;
    const float scale = (float)HID / (float)out_h;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(idx % T);
        const int i = (int)((idx / T) % out_h);
        const long long b = idx / ((long long)T * out_h);
        float src = scale * ((float)i + 0.5f) - 0.5f;
        if (src < 0.f) src = 0.f;
        int j0 = (int)src;
        if (j0 > HID - 1) j0 = HID - 1;
        const int j1 = j0 + (j0 < HID - 1 ? 1 : 0);
        const float l1 = src - (float)j0, l0 = 1.f - l1;
        const float* row = h + (b * T + t) * HID;
        out[idx] = l0 * row[j0] + l1 * row[j1];
    }
}

}  // namespace a2f

using namespace a2f;

extern "C" {

int a2f_transpose_batched(const float* x, float* y, int B, int R, int C, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(x && y && B > 0 && R > 0 && C > 0 && B < 65536, "a2f_transpose_batched: bad arguments");
    const dim3 grid((C + 31) / 32, (R + 31) / 32, B);
    A2F_CHECK_CUDA(launch_pdl(transpose_batched_kernel, grid, dim3(256), 0, as_stream(stream), x, y, R, C));
    count_launch();
    return A2F_OK;
}

int a2f_lstm_recurrence(const float* xp, const float* whh_t, float* hout, int B, int T, int hidden, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(xp && whh_t && hout && B > 0 && T > 0, "a2f_lstm_recurrence: bad arguments");
    A2F_REQUIRE(hidden == 256, "a2f_lstm_recurrence: hidden size must be 256 (one thread per unit)");
    const int bt = B >= 128 ? 4 : 2;
    const size_t smem = ((size_t)2 * bt * hidden + (size_t)(LSTM_KP - 1) * 4 * bt * hidden) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        A2F_CHECK_CUDA(cudaFuncSetAttribute(lstm_recurrence_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        A2F_CHECK_CUDA(cudaFuncSetAttribute(lstm_recurrence_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_done = true;
    }
    if (bt == 2)
        A2F_CHECK_CUDA(launch_pdl(lstm_recurrence_kernel<2>, dim3((B + 1) / 2), dim3(hidden * LSTM_KP), smem, as_stream(stream), xp,
                                  whh_t, hout, B, T, hidden));
    else
        A2F_CHECK_CUDA(launch_pdl(lstm_recurrence_kernel<4>, dim3((B + 3) / 4), dim3(hidden * LSTM_KP), smem, as_stream(stream), xp,
                                  whh_t, hout, B, T, hidden));
    count_launch();
    return A2F_OK;
}

int a2f_song2face_resize(const float* h, int B, int T, int hidden, int out_h, float* out, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(h && out && B > 0 && T > 0 && hidden > 0 && out_h > 0, "a2f_song2face_resize: bad arguments");
    const long long total = (long long)B * out_h * T;
    long long blocks = (total + 255) / 256;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    A2F_CHECK_CUDA(launch_pdl(song2face_resize_kernel, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), h, B, T, hidden,
                              out_h, out));
    count_launch();
    return A2F_OK;
}

}  // extern "C"

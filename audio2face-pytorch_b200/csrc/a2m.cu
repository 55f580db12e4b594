// Audio2Mesh (ref:src/model/audio2face.py:5-69) helpers.  The ten convolutions run as implicit GEMMs (a2f_gemm) over
// channels-last activations that carry one zero row/column of left padding, so that the im2col row of output position
// p of a (k=3, stride 2, pad 1) conv is the contiguous slice starting at padded position 2p:
//   analysis net     [B, 64, W+1, C]  conv along W (formant analysis, ref audio2face.py:13-29)
//   articulation net [B, H+1, 256]    conv along H (ref audio2face.py:31-47)
// Each GEMM writes straight into the padded layout of its successor (a2f_gemm_args.c_batch_stride).  Eval-mode
// BatchNorm that FOLLOWS a conv is folded into the packed weights/bias on the host; the two BatchNorms that PRECEDE a
// conv (layers 4 and 5 of the articulation net, ref audio2face.py:41-46) cannot be folded because the zero padding is
// applied after them, so they run as a per-channel affine pass over the un-padded rows (a2f_channel_affine).
#include "a2f_common.cuh"

namespace a2f {

// x [B,52,32] + tiled one-hot -> out [B,64,33] (single channel), column 0 = left zero pad.
// ref audio2face.py:59: one_hot.repeat(1,32).view(bs,1,-1,32)  =>  emb[r][c] = one_hot[(32 r + c) % n_onehot]
__global__ void a2m_assemble_kernel(const float* __restrict__ x, const float* __restrict__ one_hot, int n_onehot,
                                    float* __restrict__ out, int B) {
    const long long n = (long long)B * 64 * 33;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int c = (int)(i % 33);
        const int h = (int)((i / 33) % 64);
        const long long b = i / (33 * 64);
        float v = 0.f;
        if (c > 0) {
            const int w = c - 1;
            if (h < 52) v = x[(b * 52 + h) * 32 + w];
            else v = one_hot[b * n_onehot + ((32 * (h - 52) + w) % n_onehot)];
        }
        out[i] = v;
    }
}

template <typename T>
__global__ void channel_affine_kernel(T* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                      int C, long long rows_per_batch, long long ld, long long batch_stride,
                                      long long batches) {
    const long long n = batches * rows_per_batch * C;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int c = (int)(i % C);
        const long long row = i / C;
        const long long b = row / rows_per_batch, r = row % rows_per_batch;
        T* p = x + b * batch_stride + r * ld + c;
        st_from_float(p, ld_as_float(p) * scale[c] + shift[c]);
    }
}

// Tensor-core path of the conv stack (precision "bf16"): explicit im2col of a 1-D convolution over the middle axis of a
// channels-last fp32 activation [outer, L, C] into the error-compensated bf16 split [hi | lo | hi] (three Kpad-wide
// blocks per row, Kpad = taps*C rounded up to 64), which a2f_gemm multiplies with the weight split [hi | hi | lo] on
// tcgen05 -- hi*hi + lo*hi + hi*lo keeps ~2^-16 relative accuracy, so ten chained layers stay inside the 5e-4 m budget
// that a plain bf16 trunk misses (1 % of the offset scale).  An eval-mode BatchNorm that PRECEDES the conv is applied here
// (per-channel affine on in-range taps only: the zero padding comes after the BatchNorm in the reference).
// One thread per 8 consecutive k of one output row.
template <bool SPLIT>
__global__ void __launch_bounds__(256) im2col1d_split_kernel(const float* __restrict__ x, long long outer_stride, int ld, int C,
                                                             int L, int L_out, int taps, int stride, int pad,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             long long rows, int kpad, void* __restrict__ out_v) {
    pdl_sync();
    const int chunks = kpad >> 3;
    const int K = taps * C;
    const long long total = rows * chunks;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        const long long row = i / chunks;
        const int lo = (int)(row % L_out);
        const long long o = row / L_out;
        const float* xo = x + o * outer_stride;
        float v[8];
        int tap = (ch * 8) / C, c = ch * 8 - tap * C;       // one division per 8 elements; (tap, c) then advance incrementally
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = ch * 8 + e;
            float a = 0.f;
            if (k < K) {
                const int l = lo * stride - pad + tap;
                if (l >= 0 && l < L) {
                    a = __ldg(xo + (long long)l * ld + c);
                    if (scale != nullptr) a = fmaf(a, __ldg(scale + c), __ldg(shift + c));
                }
            }
            v[e] = a;
            if (++c == C) { c = 0; ++tap; }
        }
        if (!SPLIT) {                                   // fp32 rows [rows, kpad] for the SIMT parity path
            float* of = static_cast<float*>(out_v) + row * kpad + ch * 8;
            *reinterpret_cast<float4*>(of) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(of + 4) = make_float4(v[4], v[5], v[6], v[7]);
            continue;
        }
        bf16* out = static_cast<bf16*>(out_v);
        uint32_t hi[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bf16 h0 = __float2bfloat16_rn(v[2 * e]), h1 = __float2bfloat16_rn(v[2 * e + 1]);
            __nv_bfloat162 hh(h0, h1);
            hi[e] = *reinterpret_cast<uint32_t*>(&hh);
            lw[e] = pack_bf16x2(v[2 * e] - __bfloat162float(h0), v[2 * e + 1] - __bfloat162float(h1));
        }
        bf16* op = out + row * 3 * kpad + ch * 8;
        const uint4 uh = make_uint4(hi[0], hi[1], hi[2], hi[3]), ul = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        *reinterpret_cast<uint4*>(op) = uh;
        *reinterpret_cast<uint4*>(op + kpad) = ul;
        *reinterpret_cast<uint4*>(op + 2 * kpad) = uh;
    }
}

// Fused output MLP of Audio2Mesh (ref:src/model/audio2face.py:49-55,64-66 without the last Linear, which is the shared
// vertex head): z = W2 tanh(W1 (W0 [feat ; one_hot] + b0) + b1) + b2, zero-padded to ldz columns.  Four SIMT GEMM launches
// of 17-29 us each at 64 windows (profiles/r1_timeline_a2m_infer_b64.txt) become one; MLP_WPB rows per CTA, one thread
// per output neuron, activations in shared memory.
constexpr int MLP_WPB = 4;
constexpr int MLP_MAXW = 512;
__device__ __forceinline__ void mlp_layer(const float* __restrict__ in, int K, float* __restrict__ out, int N,
                                          const float* __restrict__ w, const float* __restrict__ b, bool tanh_act) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float acc[MLP_WPB];
        const float bv = __ldg(b + n);
#pragma unroll
        for (int i = 0; i < MLP_WPB; ++i) acc[i] = bv;
        const float* wr = w + (long long)n * K;
        int k = 0;
        if ((K & 3) == 0 && (reinterpret_cast<uintptr_t>(wr) & 15) == 0) {
            // 16-byte weight / activation loads, 4 independent rows in flight: the scalar loop was one dependent L1 round
            // trip per k (40 us per launch at 64 windows, profiles/r1_launches_conv_models.txt)
#pragma unroll 4
            for (; k < K; k += 4) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
                for (int i = 0; i < MLP_WPB; ++i) {
                    const float4 x4 = *reinterpret_cast<const float4*>(in + i * MLP_MAXW + k);
                    acc[i] = fmaf(w4.x, x4.x, acc[i]);
                    acc[i] = fmaf(w4.y, x4.y, acc[i]);
                    acc[i] = fmaf(w4.z, x4.z, acc[i]);
                    acc[i] = fmaf(w4.w, x4.w, acc[i]);
                }
            }
        }
        for (; k < K; ++k) {
            const float wv = __ldg(wr + k);
#pragma unroll
            for (int i = 0; i < MLP_WPB; ++i) acc[i] = fmaf(wv, in[i * MLP_MAXW + k], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < MLP_WPB; ++i) out[i * MLP_MAXW + n] = tanh_act ? tanhf(acc[i]) : acc[i];
    }
}

__global__ void __launch_bounds__(128) a2m_mlp_kernel(const float* __restrict__ feat, int ld_feat, int K0a,
                                                      const float* __restrict__ extra, int K0b, const float* __restrict__ w0,
                                                      const float* __restrict__ b0, int N0, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, int N1, const float* __restrict__ w2,
                                                      const float* __restrict__ b2, int N2, float* __restrict__ z, int ldz, int B) {
    __shared__ __align__(16) float bufA[MLP_WPB * MLP_MAXW];
    __shared__ __align__(16) float bufB[MLP_WPB * MLP_MAXW];
    pdl_sync();
    const int K0 = K0a + K0b;
    for (int r0 = blockIdx.x * MLP_WPB; r0 < B; r0 += gridDim.x * MLP_WPB) {
        for (int i = threadIdx.x; i < MLP_WPB * K0; i += blockDim.x) {
            const int wi = i / K0, k = i - wi * K0, r = r0 + wi;
            float v = 0.f;
            if (r < B) v = k < K0a ? feat[(long long)r * ld_feat + k] : extra[(long long)r * K0b + (k - K0a)];
            bufA[wi * MLP_MAXW + k] = v;
        }
        __syncthreads();
        mlp_layer(bufA, K0, bufB, N0, w0, b0, false);
        __syncthreads();
        mlp_layer(bufB, N0, bufA, N1, w1, b1, true);
        __syncthreads();
        mlp_layer(bufA, N1, bufB, N2, w2, b2, false);
        __syncthreads();
        for (int i = threadIdx.x; i < MLP_WPB * ldz; i += blockDim.x) {
            const int wi = i / ldz, j = i - wi * ldz, r = r0 + wi;
            if (r < B) z[(long long)r * ldz + j] = j < N2 ? bufB[wi * MLP_MAXW + j] : 0.f;
        }
        __syncthreads();
    }
}

}  // namespace a2f

using namespace a2f;

extern "C" {

int a2f_a2m_assemble(const float* x, const float* one_hot, int n_onehot, float* out, int B, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(x && one_hot && out && n_onehot > 0, "a2f_a2m_assemble: bad arguments");
    if (B <= 0) return A2F_OK;
    const long long n = (long long)B * 64 * 33;
    const int grid = (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
    a2m_assemble_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, one_hot, n_onehot, out, B);
    A2F_CHECK_LAUNCH("a2m_assemble_kernel");
    count_launch();
    return A2F_OK;
}

int a2f_im2col1d_split(const float* x, long long outer, long long outer_stride, int ld, int C, int L, int taps, int stride,
                        int pad, const float* scale, const float* shift, int kpad, void* out, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(x && out && outer > 0 && C > 0 && L > 0 && taps > 0 && stride > 0 && pad >= 0 && ld >= C,
                "a2f_im2col1d_split: bad arguments");
    A2F_REQUIRE((scale == nullptr) == (shift == nullptr), "a2f_im2col1d_split: scale and shift come together");
    A2F_REQUIRE(kpad >= taps * C && kpad % 8 == 0, "a2f_im2col1d_split: kpad must cover taps*C and be a multiple of 8");
    A2F_REQUIRE(reinterpret_cast<uintptr_t>(out) % 16 == 0, "a2f_im2col1d_split: out must be 16-byte aligned");
    const int L_out = (L + 2 * pad - taps) / stride + 1;
    A2F_REQUIRE(L_out > 0, "a2f_im2col1d_split: empty output");
    const long long rows = outer * L_out;
    const long long total = rows * (kpad / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    A2F_CHECK_CUDA(launch_pdl(im2col1d_split_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), x, outer_stride,
                              ld, C, L, L_out, taps, stride, pad, scale, shift, rows, kpad, out));
    count_launch();
    return A2F_OK;
}

int a2f_im2col1d(const float* x, long long outer, long long outer_stride, int ld, int C, int L, int taps, int stride, int pad,
                 const float* scale, const float* shift, int kpad, float* out, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(x && out && outer > 0 && C > 0 && L > 0 && taps > 0 && stride > 0 && pad >= 0 && ld >= C, "a2f_im2col1d: bad arguments");
    A2F_REQUIRE((scale == nullptr) == (shift == nullptr), "a2f_im2col1d: scale and shift come together");
    A2F_REQUIRE(kpad >= taps * C && kpad % 8 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0, "a2f_im2col1d: bad kpad / alignment");
    const int L_out = (L + 2 * pad - taps) / stride + 1;
    A2F_REQUIRE(L_out > 0, "a2f_im2col1d: empty output");
    const long long rows = outer * L_out;
    const long long total = rows * (kpad / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    A2F_CHECK_CUDA(launch_pdl(im2col1d_split_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), x, outer_stride,
                              ld, C, L, L_out, taps, stride, pad, scale, shift, rows, kpad, static_cast<void*>(out)));
    count_launch();
    return A2F_OK;
}

int a2f_a2m_mlp(const float* feat, int ld_feat, int k_feat, const float* extra, int k_extra, const float* w0, const float* b0,
                int n0, const float* w1, const float* b1, int n1, const float* w2, const float* b2, int n2, float* z, int ldz,
                int B, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(feat && extra && w0 && b0 && w1 && b1 && w2 && b2 && z, "a2f_a2m_mlp: NULL argument");
    A2F_REQUIRE(k_feat > 0 && k_extra >= 0 && k_feat + k_extra <= MLP_MAXW && n0 > 0 && n0 <= MLP_MAXW && n1 > 0 && n1 <= MLP_MAXW &&
                n2 > 0 && n2 <= ldz && ldz <= MLP_MAXW && ld_feat >= k_feat, "a2f_a2m_mlp: layer widths must be <= 512");
    if (B <= 0) return A2F_OK;
    int grid = (B + MLP_WPB - 1) / MLP_WPB;
    if (grid > 8 * sm_count()) grid = 8 * sm_count();
    A2F_CHECK_CUDA(launch_pdl(a2m_mlp_kernel, dim3(grid), dim3(128), 0, as_stream(stream), feat, ld_feat, k_feat, extra, k_extra, w0,
                              b0, n0, w1, b1, n1, w2, b2, n2, z, ldz, B));
    count_launch();
    return A2F_OK;
}

int a2f_channel_affine(void* x, int dtype, const float* scale, const float* shift, int C, long long rows_per_batch,
                       long long ld, long long batch_stride, long long batches, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(x && scale && shift && C > 0 && rows_per_batch > 0 && batches >= 0, "a2f_channel_affine: bad arguments");
    if (batches == 0) return A2F_OK;
    const long long n = batches * rows_per_batch * C;
    const int grid = (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
    if (dtype == A2F_BF16)
        channel_affine_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<bf16*>(x), scale, shift, C, rows_per_batch,
                                                                         ld, batch_stride, batches);
    else
        channel_affine_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<float*>(x), scale, shift, C,
                                                                          rows_per_batch, ld, batch_stride, batches);
    A2F_CHECK_LAUNCH("channel_affine_kernel");
    count_launch();
    return A2F_OK;
}

}  // extern "C"

// MFCC feature extractor on the GPU: the step immediately before Audio2Mesh / VOCA,
// ref:src/model/extractor.py:10-60 = torchaudio.transforms.MFCC (MelSpectrogram: centre-padded hann STFT, power 2,
// 128 HTK mel bands; AmplitudeToDB("power", top_db 80) with ONE cut-off for the whole 3-D batch; ortho DCT-II) followed by
// a transpose and a bilinear resize of the time axis to out_dim.
//
// Pipeline (the DFT is a GEMM on the tcgen05 / fp32 SIMT back ends of a2f_gemm; everything else is below):
//   a2f_mfcc_frames      audio [B,N] -> frame matrix [B*F, K] of raw reflect-padded samples under the window support
//                        (the window is folded into the DFT basis on the host), fp32 or the bf16x3 split [hi|lo|hi];
//   a2f_gemm             frames x basis^T -> [B*F, 2*513 (+pad)] = (Re | Im) of the one-sided spectrum, fp32;
//   a2f_mfcc_mel_db      |X|^2 -> mel bands (each band's contiguous support only) -> 10 log10(max(., 1e-10)); batch-global
//                        maximum by an order-preserving integer atomicMax;
//   a2f_mfcc_dct_resize  clamp at (max - top_db), DCT-II (128 -> n_mfcc), bilinear resize F -> out_dim (align_corners
//                        = False, ATen's source-index formula), output [B, out_dim, n_mfcc].
#include "a2f_common.cuh"

namespace a2f {

A2F_D int float_to_ordered(float v) {
    int x = __float_as_int(v);
    return x ^ ((x >> 31) & 0x7fffffff);
}
A2F_D float ordered_to_float(int x) { return __int_as_float(x ^ ((x >> 31) & 0x7fffffff)); }

// one thread per 8 consecutive K positions of one frame row
template <bool SPLIT>
__global__ void __launch_bounds__(256) mfcc_frames_kernel(const float* __restrict__ audio, int B, int N, int F, int win,
                                                          int hop, int n_fft, int kpad, void* __restrict__ Aout,
                                                          int* __restrict__ gmax_slot) {
    pdl_sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) *gmax_slot = float_to_ordered(-INFINITY);
    const int chunks = kpad >> 3;
    const long long total = (long long)B * F * chunks;
    const int lo = (n_fft - win) / 2 - n_fft / 2;       // first sample of the window support relative to the frame centre
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        const long long row = i / chunks;
        const int t = (int)(row % F), b = (int)(row / F);
        const float* x = audio + (long long)b * N;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int kk = ch * 8 + e;
            int idx = t * hop + lo + kk;                // centre=True, pad_mode="reflect"
            if (idx < 0) idx = -idx;
            if (idx >= N) idx = 2 * (N - 1) - idx;
            v[e] = (kk < win && idx >= 0 && idx < N) ? __ldg(x + idx) : 0.f;
        }
        if (SPLIT) {
            bf16* o = static_cast<bf16*>(Aout) + row * 3 * kpad + ch * 8;
            uint32_t hi[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bf16 h0 = __float2bfloat16_rn(v[2 * e]), h1 = __float2bfloat16_rn(v[2 * e + 1]);
                const float l0 = v[2 * e] - __bfloat162float(h0), l1 = v[2 * e + 1] - __bfloat162float(h1);
                __nv_bfloat162 hh(h0, h1);
                hi[e] = *reinterpret_cast<uint32_t*>(&hh);
                lw[e] = pack_bf16x2(l0, l1);
            }
            const uint4 uh = make_uint4(hi[0], hi[1], hi[2], hi[3]), ul = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            *reinterpret_cast<uint4*>(o) = uh;
            *reinterpret_cast<uint4*>(o + kpad) = ul;
            *reinterpret_cast<uint4*>(o + 2 * kpad) = uh;
        } else {
            float* o = static_cast<float*>(Aout) + row * kpad + ch * 8;
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
}

// one warp per frame: power spectrum into shared memory, then each lane sums the support of its mel bands
constexpr int MEL_WARPS = 8;
__global__ void __launch_bounds__(32 * MEL_WARPS) mfcc_mel_db_kernel(const float* __restrict__ spec, int ld_spec, int M,
                                                                     int n_freq, const float* __restrict__ fb,
                                                                     const int* __restrict__ band, int n_mels,
                                                                     float* __restrict__ db, int* __restrict__ gmax_slot) {
    extern __shared__ float mel_sm[];                   // [MEL_WARPS][n_freq]
    pdl_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* P = mel_sm + warp * n_freq;
    float lmax = -INFINITY;
    for (int row = blockIdx.x * MEL_WARPS + warp; row < M; row += gridDim.x * MEL_WARPS) {
        const float* sp = spec + (long long)row * ld_spec;
        for (int f = lane; f < n_freq; f += 32) {
            const float re = sp[f], im = sp[n_freq + f];
            P[f] = fmaf(re, re, im * im);
        }
        __syncwarp();
        for (int m = lane; m < n_mels; m += 32) {
            const int f0 = band[2 * m], f1 = band[2 * m + 1];
            float acc = 0.f;
            for (int f = f0; f < f1; ++f) acc = fmaf(P[f], __ldg(fb + (long long)f * n_mels + m), acc);
            const float d = 10.f * log10f(fmaxf(acc, 1e-10f));
            db[(long long)row * n_mels + m] = d;
            lmax = fmaxf(lmax, d);
        }
        __syncwarp();
    }
    lmax = warp_max(lmax);
    if (lane == 0 && lmax > -INFINITY) atomicMax(gmax_slot, float_to_ordered(lmax));
}

// one warp per output row (b, o): the two source frames are clamped and blended FIRST (the DCT is linear and the blend
// weights sum to one, so this equals blending the two DCT rows up to fp32 rounding), staged in shared memory, and the
// 128 x n_mfcc DCT runs with the 32 lanes split as (coefficient k, slice of the mel axis) -- all lanes busy for
// n_mfcc = 16 or 32, DCT matrix in shared memory, one shuffle round to add the slices.
constexpr int DCT_WARPS = 8;
__global__ void __launch_bounds__(32 * DCT_WARPS) mfcc_dct_resize_kernel(const float* __restrict__ db, const int* __restrict__ gmax_slot,
                                                                         float top_db, const float* __restrict__ dct, int B, int F,
                                                                         int n_mels, int n_mfcc, int out_dim, float* __restrict__ out) {
    extern __shared__ float dct_sm[];                  // [n_mels * n_mfcc] DCT matrix, then [DCT_WARPS][n_mels] blended rows
    float* sdct = dct_sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* srow = dct_sm + n_mels * n_mfcc + warp * n_mels;
    for (int i = threadIdx.x; i < n_mels * n_mfcc; i += blockDim.x) sdct[i] = __ldg(dct + i);   // module constant
    pdl_sync();
    __syncthreads();
    const float floor_db = ordered_to_float(*gmax_slot) - top_db;
    // lanes = (k, slice): KP coefficients in flight, 32/KP slices of the mel axis
    const int KP = n_mfcc >= 32 ? 32 : n_mfcc >= 16 ? 16 : n_mfcc >= 8 ? 8 : 4;
    const int NS = 32 / KP;
    const int kl = lane % KP, sl = lane / KP;
    const int mper = (n_mels + NS - 1) / NS;
    for (int row = blockIdx.x * DCT_WARPS + warp; row < B * out_dim; row += gridDim.x * DCT_WARPS) {
        const int b = row / out_dim, o = row % out_dim;
        int h0 = o, h1 = o;
        float l0 = 1.f, l1 = 0.f;
        if (out_dim != F) {
            // ATen area_pixel_compute_source_index (align_corners=False): src = scale*(dst+0.5)-0.5, clamped at 0
            const float scale = (float)F / (float)out_dim;
            float src = scale * ((float)o + 0.5f) - 0.5f;
            if (src < 0.f) src = 0.f;
            h0 = (int)src;
            if (h0 > F - 1) h0 = F - 1;
            h1 = h0 + (h0 < F - 1 ? 1 : 0);
            l1 = src - (float)h0;
            l0 = 1.f - l1;
        }
        const float* r0 = db + ((long long)b * F + h0) * n_mels;
        const float* r1 = db + ((long long)b * F + h1) * n_mels;
        for (int m = lane; m < n_mels; m += 32) {
            const float a = fmaxf(r0[m], floor_db);
            srow[m] = (out_dim != F) ? l0 * a + l1 * fmaxf(r1[m], floor_db) : a;
        }
        __syncwarp();
        for (int k0 = 0; k0 < n_mfcc; k0 += KP) {
            const int k = k0 + kl;
            // fp64 accumulation: c0 sums 128 dB values of the same sign (|c0| ~ 500), where fp32 summation-order noise
            // alone reaches 1e-3; 2048 fp64 FMAs per row are free next to the DFT
            double acc = 0.0;
            if (k < n_mfcc) {
                const int m1 = min(n_mels, (sl + 1) * mper);
                for (int m = sl * mper; m < m1; ++m) acc = fma((double)srow[m], (double)sdct[m * n_mfcc + k], acc);
            }
            for (int off = KP; off < 32; off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (sl == 0 && k < n_mfcc) out[(long long)row * n_mfcc + k] = (float)acc;
        }
        __syncwarp();
    }
}

}  // namespace a2f

using namespace a2f;

extern "C" {

int a2f_mfcc_frames(const float* audio, int B, int N, int win, int hop, int n_fft, int kpad, void* A, int a_dtype,
                    float* gmax_slot, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(audio && A && gmax_slot && B > 0 && N > 0, "a2f_mfcc_frames: bad arguments");
    A2F_REQUIRE(win > 0 && win <= n_fft && hop > 0 && kpad >= win && kpad % 8 == 0, "a2f_mfcc_frames: bad window geometry");
    A2F_REQUIRE(N > n_fft / 2, "a2f_mfcc_frames: reflect padding needs N > n_fft/2");
    A2F_REQUIRE(reinterpret_cast<uintptr_t>(A) % 16 == 0, "a2f_mfcc_frames: A must be 16-byte aligned");
    const int F = 1 + N / hop;
    const long long total = (long long)B * F * (kpad / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    cudaStream_t s = as_stream(stream);
    int* slot = reinterpret_cast<int*>(gmax_slot);
    if (a_dtype == A2F_BF16)
        A2F_CHECK_CUDA(launch_pdl(mfcc_frames_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, s, audio, B, N, F, win, hop,
                                  n_fft, kpad, A, slot));
    else if (a_dtype == A2F_F32)
        A2F_CHECK_CUDA(launch_pdl(mfcc_frames_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, s, audio, B, N, F, win, hop,
                                  n_fft, kpad, A, slot));
    else return set_error(A2F_EINVAL, "a2f_mfcc_frames: bad dtype");
    count_launch();
    return A2F_OK;
}

int a2f_mfcc_mel_db(const float* spec, int ld_spec, int M, int n_freq, const float* fb, const int* band, int n_mels,
                    float* db, float* gmax_slot, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(spec && fb && band && db && gmax_slot && M > 0 && n_freq > 0 && n_mels > 0, "a2f_mfcc_mel_db: bad arguments");
    A2F_REQUIRE(ld_spec >= 2 * n_freq, "a2f_mfcc_mel_db: spectrum rows hold (Re | Im)");
    const size_t smem = (size_t)MEL_WARPS * n_freq * sizeof(float);
    A2F_REQUIRE(smem <= 48 * 1024, "a2f_mfcc_mel_db: n_freq too large");
    int blocks = (M + MEL_WARPS - 1) / MEL_WARPS;
    if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
    A2F_CHECK_CUDA(launch_pdl(mfcc_mel_db_kernel, dim3(blocks), dim3(32 * MEL_WARPS), smem, as_stream(stream), spec, ld_spec, M,
                              n_freq, fb, band, n_mels, db, reinterpret_cast<int*>(gmax_slot)));
    count_launch();
    return A2F_OK;
}

int a2f_mfcc_dct_resize(const float* db, const float* gmax_slot, float top_db, const float* dct, int B, int F, int n_mels,
                        int n_mfcc, int out_dim, float* out, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(db && gmax_slot && dct && out && B > 0 && F > 0 && n_mels > 0 && n_mfcc > 0 && out_dim > 0,
                "a2f_mfcc_dct_resize: bad arguments");
    const size_t smem = ((size_t)n_mels * n_mfcc + (size_t)DCT_WARPS * n_mels) * sizeof(float);
    A2F_REQUIRE(smem <= 48 * 1024, "a2f_mfcc_dct_resize: n_mels * n_mfcc too large");
    long long blocks = ((long long)B * out_dim + DCT_WARPS - 1) / DCT_WARPS;
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    A2F_CHECK_CUDA(launch_pdl(mfcc_dct_resize_kernel, dim3((unsigned)blocks), dim3(32 * DCT_WARPS), smem, as_stream(stream), db,
                              reinterpret_cast<const int*>(gmax_slot), top_db, dct, B, F, n_mels, n_mfcc, out_dim, out));
    count_launch();
    return A2F_OK;
}

}  // extern "C"

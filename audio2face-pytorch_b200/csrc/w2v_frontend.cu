// wav2vec2 front end of FaceFormer's audio encoder, channels-last:
//   a2f_audio_stats     processor zero-mean/unit-variance statistics (the reference does this in numpy on the host)
//   a2f_conv0_gn_gelu   normalise + Conv1d(1->512,k10,s5) + GroupNorm(512 groups over ALL time) + GELU
//   a2f_interp_ln       linear_interpolation(align_corners=True) to frame_num + LayerNorm(512) of the projection
//   a2f_layernorm       LayerNorm over the last dim (encoder post-LN blocks)
//
// GroupNorm statistics without a pass over the 512 x L0 conv output: the conv is linear, so per channel c
//   sum_t y[c,t]   = w_c . S,          S[k]    = sum_t x[5t+k]
//   sum_t y[c,t]^2 = w_c^T R w_c,      R[k,k'] = sum_t x[5t+k] x[5t+k']
// One cheap pass over the audio builds the 10-vector S and the 10x10 matrix R per utterance (fp64), a tiny kernel
// turns them into mean / rstd per (utterance, channel), and the apply pass recomputes the conv (10 FMAs) fused with
// normalise + affine + GELU, writing channels-last so that conv1..6 are implicit GEMMs with contiguous K.
#include "a2f_common.cuh"

namespace a2f {

// ------------------------------------------------------------------------------------------------ audio stats
__global__ void __launch_bounds__(1024) audio_stats_kernel(const float* __restrict__ audio, long long N,
                                                           float* __restrict__ stats) {
    pdl_sync();   // PDL: wait for the previous kernel's results, let the next kernel's prologue start
    const float* x = audio + (long long)blockIdx.x * N;
    __shared__ double sh[32];
    __shared__ double s_mean;
    // float4 loads, 8 in flight per thread: the two passes are latency-bound (one CTA per utterance), not bandwidth-bound
    const bool vec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const long long n4 = vec ? N / 4 : 0;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    double s = 0.0;
    for (long long i0 = threadIdx.x; i0 < n4; i0 += 8LL * blockDim.x) {
        float4 a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long i = i0 + (long long)u * blockDim.x;
            a[u] = i < n4 ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) s += ((double)a[u].x + (double)a[u].y) + ((double)a[u].z + (double)a[u].w);
    }
    for (long long i = 4 * n4 + threadIdx.x; i < N; i += blockDim.x) s += (double)x[i];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = sh[threadIdx.x];
        t = warp_sum_d(t);
        if (threadIdx.x == 0) s_mean = t / (double)N;
    }
    __syncthreads();
    // numpy: mean in fp32, then var = mean(|x - mean|^2): centre with the fp32-rounded mean like the reference does
    const float meanf = (float)s_mean;
    double v = 0.0;
    for (long long i0 = threadIdx.x; i0 < n4; i0 += 8LL * blockDim.x) {
        float4 a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long i = i0 + (long long)u * blockDim.x;
            a[u] = i < n4 ? __ldg(x4 + i) : make_float4(meanf, meanf, meanf, meanf);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float d0 = a[u].x - meanf, d1 = a[u].y - meanf, d2 = a[u].z - meanf, d3 = a[u].w - meanf;
            v += ((double)(d0 * d0) + (double)(d1 * d1)) + ((double)(d2 * d2) + (double)(d3 * d3));
        }
    }
    for (long long i = 4 * n4 + threadIdx.x; i < N; i += blockDim.x) {
        const float d = x[i] - meanf;
        v += (double)(d * d);
    }
    v = warp_sum_d(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = sh[threadIdx.x];
        t = warp_sum_d(t);
        if (threadIdx.x == 0) {
            const float var = (float)(t / (double)N);
            stats[2 * blockIdx.x] = meanf;
            stats[2 * blockIdx.x + 1] = 1.0f / sqrtf(var + 1e-7f);
        }
    }
}

// ------------------------------------------------------------------------------------------------ conv0 moments
constexpr int MOM_TCH = 1024;   // conv outputs per CTA
constexpr int MOM_N = 65;       // 10 sums + 55 upper-triangular products

__global__ void __launch_bounds__(256) conv0_moments_kernel(const float* __restrict__ audio,
                                                            const float* __restrict__ stats, long long N, int L0,
                                                            int nchunk, double* __restrict__ partial) {
    pdl_sync();   // PDL: wait for the previous kernel's results, let the next kernel's prologue start
    const int b = blockIdx.y, chunk = blockIdx.x;
    const float* x = audio + (long long)b * N;
    const float mean = stats[2 * b], rstd = stats[2 * b + 1];
    float acc[MOM_N];
#pragma unroll
    for (int i = 0; i < MOM_N; ++i) acc[i] = 0.f;
    const int t_end = min(L0, (chunk + 1) * MOM_TCH);
    for (int t = chunk * MOM_TCH + threadIdx.x; t < t_end; t += blockDim.x) {
        float v[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) v[k] = (x[5LL * t + k] - mean) * rstd;
        int idx = 10;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            acc[k] += v[k];
#pragma unroll
            for (int k2 = k; k2 < 10; ++k2) {
                acc[idx] = fmaf(v[k], v[k2], acc[idx]);
                ++idx;
            }
        }
    }
    __shared__ double sh[8][MOM_N];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < MOM_N; ++i) {
        double d = warp_sum_d((double)acc[i]);
        if (lane == 0) sh[warp][i] = d;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MOM_N; i += blockDim.x) {
        double d = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) d += sh[w][i];
        partial[((long long)b * nchunk + chunk) * MOM_N + i] = d;
    }
}

// per (utterance, channel): mean and rstd of the conv output over time (biased variance, eps 1e-5)
__global__ void __launch_bounds__(512) conv0_gn_stats_kernel(const double* __restrict__ partial, int nchunk,
                                                             const float* __restrict__ w, int L0,
                                                             float2* __restrict__ gn) {
    pdl_sync();   // PDL: wait for the previous kernel's results, let the next kernel's prologue start
    const int b = blockIdx.x, c = threadIdx.x;
    __shared__ double mom[MOM_N];
    for (int i = threadIdx.x; i < MOM_N; i += blockDim.x) {
        double d = 0.0;
        for (int k = 0; k < nchunk; ++k) d += partial[((long long)b * nchunk + k) * MOM_N + i];
        mom[i] = d;
    }
    __syncthreads();
    double wc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) wc[k] = (double)w[c * 10 + k];
    double s1 = 0.0, s2 = 0.0;
    int idx = 10;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        s1 += wc[k] * mom[k];
#pragma unroll
        for (int k2 = k; k2 < 10; ++k2) {
            const double term = wc[k] * wc[k2] * mom[idx++];
            s2 += (k2 == k) ? term : 2.0 * term;
        }
    }
    const double mean = s1 / L0;
    double var = s2 / L0 - mean * mean;
    if (var < 0.0) var = 0.0;
    gn[b * 512 + c] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
}

// ------------------------------------------------------------------------------------------------ raw moments (inference)
// The processor statistics (a2f_audio_stats) and the conv0 moments in ONE pass over the RAW audio: besides the 10 lag sums
// and 55 lag products of conv0_moments_kernel, every chunk also sums x and x^2 of its samples.  conv0_gn_stats_auto_kernel
// then derives mean / rstd of the utterance and the moments of the NORMALISED audio algebraically (fp64):
//     S'[k] = r (S[k] - L0 m),     R'[k,k'] = r^2 (R[k,k'] - m (S[k] + S[k']) + L0 m^2),     x' = (x - m) r
// which removes one launch and one full pass from the dependent chain in front of conv0 (audio_stats 24 us + moments 29 us at
// the bench shape, both latency-bound).  m is the fp32-rounded mean and r = 1/sqrt(var + 1e-7) in fp32 like the processor
// (ref:src/model/faceformer.py:142-144 through Wav2Vec2Processor); var = (sum x^2 - 2 m sum x + N m^2) / N in fp64.
constexpr int RAW_N = MOM_N + 2;    // + sum x, sum x^2

__global__ void __launch_bounds__(256) conv0_raw_moments_kernel(const float* __restrict__ audio, long long N, int L0, int nchunk,
                                                                double* __restrict__ partial) {
    pdl_sync();
    const int b = blockIdx.y, chunk = blockIdx.x;
    const float* x = audio + (long long)b * N;
    __shared__ __align__(16) float xs[5 * MOM_TCH + 16];
    __shared__ double sh[8][RAW_N];
    const int t0 = chunk * MOM_TCH;
    const int tn = min(MOM_TCH, L0 - t0);                     // conv outputs of this chunk
    const long long s0 = 5LL * t0;                            // first sample of the chunk
    // samples the lag windows touch: [s0, s0 + 5 tn + 5); samples this chunk OWNS for sum x / sum x^2: [s0, s0 + 5 tn), the last
    // chunk also owns everything up to N
    const int n_win = 5 * tn + 5;
    const long long own_end = (chunk == nchunk - 1) ? N : s0 + 5LL * tn;
    float sx = 0.f, sxx = 0.f;
    for (int i = threadIdx.x; i < n_win; i += 256) {
        const long long gi = s0 + i;
        const float v = gi < N ? __ldg(x + gi) : 0.f;
        xs[i] = v;
        if (gi < own_end) { sx += v; sxx = fmaf(v, v, sxx); }
    }
    for (long long gi = s0 + n_win + threadIdx.x; gi < own_end; gi += 256) {      // tail of the last chunk (at most a few samples)
        const float v = __ldg(x + gi);
        sx += v; sxx = fmaf(v, v, sxx);
    }
    __syncthreads();
    float acc[MOM_N];
#pragma unroll
    for (int i = 0; i < MOM_N; ++i) acc[i] = 0.f;
    for (int t = threadIdx.x; t < tn; t += 256) {
        float v[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) v[k] = xs[5 * t + k];    // stride 5 over the lanes: conflict-free
        int idx = 10;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            acc[k] += v[k];
#pragma unroll
            for (int k2 = k; k2 < 10; ++k2) {
                acc[idx] = fmaf(v[k], v[k2], acc[idx]);
                ++idx;
            }
        }
    }
    // at most 4 fp32 terms per thread and 32 per warp are summed in fp32, everything above that in fp64
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < MOM_N; ++i) {
        const float d = warp_sum(acc[i]);
        if (lane == 0) sh[warp][i] = (double)d;
    }
    {
        const double d0 = warp_sum_d((double)sx), d1 = warp_sum_d((double)sxx);
        if (lane == 0) { sh[warp][MOM_N] = d0; sh[warp][MOM_N + 1] = d1; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RAW_N; i += 256) {
        double d = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) d += sh[w][i];
        partial[((long long)b * nchunk + chunk) * RAW_N + i] = d;
    }
}

// per utterance: processor statistics -> stats[2b], stats[2b+1]; per (utterance, channel): mean and rstd of the conv output
__global__ void __launch_bounds__(512) conv0_gn_stats_auto_kernel(const double* __restrict__ partial, int nchunk,
                                                                  const float* __restrict__ w, long long N, int L0,
                                                                  float* __restrict__ stats, float2* __restrict__ gn) {
    pdl_sync();
    const int b = blockIdx.x, c = threadIdx.x;
    __shared__ double raw[RAW_N];
    __shared__ double mom[MOM_N];
    for (int i = threadIdx.x; i < RAW_N; i += blockDim.x) {
        double d = 0.0;
        for (int k = 0; k < nchunk; ++k) d += partial[((long long)b * nchunk + k) * RAW_N + i];
        raw[i] = d;
    }
    __syncthreads();
    const float meanf = (float)(raw[MOM_N] / (double)N);
    const double m = (double)meanf;
    const float var = (float)((raw[MOM_N + 1] - 2.0 * m * raw[MOM_N] + (double)N * m * m) / (double)N);
    const float rstdf = 1.0f / sqrtf(fmaxf(var, 0.f) + 1e-7f);
    const double r = (double)rstdf;
    if (threadIdx.x == 0) {
        stats[2 * b] = meanf;
        stats[2 * b + 1] = rstdf;
    }
    if (threadIdx.x < MOM_N) {
        const int i = threadIdx.x;
        if (i < 10) {
            mom[i] = r * (raw[i] - (double)L0 * m);
        } else {
            // upper-triangular index i -> (k, k2) in the order of the accumulation loop
            int k = 0, base = 10;
            while (i >= base + (10 - k)) { base += 10 - k; ++k; }
            const int k2 = k + (i - base);
            mom[i] = r * r * (raw[i] - m * (raw[k] + raw[k2]) + (double)L0 * m * m);
        }
    }
    __syncthreads();
    double wc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) wc[k] = (double)w[c * 10 + k];
    double s1 = 0.0, s2 = 0.0;
    int idx = 10;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        s1 += wc[k] * mom[k];
#pragma unroll
        for (int k2 = k; k2 < 10; ++k2) {
            const double term = wc[k] * wc[k2] * mom[idx++];
            s2 += (k2 == k) ? term : 2.0 * term;
        }
    }
    const double mean = s1 / L0;
    double v2 = s2 / L0 - mean * mean;
    if (v2 < 0.0) v2 = 0.0;
    gn[b * 512 + c] = make_float2((float)mean, (float)(1.0 / sqrt(v2 + 1e-5)));
}

// ------------------------------------------------------------------------------------------------ conv0 apply
constexpr int C0_TCH = 64;   // time steps per CTA

template <typename TO>
__global__ void __launch_bounds__(256) conv0_apply_kernel(const float* __restrict__ audio,
                                                          const float* __restrict__ stats,
                                                          const float* __restrict__ w, const float2* __restrict__ gn,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, TO* __restrict__ out,
                                                          long long N, int L0, long long out_batch_stride) {
    pdl_sync();   // PDL: wait for the previous kernel's results, let the next kernel's prologue start
    const int b = blockIdx.y, t0 = blockIdx.x * C0_TCH;
    __shared__ __align__(16) float xs[5 * C0_TCH + 8];
    const float* x = audio + (long long)b * N;
    const float mean = stats[2 * b], rstd = stats[2 * b + 1];
    const int nx = 5 * C0_TCH + 8;       // 5 extra taps of the last window + float4 over-read, zero filled
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
        const long long gi = 5LL * t0 + i;
        xs[i] = (gi < N && i < 5 * C0_TCH + 5) ? (x[gi] - mean) * rstd : 0.f;
    }
    // each thread owns two adjacent channels (one packed 4-byte store in bf16, coalesced either way)
    const int c = threadIdx.x * 2;
    float2 wp[10];                        // (w[c][k], w[c+1][k]): one FFMA2 per tap for the channel pair
#pragma unroll
    for (int k = 0; k < 10; ++k) wp[k] = make_float2(w[c * 10 + k], w[(c + 1) * 10 + k]);
    const float2 g0 = gn[b * 512 + c], g1 = gn[b * 512 + c + 1];
    const float ga0 = gamma[c], ga1 = gamma[c + 1], be0 = beta[c], be1 = beta[c + 1];
    // bf16 path: GroupNorm affine folded to one FMA (y = conv*A + Bc) and the MUFU.TANH GELU; fp32 path: literal form
    const float2 Aff = make_float2(g0.y * ga0, g1.y * ga1);
    const float2 Bff = make_float2(be0 - g0.x * Aff.x, be1 - g1.x * Aff.y);
    if (sizeof(TO) == 2) {                // ... and A folded into the taps, Bc into the accumulator's start value
#pragma unroll
        for (int k = 0; k < 10; ++k) wp[k] = fmul2(wp[k], Aff);
    }
    const float2 a_init = sizeof(TO) == 2 ? Bff : make_float2(0.f, 0.f);
    __syncthreads();
    TO* o = out + (long long)b * out_batch_stride;
    const int tn = min(C0_TCH, L0 - t0);
    for (int t4 = 0; t4 < tn; t4 += 4) {
        // four output steps share one 28-float window of the audio (7 LDS.128 instead of 40 scalar LDS)
        float xw[28];
#pragma unroll
        for (int qd = 0; qd < 7; ++qd) {
            const float4 f = *reinterpret_cast<const float4*>(xs + 5 * t4 + 4 * qd);
            xw[4 * qd] = f.x; xw[4 * qd + 1] = f.y; xw[4 * qd + 2] = f.z; xw[4 * qd + 3] = f.w;
        }
        if (sizeof(TO) == 2 && t4 + 4 <= tn) {
            // complete group of four steps (all but the last group of an utterance): no per-step exit test, so the four
            // tap chains and GELUs are independent instruction streams the scheduler can interleave
            float2 a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = a_init;
#pragma unroll
            for (int k = 0; k < 10; ++k) {
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = ffma2(wp[k], make_float2(xw[5 * u + k], xw[5 * u + k]), a[u]);
            }
            TO* p = o + (long long)(t0 + t4) * 512 + c;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float2 y = gelu_fast2(a[u]);
                *reinterpret_cast<uint32_t*>(p + u * 512) = pack_bf16x2(y.x, y.y);
            }
            continue;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (t4 + u >= tn) break;
            float2 a = a_init;
#pragma unroll
            for (int k = 0; k < 10; ++k) a = ffma2(wp[k], make_float2(xw[5 * u + k], xw[5 * u + k]), a);
            TO* p = o + (long long)(t0 + t4 + u) * 512 + c;
            if (sizeof(TO) == 2) {
                const float2 y = gelu_fast2(a);
                *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(y.x, y.y);
            } else {
                const float y0 = gelu_erf((a.x - g0.x) * g0.y * ga0 + be0);
                const float y1 = gelu_erf((a.y - g1.x) * g1.y * ga1 + be1);
                *reinterpret_cast<float2*>(p) = make_float2(y0, y1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ interp + LN
// One warp per output frame.  ATen upsample_linear1d(align_corners=True): scale=(S-1)/(T-1) in fp32, src=scale*t,
// i0=trunc(src), i1=i0+(i0<S-1), l1=src-i0, out=(1-l1)*x[i0]+l1*x[i1]   (SURVEY.md A.3).
template <typename TI, typename TO, int C>
__global__ void __launch_bounds__(256) interp_ln_kernel(const TI* __restrict__ in, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        TO* __restrict__ out, int B, int S, int T) {
    pdl_sync();   // PDL: wait for the previous kernel's results, let the next kernel's prologue start
    constexpr int PER = C / 32;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * T) return;
    const int b = warp / T, t = warp % T;
    const float scale = T > 1 ? (float)(S - 1) / (float)(T - 1) : 0.f;
    const float src = scale * (float)t;
    int i0 = (int)src;
    if (i0 > S - 1) i0 = S - 1;
    const int i1 = i0 + (i0 < S - 1 ? 1 : 0);
    float l1 = src - (float)i0;
    l1 = fminf(fmaxf(l1, 0.f), 1.f);
    const float l0 = 1.f - l1;
    const TI* r0 = in + ((long long)b * S + i0) * C;
    const TI* r1 = in + ((long long)b * S + i1) * C;
    float v[PER];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int cc = j * 32 + lane;
        v[j] = l0 * ld_as_float(r0 + cc) + l1 * ld_as_float(r1 + cc);
        s += v[j];
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const float d = v[j] - mean;
        q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
    TO* o = out + ((long long)b * T + t) * C;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int cc = j * 32 + lane;
        st_from_float(o + cc, (v[j] - mean) * rstd * gamma[cc] + beta[cc]);
    }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per row, row cached in registers as float4 (C % 128 == 0, C <= 1024).
template <typename TI> A2F_D float4 ld4(const TI* p);
template <> A2F_D float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> A2F_D float4 ld4<bf16>(const bf16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
A2F_D void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
A2F_D void st4(bf16* p, float4 v) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
}

template <typename TI, typename TO, typename TO2, int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const TI* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        TO* __restrict__ out, TO2* __restrict__ out2, long long rows) {
    pdl_sync();   // PDL: wait for the previous kernel's results, let the next kernel's prologue start
    constexpr int C = NV * 128;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const TI* xr = x + row * C;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        v[j] = ld4<TI>(xr + j * 128 + lane * 4);
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int cc = j * 128 + lane * 4;
        const float4 g = *reinterpret_cast<const float4*>(gamma + cc);
        const float4 bb = *reinterpret_cast<const float4*>(beta + cc);
        float4 o;
        o.x = (v[j].x - mean) * rstd * g.x + bb.x;
        o.y = (v[j].y - mean) * rstd * g.y + bb.y;
        o.z = (v[j].z - mean) * rstd * g.z + bb.z;
        o.w = (v[j].w - mean) * rstd * g.w + bb.w;
        st4(out + row * C + cc, o);
        if (out2) st4(out2 + row * C + cc, o);
    }
}

template <typename TI, typename TO, typename TO2>
static int launch_ln(const void* x, const float* g, const float* b, float eps, void* out, void* out2, long long rows, int C,
                     cudaStream_t s) {
    const int grid = (int)((rows * 32 + 255) / 256);
#define A2F_LN_CASE(NV)                                                                                              \
    case NV:                                                                                                         \
        A2F_CHECK_CUDA(launch_pdl((layernorm_kernel<TI, TO, TO2, NV>), dim3(grid), dim3(256), 0, s,                  \
                                  static_cast<const TI*>(x), g, b, eps, static_cast<TO*>(out),                      \
                                  static_cast<TO2*>(out2), rows));                                                   \
        break;
    switch (C / 128) {
        A2F_LN_CASE(1) A2F_LN_CASE(2) A2F_LN_CASE(3) A2F_LN_CASE(4) A2F_LN_CASE(5) A2F_LN_CASE(6) A2F_LN_CASE(7) A2F_LN_CASE(8)
        default: return set_error(A2F_EINVAL, "a2f_layernorm: C must be a multiple of 128, at most 1024");
    }
#undef A2F_LN_CASE
    A2F_CHECK_LAUNCH("layernorm_kernel");
    count_launch();
    return A2F_OK;
}

static int conv0_nchunk(long long N) {
    const long long L0 = (N - 10) / 5 + 1;
    return (int)((L0 + MOM_TCH - 1) / MOM_TCH);
}

}  // namespace a2f

using namespace a2f;

extern "C" {

int a2f_audio_stats(const float* audio, int B, long long N, float* stats, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(audio && stats && B > 0 && N > 0, "a2f_audio_stats: bad arguments");
    A2F_CHECK_CUDA(launch_pdl(audio_stats_kernel, dim3(B), dim3(1024), 0, as_stream(stream), audio, N, stats));
    A2F_CHECK_LAUNCH("audio_stats_kernel");
    count_launch();
    return A2F_OK;
}

size_t a2f_conv0_workspace_bytes(int B, long long N) {
    if (B <= 0 || N < 10) return 0;
    const size_t partial = (size_t)B * conv0_nchunk(N) * MOM_N * sizeof(double);
    const size_t gn = (size_t)B * 512 * sizeof(float2);
    return partial + gn + 64;
}

size_t a2f_conv0_gn_offset(int B, long long N) {
    if (B <= 0 || N < 10) return 0;
    return (size_t)B * conv0_nchunk(N) * MOM_N * sizeof(double);
}

int a2f_conv0_gn_gelu(const float* audio, const float* stats, const float* w, const float* gamma, const float* beta,
                      void* out, int out_dtype, int B, long long N, void* workspace, size_t workspace_bytes,
                      void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(audio && stats && w && gamma && beta && out && workspace, "a2f_conv0_gn_gelu: NULL argument");
    A2F_REQUIRE(B > 0 && N >= 10, "a2f_conv0_gn_gelu: need B > 0 and N >= 10 samples");
    A2F_REQUIRE(workspace_bytes >= a2f_conv0_workspace_bytes(B, N), "a2f_conv0_gn_gelu: workspace too small");
    A2F_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 8 == 0, "a2f_conv0_gn_gelu: workspace must be 8-byte aligned");
    const long long L0ll = (N - 10) / 5 + 1;
    A2F_REQUIRE(L0ll < (1LL << 30), "a2f_conv0_gn_gelu: utterance too long");
    const int L0 = (int)L0ll;
    const int nchunk = conv0_nchunk(N);
    double* partial = static_cast<double*>(workspace);
    float2* gn = reinterpret_cast<float2*>(partial + (size_t)B * nchunk * MOM_N);
    cudaStream_t s = as_stream(stream);
    A2F_CHECK_CUDA(launch_pdl(conv0_moments_kernel, dim3(dim3(nchunk, B)), dim3(256), 0, s, audio, stats, N, L0, nchunk, partial));
    A2F_CHECK_LAUNCH("conv0_moments_kernel");
    A2F_CHECK_CUDA(launch_pdl(conv0_gn_stats_kernel, dim3(B), dim3(512), 0, s, partial, nchunk, w, L0, gn));
    A2F_CHECK_LAUNCH("conv0_gn_stats_kernel");
    const dim3 grid((L0 + C0_TCH - 1) / C0_TCH, B);
    const long long L0_pad = L0;   // dense [B,L0,512]
    if (out_dtype == A2F_BF16)
        A2F_CHECK_CUDA(launch_pdl(conv0_apply_kernel<bf16>, dim3(grid), dim3(256), 0, s, audio, stats, w, gn, gamma, beta, static_cast<bf16*>(out), N, L0,
                                                      L0_pad * 512));
    else
        A2F_CHECK_CUDA(launch_pdl(conv0_apply_kernel<float>, dim3(grid), dim3(256), 0, s, audio, stats, w, gn, gamma, beta, static_cast<float*>(out), N, L0,
                                                       L0_pad * 512));
    A2F_CHECK_LAUNCH("conv0_apply_kernel");
    count_launch(3);
    return A2F_OK;
}

/* inference: processor statistics and conv0 from the raw audio (stats_out [B,2] is written, not read) */
size_t a2f_conv0_auto_workspace_bytes(int B, long long N) {
    if (B <= 0 || N < 10) return 0;
    return (size_t)B * conv0_nchunk(N) * RAW_N * sizeof(double) + (size_t)B * 512 * sizeof(float2) + 64;
}

int a2f_conv0_gn_gelu_auto(const float* audio, float* stats_out, const float* w, const float* gamma, const float* beta,
                           void* out, int out_dtype, int B, long long N, void* workspace, size_t workspace_bytes,
                           void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(audio && stats_out && w && gamma && beta && out && workspace, "a2f_conv0_gn_gelu_auto: NULL argument");
    A2F_REQUIRE(B > 0 && N >= 10, "a2f_conv0_gn_gelu_auto: need B > 0 and N >= 10 samples");
    A2F_REQUIRE(workspace_bytes >= a2f_conv0_auto_workspace_bytes(B, N), "a2f_conv0_gn_gelu_auto: workspace too small");
    A2F_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 8 == 0, "a2f_conv0_gn_gelu_auto: workspace must be 8-byte aligned");
    const long long L0ll = (N - 10) / 5 + 1;
    A2F_REQUIRE(L0ll < (1LL << 30), "a2f_conv0_gn_gelu_auto: utterance too long");
    const int L0 = (int)L0ll;
    const int nchunk = conv0_nchunk(N);
    double* partial = static_cast<double*>(workspace);
    float2* gn = reinterpret_cast<float2*>(partial + (size_t)B * nchunk * RAW_N);
    cudaStream_t s = as_stream(stream);
    A2F_CHECK_CUDA(launch_pdl(conv0_raw_moments_kernel, dim3(nchunk, B), dim3(256), 0, s, audio, N, L0, nchunk, partial));
    A2F_CHECK_LAUNCH("conv0_raw_moments_kernel");
    A2F_CHECK_CUDA(launch_pdl(conv0_gn_stats_auto_kernel, dim3(B), dim3(512), 0, s, (const double*)partial, nchunk, w, N, L0,
                              stats_out, gn));
    A2F_CHECK_LAUNCH("conv0_gn_stats_auto_kernel");
    const dim3 grid((L0 + C0_TCH - 1) / C0_TCH, B);
    const long long L0_pad = L0;
    if (out_dtype == A2F_BF16)
        A2F_CHECK_CUDA(launch_pdl(conv0_apply_kernel<bf16>, dim3(grid), dim3(256), 0, s, audio, (const float*)stats_out, w,
                                  (const float2*)gn, gamma, beta, static_cast<bf16*>(out), N, L0, L0_pad * 512));
    else
        A2F_CHECK_CUDA(launch_pdl(conv0_apply_kernel<float>, dim3(grid), dim3(256), 0, s, audio, (const float*)stats_out, w,
                                  (const float2*)gn, gamma, beta, static_cast<float*>(out), N, L0, L0_pad * 512));
    A2F_CHECK_LAUNCH("conv0_apply_kernel");
    count_launch(3);
    return A2F_OK;
}

int a2f_interp_ln(const void* in, int in_dtype, const float* gamma, const float* beta, float eps, void* out,
                  int out_dtype, int B, int S, int T, int C, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(in && gamma && beta && out && B > 0 && S > 0 && T > 0, "a2f_interp_ln: bad arguments");
    A2F_REQUIRE(C == 512, "a2f_interp_ln: C must be 512 (wav2vec2 conv_dim)");
    const int grid = (int)(((long long)B * T * 32 + 255) / 256);
    cudaStream_t s = as_stream(stream);
    if (in_dtype == A2F_F32 && out_dtype == A2F_F32)
        A2F_CHECK_CUDA(launch_pdl((interp_ln_kernel<float, float, 512>), dim3(grid), dim3(256), 0, s, static_cast<const float*>(in), gamma, beta, eps,
                                                                  static_cast<float*>(out), B, S, T));
    else if (in_dtype == A2F_BF16 && out_dtype == A2F_BF16)
        A2F_CHECK_CUDA(launch_pdl((interp_ln_kernel<bf16, bf16, 512>), dim3(grid), dim3(256), 0, s, static_cast<const bf16*>(in), gamma, beta, eps,
                                                                static_cast<bf16*>(out), B, S, T));
    else if (in_dtype == A2F_F32 && out_dtype == A2F_BF16)
        A2F_CHECK_CUDA(launch_pdl((interp_ln_kernel<float, bf16, 512>), dim3(grid), dim3(256), 0, s, static_cast<const float*>(in), gamma, beta, eps,
                                                                 static_cast<bf16*>(out), B, S, T));
    else
        return set_error(A2F_EINVAL, "a2f_interp_ln: unsupported dtype combination");
    A2F_CHECK_LAUNCH("interp_ln_kernel");
    count_launch();
    return A2F_OK;
}

int a2f_layernorm(const void* x, int x_dtype, const float* gamma, const float* beta, float eps, void* out, int out_dtype,
                  void* out2, int out2_dtype, long long rows, int C, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(x && gamma && beta && out && rows >= 0, "a2f_layernorm: bad arguments");
    A2F_REQUIRE(C % 128 == 0 && C >= 128 && C <= 1024, "a2f_layernorm: C must be a multiple of 128, at most 1024");
    if (rows == 0) return A2F_OK;
    cudaStream_t s = as_stream(stream);
    const bool xi = x_dtype == A2F_BF16, oi = out_dtype == A2F_BF16, o2 = out2_dtype == A2F_BF16;
    if (!xi && !oi && (!out2 || o2)) return launch_ln<float, float, bf16>(x, gamma, beta, eps, out, out2, rows, C, s);
    if (!xi && oi && !out2) return launch_ln<float, bf16, bf16>(x, gamma, beta, eps, out, nullptr, rows, C, s);
    if (xi && oi && !out2) return launch_ln<bf16, bf16, bf16>(x, gamma, beta, eps, out, nullptr, rows, C, s);
    if (xi && !oi && !out2) return launch_ln<bf16, float, bf16>(x, gamma, beta, eps, out, nullptr, rows, C, s);
    return set_error(A2F_EINVAL, "a2f_layernorm: unsupported dtype combination");
}

}  // extern "C"

// Input side of every configuration: the two audio preparation steps the reference runs in its
// CPU DataLoader, as HBM-bound gather / polyphase-FIR kernels.
//   a2f_audio_fragments   ref:src/dataset/vocaset.py:408-430 get_audio_fragment (+ :64-69 normalize_audio): one
//                         `length`-second window per output frame, centred on the frame, zero-padded at the clip edges;
//                         int16 clips are scaled by 1/32768 on the fly.  All frames of a clip in ONE launch.
//   a2f_resample_sinc     torchaudio.functional.resample (ref:vocaset.py:279-283, ref:src/model/extractor.py:88): the
//                         strided conv1d against `new` polyphase windowed-sinc filters, out[j = i*new + p] =
//                         sum_k kernel[p][k] * xpad[i*orig + k], xpad = x zero-padded by (width, width + orig).
#include "a2f_common.cuh"

namespace a2f {

template <typename TI>
__global__ void __launch_bounds__(256) audio_fragments_kernel(const TI* __restrict__ audio, long long n_samples, int first_frame,
                                                              int n_frames, int sample_rate, int fps, int n_pad, int shift,
                                                              float scale, float* __restrict__ out) {
    pdl_sync();
    const int L = 2 * n_pad;
    const long long total = (long long)n_frames * L;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i / L), k = (int)(i - (long long)f * L);
        // pad_audio = [zeros(n_pad + shift), audio, zeros(2 n_pad)]; fragment = pad_audio[start : start + 2 n_pad]
        const long long start = (long long)(first_frame + f) * sample_rate / fps;
        const long long src = start + k - (n_pad + shift);
        float v = 0.f;
        if (src >= 0 && src < n_samples) v = (float)audio[src] * scale;
        out[i] = v;
    }
}

// one thread per output sample; the `new` filters (kw taps each) sit in shared memory when they fit (22 kHz <-> 16 kHz:
// 8 x 29 taps), otherwise they are read through L1/L2 (e.g. 16 kHz -> 22.05 kHz: 441 x 334 taps)
template <bool SMEM>
__global__ void __launch_bounds__(256) resample_sinc_kernel(const float* __restrict__ x, int B, long long N, int orig, int nnew,
                                                            const float* __restrict__ kernel, int kw, int width,
                                                            float* __restrict__ out, long long target_len) {
    extern __shared__ float rs_smem[];
    const float* rs_k = kernel;
    if (SMEM) {
        for (int i = threadIdx.x; i < nnew * kw; i += blockDim.x) rs_smem[i] = __ldg(kernel + i);
        rs_k = rs_smem;
    }
    pdl_sync();
    __syncthreads();
    const long long total = (long long)B * target_len;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(idx / target_len);
        const long long j = idx - (long long)b * target_len;
        const long long i = j / nnew;
        const int p = (int)(j - i * nnew);
        const float* xb = x + (long long)b * N;
        const float* kp = rs_k + p * kw;
        const long long base = i * orig - width;            // xpad[m] = x[m - width]
        float acc = 0.f;
        for (int k = 0; k < kw; ++k) {
            const long long m = base + k;
            if (m >= 0 && m < N) acc = fmaf(kp[k], __ldg(xb + m), acc);
        }
        out[idx] = acc;
    }
}

// Bilinear resize (align_corners = False, ATen's source-index formula) of the map M_b[c, t] = h[b, t, c] stored channels-last
// ([B, T, C], what the encoder kernels write) to out[b, i, j], i < out_h along the channel axis, j < out_w along time:
// the F.interpolate of ref:src/model/extractor.py:93-96 without materialising the transpose of :92.
template <typename TI>
__global__ void __launch_bounds__(256) bilinear_cl_kernel(const TI* __restrict__ h, int B, int T, int C, int out_h, int out_w,
                                                          float* __restrict__ out) {
    pdl_sync();
    const long long total = (long long)B * out_h * out_w;
    const float sh = (float)C / (float)out_h, sw = (float)T / (float)out_w;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % out_w);
        const int i = (int)((idx / out_w) % out_h);
        const long long b = idx / ((long long)out_w * out_h);
        float fy = sh * ((float)i + 0.5f) - 0.5f, fx = sw * ((float)j + 0.5f) - 0.5f;
        if (fy < 0.f) fy = 0.f;
        if (fx < 0.f) fx = 0.f;
        int y0 = (int)fy, x0 = (int)fx;
        if (y0 > C - 1) y0 = C - 1;
        if (x0 > T - 1) x0 = T - 1;
        const int y1 = y0 + (y0 < C - 1 ? 1 : 0), x1 = x0 + (x0 < T - 1 ? 1 : 0);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const TI* hb = h + b * T * C;
        const float v00 = ld_as_float(hb + (long long)x0 * C + y0), v01 = ld_as_float(hb + (long long)x1 * C + y0);
        const float v10 = ld_as_float(hb + (long long)x0 * C + y1), v11 = ld_as_float(hb + (long long)x1 * C + y1);
        out[idx] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
    }
}

}  // namespace a2f

using namespace a2f;

extern "C" {

int a2f_audio_fragments(const void* audio, int audio_dtype, long long n_samples, int first_frame, int n_frames, int sample_rate,
                        int fps, int n_pad, int shift, float* out, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(audio && out && n_samples > 0 && n_frames > 0 && first_frame >= 0, "a2f_audio_fragments: bad arguments");
    A2F_REQUIRE(sample_rate > 0 && fps > 0 && n_pad > 0 && n_pad + shift >= 0, "a2f_audio_fragments: bad window geometry");
    // the reference refuses a fragment that would end past its padded clip (vocaset.py:425-428)
    const long long last_start = (long long)(first_frame + n_frames - 1) * sample_rate / fps;
    A2F_REQUIRE(last_start + 2LL * n_pad <= (long long)n_pad + shift + n_samples + 2LL * n_pad,
                "a2f_audio_fragments: audio is not long enough for the last fragment");
    const long long total = (long long)n_frames * 2 * n_pad;
    long long blocks = (total + 255) / 256;
    if (blocks > 32LL * sm_count()) blocks = 32LL * sm_count();
    cudaStream_t s = as_stream(stream);
    if (audio_dtype == A2F_F32)
        A2F_CHECK_CUDA(launch_pdl(audio_fragments_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, s,
                                  static_cast<const float*>(audio), n_samples, first_frame, n_frames, sample_rate, fps, n_pad,
                                  shift, 1.0f, out));
    else if (audio_dtype == A2F_I16)
        A2F_CHECK_CUDA(launch_pdl(audio_fragments_kernel<short>, dim3((unsigned)blocks), dim3(256), 0, s,
                                  static_cast<const short*>(audio), n_samples, first_frame, n_frames, sample_rate, fps, n_pad,
                                  shift, 1.0f / 32768.0f, out));
    else return set_error(A2F_EINVAL, "a2f_audio_fragments: audio must be fp32 or int16");
    count_launch();
    return A2F_OK;
}

int a2f_bilinear_cl(const void* h, int h_dtype, int B, int T, int C, int out_h, int out_w, float* out, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(h && out && B > 0 && T > 0 && C > 0 && out_h > 0 && out_w > 0, "a2f_bilinear_cl: bad arguments");
    const long long total = (long long)B * out_h * out_w;
    long long blocks = (total + 255) / 256;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    cudaStream_t s = as_stream(stream);
    if (h_dtype == A2F_BF16)
        A2F_CHECK_CUDA(launch_pdl(bilinear_cl_kernel<bf16>, dim3((unsigned)blocks), dim3(256), 0, s, static_cast<const bf16*>(h), B, T,
                                  C, out_h, out_w, out));
    else if (h_dtype == A2F_F32)
        A2F_CHECK_CUDA(launch_pdl(bilinear_cl_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, s, static_cast<const float*>(h), B, T,
                                  C, out_h, out_w, out));
    else return set_error(A2F_EINVAL, "a2f_bilinear_cl: bad dtype");
    count_launch();
    return A2F_OK;
}

int a2f_resample_sinc(const float* x, int B, long long N, int orig, int nnew, const float* kernel, int kw, int width, float* out,
                      long long target_len, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(x && kernel && out && B > 0 && N > 0 && orig > 0 && nnew > 0 && kw > 0 && width >= 0 && target_len > 0,
                "a2f_resample_sinc: bad arguments");
    A2F_REQUIRE(kw == 2 * width + orig, "a2f_resample_sinc: kernel width must be 2*width + orig");
    A2F_REQUIRE(target_len <= (N + (long long)orig + 2LL * width - kw) / orig * nnew + nnew, "a2f_resample_sinc: target_len too long");
    const size_t smem = (size_t)nnew * kw * sizeof(float);
    const long long total = (long long)B * target_len;
    long long blocks = (total + 255) / 256;
    if (blocks > 32LL * sm_count()) blocks = 32LL * sm_count();
    if (smem <= 48 * 1024)
        A2F_CHECK_CUDA(launch_pdl(resample_sinc_kernel<true>, dim3((unsigned)blocks), dim3(256), smem, as_stream(stream), x, B, N,
                                  orig, nnew, kernel, kw, width, out, target_len));
    else
        A2F_CHECK_CUDA(launch_pdl(resample_sinc_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), x, B, N,
                                  orig, nnew, kernel, kw, width, out, target_len));
    count_launch();
    return A2F_OK;
}

}  // extern "C"

// C-ABI: a2f_gemm, a2f_posconv and the weight packers / casts that feed them.
#include "a2f_common.cuh"
#include "gemm_params.cuh"

namespace a2f {

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = __bfloat162float(in[i]);
}

// [Cout,Cin,taps] -> [Cout, taps*Cin] with k = tap*Cin + cin (the order in which a channels-last im2col row is laid out)
template <typename TO>
__global__ void pack_conv1d_kernel(const float* __restrict__ w, TO* __restrict__ out, int cout, int cin, int taps) {
    long long n = (long long)cout * cin * taps;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int tap = (int)(i % taps);
        long long r = i / taps;
        int ci = (int)(r % cin);
        int co = (int)(r / cin);
        st_from_float(out + ((long long)co * taps + tap) * cin + ci, w[i]);
    }
}

// weight_norm(dim=2): norm[tap] = sqrt(sum_{o,c} v[o,c,tap]^2)   (fp64 accumulation, one block per tap)
__global__ void posconv_norm_kernel(const float* __restrict__ v, float* __restrict__ norm) {
    const int tap = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < 768 * 48; i += blockDim.x) {
        double x = v[(long long)i * 128 + tap];
        s += x * x;
    }
    __shared__ double sh[32];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        t = warp_sum_d(t);
        if (threadIdx.x == 0) norm[tap] = (float)sqrt(t);
    }
}
// out[g][n][tap][c (kpad)] = g[tap] * v[g*48+n, c, tap] / norm[tap]   (c >= 48 -> 0)
// kpad == 8: the chunked layout of posconv_tc.cu, out[g][tap][c/8][n][c%8] (shared-memory image of one tap's B operand).
// torch's _weight_norm computes v * (g / norm): same association here.
template <typename TO>
__global__ void posconv_pack_kernel(const float* __restrict__ gw, const float* __restrict__ v,
                                    const float* __restrict__ norm, TO* __restrict__ out, int kpad) {
    const int kp = kpad == 8 ? 48 : kpad;
    const long long n = (long long)768 * 128 * kp;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int c = (int)(i % kp);
        long long r = i / kp;
        int tap = (int)(r % 128);
        int o = (int)(r / 128);   // g*48 + n
        float val = 0.f;
        if (c < 48) val = v[((long long)o * 48 + c) * 128 + tap] * (gw[tap] / norm[tap]);
        st_from_float(out + (kpad == 8 ? posconv_chunked_index(o / 48, o % 48, tap, c) : i), val);
    }
}

// Error-compensated bf16 split of an fp32 matrix for the tensor-core vertex head:
//   x = hi + lo (+ O(2^-17 |x|)),  hi = bf16(x), lo = bf16(x - hi)
// activations are laid out [hi | lo | hi], weights [hi | hi | lo], so that one K=3k bf16 GEMM evaluates
// a_hi*w_hi + a_lo*w_hi + a_hi*w_lo  (everything but the lo*lo term) with fp32 accumulation.
__global__ void split_bf16x3_kernel(const float* __restrict__ in, long long ld_in, bf16* __restrict__ out, long long rows,
                                    int K, int is_weight) {
    pdl_sync();   // PDL: wait for the previous kernel's results, let the next kernel's prologue start
    const long long n = rows * K;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const long long r = i / K;
        const int k = (int)(i - r * K);
        const float x = in[r * ld_in + k];
        const bf16 hi = __float2bfloat16_rn(x);
        const bf16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
        bf16* o = out + r * 3 * K;
        o[k] = hi;
        o[K + k] = is_weight ? hi : lo;
        o[2 * K + k] = is_weight ? lo : hi;
    }
}

// positional-conv implementation of the tcgen05 backend: 0 = posconv_tc.cu (chunked weight layout, kpad 8),
// 1 = legacy gemm_tc mode 2 (kpad 64).  a2f_debug_set_umma_field(9, v); the packers must be called with the matching kpad.
static int g_posconv_impl = 0;
int posconv_impl() { return g_posconv_impl; }
void set_posconv_impl(int v) { g_posconv_impl = v; }

static int fill_params(const a2f_gemm_args* a, GemmParams* p) {
    A2F_REQUIRE(a != nullptr, "a2f_gemm: args is NULL");
    A2F_REQUIRE(a->M >= 0 && a->N >= 0 && a->K > 0, "a2f_gemm: bad M/N/K");
    A2F_REQUIRE(a->A && a->W && a->C, "a2f_gemm: A, W and C must be non-NULL");
    A2F_REQUIRE(a->rows_per_batch > 0, "a2f_gemm: rows_per_batch must be positive");
    A2F_REQUIRE(a->tmpl == nullptr || a->rows_per_tmpl > 0, "a2f_gemm: rows_per_tmpl must be positive with tmpl");
    A2F_REQUIRE(a->a_dtype == A2F_F32 || a->a_dtype == A2F_BF16, "a2f_gemm: bad a_dtype");
    A2F_REQUIRE(a->c_dtype == A2F_F32 || a->c_dtype == A2F_BF16, "a2f_gemm: bad c_dtype");
    p->M = a->M; p->N = a->N; p->K = a->K;
    p->A = a->A; p->a_row_stride = a->a_row_stride; p->a_batch_stride = a->a_batch_stride;
    p->rows_per_batch = a->rows_per_batch;
    p->W = a->W; p->ldw = a->ldw;
    p->bias = a->bias; p->act = a->act;
    p->resid = a->resid; p->resid_bf16 = (a->resid_dtype == A2F_BF16); p->ldr = a->ldr;
    p->tmpl = a->tmpl; p->rows_per_tmpl = a->rows_per_tmpl > 0 ? a->rows_per_tmpl : 1;
    p->C = a->C; p->ldc = a->ldc;
    p->c_batch_stride = a->c_batch_stride > 0 ? a->c_batch_stride : (long long)a->rows_per_batch * a->ldc;
    p->a_rows = a->a_rows > 0 ? a->a_rows : a->rows_per_batch;
    A2F_REQUIRE(a->n_seg >= 0 && a->n_seg <= 4, "a2f_gemm: n_seg must be 0..4");
    p->n_seg = a->n_seg <= 1 ? 0 : a->n_seg;
    A2F_REQUIRE(p->n_seg == 0 || a->K % p->n_seg == 0, "a2f_gemm: K must be a multiple of n_seg");
    for (int i = 0; i < 4; ++i) {
        p->seg_row_off[i] = a->seg_row_off[i];
        p->seg_col_off[i] = a->seg_col_off[i];
    }
    if (a->n_seg == 1) {   // one explicit segment: still honours its row / column offset
        p->n_seg = 1;
    }
    p->r_batch_stride = a->r_batch_stride > 0 ? a->r_batch_stride : (long long)a->rows_per_batch * a->ldr;
    p->resid_mode = a->resid_mode;
    A2F_REQUIRE(a->resid_mode == A2F_RESID_ADD || (a->resid_mode == A2F_RESID_DACT && a->resid != nullptr),
                "a2f_gemm: bad resid_mode");
    p->C2 = a->C2;
    p->ldc2 = a->ldc2;
    A2F_REQUIRE(a->C2 == nullptr || (a->resid == nullptr && a->tmpl == nullptr && a->ldc2 > 0 && a->act != A2F_ACT_NONE),
                "a2f_gemm: C2 (pre-activation + activation outputs) needs an activation and excludes resid / tmpl");
    return A2F_OK;
}

}  // namespace a2f

using namespace a2f;

extern "C" {

int a2f_gemm(const a2f_gemm_args* args, int backend, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    GemmParams p;
    rc = fill_params(args, &p);
    if (rc != A2F_OK) return rc;
    if (backend == A2F_BACKEND_SIMT_F32) {
        A2F_REQUIRE(p.C2 == nullptr, "a2f_gemm: C2 is a tcgen05 back-end feature");
        return gemm_simt(p, args->a_dtype == A2F_BF16, args->c_dtype == A2F_BF16, as_stream(stream));
    } else if (backend == A2F_BACKEND_TCGEN05) {
        A2F_REQUIRE(args->a_dtype == A2F_BF16, "a2f_gemm: the tcgen05 backend takes bf16 operands");
        return gemm_tc(p, args->c_dtype == A2F_BF16, 0, as_stream(stream));
    }
    return set_error(A2F_EINVAL, "a2f_gemm: unknown backend");
}

int a2f_gemm_wgrad(const a2f_wgrad_args* a, int backend, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(a != nullptr, "a2f_gemm_wgrad: args is NULL");
    A2F_REQUIRE(a->M >= 0 && a->N > 0 && a->K > 0, "a2f_gemm_wgrad: bad M/N/K");
    A2F_REQUIRE(a->dY && a->X && a->dW, "a2f_gemm_wgrad: dY, X and dW must be non-NULL");
    A2F_REQUIRE(a->rows_per_batch > 0 && a->M % a->rows_per_batch == 0, "a2f_gemm_wgrad: M must be a multiple of rows_per_batch");
    A2F_REQUIRE(a->n_seg >= 0 && (a->n_seg <= 4 || a->x_row_step != 0), "a2f_gemm_wgrad: n_seg > 4 needs x_row_step");
    A2F_REQUIRE(a->dtype == A2F_F32 || a->dtype == A2F_BF16, "a2f_gemm_wgrad: bad dtype");
    if (a->M == 0) return A2F_OK;
    WgradParams p;
    p.M = a->M; p.N = a->N; p.K = a->K;
    p.dY = a->dY; p.dy_row_stride = a->dy_row_stride; p.dy_batch_stride = a->dy_batch_stride;
    p.X = a->X; p.x_row_stride = a->x_row_stride; p.x_batch_stride = a->x_batch_stride;
    p.rows_per_batch = a->rows_per_batch;
    p.x_rows = a->x_rows > 0 ? a->x_rows : a->rows_per_batch;
    p.n_seg = a->n_seg < 1 ? 1 : a->n_seg;
    for (int i = 0; i < 4; ++i) {
        p.x_row_off[i] = a->n_seg < 1 ? 0 : a->x_row_off[i];
        p.x_col_off[i] = a->n_seg < 1 ? 0 : a->x_col_off[i];
    }
    p.dW = a->dW; p.ldw = a->ldw;
    p.x_row_step = a->n_seg > 1 ? a->x_row_step : 0;
    if (backend == A2F_BACKEND_SIMT_F32) return wgrad_simt(p, a->dtype == A2F_BF16, as_stream(stream));
    if (backend == A2F_BACKEND_TCGEN05) {
        A2F_REQUIRE(a->dtype == A2F_BF16, "a2f_gemm_wgrad: the tcgen05 backend takes bf16 operands");
        return wgrad_tc(p, as_stream(stream));
    }
    return set_error(A2F_EINVAL, "a2f_gemm_wgrad: unknown backend");
}

int a2f_posconv(const void* h, int h_dtype, const void* Wp, const float* bias, void* out, int out_dtype, int B, int T,
                int backend, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(h && Wp && out && B > 0 && T > 0, "a2f_posconv: bad arguments");
    GemmParams p;
    p.M = B * T; p.N = 48;
    p.A = h; p.a_row_stride = 768; p.a_batch_stride = (long long)T * 768; p.rows_per_batch = T;
    p.W = Wp;
    p.bias = bias; p.act = A2F_ACT_GELU;
    p.resid = h; p.resid_bf16 = (h_dtype == A2F_BF16); p.ldr = 768;
    p.tmpl = nullptr; p.rows_per_tmpl = 1;
    p.C = out; p.ldc = 768; p.c_batch_stride = (long long)T * 768;
    p.a_rows = T; p.n_seg = 0; p.r_batch_stride = (long long)T * 768; p.resid_mode = A2F_RESID_ADD;
    for (int i = 0; i < 4; ++i) p.seg_row_off[i] = p.seg_col_off[i] = 0;
    if (backend == A2F_BACKEND_SIMT_F32) {
        p.K = 128 * 48; p.ldw = 128 * 48;
        return posconv_simt(p, h_dtype == A2F_BF16, out_dtype == A2F_BF16, as_stream(stream));
    } else if (backend == A2F_BACKEND_TCGEN05) {
        A2F_REQUIRE(h_dtype == A2F_BF16, "a2f_posconv: the tcgen05 backend takes bf16 activations");
        if (posconv_impl() == 0) {
            A2F_REQUIRE(out_dtype == A2F_BF16, "a2f_posconv: the tcgen05 backend writes bf16");
            return posconv_tc(h, Wp, bias, nullptr, 1, out, B, T, -64, A2F_ACT_GELU, as_stream(stream));
        }
        p.K = 128 * 64; p.ldw = 128 * 64;
        return gemm_tc(p, out_dtype == A2F_BF16, 2, as_stream(stream));
    }
    return set_error(A2F_EINVAL, "a2f_posconv: unknown backend");
}

static int posconv_common(GemmParams& p, const void* a, int dtype, const void* W, void* out, int B, int T, int backend,
                          cudaStream_t s) {
    p.M = B * T; p.N = 48;
    p.A = a; p.a_row_stride = 768; p.a_batch_stride = (long long)T * 768; p.rows_per_batch = T;
    p.W = W;
    p.tmpl = nullptr; p.rows_per_tmpl = 1;
    p.C = out; p.ldc = 768; p.c_batch_stride = (long long)T * 768;
    p.ldr = 768;
    if (backend == A2F_BACKEND_SIMT_F32) {
        p.K = 128 * 48; p.ldw = 128 * 48;
        return posconv_simt(p, dtype == A2F_BF16, dtype == A2F_BF16, s);
    } else if (backend == A2F_BACKEND_TCGEN05) {
        A2F_REQUIRE(dtype == A2F_BF16, "posconv: the tcgen05 backend takes bf16 activations");
        if (posconv_impl() == 0)
            return posconv_tc(a, W, p.bias, p.resid, p.resid ? 2 : 0, out, B, T, -64 + p.seg_row_off[0], p.act, s);
        p.K = 128 * 64; p.ldw = 128 * 64;
        return gemm_tc(p, 1, 2, s);
    }
    return set_error(A2F_EINVAL, "posconv: unknown backend");
}

int a2f_posconv_pre(const void* h, int h_dtype, const void* Wp, const float* bias, void* pc, int B, int T, int backend,
                    void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(h && Wp && pc && B > 0 && T > 0, "a2f_posconv_pre: bad arguments");
    GemmParams p;
    p.bias = bias; p.act = A2F_ACT_NONE; p.resid = nullptr; p.resid_bf16 = 0;
    return posconv_common(p, h, h_dtype, Wp, pc, B, T, backend, as_stream(stream));
}

int a2f_posconv_dgrad(const void* dpc, int dtype, const void* Wd, const void* dout, void* dh, int B, int T, int backend,
                      void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(dpc && Wd && dh && B > 0 && T > 0, "a2f_posconv_dgrad: bad arguments");
    GemmParams p;
    p.bias = nullptr; p.act = A2F_ACT_NONE;
    p.resid = dout; p.resid_bf16 = (dtype == A2F_BF16);
    p.seg_row_off[0] = 1;     // flipped taps: dh[s] = sum_tap' Wd[tap'] dpc[s + tap' - 63]
    return posconv_common(p, dpc, dtype, Wd, dh, B, T, backend, as_stream(stream));
}

int a2f_posconv_wgrad(const void* dpc, const void* h, int dtype, float* dWp, int B, int T, int backend, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(dpc && h && dWp && B > 0 && T > 0, "a2f_posconv_wgrad: bad arguments");
    const size_t esz = dtype == A2F_BF16 ? 2 : 4;
    for (int g = 0; g < 16; ++g) {
        WgradParams p;
        p.M = B * T; p.N = 48; p.K = 48;
        p.dY = static_cast<const char*>(dpc) + (size_t)g * 48 * esz; p.dy_row_stride = 768; p.dy_batch_stride = (long long)T * 768;
        p.X = static_cast<const char*>(h) + (size_t)g * 48 * esz; p.x_row_stride = 768; p.x_batch_stride = (long long)T * 768;
        p.rows_per_batch = T; p.x_rows = T;
        p.n_seg = 128; p.x_row_step = 1;
        for (int i = 0; i < 4; ++i) p.x_row_off[i] = p.x_col_off[i] = 0;
        p.x_row_off[0] = -64;
        p.dW = dWp + (size_t)g * 48 * 6144; p.ldw = 6144;
        if (backend == A2F_BACKEND_SIMT_F32) rc = wgrad_simt(p, dtype == A2F_BF16, as_stream(stream));
        else if (backend == A2F_BACKEND_TCGEN05) {
            A2F_REQUIRE(dtype == A2F_BF16, "a2f_posconv_wgrad: the tcgen05 backend takes bf16 operands");
            rc = wgrad_tc(p, as_stream(stream));
        } else return set_error(A2F_EINVAL, "a2f_posconv_wgrad: unknown backend");
        if (rc != A2F_OK) return rc;
    }
    return A2F_OK;
}

int a2f_pack_posconv_weight(const float* g, const float* v, void* out, int out_dtype, int kpad, float* norm_scratch,
                            void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(g && v && out && norm_scratch && (kpad == 48 || kpad == 64 || kpad == 8), "a2f_pack_posconv_weight: bad arguments");
    float* norm = norm_scratch;
    posconv_norm_kernel<<<128, 256, 0, as_stream(stream)>>>(v, norm);
    A2F_CHECK_LAUNCH("posconv_norm_kernel");
    const long long n = (long long)768 * 128 * (kpad == 8 ? 48 : kpad);
    const int grid = (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
    if (out_dtype == A2F_BF16)
        posconv_pack_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(g, v, norm, static_cast<bf16*>(out), kpad);
    else
        posconv_pack_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(g, v, norm, static_cast<float*>(out), kpad);
    A2F_CHECK_LAUNCH("posconv_pack_kernel");
    count_launch(2);
    return A2F_OK;
}

int a2f_pack_conv1d_weight(const float* w, void* out, int out_dtype, int cout, int cin, int taps, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(w && out && cout > 0 && cin > 0 && taps > 0, "a2f_pack_conv1d_weight: bad arguments");
    const long long n = (long long)cout * cin * taps;
    const int grid = (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
    if (out_dtype == A2F_BF16)
        pack_conv1d_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(w, static_cast<bf16*>(out), cout, cin, taps);
    else
        pack_conv1d_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(w, static_cast<float*>(out), cout, cin, taps);
    A2F_CHECK_LAUNCH("pack_conv1d_kernel");
    count_launch();
    return A2F_OK;
}

int a2f_cast_f32_to_bf16(const float* in, void* out, long long n, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    if (n <= 0) return A2F_OK;
    const int grid = (int)((n + 255) / 256 > 8192 ? 8192 : (n + 255) / 256);
    cast_f32_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, static_cast<bf16*>(out), n);
    A2F_CHECK_LAUNCH("cast_f32_bf16_kernel");
    count_launch();
    return A2F_OK;
}
int a2f_split_bf16x3(const float* in, long long ld_in, void* out, long long rows, int K, int is_weight, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(in && out && rows >= 0 && K > 0 && ld_in >= K, "a2f_split_bf16x3: bad arguments");
    if (rows == 0) return A2F_OK;
    const long long n = rows * K;
    const int grid = (int)((n + 255) / 256 > 8192 ? 8192 : (n + 255) / 256);
    A2F_CHECK_CUDA(launch_pdl(split_bf16x3_kernel, dim3(grid), dim3(256), 0, as_stream(stream), in, ld_in, static_cast<bf16*>(out), rows, K, is_weight));
    A2F_CHECK_LAUNCH("split_bf16x3_kernel");
    count_launch();
    return A2F_OK;
}
int a2f_cast_bf16_to_f32(const void* in, float* out, long long n, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    if (n <= 0) return A2F_OK;
    const int grid = (int)((n + 255) / 256 > 8192 ? 8192 : (n + 255) / 256);
    cast_bf16_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const bf16*>(in), out, n);
    A2F_CHECK_LAUNCH("cast_bf16_f32_kernel");
    count_launch();
    return A2F_OK;
}

}  // extern "C"

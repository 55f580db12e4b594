// Backward through FaceFormer's autoregressive rollout (ref:src/model/faceformer.py:154-185 trained by free rollout:
// the reference's forward() takes no ground truth, so dL/d(e_{i+1}) flows back into d_i through the feedback
// 64->15069->64, collapsed here to the 64x64 matrix Wc of a2f_pack_feedback).
//
// One persistent CTA per utterance walks the frames in REVERSE, carrying dL/de_{i+1} and the running dK / dV of every
// cached key (complete for key j once queries j..T-1 have been processed).  Per step it back-propagates one 64-vector
// through LN3, the FFN, LN2, LN1, the self-attention of query i (probabilities recomputed from q_i, the K cache and the
// saved log-sum-exp; sum_j p_j dp_j = dctx . ctx, so no reduction is needed for the softmax backward), and the in-proj.
// Every transposed 64-wide weight slice lives in registers for the whole walk, as in the forward kernel.  The kernel
// only emits per-step gradient VECTORS (a2f.h A2F_DECG_*); all weight gradients are batched fp32 GEMMs over [B*T] rows
// afterwards, so no outer product sits on the sequential critical path.
#include "a2f_common.cuh"
#include "gemm_params.cuh"

namespace a2f {

constexpr int DB_THREADS = 512;
constexpr int DB_LD = 68;            // padded dK / dV accumulator rows in shared memory
constexpr int DB_SMEM_T = 352;       // longest clip whose accumulators fit in shared memory

struct DecBwdW {
    const float *sa_in_w, *sa_out_w, *lin1_w, *lin2_w, *n1_w, *n2_w, *n3_w, *fb_w;
};
struct DecBwdIn {
    const float *X, *Q, *K, *V, *CTX, *Y1PRE, *Y2PRE, *HID, *Y3PRE, *LSE, *gD;
};
struct DecBwdOut {
    float *GD, *G3, *GHID, *GY2, *G2, *G1, *GQKV, *DEFB, *DSTYLE;
};

// same fast exp / division-free floor((i - j) / period) as the forward rollout (decoder.cu): the recomputed probabilities must
// be the ones the forward produced
A2F_D float dec_exp(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}
A2F_D int dec_div(int d, unsigned magic) { return (int)__umulhi((unsigned)d, magic); }

A2F_D float dot64(const float* w, const float* __restrict__ x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < 64; k += 4) {
        const float4 f = *reinterpret_cast<const float4*>(x + k);
        a0 = fmaf(w[k], f.x, a0);
        a1 = fmaf(w[k + 1], f.y, a1);
        a2 = fmaf(w[k + 2], f.z, a2);
        a3 = fmaf(w[k + 3], f.w, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

// LayerNorm(64) backward for the vector whose elements (lane, lane+32) the lanes hold.  y: LN input, g: upstream grad.
A2F_D void warp_ln64_bwd(float ya, float yb, float ga, float gb, float w0, float w1, float& oa, float& ob) {
    float s1 = ya + yb, s2 = fmaf(ya, ya, yb * yb);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float mean = s1 * (1.f / 64.f);
    const float var = fmaxf(fmaf(-mean, mean, s2 * (1.f / 64.f)), 0.f);
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    const float xa = (ya - mean) * rstd, xb = (yb - mean) * rstd;
    const float ha = ga * w0, hb = gb * w1;
    float m1 = ha + hb, m2 = fmaf(ha, xa, hb * xb);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m1 += __shfl_xor_sync(0xffffffffu, m1, o);
        m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    m1 *= (1.f / 64.f);
    m2 *= (1.f / 64.f);
    oa = rstd * (ha - m1 - xa * m2);
    ob = rstd * (hb - m1 - xb * m2);
}

__global__ void __launch_bounds__(DB_THREADS, 1)
decoder_bwd_kernel(DecBwdW w, DecBwdIn in, DecBwdOut out, int T, int period, float* __restrict__ acc_global) {
    extern __shared__ __align__(16) float bsm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    const int Tpad = (T + 3) & ~3;

    float* de_s = bsm;                 // [64]  dL/de_{i+1}
    float* gd_s = de_s + 64;           // [64]  dL/dd_i
    float* g3c = gd_s + 64;            // [4][64] per-warp copies of dL/d(LN3 input)
    float* ghid_s = g3c + 256;         // [128]
    float* dy2_s = ghid_s + 128;       // [64]
    float* g1c = dy2_s + 64;           // [2][64]
    float* dctx_s = g1c + 128;         // [64]
    float* dq_s = dctx_s + 64;         // [64]  raw-q gradient
    float* part_s = dq_s + 64;         // [3][64]
    float* q_s = part_s + 192;         // [64]  scaled query of step i
    float* ctx_s = q_s + 64;           // [64]
    float* lse_s = ctx_s + 64;         // [4] (+4 pad)
    float* pvp = lse_s + 8;            // [4][4][16]
    float* wc = pvp + 256;             // [64][64] feedback matrix, natural layout wc[r*64+k] = Wc[r][k]
    float* sc = wc + 4096;             // [4][Tpad] dS values
    float* dKa;
    float* dVa;
    int ald;
    if (acc_global) {
        dKa = acc_global + (long long)b * 2 * T * 64;
        dVa = dKa + (long long)T * 64;
        ald = 64;
    } else {
        dKa = sc + 4 * Tpad;
        dVa = dKa + (long long)T * DB_LD;
        ald = DB_LD;
        for (int idx = tid; idx < 2 * T * DB_LD; idx += DB_THREADS) dKa[idx] = 0.f;
    }

    // ---- transposed weight slices into registers ----
    float wr[64];
    if (tid < 192) {                       // in_proj^T: part p (q|k|v), input column k
        const int p = tid >> 6, k = tid & 63;
#pragma unroll
        for (int m = 0; m < 64; ++m) wr[m] = w.sa_in_w[(p * 64 + m) * 64 + k];
    } else if (tid < 256) {                // out_proj^T
        const int k = tid - 192;
#pragma unroll
        for (int t = 0; t < 64; ++t) wr[t] = w.sa_out_w[t * 64 + k];
    } else if (tid < 384) {                // linear2^T: column j of W2 [64,128]
        const int j = tid - 256;
#pragma unroll
        for (int r = 0; r < 64; ++r) wr[r] = w.lin2_w[r * 128 + j];
    } else {                               // linear1^T: output k, half of the 128 hidden units
        const int u = tid - 384, k = u >> 1, half = u & 1;
#pragma unroll
        for (int m = 0; m < 64; ++m) wr[m] = w.lin1_w[(half * 64 + m) * 64 + k];
    }
    for (int idx = tid; idx < 4096; idx += DB_THREADS) wc[idx] = w.fb_w[idx];
    if (tid < 64) de_s[tid] = 0.f;
    float lnw0 = 0.f, lnw1 = 0.f, lnv0 = 0.f, lnv1 = 0.f;     // LN weights of the lanes that run an LN backward
    if (tid >= 256 && tid < 384) { lnw0 = w.n3_w[lane]; lnw1 = w.n3_w[lane + 32]; }
    else if (tid >= 192 && tid < 256) {
        lnw0 = w.n2_w[lane]; lnw1 = w.n2_w[lane + 32];
        lnv0 = w.n1_w[lane]; lnv1 = w.n1_w[lane + 32];
    }
    float dstyle0 = 0.f, dstyle1 = 0.f;
    __syncthreads();

    const unsigned pmagic = 0xFFFFFFFFu / (unsigned)period + 1u;
    const long long bt0 = (long long)b * T;
    const float* Kg = in.K + bt0 * 64;
    const float* Vg = in.V + bt0 * 64;

    for (int i = T - 1; i >= 0; --i) {
        const long long row = bt0 + i;
        // ---- operands of this step: registers (prefetch) and shared memory ----
        float pa = 0.f, pb = 0.f, pc = 0.f, pd = 0.f;
        if (tid < 64) q_s[tid] = in.Q[row * 64 + tid];
        else if (tid < 128) ctx_s[tid - 64] = in.CTX[row * 64 + tid - 64];
        else if (tid < 132) lse_s[tid - 128] = in.LSE[row * 4 + tid - 128];
        else if (tid >= 256 && tid < 384) {
            pa = in.Y3PRE[row * 64 + lane];
            pb = in.Y3PRE[row * 64 + lane + 32];
            pc = in.HID[row * 128 + (tid - 256)];
        } else if (tid >= 192 && tid < 256) {
            pa = in.Y2PRE[row * 64 + lane];
            pb = in.Y2PRE[row * 64 + lane + 32];
            pc = in.Y1PRE[row * 64 + lane];
            pd = in.Y1PRE[row * 64 + lane + 32];
        }
        // ---------- P1 (warp 6): dL/dd_i = head gradient + Wc^T dL/de_{i+1} ----------
        if (warp == 6) {
            float a0 = in.gD[row * 64 + lane], a1 = in.gD[row * 64 + lane + 32];
#pragma unroll 8
            for (int r = 0; r < 64; ++r) {
                const float d = de_s[r];
                a0 = fmaf(wc[r * 64 + lane], d, a0);
                a1 = fmaf(wc[r * 64 + lane + 32], d, a1);
            }
            gd_s[lane] = a0;
            gd_s[lane + 32] = a1;
            out.GD[row * 64 + lane] = a0;
            out.GD[row * 64 + lane + 32] = a1;
        }
        __syncthreads();
        // ---------- P2 (warps 8..11): LN3 backward, linear2^T, ReLU mask ----------
        if (tid >= 256 && tid < 384) {
            const int wl = warp - 8;
            float oa, ob;
            warp_ln64_bwd(pa, pb, gd_s[lane], gd_s[lane + 32], lnw0, lnw1, oa, ob);
            float* gc = g3c + wl * 64;
            gc[lane] = oa;
            gc[lane + 32] = ob;
            if (wl == 0) {
                out.G3[row * 64 + lane] = oa;
                out.G3[row * 64 + lane + 32] = ob;
            }
            __syncwarp();
            const float t = dot64(wr, gc);
            const float gh = pc > 0.f ? t : 0.f;
            ghid_s[tid - 256] = gh;
            out.GHID[row * 128 + (tid - 256)] = gh;
        }
        __syncthreads();
        // ---------- P3 (warps 12..15): dL/d(LN2 output) = g3 + linear1^T ghid ----------
        if (tid >= 384) {
            const int u = tid - 384, k = u >> 1, half = u & 1;
            float part = dot64(wr, ghid_s + half * 64);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            if (half == 0) {
                const float v = g3c[k] + part;
                dy2_s[k] = v;
                out.GY2[row * 64 + k] = v;
            }
        }
        __syncthreads();
        // ---------- P4 (warps 6,7): LN2 backward, LN1 backward, out_proj^T ----------
        if (tid >= 192 && tid < 256) {
            const int wl = warp - 6;
            float g2a, g2b, g1a, g1b;
            warp_ln64_bwd(pa, pb, dy2_s[lane], dy2_s[lane + 32], lnw0, lnw1, g2a, g2b);
            warp_ln64_bwd(pc, pd, g2a, g2b, lnv0, lnv1, g1a, g1b);
            float* gc = g1c + wl * 64;
            gc[lane] = g1a;
            gc[lane + 32] = g1b;
            if (wl == 0) {
                out.G2[row * 64 + lane] = g2a;
                out.G2[row * 64 + lane + 32] = g2b;
                out.G1[row * 64 + lane] = g1a;
                out.G1[row * 64 + lane + 32] = g1b;
            }
            __syncwarp();
            dctx_s[tid - 192] = dot64(wr, gc);
        }
        __syncthreads();
        // ---------- P5 (all): self-attention backward of query i ----------
        {
            const int h = tid >> 7, u = tid & 127, wq = (tid >> 5) & 3;
            const float slope = (h == 0) ? 0.25f : (h == 1) ? 0.0625f : (h == 2) ? 0.015625f : 0.00390625f;
            float qh[16], dch[16];
            float Dh = 0.f;
#pragma unroll
            for (int d = 0; d < 16; ++d) {
                qh[d] = q_s[h * 16 + d];
                dch[d] = dctx_s[h * 16 + d];
                Dh = fmaf(dch[d], ctx_s[h * 16 + d], Dh);
            }
            const float lse = lse_s[h];
            float* sch = sc + h * Tpad;
            for (int j = u; j <= i; j += 128) {
                const float* kp = Kg + (long long)j * 64 + h * 16;
                const float* vp = Vg + (long long)j * 64 + h * 16;
                float kk[16], vv[16];
#pragma unroll
                for (int d = 0; d < 16; d += 4) {
                    const float4 f = __ldg(reinterpret_cast<const float4*>(kp + d));
                    const float4 g = __ldg(reinterpret_cast<const float4*>(vp + d));
                    kk[d] = f.x; kk[d + 1] = f.y; kk[d + 2] = f.z; kk[d + 3] = f.w;
                    vv[d] = g.x; vv[d + 1] = g.y; vv[d + 2] = g.z; vv[d + 3] = g.w;
                }
                float s = 0.f, dp = 0.f;
#pragma unroll
                for (int d = 0; d < 16; ++d) {
                    s = fmaf(qh[d], kk[d], s);
                    dp = fmaf(dch[d], vv[d], dp);
                }
                s -= slope * (float)dec_div(i - j, pmagic);
                const float p = dec_exp(s - lse);
                const float ds = p * (dp - Dh);
                sch[j] = ds;
                float* dkp = dKa + (long long)j * ald + h * 16;
                float* dvp = dVa + (long long)j * ald + h * 16;
#pragma unroll
                for (int d = 0; d < 16; d += 4) {
                    float4 a = *reinterpret_cast<float4*>(dkp + d);
                    float4 c = *reinterpret_cast<float4*>(dvp + d);
                    a.x = fmaf(ds, qh[d], a.x); a.y = fmaf(ds, qh[d + 1], a.y);
                    a.z = fmaf(ds, qh[d + 2], a.z); a.w = fmaf(ds, qh[d + 3], a.w);
                    c.x = fmaf(p, dch[d], c.x); c.y = fmaf(p, dch[d + 1], c.y);
                    c.z = fmaf(p, dch[d + 2], c.z); c.w = fmaf(p, dch[d + 3], c.w);
                    *reinterpret_cast<float4*>(dkp + d) = a;
                    *reinterpret_cast<float4*>(dvp + d) = c;
                }
            }
            named_bar_sync(1 + h, 128);
            // dq_h[d] = 0.25 * sum_j ds_j K_j[d]  (the saved query is already scaled by 0.25)
            const int d = u & 15, jg = u >> 4;
            float acc = 0.f;
            for (int j = jg; j <= i; j += 8) acc = fmaf(sch[j], __ldg(Kg + (long long)j * 64 + h * 16 + d), acc);
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            if (lane < 16) pvp[(h * 4 + wq) * 16 + lane] = acc;
            named_bar_sync(1 + h, 128);
            if (u < 16) {
                const float* pp = pvp + h * 64 + u;
                dq_s[h * 16 + u] = 0.25f * ((pp[0] + pp[16]) + (pp[32] + pp[48]));
            }
        }
        __syncthreads();
        // ---------- P6 (warps 0..5): in_proj^T [dq | dK_i | dV_i] ----------
        if (tid < 192) {
            const int p = tid >> 6, k = tid & 63;
            const float* src = (p == 0) ? dq_s : (p == 1) ? dKa + (long long)i * ald : dVa + (long long)i * ald;
            out.GQKV[row * 192 + tid] = src[k];
            part_s[p * 64 + k] = dot64(wr, src);
        }
        __syncthreads();
        // ---------- P7 (warp 6): dL/de_i; carried to step i-1 by the same warp ----------
        if (warp == 6) {
            const float e0 = g1c[lane] + ((part_s[lane] + part_s[64 + lane]) + part_s[128 + lane]);
            const float e1 = g1c[lane + 32] + ((part_s[lane + 32] + part_s[96 + lane]) + part_s[160 + lane]);
            de_s[lane] = e0;
            de_s[lane + 32] = e1;
            dstyle0 += e0;
            dstyle1 += e1;
            out.DEFB[row * 64 + lane] = i > 0 ? e0 : 0.f;
            out.DEFB[row * 64 + lane + 32] = i > 0 ? e1 : 0.f;
            __syncwarp();
        }
    }
    if (warp == 6) {
        out.DSTYLE[b * 64 + lane] = dstyle0;
        out.DSTYLE[b * 64 + lane + 32] = dstyle1;
    }
}

// dgamma += sum_rows dy*xhat, dbeta += sum_rows dy for LayerNorm(64); one warp per row, rows strided over the grid.
__global__ void __launch_bounds__(256) ln64_param_grad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                              long long rows, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ga = 0.f, gb = 0.f, ba = 0.f, bb = 0.f;
    for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += (long long)gridDim.x * 8) {
        const float ya = x[r * 64 + lane], yb = x[r * 64 + lane + 32];
        float s1 = ya + yb, s2 = fmaf(ya, ya, yb * yb);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float mean = s1 * (1.f / 64.f);
        const float var = fmaxf(fmaf(-mean, mean, s2 * (1.f / 64.f)), 0.f);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        const float da = dy[r * 64 + lane], db = dy[r * 64 + lane + 32];
        ga = fmaf(da, (ya - mean) * rstd, ga);
        gb = fmaf(db, (yb - mean) * rstd, gb);
        ba += da;
        bb += db;
    }
    __shared__ float sh[8][128];
    sh[warp][lane] = ga; sh[warp][lane + 32] = gb; sh[warp][64 + lane] = ba; sh[warp][96 + lane] = bb;
    __syncthreads();
    if (threadIdx.x < 128) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x];
        atomicAdd((threadIdx.x < 64 ? dgamma : dbeta) + (threadIdx.x & 63), t);
    }
}

static size_t db_smem_bytes(int T, bool acc_in_smem) {
    const int Tpad = (T + 3) & ~3;
    size_t fl = 64 * 2 + 256 + 128 + 64 + 128 + 64 + 64 + 192 + 64 + 64 + 8 + 256 + 4096 + (size_t)4 * Tpad;
    if (acc_in_smem) fl += (size_t)2 * T * DB_LD;
    return fl * sizeof(float);
}

}  // namespace a2f

using namespace a2f;

extern "C" {

int a2f_decoder_grad_offset(int field) {
    static const int width[A2F_DECG_NFIELDS] = {64, 64, 128, 64, 64, 64, 192, 64};
    if (field < 0 || field > A2F_DECG_NFIELDS) return -1;
    int o = 0;
    for (int i = 0; i < field; ++i) o += width[i];
    return o;
}

size_t a2f_decoder_bwd_workspace_bytes(int B, int T) {
    if (B <= 0 || T <= 0) return 0;
    return T > DB_SMEM_T ? (size_t)2 * B * T * 64 * sizeof(float) : 16;
}

int a2f_decoder_rollout_bwd(const a2f_decoder_weights* w, const float* saves, const float* gD, int period, float* grads,
                            int B, int T, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(w && saves && gD && grads && workspace, "a2f_decoder_rollout_bwd: NULL argument");
    A2F_REQUIRE(B > 0 && T > 0 && period > 0 && T <= 8192, "a2f_decoder_rollout_bwd: bad sizes");
    A2F_REQUIRE(workspace_bytes >= a2f_decoder_bwd_workspace_bytes(B, T), "a2f_decoder_rollout_bwd: workspace too small");
    A2F_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "a2f_decoder_rollout_bwd: workspace must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    const size_t bt = (size_t)B * T;
    DecBwdW dw;
    dw.sa_in_w = w->sa_in_w; dw.sa_out_w = w->sa_out_w; dw.lin1_w = w->lin1_w; dw.lin2_w = w->lin2_w;
    dw.n1_w = w->n1_w; dw.n2_w = w->n2_w; dw.n3_w = w->n3_w; dw.fb_w = w->fb_w;
    DecBwdIn in;
    in.X = saves + bt * a2f_decoder_save_offset(A2F_DEC_X); in.Q = saves + bt * a2f_decoder_save_offset(A2F_DEC_Q);
    in.K = saves + bt * a2f_decoder_save_offset(A2F_DEC_K); in.V = saves + bt * a2f_decoder_save_offset(A2F_DEC_V);
    in.CTX = saves + bt * a2f_decoder_save_offset(A2F_DEC_CTX); in.Y1PRE = saves + bt * a2f_decoder_save_offset(A2F_DEC_Y1PRE);
    in.Y2PRE = saves + bt * a2f_decoder_save_offset(A2F_DEC_Y2PRE); in.HID = saves + bt * a2f_decoder_save_offset(A2F_DEC_HID);
    in.Y3PRE = saves + bt * a2f_decoder_save_offset(A2F_DEC_Y3PRE); in.LSE = saves + bt * a2f_decoder_save_offset(A2F_DEC_LSE);
    in.gD = gD;
    DecBwdOut out;
    out.GD = grads + bt * a2f_decoder_grad_offset(A2F_DECG_GD); out.G3 = grads + bt * a2f_decoder_grad_offset(A2F_DECG_G3);
    out.GHID = grads + bt * a2f_decoder_grad_offset(A2F_DECG_GHID); out.GY2 = grads + bt * a2f_decoder_grad_offset(A2F_DECG_GY2);
    out.G2 = grads + bt * a2f_decoder_grad_offset(A2F_DECG_G2); out.G1 = grads + bt * a2f_decoder_grad_offset(A2F_DECG_G1);
    out.GQKV = grads + bt * a2f_decoder_grad_offset(A2F_DECG_GQKV); out.DEFB = grads + bt * a2f_decoder_grad_offset(A2F_DECG_DEFB);
    out.DSTYLE = grads + bt * a2f_decoder_grad_offset(A2F_DECG_NFIELDS);
    float* acc = nullptr;
    if (T > DB_SMEM_T) {
        acc = static_cast<float*>(workspace);
        A2F_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)2 * B * T * 64 * sizeof(float), s));
    }
    const size_t smem = db_smem_bytes(T, acc == nullptr);
    A2F_CHECK_CUDA(cudaFuncSetAttribute(decoder_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    decoder_bwd_kernel<<<B, DB_THREADS, smem, s>>>(dw, in, out, T, period, acc);
    A2F_CHECK_LAUNCH("decoder_bwd_kernel");
    count_launch();
    return A2F_OK;
}

int a2f_ln64_param_grad(const float* dy, const float* x, long long rows, float* dgamma, float* dbeta, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(dy && x && dgamma && dbeta && rows >= 0, "a2f_ln64_param_grad: bad arguments");
    if (rows == 0) return A2F_OK;
    long long blocks = (rows + 7) / 8;
    if (blocks > 2LL * sm_count()) blocks = 2LL * sm_count();
    ln64_param_grad_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(dy, x, rows, dgamma, dbeta);
    A2F_CHECK_LAUNCH("ln64_param_grad_kernel");
    count_launch();
    return A2F_OK;
}

}  // extern "C"

// FaceFormer's autoregressive decoder (ref:src/model/faceformer.py:154-185) as one persistent kernel.
//
// The reference re-runs nn.TransformerDecoder on the whole prefix (and the 64->15069 head on the whole prefix) for
// every new frame, uploading a fresh [4,t,t] bias mask and a python-built [t,T] memory mask each step.  In eval mode
// this is mathematically a KV-cached decode (causal self-attention => token i only depends on e_0..e_i), and the
// diagonal memory mask (ref:faceformer.py:58-66, dataset "vocaset") leaves one visible key per query so that
// cross-attention collapses to  out_proj(v_proj(memory_i))  -- computed for all frames up front by two small GEMMs.
//
// One CTA per utterance, 512 threads.  Every thread keeps one 64-wide weight row slice in REGISTERS for the whole
// rollout (50k parameters live in the register file, none are re-read from memory inside the T-step loop):
//   tid   0..191  self-attn in_proj rows (q | k | v)          tid 192..255  self-attn out_proj rows
//   tid 256..383  linear1 rows (ffn 64->128, ReLU)            tid 384..511  linear2 rows, two 64-wide halves per row
// The 64x64 feedback matrix Wc = vertice_map.weight @ vertice_map_r.weight (a2f_pack_feedback) sits transposed in
// shared memory and is applied by threads 192..255.
// K/V cache rows live in shared memory (padded to 68 floats: conflict-free float4 reads) when T <= 360, otherwise in
// the L2-resident workspace.
// Long clips (T > 360, inference): one CTA reading a 1.8 MB K/V prefix per step is bound by a single SM's L2 read
// bandwidth (15 us per step at T = 3600, profiles/r1_sweep_long.txt).  The CLUSTER instantiation spreads the keys of an
// utterance over a thread-block cluster of C CTAs (key j belongs to rank j mod C): rank 0 runs the step as above,
// broadcasts the scaled query into every CTA's shared memory (DSMEM), all ranks attend over their own keys straight
// from L2 and hand (max, sum, P.V) partials back through rank 0's shared memory; two cluster barriers per step.  The temporal bias -2^{-2(h+1)} * floor((i-j)/period) (ref:faceformer.py:22-54) and the
// periodic positional encoding row (i mod period) (ref:faceformer.py:70-88) are generated from indices.
#include "a2f_common.cuh"
#include "gemm_params.cuh"
#include <cooperative_groups.h>

namespace a2f {
namespace cg = cooperative_groups;

constexpr int DEC_THREADS = 512;
constexpr int KV_LD = 68;
constexpr int DEC_SMEM_T = 360;     // longest clip whose K/V cache fits in shared memory
constexpr int DEC_CLUSTER_T = 900;  // automatic mode: clips from this length on spread their keys over a CTA cluster

struct DecW {
    const float *sa_in_w, *sa_in_b, *sa_out_w, *sa_out_b, *lin1_w, *lin1_b, *lin2_w, *lin2_b;
    const float *n1_w, *n1_b, *n2_w, *n2_b, *n3_w, *n3_b, *fb_w, *fb_b, *obj_w, *pe;
    const float *fold_w, *fold_pe;   // a2f_pack_decoder_fold: in_proj @ Wc [192,64], pe @ in_proj^T [period,192]
};

// exp() on the sequential critical path of the rollout: MUFU.EX2 on x * log2(e) (2 instructions, relative error 2^-22 -- three
// orders of magnitude inside the fp32 path's 1e-5 m budget) instead of expf()'s ~15-instruction sequence; exp(-inf) = 0.
A2F_D float dec_exp(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}
// floor(d / period) for 0 <= d < 2^16 without an integer division: high word of d * ceil(2^32 / period)
A2F_D int dec_div(int d, unsigned magic) { return (int)__umulhi((unsigned)d, magic); }

A2F_D float dot64_smem(const float* w, const float* __restrict__ x) {
    // four independent accumulators: the 64-term dependent FMA chain (64 x 4 cycles) was the matvec critical path
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

// LayerNorm(64) of the vector whose elements (lane, lane+32) this warp's lanes hold; eps 1e-5, biased variance.
// Sum and sum of squares are reduced in the same five shuffle rounds (two independent chains); var = E[x^2] - mean^2.
A2F_D void warp_ln64(float& a, float& b, float g0, float g1, float b0, float b1) {
    float s1 = a + b, s2 = fmaf(a, a, b * b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float mean = s1 * (1.f / 64.f);
    const float var = fmaxf(fmaf(-mean, mean, s2 * (1.f / 64.f)), 0.f);
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    a = (a - mean) * rstd * g0 + b0;
    b = (b - mean) * rstd * g1 + b1;
}

// Activations the backward pass (decoder_bwd.cu) needs, all [B,T,width] fp32; written only by the TRAIN instantiation.
struct DecSaves {
    float *X, *Q, *K, *V, *CTX, *Y1PRE, *Y2PRE, *Y2, *HID, *Y3PRE, *LSE;
    // inference, streaming hand-over to the vertex head (a2f_decoder_rollout_stream): d_i is ALSO written as the head's
    // bf16 operand (hi | lo | hi, 192 columns) at row i * zB + b of zs ("frame-major"), and frames_done[i] is incremented
    // (release) once per utterance -- the head kernel, running at the same time on the idle SMs, waits on it
    bf16* zs;
    unsigned* frames_done;
    int zB;
};

// K/V rows written by another SM of the cluster: read through L2 (L1 is not coherent across SMs)
template <bool CL> A2F_D float4 ld_kv4(const float* p) {
    return CL ? __ldcg(reinterpret_cast<const float4*>(p)) : *reinterpret_cast<const float4*>(p);
}

// debug (a2f_debug_set_decoder_timing): when non-NULL, thread 0 of CTA 0 accumulates the clock64() cycles between the block
// barriers of a step into [0..4] = attention, out_proj, LN1/LN2/linear1, linear2, LN3/feedback/in-projection; [5] = steps
__device__ unsigned long long* g_dec_timing_dev = nullptr;

// Streaming hand-over to the vertex head (a2f_decoder_rollout_stream), spread over two warps that idle through phase 6 so that
// nothing is added to the step's critical path (in warp 6, behind the D store, the gpu-scope release alone cost 0.6 us per
// step; one idle warp doing load + stores + release in one go still overran the phase by 0.3 us):
//   dec_store_frame   (warp 9, step i)    decoder state of frame i-1 (written to D by warp 6 a step ago, re-read from L2) ->
//                                         row (i-1) * zB + b of the frame-major bf16 operand: hi | lo | hi with the rounding
//                                         of split_bf16x3_kernel (api_gemm.cu): hi = bf16(x), lo = bf16(x - hi)
//   dec_release_frame (warp 8, step i+1)  frames_done[i-1] += 1, release.gpu: warp 9's stores were ordered before this thread
//                                         by the barrier that ended step i (the release is cumulative)
// The head sees a frame two steps (5 us) late.
A2F_D void dec_store_frame(const DecSaves& sv, const float* D_b, int f, int b, int lane) {
    const float a = __ldcg(D_b + (long long)f * 64 + lane), c = __ldcg(D_b + (long long)f * 64 + lane + 32);
    bf16* zr = sv.zs + ((long long)f * sv.zB + b) * 192;
    const bf16 ha = __float2bfloat16_rn(a), hc = __float2bfloat16_rn(c);
    const bf16 la = __float2bfloat16_rn(a - __bfloat162float(ha)), lc = __float2bfloat16_rn(c - __bfloat162float(hc));
    zr[lane] = ha; zr[lane + 32] = hc;
    zr[64 + lane] = la; zr[96 + lane] = lc;
    zr[128 + lane] = ha; zr[160 + lane] = hc;
}
A2F_D void dec_release_frame(const DecSaves& sv, int f) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(sv.frames_done + f) : "memory");
}

template <bool TRAIN, bool CL, bool TIMING = false>
__global__ void __launch_bounds__(DEC_THREADS, 1)
decoder_rollout_kernel(DecW w, const float* __restrict__ ca /*[B,T,64] cross-attn vectors*/,
                       const float* __restrict__ one_hot, int n_onehot, int period, float* __restrict__ D, int T,
                       float* __restrict__ kv_global /* [B][2][T][KV_LD] or NULL */, DecSaves sv) {
    extern __shared__ __align__(16) float dsm[];
    pdl_sync();   // PDL: wait for the previous kernel's results, let the next kernel's prologue start
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = CL ? (int)cluster.num_blocks() : 1;          // CTAs per utterance
    const int rank = CL ? (int)cluster.block_rank() : 0;        // this CTA owns the keys j with j % CS == rank
    const int b = CL ? (int)blockIdx.x / CS : (int)blockIdx.x;
    const bool lead = rank == 0;
    const int Tpad = CL ? ((((T + CS - 1) / CS) + 3) & ~3) : ((T + 3) & ~3);    // score slots per head in this CTA

    // ---- shared memory carve-up ----
    float* xs = dsm;                 // [64]  decoder input of the current step (e_i + pe)
    float* qs = xs + 64;             // [64]  scaled query
    float* os = qs + 64;             // [64]  attention output (heads concatenated)
    float* ys = os + 64;             // [64]  x + self_attn (pre-LN1)
    float* x2 = ys + 64;             // [4][64] per-FFN1-warp copies of LN2 output
    float* f1 = x2 + 256;            // [128] relu(linear1)
    float* y3 = f1 + 128;            // [64]  pre-LN3
    float* dcp = y3 + 64;            // [8][64] per-warp copies of d_i (warps 0..5: q|k|v of the next token, 6..7: feedback)
    float* style = dcp + 512;        // [64]
    float* red = style + 64;         // [4][8] group reductions (max, sum) ; [4][4][16] PV partials after it
    float* pvp = red + 32;           // [4][4][16]
    float* wct = pvp + 256;          // [64][64] feedback matrix transposed: wct[k*64+r] = Wc[r][k]
    float* cpart = wct + 4096;       // [8][4][20] cluster partials (max, sum, P.V[16]) per rank and head (rank 0's copy is read)
    float* sc = cpart + 640;         // [4][Tpad] scores / probabilities
    float* Kc;
    float* Vc;
    if (kv_global) {
        Kc = kv_global + (long long)b * 2 * T * KV_LD;
        Vc = Kc + (long long)T * KV_LD;
    } else {
        Kc = sc + 4 * Tpad;
        Vc = Kc + (long long)T * KV_LD;
    }

    // ---- weights into registers ----
    float wr[64];
    float bias_r = 0.f;
    if (lead) {
        const float* src;
        if (tid < 192) { src = w.fold_w + tid * 64; }      // bias_r: set below (depends on the style embedding)
        else if (tid < 256) { src = w.sa_out_w + (tid - 192) * 64; bias_r = w.sa_out_b[tid - 192]; }
        else if (tid < 384) { src = w.lin1_w + (tid - 256) * 64; bias_r = w.lin1_b[tid - 256]; }
        else {
            const int u = tid - 384, row = u >> 1, half = u & 1;
            src = w.lin2_w + row * 128 + half * 64;
            bias_r = half == 0 ? w.lin2_b[row] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 64; k += 4) {
            const float4 f = *reinterpret_cast<const float4*>(src + k);
            wr[k] = f.x; wr[k + 1] = f.y; wr[k + 2] = f.z; wr[k + 3] = f.w;
        }
    }
    // feedback matrix (a2f_pack_feedback) transposed into shared memory; its rows are applied by threads 192..255
    if (lead)
        for (int idx = tid; idx < 4096; idx += DEC_THREADS) wct[(idx & 63) * 64 + (idx >> 6)] = w.fb_w[idx];
    const float fb_bias = (tid >= 192 && tid < 256) ? w.fb_b[tid - 192] : 0.f;
    // style embedding: obj_vector(one_hot), no bias (ref:faceformer.py:131,148); token 0 = style
    if (tid < 64) {
        float acc = 0.f;
        for (int k = 0; k < n_onehot; ++k) acc = fmaf(w.obj_w[tid * n_onehot + k], one_hot[(long long)b * n_onehot + k], acc);
        style[tid] = acc;
        xs[tid] = acc + w.pe[tid];       // position 0
    }
    __syncthreads();

    const float* ca_b = ca + (long long)b * T * 64;
    float* D_b = D + (long long)b * T * 64;
    const unsigned pmagic = 0xFFFFFFFFu / (unsigned)period + 1u;      // dec_div: floor((i - j) / period) without a division

    // The feedback e_{i+1} = Wc d_i + bc + style and the in-projection of token i+1 are two Linear layers in a row:
    //     [q|k|v]_{i+1} = in_proj(e_{i+1} + pe_{i+1}) = (in_proj Wc) d_i + in_proj (bc + style) + b_in + in_proj pe_{i+1}
    // so q, k, v of the NEXT token come straight from d_i (threads 0..191 hold rows of in_proj Wc, a2f_pack_decoder_fold)
    // in the same phase that produces e_{i+1}: one matvec + one barrier less on the sequential critical path of every step.
    // bias_r (threads 0..191) = in_proj (bc + style) + b_in, utterance constant; in_proj pe_pos comes from the packed table.
    // Token 0 (= style + pe_0, no d yet) takes the plain in-projection once, here.
    if (lead && tid < 192) {
        const float* wrow = w.sa_in_w + tid * 64;
        float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll 8
        for (int k = 0; k < 64; k += 2) {
            const float w0 = __ldg(wrow + k), w1 = __ldg(wrow + k + 1);
            a0 = fmaf(w0, xs[k], a0);
            a1 = fmaf(w1, xs[k + 1], a1);
            c0 = fmaf(w0, w.fb_b[k] + style[k], c0);
            c1 = fmaf(w1, w.fb_b[k + 1] + style[k + 1], c1);
        }
        const float bin = w.sa_in_b[tid];
        bias_r = bin + (c0 + c1);
        const float acc = bin + (a0 + a1);
        if (tid < 64) {
            qs[tid] = acc * 0.25f;                                // 1/sqrt(head_dim 16), exact power of two
            if (CL)
                for (int rr = 1; rr < CS; ++rr) cluster.map_shared_rank(qs, rr)[tid] = acc * 0.25f;
        }
        else if (tid < 128) Kc[tid - 64] = acc;
        else Vc[tid - 128] = acc;
        if (TRAIN) {
            const long long o = ((long long)b * T) * 64 + (tid & 63);
            if (tid < 64) { sv.Q[o] = acc * 0.25f; sv.X[o] = xs[tid]; }
            else if (tid < 128) sv.K[o] = acc;
            else sv.V[o] = acc;
        }
    }
    if (CL) cluster.sync();          // q_0 has landed in every rank's shared memory, K/V row 0 is visible cluster-wide
    else __syncthreads();

    // loop-invariant LayerNorm parameters of the lanes that apply them (norm1/norm2: FFN1 warps, norm3: feedback warps)
    float lnA[4] = {0.f, 0.f, 0.f, 0.f}, lnB[4] = {0.f, 0.f, 0.f, 0.f};
    if (tid >= 256 && tid < 384) {
        lnA[0] = w.n1_w[lane]; lnA[1] = w.n1_w[lane + 32]; lnA[2] = w.n1_b[lane]; lnA[3] = w.n1_b[lane + 32];
        lnB[0] = w.n2_w[lane]; lnB[1] = w.n2_w[lane + 32]; lnB[2] = w.n2_b[lane]; lnB[3] = w.n2_b[lane + 32];
    } else if (tid < 256) {          // warps 0..7 each normalise y3 -> d_i for themselves (phase 6)
        lnA[0] = w.n3_w[lane]; lnA[1] = w.n3_w[lane + 32]; lnA[2] = w.n3_b[lane]; lnA[3] = w.n3_b[lane + 32];
    }

    unsigned long long* const tl = (TIMING && blockIdx.x == 0 && tid == 0) ? g_dec_timing_dev : nullptr;
    unsigned long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = (TIMING && tl) ? clock64() : 0;
#define DEC_TSTAMP(k) do { if (TIMING && tl) { const long long tn = clock64(); tacc[k] += (unsigned long long)(tn - tprev); tprev = tn; } } while (0)
    for (int i = 0; i < T; ++i) {
        // prefetch this step's global operands so that their latency hides behind phases 1-3
        float pre0 = 0.f, pre1 = 0.f;
        if (!lead) {
        } else if (tid >= 256 && tid < 384) {
            pre0 = __ldg(ca_b + (long long)i * 64 + lane);
            pre1 = __ldg(ca_b + (long long)i * 64 + lane + 32);
        } else if (tid >= 192 && tid < 256) {
            pre0 = __ldg(w.pe + ((i + 1) % period) * 64 + (tid - 192));
        } else if (tid < 192) {
            pre0 = __ldg(w.fold_pe + ((i + 1) % period) * 192 + tid);      // in_proj pe_{i+1}
        }
        // ---------- phase 2: biased causal attention over keys 0..i (threads 0..511, 4 warps per head) ----------
        // Flash-style split: warp wq of head h owns the keys {128m + 32wq + lane} and runs its own softmax
        // (max by one REDUX on order-preserving integers, probabilities and P.V warp-synchronously); the four
        // (max, sum, P.V) partials of a head are merged after ONE named barrier.  (The first version shared max and
        // sum through shared memory with three barriers per step: 40 % of the step, profiles/r1_decoder_lines.txt.)
        {
            const int h = tid >> 7, u = tid & 127, wq = (tid >> 5) & 3;
            const float slope = (h == 0) ? 0.25f : (h == 1) ? 0.0625f : (h == 2) ? 0.015625f : 0.00390625f;
            float qh[16];
#pragma unroll
            for (int d = 0; d < 16; d += 4) {
                const float4 f = *reinterpret_cast<const float4*>(qs + h * 16 + d);
                qh[d] = f.x; qh[d + 1] = f.y; qh[d + 2] = f.z; qh[d + 3] = f.w;
            }
            float* sch = sc + h * Tpad;
            // local key slot n <-> key j = rank + CS * n (one CTA: j = n); n_max = last slot that is <= i
            const int n_max = (i >= rank) ? (i - rank) / CS : -1;
            float lmax = -INFINITY;
            for (int n = u; n <= n_max; n += 128) {
                const int j = rank + CS * n;
                const float* kp = Kc + (long long)j * KV_LD + h * 16;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int d = 0; d < 16; d += 4) {
                    const float4 f = ld_kv4<CL>(kp + d);
                    a0 = fmaf(qh[d], f.x, a0);
                    a1 = fmaf(qh[d + 1], f.y, a1);
                    a2 = fmaf(qh[d + 2], f.z, a2);
                    a3 = fmaf(qh[d + 3], f.w, a3);
                }
                const float s = ((a0 + a1) + (a2 + a3)) - slope * (float)dec_div(i - j, pmagic);
                sch[n] = s;
                lmax = fmaxf(lmax, s);
            }
            DEC_TSTAMP(0);
            const float wmax = warp_max_redux(lmax);                  // -inf when this warp owns no key yet
            float lsum = 0.f;
            for (int n = u; n <= n_max; n += 128) {
                const float pj = dec_exp(sch[n] - wmax);
                sch[n] = pj;
                lsum += pj;
            }
            __syncwarp();
            DEC_TSTAMP(1);
            // P.V over this warp's keys: lane = (key slot kg, 4-wide column group dg); conflict-free float4 reads
            const int kg = lane & 7, dg = lane >> 3;
            float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
            for (int base = 32 * wq; base <= n_max; base += 128) {
                const int nvalid = min(32, n_max - base + 1);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int jj = kg + 8 * t;
                    if (jj < nvalid) {
                        const int n = base + jj;
                        const int j = rank + CS * n;
                        const float pj = sch[n];
                        const float4 v = ld_kv4<CL>(Vc + (long long)j * KV_LD + h * 16 + dg * 4);
                        o0 = fmaf(pj, v.x, o0);
                        o1 = fmaf(pj, v.y, o1);
                        o2 = fmaf(pj, v.z, o2);
                        o3 = fmaf(pj, v.w, o3);
                    }
                }
            }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                o0 += __shfl_xor_sync(0xffffffffu, o0, o);
                o1 += __shfl_xor_sync(0xffffffffu, o1, o);
                o2 += __shfl_xor_sync(0xffffffffu, o2, o);
                o3 += __shfl_xor_sync(0xffffffffu, o3, o);
            }
            lsum = warp_sum(lsum);
            if (kg == 0) *reinterpret_cast<float4*>(pvp + (h * 4 + wq) * 16 + dg * 4) = make_float4(o0, o1, o2, o3);
            if (lane == 0) {
                red[h * 8 + wq] = wmax;
                red[h * 8 + 4 + wq] = lsum;
            }
            DEC_TSTAMP(2);
            named_bar_sync(1 + h, 128);
            DEC_TSTAMP(3);
            if (CL) {
                // this rank's (max, sum, unnormalised P.V) of head h -> rank 0's shared memory
                if (u < 16) {
                    const float m0 = red[h * 8], m1 = red[h * 8 + 1], m2 = red[h * 8 + 2], m3 = red[h * 8 + 3];
                    const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    const float e0 = m0 == -INFINITY ? 0.f : dec_exp(m0 - m), e1 = m1 == -INFINITY ? 0.f : dec_exp(m1 - m);
                    const float e2 = m2 == -INFINITY ? 0.f : dec_exp(m2 - m), e3 = m3 == -INFINITY ? 0.f : dec_exp(m3 - m);
                    const float l = (e0 * red[h * 8 + 4] + e1 * red[h * 8 + 5]) + (e2 * red[h * 8 + 6] + e3 * red[h * 8 + 7]);
                    const float* pp = pvp + h * 64 + u;
                    float* dst = cluster.map_shared_rank(cpart, 0) + (rank * 4 + h) * 20;
                    dst[2 + u] = (e0 * pp[0] + e1 * pp[16]) + (e2 * pp[32] + e3 * pp[48]);
                    if (u == 0) { dst[0] = m; dst[1] = l; }
                }
                cluster.sync();
                if (!lead) {                 // ranks > 0 only attend; rank 0 now runs phases 3..6 and publishes q_{i+1}, k/v row i+1
                    cluster.sync();
                    continue;
                }
                if (u < 16) {
                    float m = -INFINITY;
                    for (int rr = 0; rr < CS; ++rr) m = fmaxf(m, cpart[(rr * 4 + h) * 20]);
                    float l = 0.f, cv = 0.f;
                    for (int rr = 0; rr < CS; ++rr) {
                        const float* cp = cpart + (rr * 4 + h) * 20;
                        const float e = cp[0] == -INFINITY ? 0.f : dec_exp(cp[0] - m);
                        l = fmaf(e, cp[1], l);
                        cv = fmaf(e, cp[2 + u], cv);
                    }
                    os[h * 16 + u] = cv / l;
                }
            } else if (u < 16) {
                const float m0 = red[h * 8], m1 = red[h * 8 + 1], m2 = red[h * 8 + 2], m3 = red[h * 8 + 3];
                const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                const float e0 = dec_exp(m0 - m), e1 = dec_exp(m1 - m), e2 = dec_exp(m2 - m), e3 = dec_exp(m3 - m);
                const float l = (e0 * red[h * 8 + 4] + e1 * red[h * 8 + 5]) + (e2 * red[h * 8 + 6] + e3 * red[h * 8 + 7]);
                const float* pp = pvp + h * 64 + u;
                const float cv = ((e0 * pp[0] + e1 * pp[16]) + (e2 * pp[32] + e3 * pp[48])) / l;
                os[h * 16 + u] = cv;
                if (TRAIN) {
                    sv.CTX[((long long)b * T + i) * 64 + h * 16 + u] = cv;
                    if (u == 0) sv.LSE[((long long)b * T + i) * 4 + h] = m + logf(l);
                }
            }
        }
        DEC_TSTAMP(4);
        __syncthreads();
        DEC_TSTAMP(5);

        // ---------- phase 3: self-attn out_proj + residual ----------
        if (tid >= 192 && tid < 256) {
            const int r = tid - 192;
            const float yv = xs[r] + (bias_r + dot64_smem(wr, os));
            ys[r] = yv;
            if (TRAIN) sv.Y1PRE[((long long)b * T + i) * 64 + r] = yv;
        }
        __syncthreads();
        DEC_TSTAMP(6);

        // ---------- phase 4: LN1, + cross-attention vector, LN2 (each FFN1 warp redundantly), linear1 + ReLU ----------
        if (tid >= 256 && tid < 384) {
            const int wl = warp - 8;
            float a = ys[lane], c = ys[lane + 32];
            warp_ln64(a, c, lnA[0], lnA[1], lnA[2], lnA[3]);
            a += pre0;
            c += pre1;
            if (TRAIN && wl == 0) {
                sv.Y2PRE[((long long)b * T + i) * 64 + lane] = a;
                sv.Y2PRE[((long long)b * T + i) * 64 + lane + 32] = c;
            }
            warp_ln64(a, c, lnB[0], lnB[1], lnB[2], lnB[3]);
            float* xc = x2 + wl * 64;
            xc[lane] = a;
            xc[lane + 32] = c;
            if (TRAIN && wl == 0) {
                sv.Y2[((long long)b * T + i) * 64 + lane] = a;
                sv.Y2[((long long)b * T + i) * 64 + lane + 32] = c;
            }
            __syncwarp();
            const float hv = relu(bias_r + dot64_smem(wr, xc));
            f1[tid - 256] = hv;
            if (TRAIN) sv.HID[((long long)b * T + i) * 128 + (tid - 256)] = hv;
        }
        __syncthreads();
        DEC_TSTAMP(7);

        // ---------- phase 5: linear2 (two half-rows per output) + residual ----------
        if (tid >= 384 && tid < 512) {
            const int u = tid - 384, row = u >> 1, half = u & 1;
            float part = bias_r + dot64_smem(wr, f1 + half * 64);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            if (half == 0) {
                const float yv = x2[row] + part;
                y3[row] = yv;
                if (TRAIN) sv.Y3PRE[((long long)b * T + i) * 64 + row] = yv;
            }
        }
        __syncthreads();
        DEC_TSTAMP(8);

        // ---------- phase 6: LN3 -> d_i (warps 0..7, each for itself), store, then IN ONE STEP the next token's
        //            decoder input e_{i+1} + pe (warps 6,7: Wc) and its q | k | v (warps 0..5: in_proj Wc) ----------
        if (tid < 256) {
            float a = y3[lane], c = y3[lane + 32];
            warp_ln64(a, c, lnA[0], lnA[1], lnA[2], lnA[3]);
            float* dc = dcp + warp * 64;
            dc[lane] = a;
            dc[lane + 32] = c;
            if (warp == 6) {
                D_b[(long long)i * 64 + lane] = a;
                D_b[(long long)i * 64 + lane + 32] = c;
            }
            __syncwarp();
            DEC_TSTAMP(9);
            if (tid >= 192) {
                const int r = tid - 192;
                float f0 = 0.f, f1v = 0.f, f2 = 0.f, f3 = 0.f;
#pragma unroll
                for (int k = 0; k < 64; k += 4) {
                    const float4 dv = *reinterpret_cast<const float4*>(dc + k);
                    f0 = fmaf(wct[k * 64 + r], dv.x, f0);
                    f1v = fmaf(wct[(k + 1) * 64 + r], dv.y, f1v);
                    f2 = fmaf(wct[(k + 2) * 64 + r], dv.z, f2);
                    f3 = fmaf(wct[(k + 3) * 64 + r], dv.w, f3);
                }
                const float e = (fb_bias + ((f0 + f1v) + (f2 + f3))) + style[r];
                xs[r] = e + pre0;
                if (TRAIN && i + 1 < T) sv.X[((long long)b * T + i + 1) * 64 + r] = e + pre0;
            } else if (i + 1 < T) {
                const float acc = (bias_r + pre0) + dot64_smem(wr, dc);
                if (tid < 64) {
                    qs[tid] = acc * 0.25f;
                    if (CL)
                        for (int rr = 1; rr < CS; ++rr) cluster.map_shared_rank(qs, rr)[tid] = acc * 0.25f;
                }
                else if (tid < 128) Kc[(long long)(i + 1) * KV_LD + (tid - 64)] = acc;
                else Vc[(long long)(i + 1) * KV_LD + (tid - 128)] = acc;
                if (TRAIN) {
                    const long long o = ((long long)b * T + i + 1) * 64 + (tid & 63);
                    if (tid < 64) sv.Q[o] = acc * 0.25f;
                    else if (tid < 128) sv.K[o] = acc;
                    else sv.V[o] = acc;
                }
            }
        }
        else if (!TRAIN && !CL && sv.zs != nullptr) {
            if (warp == 9 && i > 0) dec_store_frame(sv, D_b, i - 1, b, lane);
            else if (tid == 256 && i > 1) dec_release_frame(sv, i - 2);
        }
        DEC_TSTAMP(10);
        if (CL) cluster.sync();      // q_{i+1} has landed in every rank's shared memory, K/V row i+1 is visible cluster-wide
        else __syncthreads();
        DEC_TSTAMP(11);
    }
#undef DEC_TSTAMP
    if (!TRAIN && !CL && sv.zs != nullptr) {                    // the last two frames (uniform branch: every thread takes it)
        if (warp == 9) dec_store_frame(sv, D_b, T - 1, b, lane);
        __syncthreads();
        if (tid == 256) {
            if (T > 1) dec_release_frame(sv, T - 2);
            dec_release_frame(sv, T - 1);
        }
    }
    if (TIMING && tl) {
        for (int k = 0; k < 12; ++k) tl[k] = tacc[k];
        tl[12] = (unsigned long long)T;
    }
}

// Operands of the folded step (see the kernel): fold_w = in_proj_weight @ Wc  [192,64],  fold_pe[pos] = in_proj_weight @ pe[pos]
// [period,192]; fp64 accumulation.  Runs with a2f_pack_feedback whenever a decoder weight changes (Wc must be current).
__global__ void __launch_bounds__(192) pack_decoder_fold_kernel(const float* __restrict__ in_w, const float* __restrict__ wc,
                                                                const float* __restrict__ pe, int period,
                                                                float* __restrict__ fold_w, float* __restrict__ fold_pe) {
    const int r = threadIdx.x;                 // row of in_proj_weight
    const int blk = blockIdx.x;                // blocks 0..63: column blk of fold_w; blocks 64..64+period-1: position
    const float* wrow = in_w + r * 64;
    double acc = 0.0;
    if (blk < 64) {
        for (int k = 0; k < 64; ++k) acc = fma((double)wrow[k], (double)wc[k * 64 + blk], acc);
        fold_w[r * 64 + blk] = (float)acc;
    } else {
        const float* pp = pe + (blk - 64) * 64;
        for (int k = 0; k < 64; ++k) acc = fma((double)wrow[k], (double)pp[k], acc);
        fold_pe[(blk - 64) * 192 + r] = (float)acc;
    }
}

// Wc = Wm @ Wr ([64,V3] @ [V3,64]) and bc = Wm @ br + bm, fp64 accumulation.  One 8-CTA cluster per output row of Wc:
// CTA `cx` of the cluster covers the cx-th eighth of the 15069-long contraction with 4 interleaved partial sums per
// column, and the leader CTA adds the eight per-CTA sums through distributed shared memory in a fixed order
// (deterministic; this runs once per weight version in inference but once per STEP in training).
constexpr int FB_SPLIT = 8;

__global__ void __cluster_dims__(FB_SPLIT, 1, 1) __launch_bounds__(256)
pack_feedback_kernel(const float* __restrict__ vm_w, const float* __restrict__ vm_b, const float* __restrict__ vmr_w,
                     const float* __restrict__ vmr_b, int V3, float* __restrict__ Wc, float* __restrict__ bc) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int r = blockIdx.y;                  // row of Wc
    const int cx = blockIdx.x;                 // slice of the contraction (== rank in the cluster)
    const int c = threadIdx.x & 63, part = threadIdx.x >> 6;   // 4 partial sums per column
    const int chunk = (V3 + FB_SPLIT - 1) / FB_SPLIT;
    const int v0 = cx * chunk, v1 = min(V3, v0 + chunk);
    __shared__ double sh[4][65];
    __shared__ double tot[65];
    const float* a_row = vm_w + (long long)r * V3;
    double acc0 = 0.0, acc1 = 0.0, accb = 0.0;
    int v = v0 + part;
    for (; v + 4 < v1; v += 8) {               // two independent chains, loads issued together
        const float a0 = __ldg(a_row + v), a1 = __ldg(a_row + v + 4);
        const float w0 = __ldg(vmr_w + (long long)v * 64 + c), w1 = __ldg(vmr_w + (long long)(v + 4) * 64 + c);
        acc0 = fma((double)a0, (double)w0, acc0);
        acc1 = fma((double)a1, (double)w1, acc1);
        if (c == 0) accb += (double)a0 * (double)__ldg(vmr_b + v) + (double)a1 * (double)__ldg(vmr_b + v + 4);
    }
    for (; v < v1; v += 4) {
        const float a0 = __ldg(a_row + v);
        acc0 = fma((double)a0, (double)__ldg(vmr_w + (long long)v * 64 + c), acc0);
        if (c == 0) accb += (double)a0 * (double)__ldg(vmr_b + v);
    }
    sh[part][c] = acc0 + acc1;
    if (c == 0) sh[part][64] = accb;
    __syncthreads();
    if (threadIdx.x < 65) tot[threadIdx.x] = (sh[0][threadIdx.x] + sh[1][threadIdx.x]) + (sh[2][threadIdx.x] + sh[3][threadIdx.x]);
    cluster.sync();                            // every CTA's `tot` is complete and visible cluster-wide
    if (cx == 0 && threadIdx.x < 65) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < FB_SPLIT; ++k) s += *cluster.map_shared_rank(&tot[threadIdx.x], k);
        if (threadIdx.x < 64) Wc[r * 64 + threadIdx.x] = (float)s;
        else bc[r] = (float)(s + (double)vm_b[r]);
    }
    cluster.sync();                            // peers keep their shared memory alive until the leader has read it
}

// Cross-attention under the diagonal memory mask is out_proj(v_proj(audio_feature_map(h))): three Linear layers in a row.
// This folds them (fp64) into ONE [64, Kin] operand + bias, so that the encoder states go straight to the vectors the
// decoder adds:  W = Wo Wv Wa,  b = Wo (Wv ba + bv) + bo.  One CTA per 64 input columns; every CTA first forms
// M = Wo Wv (64 x 64, fp64, in shared memory), CTA 0 also the bias.  Replaces three torch fp64 matmuls (a cuBLAS DGEMM
// chain) that ran on every weight refresh, i.e. on every training step (VERDICT r1 hygiene #13).
template <typename TO>
__global__ void __launch_bounds__(256)
pack_cross_attention_kernel(const float* __restrict__ wv, const float* __restrict__ bv, const float* __restrict__ wo,
                            const float* __restrict__ bo, const float* __restrict__ wa, const float* __restrict__ ba,
                            int Kin, TO* __restrict__ W, float* __restrict__ b) {
    __shared__ double Msh[64][64];
    __shared__ float wa_sh[64][64];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < 4096; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        double acc = 0.0;
        for (int k = 0; k < 64; ++k) acc = fma((double)wo[r * 64 + k], (double)wv[k * 64 + c], acc);
        Msh[r][c] = acc;
    }
    const int c0 = blockIdx.x * 64;
    for (int idx = tid; idx < 4096; idx += 256) {
        const int k = idx >> 6, c = idx & 63;
        wa_sh[k][c] = (c0 + c < Kin) ? wa[(long long)k * Kin + c0 + c] : 0.f;
    }
    __syncthreads();
    for (int idx = tid; idx < 4096; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        if (c0 + c >= Kin) continue;
        double acc = 0.0;
#pragma unroll 8
        for (int k = 0; k < 64; ++k) acc = fma(Msh[r][k], (double)wa_sh[k][c], acc);
        st_from_float(W + (long long)r * Kin + c0 + c, (float)acc);
    }
    if (blockIdx.x == 0 && tid < 64) {
        double acc = 0.0;
        for (int k = 0; k < 64; ++k) acc = fma(Msh[tid][k], (double)ba[k], acc);
        for (int k = 0; k < 64; ++k) acc = fma((double)wo[tid * 64 + k], (double)bv[k], acc);
        b[tid] = (float)(acc + (double)bo[tid]);
    }
}

static bool g_dec_timing_on = false;   // debug (a2f_debug_set_decoder_timing): launch the instantiations with cycle counters
static int g_dec_cluster = 0;      // debug (a2f_debug_set_umma_field 8): 0 = automatic, 1 = never, 2/4/8 = forced cluster size
void set_dec_cluster(int v) { g_dec_cluster = v; }

static size_t dec_smem_bytes(int T, bool kv_in_smem, int cluster = 1) {
    const int Tpad = (((T + cluster - 1) / cluster) + 3) & ~3;
    size_t fl = 64 * 4 + 256 + 128 + 64 + 512 + 64 + 32 + 256 + 4096 + 640 + (size_t)4 * Tpad;
    if (kv_in_smem) fl += (size_t)2 * T * KV_LD;
    return fl * sizeof(float);
}

}  // namespace a2f

using namespace a2f;

extern "C" {

size_t a2f_decoder_workspace_bytes(int B, int T) {
    if (B <= 0 || T <= 0) return 0;
    size_t n = (size_t)2 * B * T * 64;                        // v_proj(mem), cross-attn vectors
    if (T > DEC_SMEM_T) n += (size_t)2 * B * T * KV_LD;       // K/V cache
    return n * sizeof(float);
}

int a2f_decoder_save_offset(int field) {
    // floats per (utterance, frame) before `field` in the saves buffer of a2f_decoder_rollout_train
    static const int width[A2F_DEC_NFIELDS] = {64, 64, 64, 64, 64, 64, 64, 64, 128, 64, 4};
    if (field < 0 || field > A2F_DEC_NFIELDS) return -1;
    int o = 0;
    for (int i = 0; i < field; ++i) o += width[i];
    return o;
}

static int decoder_rollout_impl(const a2f_decoder_weights* w, const float* memory, int memory_is_ca, const float* one_hot,
                                int n_onehot, int period, float* D, int B, int T, void* workspace, size_t workspace_bytes,
                                float* saves, void* stream, void* zs = nullptr, unsigned* frames_done = nullptr);

int a2f_decoder_rollout_train(const a2f_decoder_weights* w, const float* memory, const float* one_hot, int n_onehot,
                              int period, float* D, int B, int T, void* workspace, size_t workspace_bytes, float* saves,
                              void* stream) {
    return decoder_rollout_impl(w, memory, 0, one_hot, n_onehot, period, D, B, T, workspace, workspace_bytes, saves, stream);
}

int a2f_decoder_rollout_ca(const a2f_decoder_weights* w, const float* ca, const float* one_hot, int n_onehot, int period,
                           float* D, int B, int T, void* workspace, size_t workspace_bytes, void* stream) {
    return decoder_rollout_impl(w, ca, 1, one_hot, n_onehot, period, D, B, T, workspace, workspace_bytes, nullptr, stream);
}

int a2f_decoder_rollout_stream(const a2f_decoder_weights* w, const float* ca, const float* one_hot, int n_onehot, int period,
                               float* D, int B, int T, void* workspace, size_t workspace_bytes, void* z3_frame_major,
                               unsigned* frames_done, void* stream) {
    A2F_REQUIRE(z3_frame_major && frames_done, "a2f_decoder_rollout_stream: NULL z3 / frames_done");
    A2F_REQUIRE(reinterpret_cast<uintptr_t>(z3_frame_major) % 16 == 0 && reinterpret_cast<uintptr_t>(frames_done) % 4 == 0,
                "a2f_decoder_rollout_stream: z3 must be 16-byte aligned");
    return decoder_rollout_impl(w, ca, 1, one_hot, n_onehot, period, D, B, T, workspace, workspace_bytes, nullptr, stream,
                                z3_frame_major, frames_done);
}

static int decoder_rollout_impl(const a2f_decoder_weights* w, const float* memory, int memory_is_ca, const float* one_hot,
                                int n_onehot, int period, float* D, int B, int T, void* workspace, size_t workspace_bytes,
                                float* saves, void* stream, void* zs, unsigned* frames_done) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(w && memory && one_hot && D && workspace, "a2f_decoder_rollout: NULL argument");
    A2F_REQUIRE(B > 0 && T > 0 && n_onehot > 0 && period > 0, "a2f_decoder_rollout: bad sizes");
    A2F_REQUIRE(T <= 8192, "a2f_decoder_rollout: T above 8192 frames is not supported");
    A2F_REQUIRE(workspace_bytes >= a2f_decoder_workspace_bytes(B, T), "a2f_decoder_rollout: workspace too small");
    A2F_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "a2f_decoder_rollout: workspace must be 16-byte aligned");
    const void* all[] = {w->sa_in_w, w->sa_in_b, w->sa_out_w, w->sa_out_b, w->ca_in_w, w->ca_in_b, w->ca_out_w, w->ca_out_b,
                         w->lin1_w, w->lin1_b, w->lin2_w, w->lin2_b, w->n1_w, w->n1_b, w->n2_w, w->n2_b, w->n3_w, w->n3_b,
                         w->fb_w, w->fb_b, w->obj_w, w->pe};
    for (const void* p : all) A2F_REQUIRE(p != nullptr, "a2f_decoder_rollout: NULL weight pointer");
    cudaStream_t s = as_stream(stream);
    float* tmp = static_cast<float*>(workspace);
    float* ca = tmp + (size_t)B * T * 64;
    float* kv = (T > DEC_SMEM_T) ? ca + (size_t)B * T * 64 : nullptr;

    // cross-attention with the diagonal memory mask: ca_t = out_proj(v_proj(memory_t))   (SURVEY.md fact 0.6);
    // a2f_decoder_rollout_ca: the caller already folded both projections (and audio_feature_map) into the GEMM that
    // produced `memory`, which then IS the cross-attention vector
    const float* ca_in = memory_is_ca ? memory : nullptr;
    GemmParams g;
    if (!memory_is_ca) {
    g.M = B * T; g.N = 64; g.K = 64;
    g.A = memory; g.a_row_stride = 64; g.a_batch_stride = 0; g.rows_per_batch = B * T;
    g.W = w->ca_in_w + 128 * 64; g.ldw = 64; g.bias = w->ca_in_b + 128; g.act = A2F_ACT_NONE;
    g.resid = nullptr; g.resid_bf16 = 0; g.ldr = 0; g.tmpl = nullptr; g.rows_per_tmpl = 1;
    g.C = tmp; g.ldc = 64; g.c_batch_stride = (long long)B * T * 64;
    rc = gemm_simt(g, 0, 0, s);
    if (rc != A2F_OK) return rc;
    g.A = tmp; g.W = w->ca_out_w; g.bias = w->ca_out_b; g.C = ca;
    rc = gemm_simt(g, 0, 0, s);
    if (rc != A2F_OK) return rc;
    }
    const float* ca_use = memory_is_ca ? ca_in : ca;

    DecW dw;
    dw.sa_in_w = w->sa_in_w; dw.sa_in_b = w->sa_in_b; dw.sa_out_w = w->sa_out_w; dw.sa_out_b = w->sa_out_b;
    dw.lin1_w = w->lin1_w; dw.lin1_b = w->lin1_b; dw.lin2_w = w->lin2_w; dw.lin2_b = w->lin2_b;
    dw.n1_w = w->n1_w; dw.n1_b = w->n1_b; dw.n2_w = w->n2_w; dw.n2_b = w->n2_b; dw.n3_w = w->n3_w; dw.n3_b = w->n3_b;
    dw.fb_w = w->fb_w; dw.fb_b = w->fb_b; dw.obj_w = w->obj_w; dw.pe = w->pe;
    A2F_REQUIRE(w->fold_w != nullptr && w->fold_pe != nullptr, "a2f_decoder_rollout: fold_w / fold_pe (a2f_pack_decoder_fold) missing");
    dw.fold_w = w->fold_w; dw.fold_pe = w->fold_pe;
    const size_t smem = dec_smem_bytes(T, kv == nullptr);
    // long clips, inference: keys of one utterance spread over a cluster of CS CTAs (largest power of two that still
    // gives every utterance its own cluster in one wave, at most 8 = the portable cluster size)
    // Measured (profiles/r1_sweep_long.txt): the two cluster barriers cost ~0.8 us per step, the single CTA's L2-bound
    // attention ~0.5 us per 100 keys: the cluster wins from T ~ 800 on.
    int CS = 1;
    if (saves == nullptr && kv != nullptr && g_dec_cluster != 1 && zs == nullptr) {
        if (g_dec_cluster > 1) CS = g_dec_cluster;
        else if (T >= DEC_CLUSTER_T) {
            // largest cluster size whose B clusters are all co-resident: a cluster needs CS free SMs inside ONE GPC, so
            // B*CS <= #SMs is not enough (16 clusters of 8 do not fit the 148 SMs of a B200: the B=16 points of the long
            // sweep ran in two waves, 2x the B=8 time) -- ask the occupancy calculator for the real limit
            auto kern = decoder_rollout_kernel<false, true>;
            A2F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            for (int cs = 8; cs >= 2; cs >>= 1) {
                if ((long long)cs * B > sm_count()) continue;
                cudaLaunchConfig_t q;
                memset(&q, 0, sizeof(q));
                q.gridDim = dim3(B * cs, 1, 1);
                q.blockDim = dim3(DEC_THREADS, 1, 1);
                q.dynamicSmemBytes = dec_smem_bytes(T, false, cs);
                cudaLaunchAttribute qa[1];
                qa[0].id = cudaLaunchAttributeClusterDimension;
                qa[0].val.clusterDim.x = cs;
                qa[0].val.clusterDim.y = 1;
                qa[0].val.clusterDim.z = 1;
                q.attrs = qa;
                q.numAttrs = 1;
                int n_clusters = 0;
                if (cudaOccupancyMaxActiveClusters(&n_clusters, kern, &q) != cudaSuccess) {
                    cudaGetLastError();
                    continue;
                }
                if (n_clusters >= B) { CS = cs; break; }
            }
        }
    }
    if (CS > 1) {
        DecSaves none = {};
        auto kern = decoder_rollout_kernel<false, true>;
        const size_t csmem = dec_smem_bytes(T, false, CS);
        A2F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(B * CS, 1, 1);
        cfg.blockDim = dim3(DEC_THREADS, 1, 1);
        cfg.dynamicSmemBytes = csmem;
        cfg.stream = s;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        A2F_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, dw, ca_use, one_hot, n_onehot, period, D, T, kv, none));
    } else if (saves == nullptr) {
        DecSaves none = {};
        none.zs = static_cast<bf16*>(zs);
        none.frames_done = frames_done;
        none.zB = B;
        if (g_dec_timing_on) {      // debug instantiation with the per-phase cycle counters (tools/decoder_phases.py)
            A2F_CHECK_CUDA(cudaFuncSetAttribute(decoder_rollout_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            A2F_CHECK_CUDA(launch_pdl(decoder_rollout_kernel<false, false, true>, dim3(B), dim3(DEC_THREADS), smem, s, dw, ca_use, one_hot, n_onehot, period, D, T, kv, none));
        } else {
        A2F_CHECK_CUDA(cudaFuncSetAttribute(decoder_rollout_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        A2F_CHECK_CUDA(launch_pdl(decoder_rollout_kernel<false, false>, dim3(B), dim3(DEC_THREADS), smem, s, dw, ca_use, one_hot, n_onehot, period, D, T, kv, none));
        }
    } else {
        DecSaves sv = {};
        const size_t bt = (size_t)B * T;
        float* f = saves;
        sv.X = f + bt * a2f_decoder_save_offset(A2F_DEC_X); sv.Q = f + bt * a2f_decoder_save_offset(A2F_DEC_Q);
        sv.K = f + bt * a2f_decoder_save_offset(A2F_DEC_K); sv.V = f + bt * a2f_decoder_save_offset(A2F_DEC_V);
        sv.CTX = f + bt * a2f_decoder_save_offset(A2F_DEC_CTX); sv.Y1PRE = f + bt * a2f_decoder_save_offset(A2F_DEC_Y1PRE);
        sv.Y2PRE = f + bt * a2f_decoder_save_offset(A2F_DEC_Y2PRE); sv.Y2 = f + bt * a2f_decoder_save_offset(A2F_DEC_Y2);
        sv.HID = f + bt * a2f_decoder_save_offset(A2F_DEC_HID); sv.Y3PRE = f + bt * a2f_decoder_save_offset(A2F_DEC_Y3PRE);
        sv.LSE = f + bt * a2f_decoder_save_offset(A2F_DEC_LSE);
        if (g_dec_timing_on) {
            A2F_CHECK_CUDA(cudaFuncSetAttribute(decoder_rollout_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            A2F_CHECK_CUDA(launch_pdl(decoder_rollout_kernel<true, false, true>, dim3(B), dim3(DEC_THREADS), smem, s, dw, ca_use, one_hot, n_onehot, period, D, T, kv, sv));
        } else {
        A2F_CHECK_CUDA(cudaFuncSetAttribute(decoder_rollout_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        A2F_CHECK_CUDA(launch_pdl(decoder_rollout_kernel<true, false>, dim3(B), dim3(DEC_THREADS), smem, s, dw, ca_use, one_hot, n_onehot, period, D, T, kv, sv));
        }
    }
    A2F_CHECK_LAUNCH("decoder_rollout_kernel");
    count_launch();
    return A2F_OK;
}

int a2f_decoder_rollout(const a2f_decoder_weights* w, const float* memory, const float* one_hot, int n_onehot,
                        int period, float* D, int B, int T, void* workspace, size_t workspace_bytes, void* stream) {
    return a2f_decoder_rollout_train(w, memory, one_hot, n_onehot, period, D, B, T, workspace, workspace_bytes, nullptr,
                                     stream);
}

int a2f_pack_feedback(const float* vm_w, const float* vm_b, const float* vmr_w, const float* vmr_b, int V3, float* Wc,
                      float* bc, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(vm_w && vm_b && vmr_w && vmr_b && Wc && bc && V3 > 0, "a2f_pack_feedback: bad arguments");
    pack_feedback_kernel<<<dim3(FB_SPLIT, 64), 256, 0, as_stream(stream)>>>(vm_w, vm_b, vmr_w, vmr_b, V3, Wc, bc);
    A2F_CHECK_LAUNCH("pack_feedback_kernel");
    count_launch();
    return A2F_OK;
}

int a2f_debug_set_decoder_timing(void* dev_ptr) {
    unsigned long long* p = static_cast<unsigned long long*>(dev_ptr);
    g_dec_timing_on = p != nullptr;
    cudaError_t e = cudaMemcpyToSymbol(g_dec_timing_dev, &p, sizeof(p));
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemcpyToSymbol(g_dec_timing_dev)");
    return A2F_OK;
}

int a2f_pack_decoder_fold(const float* sa_in_w, const float* wc, const float* pe, int period, float* fold_w, float* fold_pe,
                          void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(sa_in_w && wc && pe && fold_w && fold_pe && period > 0, "a2f_pack_decoder_fold: bad arguments");
    pack_decoder_fold_kernel<<<64 + period, 192, 0, as_stream(stream)>>>(sa_in_w, wc, pe, period, fold_w, fold_pe);
    A2F_CHECK_LAUNCH("pack_decoder_fold_kernel");
    count_launch();
    return A2F_OK;
}

int a2f_pack_cross_attention(const float* wv, const float* bv, const float* wo, const float* bo, const float* wa,
                             const float* ba, int Kin, void* W, int w_dtype, float* b, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(wv && bv && wo && bo && wa && ba && W && b && Kin > 0, "a2f_pack_cross_attention: bad arguments");
    A2F_REQUIRE(w_dtype == A2F_F32 || w_dtype == A2F_BF16, "a2f_pack_cross_attention: bad dtype");
    const unsigned grid = (unsigned)((Kin + 63) / 64);
    if (w_dtype == A2F_BF16)
        pack_cross_attention_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(wv, bv, wo, bo, wa, ba, Kin, (bf16*)W, b);
    else
        pack_cross_attention_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(wv, bv, wo, bo, wa, ba, Kin, (float*)W, b);
    A2F_CHECK_LAUNCH("pack_cross_attention_kernel");
    count_launch();
    return A2F_OK;
}

}  // extern "C"

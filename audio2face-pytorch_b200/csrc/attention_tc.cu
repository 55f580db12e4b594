// Encoder self-attention on tcgen05 / TMEM / TMA for sm_100a (inference forward, bf16, head_dim 64, no mask):
//   out = softmax(Q K^T * scale) V     per (utterance, head)       -- HF modeling_wav2vec2.py:438-463 (eager attention)
//
// The mma.sync kernel of attention.cu is fine at T = 150..300 (attention is 4 % of the FLOPs there) but it is what
// bounds the long-sequence sweep (BASELINE.json configs[4]: T = 600..3600, where attention is up to 35 % of the
// FLOPs): it ran at ~65 TFLOP/s at T = 3600 (profiles/r1_sweep_long.txt).  This kernel keeps both GEMMs of the flash
// loop on the 5th-generation tensor cores:
//
//   CTA = one (utterance, head, 128-query tile); 192 threads; two CTAs resident per SM (one CTA's softmax overlaps the
//   other's MMAs; 2 x 256 TMEM columns, 2 x 112 KB of shared memory).
//     warp 0       TMA producer: Q once, then K / V tiles of 128 keys through a 2-stage ring.  One 3-D tensor map over
//                  qkv [B][T][3*H*64] serves all three (box 64 columns x 128 rows); rows past T are zero-filled by TMA.
//     warp 1       MMA issuer (one lane): S = Q K^T  -> TMEM columns   0..127 (M=128, N=128, K=64: 4 tcgen05.mma)
//                                         O_j = P V  -> TMEM columns 128..191 (M=128, N=64, K=128: 8 tcgen05.mma),
//                  V is consumed IN PLACE as an MN-major operand (keys are the reduction rows), P from shared memory.
//     warps 2..5   softmax: thread = query row = TMEM lane.  Pass 1 over S finds the row maximum, pass 2 exponentiates
//                  (ex2.approx on pre-scaled logits), accumulates the row sum and writes bf16 P into the K-major
//                  128B-swizzled layout the P V MMA reads.  O_j comes back from TMEM once per key tile and is folded into
//                  a register accumulator with the usual online-softmax rescale (no TMEM read-modify-write).
//   Probabilities never leave the SM (the reference materialises and returns [B,12,T,T] for 12 layers,
//   ref:src/model/wav2vec.py:101).
//
// Bound: with head_dim 64 a 128x128 tile needs 16384 exponentials (1024 cycles of the SM's 16/clk MUFU) against 512
// cycles of tensor work, so the kernel's ceiling is ~50 % of the dense bf16 peak; the MUFU pipe is the roofline here.
#include "a2f_common.cuh"

namespace a2f {

constexpr int FT_BM = 128;            // queries per CTA (UMMA M)
constexpr int FT_BN = 128;            // keys per tile
constexpr int FT_D = 64;              // head dim
constexpr int FT_THREADS = 192;
constexpr int FT_TILE_BYTES = 128 * 64 * 2;                 // one [128 rows x 64 bf16] box = 16 KB
constexpr int FT_SMEM = 5 * FT_TILE_BYTES + 2 * FT_TILE_BYTES + 256;    // Q, K[2], V[2], P (two 64-key blocks), barriers
constexpr int FT_TMEM_COLS = 256;     // S: 128 fp32 columns, O_j: 64

A2F_D float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// K-major SWIZZLE_128B operand (rows of 128 B, 8-row groups 1024 B apart)
A2F_D uint64_t ft_desc_kmajor(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B operand: rows = reduction index (keys), 64 contiguous MN elements (head dim) per 128-B row
A2F_D uint64_t ft_desc_mnmajor(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(FT_THREADS, 2)
mha_tc_kernel(const __grid_constant__ CUtensorMap qkv_map, bf16* __restrict__ out, int T, int H, float c /* scale*log2(e) */) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + FT_TILE_BYTES;
    uint8_t* sV = sK + 2 * FT_TILE_BYTES;
    uint8_t* sP = sV + 2 * FT_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * FT_TILE_BYTES);
    uint64_t* q_full = bars;            // [1]
    uint64_t* k_full = bars + 1;        // [2]
    uint64_t* v_full = bars + 3;        // [2]
    uint64_t* k_empty = bars + 5;       // [2]  K slot free: S = Q K^T of that slot has retired
    uint64_t* v_empty = bars + 7;       // [2]  V slot free: O = P V of that slot has retired
    uint64_t* s_full = bars + 9;        // S_j complete in TMEM
    uint64_t* s_free = bars + 10;       // 128 softmax threads have read S_j
    uint64_t* p_full = bars + 11;       // 128 softmax threads have written P_j (and read O_{j-1})
    uint64_t* o_full = bars + 12;       // O_j complete in TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * FT_BM, h = blockIdx.y, b = blockIdx.z;
    const int n_kv = (T + FT_BN - 1) / FT_BN;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == 0 && lane == 0) tma_prefetch_desc(&qkv_map);
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 128);
        mbar_init(p_full, 128);
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<FT_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(q_full, FT_TILE_BYTES);
            tma_load_3d(sQ, &qkv_map, q_full, h * FT_D, q0, b);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j & 1;
                const uint32_t ph = (uint32_t)(j >> 1) & 1u;
                mbar_wait(&k_empty[s], ph ^ 1);
                mbar_expect_tx(&k_full[s], FT_TILE_BYTES);
                tma_load_3d(sK + s * FT_TILE_BYTES, &qkv_map, &k_full[s], (H + h) * FT_D, j * FT_BN, b);
                mbar_wait(&v_empty[s], ph ^ 1);
                mbar_expect_tx(&v_full[s], FT_TILE_BYTES);
                tma_load_3d(sV + s * FT_TILE_BYTES, &qkv_map, &v_full[s], (2 * H + h) * FT_D, j * FT_BN, b);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // D=f32, A=B=bf16; S: both K-major, N=128; O: A (P) K-major, B (V) MN-major (bit 16), N=64; M=128
            const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FT_BN >> 3) << 17) |
                                      ((uint32_t)(FT_BM >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(FT_D >> 3) << 17) |
                                      ((uint32_t)(FT_BM >> 4) << 24);
            const uint32_t t_s = tmem_base, t_o = tmem_base + 128;
            const uint64_t qdesc = ft_desc_kmajor(smem_u32(sQ));
            const uint64_t pdesc0 = ft_desc_kmajor(smem_u32(sP));
            const uint64_t pdesc1 = ft_desc_kmajor(smem_u32(sP + FT_TILE_BYTES));
            auto issue_qk = [&](int j) {
                const int s = j & 1;
                mbar_wait(&k_full[s], (uint32_t)(j >> 1) & 1u);
                if (j > 0) mbar_wait(s_free, (uint32_t)(j - 1) & 1u);     // softmax has read S_{j-1}
                tc_fence_after();
                const uint64_t kdesc = ft_desc_kmajor(smem_u32(sK + s * FT_TILE_BYTES));
#pragma unroll
                for (int k = 0; k < FT_D / 16; ++k)
                    umma_f16(t_s, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_qk, k != 0 ? 1u : 0u);
                umma_commit(&k_empty[s]);                  // K slot reusable as soon as this S has retired
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            issue_qk(0);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j & 1;
                if (j + 1 < n_kv) issue_qk(j + 1);        // overlaps the tail of softmax j / runs ahead of P V_j
                mbar_wait(p_full, (uint32_t)j & 1u);       // P_j is in shared memory, O_{j-1} has been read
                mbar_wait(&v_full[s], (uint32_t)(j >> 1) & 1u);
                tc_fence_after();
                const uint64_t vdesc = ft_desc_mnmajor(smem_u32(sV + s * FT_TILE_BYTES));
#pragma unroll
                for (int kk = 0; kk < FT_BN / 16; ++kk) {
                    const uint64_t pd = (kk < 4 ? pdesc0 : pdesc1) + (uint64_t)(2 * (kk & 3));
                    // 16 keys = 2 groups of 8 rows x 128 B = 2048 B: +128 in the >>4 address field
                    umma_f16(t_o, pd, vdesc + (uint64_t)(128 * kk), idesc_pv, kk != 0 ? 1u : 0u);
                }
                umma_commit(&v_empty[s]);
                umma_commit(o_full);
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax + output: 4 warps, thread = query row =====================
        const int q = warp & 3;                            // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;                       // row within the tile
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        uint8_t* prow = sP + r * 128;
        const int sw = r & 7;
        float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
        float o[FT_D];
#pragma unroll
        for (int i = 0; i < FT_D; ++i) o[i] = 0.f;

        auto fold_o = [&](float a) {                       // o = o * a + O_j (TMEM columns 128..191)
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                float v[32];
                tmem_ld_32x32(t_row + 128 + cc * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[cc * 32 + i] = fmaf(o[cc * 32 + i], a, v[i]);
            }
        };

        for (int j = 0; j < n_kv; ++j) {
            const int valid = min(FT_BN, T - j * FT_BN);   // live keys of this tile (uniform)
            mbar_wait(s_full, (uint32_t)j & 1u);
            tc_fence_after();
            // ---- pass 1: row maximum of the raw logits ----
            float mx = -INFINITY;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                if (cc * 32 >= valid) break;
                float v[32];
                tmem_ld_32x32(t_row + cc * 32, v);
                tmem_ld_wait();
                if (cc * 32 + 32 <= valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, v[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (cc * 32 + i < valid) mx = fmaxf(mx, v[i]);
                }
            }
            const float m_new = fmaxf(m, mx);
            const float alpha = ex2_approx((m - m_new) * c);        // first tile: exp2(-inf) = 0
            const float mc = m_new * c;
            // ---- O_{j-1} back from TMEM (also proves P V_{j-1} is done with the P buffer) ----
            if (j > 0) {
                mbar_wait(o_full, (uint32_t)(j - 1) & 1u);
                tc_fence_after();
                fold_o(alpha_prev);
            }
            // ---- pass 2: p = 2^(s*c - m*c), row sum, bf16 P into the swizzled K-major layout ----
            float sum = 0.f;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t pk[16];
                if (cc * 32 < valid) {
                    float v[32];
                    tmem_ld_32x32(t_row + cc * 32, v);
                    tmem_ld_wait();
                    if (cc * 32 + 32 <= valid) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = ex2_approx(fmaf(v[i], c, -mc));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = (cc * 32 + i < valid) ? ex2_approx(fmaf(v[i], c, -mc)) : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        sum += v[i] + v[i + 1];
                        pk[i >> 1] = pack_bf16x2(v[i], v[i + 1]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = 0u;
                }
                uint8_t* blk = prow + (cc >> 1) * FT_TILE_BYTES;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const int pch = ((cc & 1) * 4 + ch) ^ sw;
                    *reinterpret_cast<uint4*>(blk + pch * 16) = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
                }
            }
            l = fmaf(l, alpha, sum);
            m = m_new;
            alpha_prev = alpha;
            fence_proxy_async_smem();          // P (generic-proxy stores) visible to the tensor core's async proxy
            tc_fence_before();
            mbar_arrive(s_free);
            mbar_arrive(p_full);
        }
        mbar_wait(o_full, (uint32_t)(n_kv - 1) & 1u);
        tc_fence_after();
        fold_o(alpha_prev);
        const int row = q0 + r;
        if (row < T) {
            const float inv = 1.0f / l;
            bf16* dst = out + ((long long)b * T + row) * (H * FT_D) + h * FT_D;
#pragma unroll
            for (int i = 0; i < FT_D; i += 8) {
                uint4 u;
                u.x = pack_bf16x2(o[i] * inv, o[i + 1] * inv);
                u.y = pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
                u.z = pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv);
                u.w = pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
                *reinterpret_cast<uint4*>(dst + i) = u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<FT_TMEM_COLS>(tmem_base);
    }
}

// qkv [B,T,3*H*64] bf16 (16-byte aligned), out [B,T,H*64] bf16 (16-byte aligned)
int mha_tc_fwd(const void* qkv, void* out, int B, int T, int H, float scale, cudaStream_t s) {
    CUtensorMap map;
    const uint64_t ld = (uint64_t)3 * H * FT_D;
    uint64_t dims[3] = {ld, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {ld * 2, ld * 2 * (uint64_t)T};
    uint32_t box[3] = {FT_D, FT_BN, 1};
    int rc = encode_tmap_bf16(&map, qkv, 3, dims, strides, box, 1);
    if (rc != A2F_OK) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        A2F_CHECK_CUDA(cudaFuncSetAttribute(mha_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
        attr_done = true;
    }
    const dim3 grid((T + FT_BM - 1) / FT_BM, H, B);
    A2F_CHECK_CUDA(launch_pdl(mha_tc_kernel, grid, dim3(FT_THREADS), (size_t)FT_SMEM, s, map, static_cast<bf16*>(out), T, H,
                              scale * 1.4426950408889634f));
    return A2F_OK;
}

}  // namespace a2f

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
//     warps 2..5   softmax: thread = query row = TMEM lane.  The S row comes into registers with one TMEM round trip
//                  (S is released at once, so Q K^T of the next tile overlaps the exponentials), ex2.approx on
//                  pre-scaled logits, row sum, bf16 P into the K-major 128B-swizzled layout the P V MMA reads.
//                  O stays in TMEM across key tiles (accumulating MMAs); it is rescaled only when the running row
//                  maximum has grown by more than 2^8 (lazy online softmax), one tcgen05.ld / st round trip.
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
                    umma_f16(t_o, pd, vdesc + (uint64_t)(128 * kk), idesc_pv, (j | kk) != 0 ? 1u : 0u);
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
        // O accumulates in TMEM across key tiles (the P V MMAs run with accumulate = 1).  `m_used` is the row maximum
        // that O and the row sum l are currently expressed against; it only moves (and O is only rescaled, a
        // tcgen05.ld / st round trip) when the running maximum has grown by more than 2^8 -- probabilities stay
        // below 256, far inside bf16 / fp32 range, and after the first tiles the rescale almost never triggers.
        float m_used = 0.f, l = 0.f;

        for (int j = 0; j < n_kv; ++j) {
            const int valid = min(FT_BN, T - j * FT_BN);   // live keys of this tile (uniform)
            mbar_wait(s_full, (uint32_t)j & 1u);
            tc_fence_after();
            // ---- the whole S row into registers with one TMEM round trip, then release S for Q K^T_{j+1} ----
            float v[FT_BN];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32(t_row + cc * 32, v + cc * 32);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free);
            if (valid < FT_BN) {
#pragma unroll
                for (int i = 0; i < FT_BN; ++i)
                    if (i >= valid) v[i] = -INFINITY;     // keys past T (zero-filled by TMA): exp2(-inf) = 0
            }
            float mx0 = v[0], mx1 = v[1], mx2 = v[2], mx3 = v[3];
#pragma unroll
            for (int i = 4; i < FT_BN; i += 4) {
                mx0 = fmaxf(mx0, v[i]);
                mx1 = fmaxf(mx1, v[i + 1]);
                mx2 = fmaxf(mx2, v[i + 2]);
                mx3 = fmaxf(mx3, v[i + 3]);
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            float alpha = 1.f;
            bool rescale = false;
            if (j == 0) {
                m_used = mx;
            } else if ((mx - m_used) * c > 8.0f) {
                alpha = ex2_approx((m_used - mx) * c);
                m_used = mx;
                rescale = true;
            }
            const float mc = m_used * c;
            // ---- p = 2^(s*c - m*c), row sum, packed to bf16 ----
            uint32_t pk[FT_BN / 2];
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < FT_BN; i += 2) {
                const float p0 = ex2_approx(fmaf(v[i], c, -mc));
                const float p1 = ex2_approx(fmaf(v[i + 1], c, -mc));
                s0 += p0;
                s1 += p1;
                pk[i >> 1] = pack_bf16x2(p0, p1);
            }
            l = fmaf(l, alpha, s0 + s1);
            // ---- P V_{j-1} must have retired before P is overwritten / O is rescaled ----
            if (j > 0) {
                mbar_wait(o_full, (uint32_t)(j - 1) & 1u);
                tc_fence_after();
                if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        float o[32];
                        tmem_ld_32x32(t_row + 128 + cc * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] *= alpha;
                        tmem_st_32x32(t_row + 128 + cc * 32, o);
                    }
                    tmem_st_wait();
                }
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                uint8_t* blk = prow + (cc >> 1) * FT_TILE_BYTES;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const int pch = ((cc & 1) * 4 + ch) ^ sw;
                    *reinterpret_cast<uint4*>(blk + pch * 16) =
                        make_uint4(pk[cc * 16 + ch * 4], pk[cc * 16 + ch * 4 + 1], pk[cc * 16 + ch * 4 + 2], pk[cc * 16 + ch * 4 + 3]);
                }
            }
            fence_proxy_async_smem();          // P (generic-proxy stores) visible to the tensor core's async proxy
            tc_fence_before();
            mbar_arrive(p_full);
        }
        mbar_wait(o_full, (uint32_t)(n_kv - 1) & 1u);
        tc_fence_after();
        const int row = q0 + r;
        const float inv = 1.0f / l;
        bf16* dst = out + ((long long)b * T + row) * (H * FT_D) + h * FT_D;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            float o[32];
            tmem_ld_32x32(t_row + 128 + cc * 32, o);
            tmem_ld_wait();
            if (row < T) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 u;
                    u.x = pack_bf16x2(o[i] * inv, o[i + 1] * inv);
                    u.y = pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
                    u.z = pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv);
                    u.w = pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
                    *reinterpret_cast<uint4*>(dst + cc * 32 + i) = u;
                }
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

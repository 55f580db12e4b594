// tcgen05 kernel for wav2vec2's positional convolution (HF:modeling_wav2vec2.py:326-379): Conv1d(768, 768, k=128,
// padding=64, groups=16) over channels-last activations, i.e. per group an implicit GEMM
//     out[t, n] = sum_{tap < 128} sum_{c < 48} h[t + tap + row_off, 48 g + c] * W[g][tap][n][c]
// whose A operand for consecutive taps is the SAME activation slab shifted by one time step.
//
// The generic GEMM kernel (gemm_tc.cu, legacy mode 2) fetched one 128 x 64 TMA box per tap: 2 MB of L2 -> SMEM traffic
// per 128-row tile (2.9 GB per forward at B=32 x 150 frames) and 25 % zero K-padding -- 161 TFLOP/s.  Here
//   * the activation slab of a work item ((MT*128 + 128) time steps x 48 channels) is loaded ONCE into shared memory in
//     the no-swizzle K-major layout [8-channel chunk][time][8]: a row is one 16-byte piece and consecutive rows are
//     consecutive pieces, so "shift by one tap" is +16 bytes on the UMMA descriptor's start address
//     (SBO = 128 B between 8-row groups, LBO = R*16 B between the two K chunks of one K=16 MMA);
//   * K is exactly 48 per tap (three K=16 MMAs, no padding);
//   * the weights stream through a 4-stage ring of 4-tap blocks; they are packed in global memory as the shared-memory
//     image of the B operand ([group][tap][chunk 6][n 48][8], 4608 contiguous bytes per tap) so that one stage is one
//     1-D cp.async.bulk;
//   * one CTA owns MT (<= 3) consecutive 128-row tiles of one utterance and one group: every weight block fetched from L2
//     feeds MT accumulators in TMEM;
//   * epilogue (4 warps): bias, GELU, residual (read back from the slab: the residual IS the input), bf16 stores.
// ~100 KB of shared memory and <= 256 TMEM columns per CTA: two CTAs per SM overlap one's prologue / epilogue with the
// other's MMAs.  Serves a2f_posconv (inference), a2f_posconv_pre and a2f_posconv_dgrad (training).
#include "a2f_common.cuh"
#include "gemm_params.cuh"

namespace a2f {

constexpr int PC_TPS = 4;                         // taps per weight stage
constexpr int PC_TAP_BYTES = 6 * 48 * 16;         // one tap of one group: [6 chunks][48 n][8 c] bf16
constexpr int PC_STAGE_BYTES = PC_TPS * PC_TAP_BYTES;
constexpr int PC_STAGES = 4;
constexpr int PC_THREADS = 192;                   // warp 0 weight producer, warp 1 MMA issuer, warps 2..5 epilogue
constexpr int PC_ITERS = 128 / PC_TPS;

struct PosconvTcParams {
    const bf16* A;        // [B, T, 768]
    const bf16* W;        // chunked packed weight
    const float* bias;    // [768] or NULL
    const bf16* resid;    // external residual [B, T, 768] (resid_mode 2)
    bf16* C;              // [B, T, 768]
    int B, T;
    int row_off;          // -64 forward, -63 data gradient (flipped taps)
    int act;              // A2F_ACT_NONE / A2F_ACT_GELU
    int resid_mode;       // 0 none, 1 the input itself (read from the slab), 2 external
    int chunks_per_utt;   // ceil(T / (MT*128))
    int swap_strides;     // debug: exchange the descriptor's LBO / SBO fields
};

template <int MT> struct PcCfg {
    static constexpr int R = MT * 128 + 128;                       // slab rows (time steps)
    static constexpr int SLAB_BYTES = 6 * R * 16;
    static constexpr int TMEM_COLS = MT == 1 ? 64 : MT == 2 ? 128 : 256;
    static constexpr size_t SMEM_BYTES = (size_t)SLAB_BYTES + (size_t)PC_STAGES * PC_STAGE_BYTES + 128;
};

A2F_D void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// K-major, no swizzle: 8 rows x 16 B core matrices; lbo = distance between the two K chunks, sbo = between 8-row groups
A2F_D uint64_t pc_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo16 & 0x3FFFu) << 16) | ((uint64_t)(sbo16 & 0x3FFFu) << 32) |
           (1ull << 46);
}

template <int MT>
__global__ void __launch_bounds__(PC_THREADS, 2) posconv_tc_kernel(const PosconvTcParams p) {
    using Cfg = PcCfg<MT>;
    constexpr int R = Cfg::R;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* slab = smem;
    uint8_t* sB = smem + Cfg::SLAB_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)PC_STAGES * PC_STAGE_BYTES);
    uint64_t* full_bar = bars;                  // [PC_STAGES]
    uint64_t* empty_bar = bars + PC_STAGES;     // [PC_STAGES]
    uint64_t* tfull_bar = bars + 2 * PC_STAGES; // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = blockIdx.x & 15;
    const int item = blockIdx.x >> 4;
    const int b = item / p.chunks_per_utt;
    const int t0 = (item - b * p.chunks_per_utt) * (MT * 128);

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < PC_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(tfull_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    if (warp == 0) {
        // ===================== weight producer =====================
        if (lane == 0) {
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.W) + (size_t)grp * 128 * PC_TAP_BYTES;
            for (int it = 0; it < PC_ITERS; ++it) {
                const int s = it % PC_STAGES;
                mbar_wait(&empty_bar[s], (((uint32_t)(it / PC_STAGES)) & 1u) ^ 1u);
                mbar_expect_tx(&full_bar[s], PC_STAGE_BYTES);
                bulk_load_1d(sB + (size_t)s * PC_STAGE_BYTES, wsrc + (size_t)it * PC_STAGE_BYTES, PC_STAGE_BYTES, &full_bar[s]);
            }
        }
        __syncwarp();
    } else {
        // ===================== activation slab: rows t0 + row_off + [0, R), zero outside [0, T) =====================
        {
            const bf16* abase = p.A + (size_t)b * p.T * 768 + grp * 48;
            const int tbase = t0 + p.row_off;
#pragma unroll 4
            for (int pc = threadIdx.x - 32; pc < 6 * R; pc += PC_THREADS - 32) {
                const int c = pc / R, r = pc - c * R;
                const int t = tbase + r;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (t >= 0 && t < p.T) v = *reinterpret_cast<const uint4*>(abase + (size_t)t * 768 + c * 8);
                *reinterpret_cast<uint4*>(slab + (size_t)pc * 16) = v;
            }
            fence_proxy_async_smem();            // generic-proxy writes -> visible to the tensor core (async proxy)
            named_bar_sync(1, PC_THREADS - 32);
        }
        if (warp == 1) {
            // ===================== MMA issuer =====================
            if (lane == 0) {
                tc_fence_after();
                // instruction descriptor: D=f32, A=B=bf16, both K-major, N=48, M=128
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(48 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                const uint32_t slab_addr = smem_u32(slab), b_addr = smem_u32(sB);
                const uint32_t a_lbo = p.swap_strides ? 8u : (uint32_t)R, a_sbo = p.swap_strides ? (uint32_t)R : 8u;
                const uint32_t b_lbo = p.swap_strides ? 8u : 48u, b_sbo = p.swap_strides ? 48u : 8u;
                for (int it = 0; it < PC_ITERS; ++it) {
                    const int s = it % PC_STAGES;
                    mbar_wait(&full_bar[s], ((uint32_t)(it / PC_STAGES)) & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int tp = 0; tp < PC_TPS; ++tp) {
                        const int tap = it * PC_TPS + tp;
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const uint64_t bdesc = pc_desc(b_addr + (uint32_t)(s * PC_STAGE_BYTES + tp * PC_TAP_BYTES + k * 2 * 768),
                                                           b_lbo, b_sbo);
#pragma unroll
                            for (int i = 0; i < MT; ++i) {
                                const uint64_t adesc = pc_desc(slab_addr + (uint32_t)((2 * k * R + i * 128 + tap) * 16), a_lbo, a_sbo);
                                umma_f16(tmem_base + (uint32_t)(i * 64), adesc, bdesc, idesc, (tap | k) != 0 ? 1u : 0u);
                            }
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(tfull_bar);
            }
            __syncwarp();
        } else {
            // ===================== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====================
            const int q = warp & 3;
            const int r = q * 32 + lane;
            const float* __restrict__ bias = p.bias ? p.bias + grp * 48 : nullptr;
            mbar_wait(tfull_bar, 0);
            tc_fence_after();
#pragma unroll 1
            for (int i = 0; i < MT; ++i) {
                const int t = t0 + i * 128 + r;
                float v[48];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(i * 64);
                tmem_ld_32x16(taddr, v);
                tmem_ld_32x16(taddr + 16, v + 16);
                tmem_ld_32x16(taddr + 32, v + 32);
                tmem_ld_wait();
                if (t < p.T) {
                    if (bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 48; j += 4) {
                            const float4 f = __ldg(reinterpret_cast<const float4*>(bias + j));
                            v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
                        }
                    }
                    if (p.act == A2F_ACT_GELU) {
#pragma unroll
                        for (int j = 0; j < 48; j += 2) {
                            const float2 gl = gelu_fast2(make_float2(v[j], v[j + 1]));
                            v[j] = gl.x;
                            v[j + 1] = gl.y;
                        }
                    }
                    const size_t goff = ((size_t)b * p.T + t) * 768 + grp * 48;
                    const int srow = i * 128 + r - p.row_off;      // slab row that holds the input at time t
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        if (p.resid_mode != 0) {
                            uint4 u;
                            if (p.resid_mode == 1) u = *reinterpret_cast<const uint4*>(slab + ((size_t)c * R + srow) * 16);
                            else u = *reinterpret_cast<const uint4*>(p.resid + goff + c * 8);
                            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __bfloat1622float2(h2[e]);
                                v[c * 8 + 2 * e] += f.x;
                                v[c * 8 + 2 * e + 1] += f.y;
                            }
                        }
                        uint4 o;
                        o.x = pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]);
                        o.y = pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]);
                        o.z = pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]);
                        o.w = pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]);
                        *reinterpret_cast<uint4*>(p.C + goff + c * 8) = o;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

static int g_pc_swap = 0;
void set_posconv_swap(int v) { g_pc_swap = v; }

template <int MT> static int launch_posconv(PosconvTcParams& p, cudaStream_t s) {
    using Cfg = PcCfg<MT>;
    static int configured[64] = {0};
    int dev = 0;
    A2F_CHECK_CUDA(cudaGetDevice(&dev));
    auto kern = posconv_tc_kernel<MT>;
    if (dev < 64 && !configured[dev]) {
        A2F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
        configured[dev] = 1;
    }
    p.chunks_per_utt = (p.T + MT * 128 - 1) / (MT * 128);
    const long long grid = (long long)p.B * p.chunks_per_utt * 16;
    A2F_REQUIRE(grid < (1LL << 31), "posconv_tc: grid too large");
    A2F_CHECK_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(PC_THREADS), Cfg::SMEM_BYTES, s, p));
    count_launch();
    return A2F_OK;
}

// A, C (and resid) are bf16 channels-last [B, T, 768]; W is the chunked layout a2f_pack_posconv_weight writes for kpad = 8.
int posconv_tc(const void* A, const void* W, const float* bias, const void* resid, int resid_mode, void* C, int B, int T,
               int row_off, int act, cudaStream_t s) {
    A2F_REQUIRE(act == A2F_ACT_NONE || act == A2F_ACT_GELU, "posconv_tc: activation must be none or GELU");
    A2F_REQUIRE((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C) |
                 reinterpret_cast<uintptr_t>(resid)) % 16 == 0, "posconv_tc: operands must be 16-byte aligned");
    PosconvTcParams p;
    p.A = static_cast<const bf16*>(A);
    p.W = static_cast<const bf16*>(W);
    p.bias = bias;
    p.resid = static_cast<const bf16*>(resid);
    p.C = static_cast<bf16*>(C);
    p.B = B; p.T = T; p.row_off = row_off; p.act = act; p.resid_mode = resid_mode;
    p.swap_strides = g_pc_swap;
    // tiles per CTA: minimise max(MMA time of the 128-row tiles issued, weight-stream time of the CTAs launched) per
    // (utterance, group) -- ~12k cycles per tile, ~14k cycles per 590 KB weight stream; ties go to the larger MT
    const int tiles = (T + 127) / 128;
    int best = 1, best_cost = 1 << 30;
    for (int mt = 1; mt <= 3; ++mt) {
        const int chunks = (tiles + mt - 1) / mt;
        const int cost = max(chunks * mt * 12, chunks * 14);
        if (cost <= best_cost) { best = mt; best_cost = cost; }
    }
    if (best == 1) return launch_posconv<1>(p, s);
    if (best == 2) return launch_posconv<2>(p, s);
    return launch_posconv<3>(p, s);
}

}  // namespace a2f

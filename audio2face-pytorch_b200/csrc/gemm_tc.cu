// tcgen05 / TMEM / TMA GEMM for sm_100a: C = act(A W^T + bias) + resid + tmpl, bf16 operands, fp32 accumulate.
//
// Persistent, warp-specialised, one CTA per SM (320 threads):
//   warp 0      TMA producer   (cp.async.bulk.tensor into a STAGES-deep 128B-swizzled smem ring)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma; accumulators double-buffered in TMEM)
//   warps 2..9  epilogue       (tcgen05.ld -> bias/act/residual -> smem -> global), overlaps the next tile's mainloop;
//                               warp%4 selects the TMEM lane quarter, (warp-2)/4 the column half of the tile
// Two epilogue flavours:
//   TMA store   16-byte-aligned outputs: results are staged in 128B-swizzled smem blocks and written with
//               cp.async.bulk.tensor stores (full-line writes, ragged M/N edges clipped by the tensor map)
//   scalar      the 15069-wide vertex head (rows only 4-byte aligned) and the template-add epilogue: 32x32
//               transposes through smem so that each warp store covers 32 consecutive floats of one row.
// A is addressed through <=3-D tensor maps so that the same mainloop serves
//   mode 0: plain [M,K] matrices and the stride-2 Conv1d stack as an implicit GEMM over channels-last activations
//           (rows of the im2col matrix are strided views of the activation, split into K segments so that every
//            tensor map has non-overlapping rows),
//   mode 2: the k=128 grouped positional conv (K block = one tap, A box shifted by one time step per block).
#include "a2f_common.cuh"
#include "gemm_params.cuh"

namespace a2f {

constexpr int TBM = 128;          // tile rows (UMMA M)
constexpr int TBK = 64;           // K per stage: 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int TC_THREADS = 320;
constexpr int A_STAGE_BYTES = TBM * TBK * 2;
constexpr int MAX_SEGS = 4;
constexpr int EPI_STAGE_BYTES = 16384;   // one 128-row x 128-byte store block (or 4 warps x 32x32 fp32 transposes)

struct TmapSet {
    CUtensorMap a[MAX_SEGS];
    CUtensorMap b;
    CUtensorMap bs;                 // pair kernel: B map of the split tail tiles (box rows = BN / 2 / tail_split)
    CUtensorMap c;
    CUtensorMap c2;                 // pair kernel: second output act(z) (GemmParams::C2)
    CUtensorMap r;                  // residual / saved pre-activation (pair kernel, bf16, 16-byte aligned rows)
};

struct TcParams {
    GemmParams g;
    int mode;
    int tiles_m_per_batch, num_batches, tiles_n, num_k_blocks, kb_per_seg;
    int resid_vec_ok;               // 16-byte vector residual loads are legal
    int bias_vec_ok;                // 16-byte vector bias loads are legal
    int fast_gelu;                  // bf16 output: A-S erf approximation instead of erff
    int resid_tma;                  // pair kernel: the residual tile arrives by TMA in the store-staging layout
    // pair kernel, wave quantisation: the first `full_tiles` tiles run at full width; every later tile is cut into
    // `tail_split` column slices (256 / tail_split wide) that are scheduled as tiles of their own, so the last, partly
    // filled wave costs 1 / tail_split of a tile time (FFN1 at M=4800: 228 tiles on 74 pairs = 3.08 waves -> 3.25)
    int full_tiles, tail_split;
    uint32_t lbo_enc, sbo_enc, desc_version, desc_layout;
    unsigned long long* timeline;   // debug (a2f_debug_set_timeline): 8 globaltimer stamps per CTA, else NULL
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TL_STAMP(slot) do { if (p.timeline) { p.timeline[blockIdx.x * 8 + (slot)] = gtimer(); \
                                               p.timeline[(gridDim.x + blockIdx.x) * 8 + (slot)] = (unsigned long long)clock64(); } } while (0)

// debug knobs (a2f_debug_set_umma_field): defaults are the CUTLASS-documented K-major SWIZZLE_128B values
static uint32_t g_umma_fields[4] = {1u, 64u, 1u, 2u};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, const TcParams& p) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(p.lbo_enc & 0x3FFFu) << 16) |
           ((uint64_t)(p.sbo_enc & 0x3FFFu) << 32) | ((uint64_t)(p.desc_version & 3u) << 46) |
           ((uint64_t)(p.desc_layout & 7u) << 61);
}

// WIDE: the scalar epilogue of a 256-column fp32 tile gives every epilogue warp 32 rows x 128 columns, staged 16 rows at
// a time (8 KB per warp), so that every output / template row is touched in 512-byte runs.  The operand ring shrinks to
// 2 stages (the vertex heads have K <= 192 and are bound by the HBM traffic of the epilogue, not by the mainloop): 160 KB
// of shared memory in total, which leaves the L1 large enough for the template loads in flight.
template <int BN, typename TC, bool SCALAR> constexpr bool tc_wide_v = SCALAR && BN == 256 && sizeof(TC) == 4;
constexpr int WIDE_WARP_FLOATS = 16 * 128;      // staging block of one epilogue warp: 16 rows x 128 columns
constexpr int WIDE_WR = 8;               // rows of the rolling template window (WIDE_WR x 4 loads per lane in flight)

template <int BN, typename TC, bool WIDE = false> struct TcCfg {
    static constexpr int ACC_COLS = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;
    static constexpr int TMEM_COLS = 2 * ACC_COLS;
    static constexpr int B_STAGE_BYTES = BN * TBK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = WIDE ? 2 : (BN >= 256) ? 4 : (BN >= 128) ? 6 : 8;
    static constexpr int EPI_BYTES = WIDE ? 8 * WIDE_WARP_FLOATS * 4 : 2 * EPI_STAGE_BYTES;
    // TMA-store block width (columns): 128 bytes of output per row, except for the 48-wide posconv tiles
    static constexpr int SBW = (BN % 32 != 0) ? ((sizeof(TC) == 2) ? BN : 16) : (int)(128 / sizeof(TC));
    static constexpr int NBLK = BN / SBW;
    static constexpr int ROW_PITCH = SBW * (int)sizeof(TC);
    static constexpr bool SWZ = (ROW_PITCH == 128);
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 256 /*barriers*/;
};

// ---- epilogue helpers: every data-independent condition is tested once per chunk, never per element ----
template <int CH>
__device__ __forceinline__ void epi_bias_act(float* v, const float* __restrict__ bias, int bias_vec_ok, int nc, int n_end,
                                             int act, int fast_gelu) {
    if (bias != nullptr) {
        if (bias_vec_ok && nc + CH <= n_end) {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
                const float4 f = __ldg(reinterpret_cast<const float4*>(bias + nc + j));
                v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < CH; ++j)
                if (nc + j < n_end) v[j] += __ldg(bias + nc + j);
        }
    }
    if (act == A2F_ACT_GELU) {
        if (fast_gelu) {
#pragma unroll
            for (int j = 0; j < CH; j += 2) {       // packed fp32 (FFMA2 / FMUL2): half the issue slots of the scalar form
                const float2 r = gelu_fast2(make_float2(v[j], v[j + 1]));
                v[j] = r.x;
                v[j + 1] = r.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = gelu_erf(v[j]);
        }
    } else if (act == A2F_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = relu(v[j]);
    } else if (act == A2F_ACT_TANH) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = tanhf(v[j]);
    }
}

template <int CH>
__device__ __forceinline__ void epi_resid(float* v, const void* __restrict__ resid, int resid_bf16, long long off, int nc,
                                          int n_end, int vec_ok) {
    const bool full = vec_ok && (nc + CH <= n_end);
    if (resid_bf16) {
        const bf16* rp = static_cast<const bf16*>(resid) + off;
        if (full) {
#pragma unroll
            for (int j = 0; j < CH; j += 8) {
                const uint4 u = *reinterpret_cast<const uint4*>(rp + j);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h2[e]);
                    v[j + 2 * e] += f.x;
                    v[j + 2 * e + 1] += f.y;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < CH; ++j)
                if (nc + j < n_end) v[j] += __bfloat162float(rp[j]);
        }
    } else {
        const float* rp = static_cast<const float*>(resid) + off;
        if (full) {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
                const float4 f = *reinterpret_cast<const float4*>(rp + j);
                v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < CH; ++j)
                if (nc + j < n_end) v[j] += rp[j];
        }
    }
}

// ---- wide scalar epilogue helpers (gemm_tc_kernel, WIDE) ----
// Position the template refill stream on output row m0 (template row m0 / rows_per_tmpl), column ncol.
A2F_D void wide_fill_start(const float* tmpl, long long m0, int ncol, int rpt, int N, const float*& tp, int& rem) {
    if (tmpl == nullptr) return;
    const unsigned tr = (unsigned)m0 / (unsigned)rpt;
    rem = (int)((unsigned)m0 - tr * (unsigned)rpt);
    tp = tmpl + (long long)tr * N + ncol;
}
// WIDE_WR output rows [row_base, row_base+WIDE_WR) of a warp's 32 x 128 slice: out = (acc + bias) + template, where the template
// values come from the rolling window `tadd`; every consumed window slot is refilled at once from the refill stream
// (WIDE_WR rows ahead).  PRED = false: slice and refill target are complete, no predicates.
// HAS_T is a template parameter on purpose: with a run-time flag the compiler loads into a temporary and moves it into
// the window under a predicate, and that move waits for the load at once (profiles/r1_voca_head_wide.txt: the hottest
// instructions were those moves), which serialises the window.
template <bool PRED, bool HAS_T>
A2F_D void wide_rows(float (&tadd)[WIDE_WR * 4], const float* trw, const float (&bj)[4], float* cp, int ldc, int row_base, int rows,
                     const bool (&cok)[4], const float*& tp, int& rem, int rpt, int N, int fill_rows,
                     const bool (&fok)[4], int lane) {
#pragma unroll
    for (int r = 0; r < WIDE_WR; ++r) {
        const int rr = row_base + r;
        float sv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) sv[c] = trw[c * 512 + (rr & 15) * 32 + ((lane + rr) & 31)];
        float* cpr = cp + (long long)(rr * ldc);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float o = (sv[c] + bj[c]) + tadd[r * 4 + c];
            if (!PRED || (rr < rows && cok[c])) __stcs(cpr + c * 32, o);     // written once, never re-read here
        }
        if (HAS_T) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (PRED) tadd[r * 4 + c] = (r < fill_rows && fok[c]) ? __ldg(tp + c * 32) : 0.f;
                else tadd[r * 4 + c] = __ldg(tp + c * 32);
            }
            if (++rem == rpt) { rem = 0; tp += N; }
        }
    }
}

// LOSS: the fused vertex-head + loss epilogue (a2f_vertex_head_loss) is its own instantiation, so that its registers and code
// do not weigh on the inference head (an earlier version shared one kernel: the VOCA head lost 7 % to spills).
template <int BN, typename TC, bool SCALAR, bool LOSS = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ TmapSet maps, const TcParams p) {
    constexpr bool WIDE = tc_wide_v<BN, TC, SCALAR>;
    using Cfg = TcCfg<BN, TC, WIDE>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)STAGES * A_STAGE_BYTES;
    uint8_t* sEpi = smem + (size_t)STAGES * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + Cfg::EPI_BYTES);
    uint64_t* full_bar = bars;                 // [STAGES]
    uint64_t* empty_bar = bars + STAGES;       // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
    uint64_t* tempty_bar = tfull_bar + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GemmParams& g = p.g;
    if (threadIdx.x == 0) TL_STAMP(0);                                  // kernel entry

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B needs 1024-byte alignment
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a[0]);
        tma_prefetch_desc(&maps.b);
        if (!SCALAR) tma_prefetch_desc(&maps.c);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 8);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();     // PDL: everything above overlapped the previous kernel's tail; global memory is touched below
    if (threadIdx.x == 0) TL_STAMP(1);                                  // setup done (barriers, TMEM)

    const int tiles_m = p.tiles_m_per_batch * p.num_batches;
    const int total_tiles = tiles_m * p.tiles_n;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int waited = -1;                      // last counter index known to have reached its target
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nb = tile % p.tiles_n, mb = tile / p.tiles_n;
                const int b = mb / p.tiles_m_per_batch, lt = mb % p.tiles_m_per_batch;
                if (p.g.wait_counters != nullptr) {
                    // rows of this batch are being produced by a kernel that runs at the same time (decoder rollout):
                    // generic-proxy stores + red.release there, ld.acquire here, then the async proxy (TMA) may read
                    const int f = min(p.g.wait_n, (b + 1) * p.g.wait_per_batch) - 1;
                    if (f > waited) {
                        const unsigned* cnt = p.g.wait_counters + f;
                        uint32_t spins = 0;
                        for (;;) {
                            unsigned v;
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
                            if (v >= (unsigned)p.g.wait_target) break;
                            __nanosleep(200);
                            if (++spins > (1u << 24)) __trap();     // ~3 s: a missing producer becomes a launch error
                        }
                        asm volatile("fence.proxy.async;" ::: "memory");
                        waited = f;
                    }
                }
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    uint8_t* dstA = sA + (size_t)stage * A_STAGE_BYTES;
                    uint8_t* dstB = sB + (size_t)stage * Cfg::B_STAGE_BYTES;
                    if (p.mode == 2) {
                        tma_load_3d(dstA, &maps.a[0], &full_bar[stage], nb * 48, lt * TBM + kb - 64 + p.g.seg_row_off[0], b);
                        tma_load_2d(dstB, &maps.b, &full_bar[stage], kb * TBK, nb * 48);
                    } else if (p.g.n_seg > 0) {
                        // explicit K segments (data gradient of the strided convs): one map, per-segment row / column
                        // offsets; rows outside [0, a_rows) are zero-filled by the TMA unit
                        const int seg = kb / p.kb_per_seg, kin = (kb - seg * p.kb_per_seg) * TBK;
                        tma_load_3d(dstA, &maps.a[0], &full_bar[stage], p.g.seg_col_off[seg] + kin,
                                    lt * TBM + p.g.seg_row_off[seg], b);
                        tma_load_2d(dstB, &maps.b, &full_bar[stage], kb * TBK, nb * BN);
                    } else {
                        const int seg = kb / p.kb_per_seg, kin = (kb - seg * p.kb_per_seg) * TBK;
                        tma_load_3d(dstA, &maps.a[seg], &full_bar[stage], kin, lt * TBM, b);
                        tma_load_2d(dstB, &maps.b, &full_bar[stage], kb * TBK, nb * BN);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(TBM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS);
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    if (tile == (int)blockIdx.x && kb == 0) TL_STAMP(2);   // first operands landed
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_u32(sA + (size_t)stage * A_STAGE_BYTES), p);
                    const uint64_t bdesc = make_smem_desc(smem_u32(sB + (size_t)stage * Cfg::B_STAGE_BYTES), p);
#pragma unroll
                    for (int k = 0; k < TBK / UMMA_K; ++k) {
                        // advance 32 bytes (16 bf16) along K inside the 128B swizzle row: +2 in the >>4 address field
                        umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                 (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);   // smem slot reusable once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);         // accumulator complete -> epilogue
                if (tile == (int)blockIdx.x) TL_STAMP(3);                  // first tile's MMAs all issued
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: 8 warps =====================
        const int ew = warp - 2;
        const int q = warp & 3;            // TMEM lane quarter this warp may read
        const int half = ew >> 2;          // column half of the tile
        const bool leader = ((ew & 3) == 0) && lane == 0;
        const int bar_id = 1 + half;
        uint8_t* stage_buf = sEpi + half * EPI_STAGE_BYTES;
        int acc = 0;
        uint32_t acc_phase = 0;
        TC* __restrict__ C = static_cast<TC*>(g.C);
        // kernel parameters used per element live in registers (constant-bank reads showed up as stalls, profiles/r1_ffn1)
        const float* __restrict__ e_bias = g.bias;
        const void* __restrict__ e_resid = g.resid;
        const int e_act = g.act;
        const bool e_dact = g.resid_mode == A2F_RESID_DACT;
        // ---- wide scalar epilogue (vertex heads): this warp owns 32 rows x 128 columns of every tile ----
        const bool wide = !LOSS && WIDE && g.resid == nullptr && g.act == A2F_ACT_NONE && p.mode != 2;
        double loss_rec = 0.0, loss_vel = 0.0;   // fused-loss epilogue: this thread's share of the two sums
        float tadd[WIDE_WR * 4];    // template values of the NEXT WIDE_WR rows x 4 chunks, always in flight
        bool tadd_primed = false;
        const float* fill_tp = nullptr;         // template row the refill stream reads next
        int fill_rem = 0;                       // output rows already served by that template row
        const float* __restrict__ e_tmpl = g.tmpl;
        const int e_N = g.N;
        // slice of tile `t_` owned by this warp: output pointer of (row 0, column lane), first global row, live rows
        auto wslice = [&](int t_, float*& cp_, long long& m0_, int& rows_, int& ncol_) {
            if (t_ >= total_tiles) { cp_ = nullptr; m0_ = 0; rows_ = 0; ncol_ = 0; return; }
            const int nb_ = t_ % p.tiles_n, mb_ = t_ / p.tiles_n;
            const int b_ = mb_ / p.tiles_m_per_batch, lt_ = mb_ % p.tiles_m_per_batch;
            const int r0_ = lt_ * TBM + q * 32;
            rows_ = min(32, max(0, g.rows_per_batch - r0_));
            ncol_ = nb_ * BN + half * 128 + lane;
            m0_ = (long long)b_ * g.rows_per_batch + r0_;
            cp_ = reinterpret_cast<float*>(g.C) + (long long)b_ * g.c_batch_stride + (long long)r0_ * g.ldc + ncol_;
            if (g.perm_rows != 0) {
                // frame-major rows (a2f_vertex_head_stream): this 32-row slice = utterances ut..ut+31 of frame fr of the group
                const int fr = r0_ / g.perm_rows, ut = r0_ - fr * g.perm_rows;
                if (m0_ >= g.live_rows) rows_ = 0;
                cp_ = reinterpret_cast<float*>(g.C) + (long long)b_ * g.c_batch_stride + (long long)fr * g.perm_stride +
                      (long long)ut * g.ldc + ncol_;
                m0_ = ut;                                     // template row = utterance
            }
        };
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nb = tile % p.tiles_n, mb = tile / p.tiles_n;
            const int b = mb / p.tiles_m_per_batch, lt = mb % p.tiles_m_per_batch;
            const int n_tile0 = (p.mode == 2) ? nb * 48 : nb * BN;          // first global column of this tile
            const int n_lim = (p.mode == 2) ? 48 : min(BN, g.N - n_tile0);   // live columns in this tile
            const int n_end = n_tile0 + n_lim;

            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if (threadIdx.x == 64 && tile == (int)blockIdx.x) TL_STAMP(4);  // first accumulator complete

            const int r_tile = q * 32 + lane;
            const int r_in_batch = lt * TBM + r_tile;
            const bool row_ok = r_in_batch < g.rows_per_batch;
            const long long m = (long long)b * g.rows_per_batch + r_in_batch;
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::ACC_COLS);

            if (!SCALAR) {
                constexpr int SBW = Cfg::SBW;
                constexpr int CH = (SBW % 32 == 0) ? 32 : 16;
                constexpr int EPC = 16 / (int)sizeof(TC);      // elements per 16-byte chunk
                uint8_t* rowp = stage_buf + r_tile * Cfg::ROW_PITCH;
#pragma unroll 1
                for (int blk = half; blk < Cfg::NBLK; blk += 2) {
                    const int col0 = blk * SBW;          // within the tile
                    if (col0 >= n_lim) break;            // uniform across the CTA
                    const int ncol0 = n_tile0 + col0;    // global column of the block
                    // accumulator -> registers, bias / activation / residual (overlaps the previous block's TMA store)
                    float v[SBW];
#pragma unroll
                    for (int cc = 0; cc < SBW / CH; ++cc) {
                        if (CH == 32) tmem_ld_32x32(t_row + col0 + cc * CH, v + cc * CH);
                        else tmem_ld_32x16(t_row + col0 + cc * CH, v + cc * CH);
                    }
                    tmem_ld_wait();
                    const long long r_off = (long long)b * g.r_batch_stride + (long long)r_in_batch * g.ldr + ncol0;
                    if (e_dact) {
                        // activation backward fused into the data gradient: v *= act'(z), z = saved pre-activation
                        if (row_ok) {
                            float z[SBW];
#pragma unroll
                            for (int j = 0; j < SBW; ++j) z[j] = 0.f;
                            epi_resid<SBW>(z, e_resid, g.resid_bf16, r_off, ncol0, n_end, p.resid_vec_ok);
#pragma unroll
                            for (int j = 0; j < SBW; ++j) v[j] *= act_grad(z[j], e_act);
                        }
                    } else {
                        epi_bias_act<SBW>(v, e_bias, p.bias_vec_ok, ncol0, n_end, e_act, p.fast_gelu);
                        if (e_resid != nullptr && row_ok)
                            epi_resid<SBW>(v, e_resid, g.resid_bf16, r_off, ncol0, n_end, p.resid_vec_ok);
                    }
                    // the previous TMA store of this half must have finished reading the staging block
                    if (leader) tma_store_wait_read();
                    named_bar_sync(bar_id, 128);
#pragma unroll
                    for (int ch = 0; ch < SBW / EPC; ++ch) {
                        const int pch = Cfg::SWZ ? (ch ^ (r_tile & 7)) : ch;
                        uint4 u;
                        if (sizeof(TC) == 2) {
                            u.x = pack_bf16x2(v[ch * 8 + 0], v[ch * 8 + 1]);
                            u.y = pack_bf16x2(v[ch * 8 + 2], v[ch * 8 + 3]);
                            u.z = pack_bf16x2(v[ch * 8 + 4], v[ch * 8 + 5]);
                            u.w = pack_bf16x2(v[ch * 8 + 6], v[ch * 8 + 7]);
                        } else {
                            u.x = __float_as_uint(v[ch * EPC + 0]);
                            u.y = __float_as_uint(v[ch * EPC + 1]);
                            u.z = __float_as_uint(v[ch * EPC + 2]);
                            u.w = __float_as_uint(v[ch * EPC + 3]);
                        }
                        *reinterpret_cast<uint4*>(rowp + pch * 16) = u;
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(bar_id, 128);
                    if (leader) {
                        tma_store_3d(&maps.c, stage_buf, ncol0, lt * TBM, b);
                        tma_store_commit();
                    }
                }
            } else if (WIDE && wide) {
                // 512-byte runs per row: HBM pages are touched in 4x longer bursts than by the 32x32 chunks below, and
                // 32 template loads per lane stay in flight through the whole tile (rolling WIDE_WR-row window that runs
                // ahead into the next tile), profiles/r1_vertex_head_access_pattern.txt
                float* cp; long long m0; int rows, ncol;
                wslice(tile, cp, m0, rows, ncol);
                bool cok[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) cok[c] = ncol + c * 32 < e_N;
                if (!tadd_primed) {
                    wide_fill_start(e_tmpl, m0, ncol, g.rows_per_tmpl, e_N, fill_tp, fill_rem);
#pragma unroll
                    for (int r = 0; r < WIDE_WR; ++r) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            tadd[r * 4 + c] = (e_tmpl && r < rows && cok[c]) ? __ldg(fill_tp + c * 32) : 0.f;
                        if (++fill_rem == g.rows_per_tmpl) { fill_rem = 0; fill_tp += e_N; }
                    }
                    tadd_primed = true;
                }
                float* trw = reinterpret_cast<float*>(sEpi) + ew * WIDE_WARP_FLOATS;
                float bj[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) bj[c] = (e_bias && cok[c]) ? __ldg(e_bias + ncol + c * 32) : 0.f;
                const int ldc_i = (int)g.ldc;
                float* cpn; long long m0n; int rowsn, ncoln;
                wslice(tile + (int)gridDim.x, cpn, m0n, rowsn, ncoln);
                bool cokn[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) cokn[c] = ncoln + c * 32 < e_N;
                // interior tiles (this slice and the next one complete): no predicates in the 128-store body
                const bool interior = rows == 32 && rowsn == 32 && n_tile0 + half * 128 + 128 <= e_N &&
                                      __all_sync(0xffffffffu, cokn[3]);
                // Two passes of 16 rows: the staging block holds 16 rows x 128 columns (8 KB per warp), so that the
                // kernel's shared memory stays at 160 KB and the L1 keeps enough lines for the template loads in flight
                // (with the 228 KB carve-out a pure streaming kernel of this pattern drops from 6.0 to 4.1 TB/s,
                // profiles/r1_vertex_head_access_pattern.txt).  Each pass re-reads the accumulator chunks from TMEM and
                // the lanes of its row half write them transposed.
#pragma unroll
                for (int hp = 0; hp < 2; ++hp) {
                    __syncwarp();
#pragma unroll 1
                    for (int cc = 0; cc < 4; ++cc) {
                        if (half * 128 + cc * 32 >= n_lim) break;
                        float v[32];
                        tmem_ld_32x32(t_row + half * 128 + cc * 32, v);
                        tmem_ld_wait();
                        if ((lane >> 4) == hp) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) trw[cc * 512 + (lane & 15) * 32 + ((j + lane) & 31)] = v[j];
                        }
                    }
                    if (hp == 1) tc_fence_before();
                    __syncwarp();
                    // second pass read: hand the TMEM buffer back before the remaining stores
                    if (hp == 1 && lane == 0) mbar_arrive(&tempty_bar[acc]);
                    // each phase consumes the window (WIDE_WR rows) and refills it with the rows WIDE_WR further on:
                    // the last phase of a tile fetches the first rows of the next tile of this CTA
                    if (interior) {
#pragma unroll
                        for (int ph = hp * (16 / WIDE_WR); ph < (hp + 1) * (16 / WIDE_WR); ++ph) {
                            if (ph == 32 / WIDE_WR - 1)
                                wide_fill_start(e_tmpl, m0n, ncoln, g.rows_per_tmpl, e_N, fill_tp, fill_rem);
                            if (e_tmpl != nullptr)
                                wide_rows<false, true>(tadd, trw, bj, cp, ldc_i, ph * WIDE_WR, 32, cok, fill_tp, fill_rem,
                                                       g.rows_per_tmpl, e_N, WIDE_WR, cok, lane);
                            else
                                wide_rows<false, false>(tadd, trw, bj, cp, ldc_i, ph * WIDE_WR, 32, cok, fill_tp, fill_rem,
                                                        g.rows_per_tmpl, e_N, WIDE_WR, cok, lane);
                        }
                    } else {
#pragma unroll 1
                        for (int ph = hp * (16 / WIDE_WR); ph < (hp + 1) * (16 / WIDE_WR); ++ph) {
                            const bool last = ph == 32 / WIDE_WR - 1;
                            if (last) wide_fill_start(e_tmpl, m0n, ncoln, g.rows_per_tmpl, e_N, fill_tp, fill_rem);
                            bool fok[4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) fok[c] = last ? cokn[c] : cok[c];
                            const int fill_rows = last ? rowsn : rows - (ph + 1) * WIDE_WR;
                            if (e_tmpl != nullptr)
                                wide_rows<true, true>(tadd, trw, bj, cp, ldc_i, ph * WIDE_WR, rows, cok, fill_tp, fill_rem,
                                                      g.rows_per_tmpl, e_N, fill_rows, fok, lane);
                            else
                                wide_rows<true, false>(tadd, trw, bj, cp, ldc_i, ph * WIDE_WR, rows, cok, fill_tp, fill_rem,
                                                       g.rows_per_tmpl, e_N, fill_rows, fok, lane);
                        }
                    }
                }
            } else {
                // scalar epilogue: 32x32 fp32 transposes, coalesced 4-byte stores, template add
                float* tr = reinterpret_cast<float*>(sEpi) + ew * 1024;
                constexpr int NCH = BN / 32;
                const int rows_left = g.rows_per_batch - (lt * TBM + q * 32);   // rows of this warp that exist
                const long long m0w = (long long)b * g.rows_per_batch + lt * TBM + q * 32;
                const long long c_row0 = (long long)b * g.c_batch_stride + (long long)(lt * TBM + q * 32) * g.ldc;
                // template row of the warp's first output row; later rows advance it incrementally (no per-element
                // 64-bit division in the store loop)
                const long long trow0 = g.tmpl ? m0w / g.rows_per_tmpl : 0;
                const int trem0 = g.tmpl ? (int)(m0w - trow0 * g.rows_per_tmpl) : 0;
#pragma unroll 1
                for (int c = half; c < NCH; c += 2) {
                    if (c * 32 >= n_lim) break;
                    float v[32];
                    tmem_ld_32x32(t_row + c * 32, v);
                    tmem_ld_wait();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; ++j) tr[lane * 32 + ((j + lane) & 31)] = v[j];
                    __syncwarp();
                    const int ncol = n_tile0 + c * 32 + lane;
                    const bool col_ok = ncol < n_end;
                    const float bj = (g.bias && col_ok) ? __ldg(g.bias + ncol) : 0.f;
                    if (LOSS && SCALAR) {
                        // ---- vertex head + reconstruction / velocity loss (ref:src/loss/loss.py:29-55) in one pass ----
                        // lane = column, r = row of this warp's 32-row slice; rows (2k, 2k+1) of a velocity pair are
                        // consecutive r (slices start on even rows and M is even), so the pair never leaves the thread.
                        // y = W d + b + template is used where it sits in registers: gt is read ONCE, dL/dy leaves as
                        // bf16 in the layout the backward GEMMs read, y itself is stored only on request.
                        const float* gp = g.loss_gt + m0w * (long long)g.N + ncol;
                        bf16* dyp = static_cast<bf16*>(g.loss_dy) + m0w * g.ld_dy + ncol;
                        // all 32 ground-truth values (and the template values) of this lane's column are requested before
                        // anything is consumed: 64 independent loads in flight per lane instead of 4 dependent rounds
                        float gv[32], tv[32];
#pragma unroll
                        for (int r = 0; r < 32; ++r) gv[r] = (col_ok && r < rows_left) ? __ldg(gp + (long long)r * g.N) : 0.f;
                        if (g.tmpl != nullptr) {
                            const float* tp = g.tmpl + trow0 * (long long)g.N + ncol;
                            int trem = trem0;
#pragma unroll
                            for (int r = 0; r < 32; ++r) {
                                tv[r] = (col_ok && r < rows_left) ? __ldg(tp) : 0.f;
                                if (++trem == g.rows_per_tmpl) { trem = 0; tp += g.N; }
                            }
                        } else {
#pragma unroll
                            for (int r = 0; r < 32; ++r) tv[r] = 0.f;
                        }
                        float racc = 0.f, vacc = 0.f;
#pragma unroll
                        for (int r = 0; r < 32; r += 2) {
                            if (col_ok && r + 1 < rows_left) {
                                const float y0 = (tr[r * 32 + ((lane + r) & 31)] + bj) + tv[r];
                                const float y1 = (tr[(r + 1) * 32 + ((lane + r + 1) & 31)] + bj) + tv[r + 1];
                                const float d0 = y0 - gv[r], d1 = y1 - gv[r + 1];
                                const float dv = (y1 - y0) - (gv[r + 1] - gv[r]);
                                racc = fmaf(d0, d0, fmaf(d1, d1, racc));
                                vacc = fmaf(dv, dv, vacc);
                                dyp[(long long)r * g.ld_dy] = __float2bfloat16_rn(g.c_rec * d0 - g.c_vel * dv);
                                dyp[(long long)(r + 1) * g.ld_dy] = __float2bfloat16_rn(g.c_rec * d1 + g.c_vel * dv);
                                if (g.C != nullptr) {
                                    reinterpret_cast<float*>(C)[c_row0 + (long long)r * g.ldc + ncol] = y0;
                                    reinterpret_cast<float*>(C)[c_row0 + (long long)(r + 1) * g.ldc + ncol] = y1;
                                }
                            }
                        }
                        loss_rec += (double)racc;      // 32 fp32 terms per chunk, then fp64 (deterministic order)
                        loss_vel += (double)vacc;
                        continue;
                    }
                    const bool fast = (n_tile0 + c * 32 + 32 <= n_end) && rows_left >= 32 && g.resid == nullptr &&
                                      g.act == A2F_ACT_NONE;      // warp-uniform: whole 32x32 chunk live, plain epilogue
                    if (fast) {
                        // common case of the vertex head: no predicates, one 64-bit address per chunk, 32 independent
                        // rows in flight.  Template rows: one shared row (FaceFormer: T frames per utterance) or one per
                        // output row (VOCA / Audio2Mesh).
                        float* cp = reinterpret_cast<float*>(C) + c_row0 + ncol;
                        const int ldc_i = (int)g.ldc;
                        if (g.tmpl == nullptr) {
#pragma unroll
                            for (int r = 0; r < 32; ++r) cp[(long long)(r * ldc_i)] = tr[r * 32 + ((lane + r) & 31)] + bj;
                        } else if (trem0 + 32 <= g.rows_per_tmpl) {
                            const float tb = __ldg(g.tmpl + trow0 * (long long)g.N + ncol) + bj;
#pragma unroll
                            for (int r = 0; r < 32; ++r) cp[(long long)(r * ldc_i)] = tr[r * 32 + ((lane + r) & 31)] + tb;
                        } else if (g.rows_per_tmpl == 1) {
                            const float* tp = g.tmpl + m0w * (long long)g.N + ncol;
                            float add[32];
#pragma unroll
                            for (int r = 0; r < 32; ++r) add[r] = __ldg(tp + (long long)(r * g.N));
#pragma unroll
                            for (int r = 0; r < 32; ++r)
                                cp[(long long)(r * ldc_i)] = (tr[r * 32 + ((lane + r) & 31)] + bj) + add[r];
                        } else {
                            const float* tp = g.tmpl + trow0 * (long long)g.N + ncol;
                            int trem = trem0;
                            float add[32];
#pragma unroll
                            for (int r = 0; r < 32; ++r) {
                                add[r] = __ldg(tp);
                                if (++trem == g.rows_per_tmpl) { trem = 0; tp += g.N; }
                            }
#pragma unroll
                            for (int r = 0; r < 32; ++r)
                                cp[(long long)(r * ldc_i)] = (tr[r * 32 + ((lane + r) & 31)] + bj) + add[r];
                        }
                        continue;
                    }
                    // general path (ragged edges, residual, activation)
                    float add[32];
#pragma unroll
                    for (int r = 0; r < 32; ++r) add[r] = 0.f;
                    if (g.tmpl != nullptr) {
                        const float* tp = g.tmpl + trow0 * (long long)g.N + ncol;
                        int trem = trem0;
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            if (col_ok && r < rows_left) add[r] = __ldg(tp);
                            if (++trem == g.rows_per_tmpl) {   // next output row belongs to the next template
                                trem = 0;
                                tp += g.N;
                            }
                        }
                    }
                    if (g.resid != nullptr) {
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            if (col_ok && r < rows_left) {
                                const long long ri = (long long)b * g.r_batch_stride +
                                                     (long long)(lt * TBM + q * 32 + r) * g.ldr + ncol;
                                add[r] += g.resid_bf16 ? __bfloat162float(static_cast<const bf16*>(g.resid)[ri])
                                                       : static_cast<const float*>(g.resid)[ri];
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 32; ++r)
                        if (col_ok && r < rows_left)
                            st_from_float(C + c_row0 + (long long)r * g.ldc + ncol,
                                          apply_act_rt(tr[r * 32 + ((lane + r) & 31)] + bj, g.act) + add[r]);
                }
            }
            if (!(WIDE && wide)) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
            if (threadIdx.x == 64 && tile == (int)blockIdx.x) TL_STAMP(5);  // first tile's epilogue issued
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (threadIdx.x == 64) TL_STAMP(6);                                 // all tiles' epilogues issued
        if (!SCALAR && leader) tma_store_wait_all();
        if (threadIdx.x == 64) TL_STAMP(7);                                 // stores drained
        if (LOSS && SCALAR && g.loss_partial != nullptr) {
            loss_rec = warp_sum_d(loss_rec);
            loss_vel = warp_sum_d(loss_vel);
            if (lane == 0) {
                g.loss_partial[((long long)blockIdx.x * 8 + ew) * 2] = loss_rec;
                g.loss_partial[((long long)blockIdx.x * 8 + ew) * 2 + 1] = loss_vel;
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

// ====================================================================================================================
// CTA-pair variant (cta_group::2): the two SMs of a TPC run ONE 256 x BN UMMA tile.  Each CTA stages its own 128 A rows
// and only HALF of the B tile, so a k-block costs (128 + BN/2) x 128 B of L2->SM traffic per SM instead of
// (128 + BN) x 128 B -- the 1-CTA mainloop above is bound by exactly that traffic (measured: ~750 cycles per k-block
// against 512 of MMA time, profiles/r1_gemm_timeline.txt).  The leader CTA (even rank) issues every MMA and owns the
// full / accumulator-empty barriers; smem-empty and accumulator-full arrive in both CTAs through multicast commits.
// TMA-store epilogue only (16-byte aligned outputs), mode 0.
template <int BN, typename TC> struct Tc2Cfg {
    static constexpr int ACC_COLS = 256;
    static constexpr int TMEM_COLS = 512;
    static constexpr int B_HALF_BYTES = (BN / 2) * TBK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_HALF_BYTES;        // per CTA
    static constexpr int STAGES = (BN >= 256) ? 5 : 8;
    static constexpr int SBW = (int)(128 / sizeof(TC));
    static constexpr int NBLK = BN / SBW;
    static constexpr int N_EPI_BUF = 4;                                     // two column halves x double buffering
    static constexpr int BIAS_BYTES = 2 * BN * 4;                           // bias of the tile, double-buffered by accumulator
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + N_EPI_BUF * EPI_STAGE_BYTES + BIAS_BYTES + 256;
};

// virtual tile `t` of the pair kernel -> (m block, first column, width); see TcParams::tail_split
struct PairTile { int mb, n_off, n_cols; bool split; };
A2F_D PairTile pair_tile(int t, const TcParams& p, int BN) {
    PairTile r;
    int tile = t, sub = 0;
    r.split = t >= p.full_tiles;
    r.n_cols = BN;
    if (r.split) {
        const int u = t - p.full_tiles;
        tile = p.full_tiles + u / p.tail_split;
        sub = u - (u / p.tail_split) * p.tail_split;
        r.n_cols = BN / p.tail_split;
    }
    r.mb = tile / p.tiles_n;
    r.n_off = (tile - r.mb * p.tiles_n) * BN + sub * r.n_cols;
    return r;
}

template <int BN, typename TC>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ TmapSet maps, const TcParams p) {
    using Cfg = Tc2Cfg<BN, TC>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int PBM = 2 * TBM;              // rows of a pair tile

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)STAGES * A_STAGE_BYTES;
    uint8_t* sEpi = smem + (size_t)STAGES * Cfg::STAGE_BYTES;
    float* sbias = reinterpret_cast<float*>(sEpi + Cfg::N_EPI_BUF * EPI_STAGE_BYTES);      // [2][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + Cfg::N_EPI_BUF * EPI_STAGE_BYTES + Cfg::BIAS_BYTES);
    uint64_t* full_bar = bars;                 // [STAGES]  used in the leader only
    uint64_t* empty_bar = bars + STAGES;       // [STAGES]  both CTAs (multicast commit)
    uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]       both CTAs (multicast commit)
    uint64_t* tempty_bar = tfull_bar + 2;      // [2]       leader only: 8 epilogue warps x 2 CTAs
    uint64_t* rbar = tempty_bar + 2;           // [2]       residual tile landed (one per column half), CTA-local
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GemmParams& g = p.g;
    const int rank = (int)cluster_ctarank();
    const bool is_leader = rank == 0;
    const int pair = (int)cluster_id_x(), n_pairs = (int)cluster_count_x();
    if (threadIdx.x == 0) TL_STAMP(0);

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a[0]);
        tma_prefetch_desc(&maps.b);
        tma_prefetch_desc(&maps.c);
        if (p.resid_tma) tma_prefetch_desc(&maps.r);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 16);
            mbar_init(&rbar[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();                        // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();
    if (threadIdx.x == 0) TL_STAMP(1);

    const int tiles_m = p.tiles_m_per_batch * p.num_batches;     // pair tiles (256 rows)
    const int total_tiles = p.full_tiles + (tiles_m * p.tiles_n - p.full_tiles) * p.tail_split;   // virtual tiles

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            if (p.tail_split > 1) tma_prefetch_desc(&maps.bs);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < total_tiles; tile += n_pairs) {
                const PairTile pt = pair_tile(tile, p, BN);
                const int mb = pt.mb;
                const int b = mb / p.tiles_m_per_batch, lt = mb % p.tiles_m_per_batch;
                const int row0 = lt * PBM + rank * TBM;           // this CTA's 128 A rows
                const int wrow0 = pt.n_off + rank * (pt.n_cols / 2);      // this CTA's half of the B tile
                const CUtensorMap* bmap = pt.split ? &maps.bs : &maps.b;
                const uint32_t tx = 2u * (uint32_t)(A_STAGE_BYTES + (pt.n_cols / 2) * TBK * 2);
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (is_leader) mbar_expect_tx(&full_bar[stage], tx);
                    uint8_t* dstA = sA + (size_t)stage * A_STAGE_BYTES;
                    uint8_t* dstB = sB + (size_t)stage * Cfg::B_HALF_BYTES;
                    const int seg = kb / p.kb_per_seg, kin = (kb - seg * p.kb_per_seg) * TBK;
                    if (p.g.n_seg > 0)
                        tma_load_3d_2sm(dstA, &maps.a[0], &full_bar[stage], p.g.seg_col_off[seg] + kin,
                                        row0 + p.g.seg_row_off[seg], b);
                    else
                        tma_load_3d_2sm(dstA, &maps.a[seg], &full_bar[stage], kin, row0, b);
                    tma_load_2d_2sm(dstB, bmap, &full_bar[stage], kb * TBK, wrow0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (is_leader && lane == 0) {
            // D=f32, A=B=bf16, K-major, N = tile width (BN, or BN / tail_split for a tail slice), M=256 (128 rows per CTA)
            const uint32_t idesc_full = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                        ((uint32_t)(PBM >> 4) << 24);
            const uint32_t idesc_tail = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((BN / p.tail_split) >> 3) << 17) |
                                        ((uint32_t)(PBM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < total_tiles; tile += n_pairs) {
                const uint32_t idesc = tile >= p.full_tiles ? idesc_tail : idesc_full;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS);
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    if (tile == pair && kb == 0) TL_STAMP(2);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_u32(sA + (size_t)stage * A_STAGE_BYTES), p);
                    const uint64_t bdesc = make_smem_desc(smem_u32(sB + (size_t)stage * Cfg::B_HALF_BYTES), p);
#pragma unroll
                    for (int k = 0; k < TBK / UMMA_K; ++k)
                        umma_f16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                     (kb | k) != 0 ? 1u : 0u);
                    umma_commit_2sm(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&tfull_bar[acc]);
                if (tile == pair) TL_STAMP(3);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: 8 warps per CTA, each CTA drains its own 128 accumulator rows ===============
        // Latency plan (per-CTA timeline, profiles/r1_gemm_timeline.txt: 2.9 us per tile, 6.2 us with a residual, against
        // 4.3 us of mainloop at K=768): the tile's bias sits in shared memory before the accumulator is complete; the
        // residual / saved pre-activation tile is fetched by TMA straight into the (128B-swizzled) store-staging buffer,
        // so each thread reads its own row with conflict-free 16-byte loads instead of 32 uncoalesced global rows per
        // warp instruction; staging is double-buffered so that a block never waits for the previous block's store.
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        const bool leader = ((ew & 3) == 0) && lane == 0;
        const int bar_id = 1 + half;
        uint8_t* stage_base = sEpi + half * 2 * EPI_STAGE_BYTES;       // two staging buffers per column half
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t sb = 0;                                               // blocks this half has processed
        const float* __restrict__ e_bias = g.bias;
        const void* __restrict__ e_resid = g.resid;
        const int e_act = g.act;
        const bool e_dact = g.resid_mode == A2F_RESID_DACT;
        const bool rtma = p.resid_tma != 0;
        const bool dual = g.C2 != nullptr;                             // C = z (pre-activation), C2 = act(z)
        const bool fast_dgelu = e_dact && p.fast_gelu && e_act == A2F_ACT_GELU;    // bf16 output: MUFU-based GELU'
        constexpr int SBW = Cfg::SBW;
        constexpr int EPC = 16 / (int)sizeof(TC);
        for (int tile = pair; tile < total_tiles; tile += n_pairs) {
            const PairTile pt = pair_tile(tile, p, BN);
            const int mb = pt.mb;
            const int b = mb / p.tiles_m_per_batch, lt = mb % p.tiles_m_per_batch;
            const int n_tile0 = pt.n_off;
            const int n_lim = min(pt.n_cols, g.N - n_tile0);
            const int n_end = n_tile0 + n_lim;
            const int row_base = lt * PBM + rank * TBM;

            // while the mainloop runs: bias of this tile -> smem, first residual block -> staging buffer
            float* tb = sbias + acc * BN;
            if (e_bias != nullptr) {
                const int c = ew * 32 + lane;                          // 256 epilogue threads, BN == 256 columns
                tb[c] = (n_tile0 + c < g.N) ? __ldg(e_bias + n_tile0 + c) : 0.f;
            }
            if (rtma && leader && half * SBW < n_lim) {
                tma_store_wait_read1();                                // the store that used this buffer two blocks ago
                mbar_expect_tx(&rbar[half], EPI_STAGE_BYTES);
                tma_load_3d(stage_base + (sb & 1) * EPI_STAGE_BYTES, &maps.r, &rbar[half], n_tile0 + half * SBW, row_base, b);
            }
            named_bar_sync(3, 256);                                    // bias visible to all epilogue warps

            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if (threadIdx.x == 64 && tile == pair) TL_STAMP(4);

            const int r_tile = q * 32 + lane;
            const int r_in_batch = row_base + r_tile;
            const bool row_ok = r_in_batch < g.rows_per_batch;
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::ACC_COLS);
#pragma unroll 1
            for (int blk = half; blk < Cfg::NBLK; blk += 2) {
                const int col0 = blk * SBW;
                if (col0 >= n_lim) break;
                const int ncol0 = n_tile0 + col0;
                uint8_t* stage_buf = stage_base + (sb & 1) * EPI_STAGE_BYTES;
                uint8_t* rowp = stage_buf + r_tile * 128;
                float v[SBW];
#pragma unroll
                for (int cc = 0; cc < SBW / 32; ++cc) tmem_ld_32x32(t_row + col0 + cc * 32, v + cc * 32);
                tmem_ld_wait();
                if (!e_dact) {
                    if (e_bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < SBW; j += 4) {
                            const float4 f = *reinterpret_cast<const float4*>(tb + col0 + j);
                            v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
                        }
                    }
                    if (!dual) epi_bias_act<SBW>(v, nullptr, 0, ncol0, n_end, e_act, p.fast_gelu);
                }
                if (rtma) {
                    mbar_wait(&rbar[half], sb & 1);                    // residual block landed (and the buffer is ours)
                    if (sizeof(TC) == 2) {
#pragma unroll
                        for (int ch = 0; ch < SBW / 8; ++ch) {
                            const uint4 u = *reinterpret_cast<const uint4*>(rowp + ((ch ^ (r_tile & 7)) * 16));
                            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __bfloat1622float2(h2[e]);
                                if (e_dact) {
                                    if (fast_dgelu) {
                                        v[ch * 8 + 2 * e] *= gelu_grad_fast(f.x);
                                        v[ch * 8 + 2 * e + 1] *= gelu_grad_fast(f.y);
                                    } else {
                                        v[ch * 8 + 2 * e] *= act_grad(f.x, e_act);
                                        v[ch * 8 + 2 * e + 1] *= act_grad(f.y, e_act);
                                    }
                                } else {
                                    v[ch * 8 + 2 * e] += f.x;
                                    v[ch * 8 + 2 * e + 1] += f.y;
                                }
                            }
                        }
                    }
                } else {
                    const long long r_off = (long long)b * g.r_batch_stride + (long long)r_in_batch * g.ldr + ncol0;
                    if (e_dact) {
                        if (row_ok) {
                            float z[SBW];
#pragma unroll
                            for (int j = 0; j < SBW; ++j) z[j] = 0.f;
                            epi_resid<SBW>(z, e_resid, g.resid_bf16, r_off, ncol0, n_end, p.resid_vec_ok);
#pragma unroll
                            for (int j = 0; j < SBW; ++j) v[j] *= act_grad(z[j], e_act);
                        }
                    } else if (e_resid != nullptr && row_ok) {
                        epi_resid<SBW>(v, e_resid, g.resid_bf16, r_off, ncol0, n_end, p.resid_vec_ok);
                    }
                    // the store that used this staging buffer two blocks ago must have finished reading it
                    if (leader) tma_store_wait_read1();
                    named_bar_sync(bar_id, 128);
                }
#pragma unroll
                for (int ch = 0; ch < SBW / EPC; ++ch) {
                    const int pch = ch ^ (r_tile & 7);
                    uint4 u;
                    if (sizeof(TC) == 2) {
                        u.x = pack_bf16x2(v[ch * 8 + 0], v[ch * 8 + 1]);
                        u.y = pack_bf16x2(v[ch * 8 + 2], v[ch * 8 + 3]);
                        u.z = pack_bf16x2(v[ch * 8 + 4], v[ch * 8 + 5]);
                        u.w = pack_bf16x2(v[ch * 8 + 6], v[ch * 8 + 7]);
                    } else {
                        u.x = __float_as_uint(v[ch * EPC + 0]);
                        u.y = __float_as_uint(v[ch * EPC + 1]);
                        u.z = __float_as_uint(v[ch * EPC + 2]);
                        u.w = __float_as_uint(v[ch * EPC + 3]);
                    }
                    *reinterpret_cast<uint4*>(rowp + pch * 16) = u;
                }
                fence_proxy_async_smem();
                named_bar_sync(bar_id, 128);
                ++sb;
                if (leader) {
                    tma_store_3d(&maps.c, stage_buf, ncol0, row_base, b);
                    tma_store_commit();
                }
                if (dual) {
                    // second output: act(z) through the OTHER staging buffer (the block behaves like two blocks in a row)
                    epi_bias_act<SBW>(v, nullptr, 0, ncol0, n_end, e_act, p.fast_gelu);
                    uint8_t* rowp2 = stage_base + (sb & 1) * EPI_STAGE_BYTES + r_tile * 128;
                    if (leader) tma_store_wait_read1();
                    named_bar_sync(bar_id, 128);
#pragma unroll
                    for (int ch = 0; ch < SBW / EPC; ++ch) {
                        const int pch = ch ^ (r_tile & 7);
                        uint4 u;
                        if (sizeof(TC) == 2) {
                            u.x = pack_bf16x2(v[ch * 8 + 0], v[ch * 8 + 1]);
                            u.y = pack_bf16x2(v[ch * 8 + 2], v[ch * 8 + 3]);
                            u.z = pack_bf16x2(v[ch * 8 + 4], v[ch * 8 + 5]);
                            u.w = pack_bf16x2(v[ch * 8 + 6], v[ch * 8 + 7]);
                        } else {
                            u.x = __float_as_uint(v[ch * EPC + 0]);
                            u.y = __float_as_uint(v[ch * EPC + 1]);
                            u.z = __float_as_uint(v[ch * EPC + 2]);
                            u.w = __float_as_uint(v[ch * EPC + 3]);
                        }
                        *reinterpret_cast<uint4*>(rowp2 + pch * 16) = u;
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(bar_id, 128);
                    if (leader) {
                        tma_store_3d(&maps.c2, stage_base + (sb & 1) * EPI_STAGE_BYTES, ncol0, row_base, b);
                        tma_store_commit();
                    }
                    ++sb;
                }
                if (leader) {
                    if (rtma && col0 + 2 * SBW < n_lim) {              // next block of this tile: fetch its residual now
                        tma_store_wait_read1();
                        mbar_expect_tx(&rbar[half], EPI_STAGE_BYTES);
                        tma_load_3d(stage_base + (sb & 1) * EPI_STAGE_BYTES, &maps.r, &rbar[half], ncol0 + 2 * SBW, row_base, b);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
            if (threadIdx.x == 64 && tile == pair) TL_STAMP(5);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (threadIdx.x == 64) TL_STAMP(6);
        if (leader) tma_store_wait_all();
        if (threadIdx.x == 64) TL_STAMP(7);
    }

    tc_fence_before();
    cluster_sync_all();                        // both CTAs are done with TMEM and with each other's barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
    }
}

static int g_tail_split = 1;   // debug (a2f_debug_set_umma_field 11): 0 = never split the tail tiles of the pair kernel

template <int BN, typename TC>
static int launch_tc2(TmapSet& maps, const TcParams& p_in, cudaStream_t s) {
    using Cfg = Tc2Cfg<BN, TC>;
    const GemmParams& g = p_in.g;
    uint64_t dims[3] = {(uint64_t)g.N, (uint64_t)g.rows_per_batch, (uint64_t)p_in.num_batches};
    uint64_t strides[2] = {(uint64_t)g.ldc * sizeof(TC), (uint64_t)g.c_batch_stride * sizeof(TC)};
    uint32_t box[3] = {(uint32_t)Cfg::SBW, TBM, 1};
    int rc = encode_tmap(&maps.c, g.C, (int)sizeof(TC), 3, dims, strides, box, 1);
    if (rc != A2F_OK) return rc;
    TcParams p = p_in;
    p.resid_tma = 0;
    if (g.C2 != nullptr) {
        A2F_REQUIRE(reinterpret_cast<uintptr_t>(g.C2) % 16 == 0 && (g.ldc2 * (long long)sizeof(TC)) % 16 == 0 && g.ldc > 0 &&
                    (g.c_batch_stride * g.ldc2) % g.ldc == 0, "gemm_tc: C2 must be 16-byte aligned with a layout proportional to C");
        const long long c2_batch = g.c_batch_stride / g.ldc * g.ldc2;
        A2F_REQUIRE((c2_batch * (long long)sizeof(TC)) % 16 == 0, "gemm_tc: C2 batch stride must be a multiple of 16 bytes");
        uint64_t strides2[2] = {(uint64_t)g.ldc2 * sizeof(TC), (uint64_t)c2_batch * sizeof(TC)};
        rc = encode_tmap(&maps.c2, g.C2, (int)sizeof(TC), 3, dims, strides2, box, 1);
        if (rc != A2F_OK) return rc;
    }
    if (g.resid != nullptr && g.resid_bf16 && sizeof(TC) == 2 && reinterpret_cast<uintptr_t>(g.resid) % 16 == 0 &&
        (g.ldr * 2) % 16 == 0 && (g.r_batch_stride * 2) % 16 == 0) {
        uint64_t rdims[3] = {(uint64_t)g.N, (uint64_t)g.rows_per_batch, (uint64_t)p.num_batches};
        uint64_t rstr[2] = {(uint64_t)g.ldr * 2, (uint64_t)g.r_batch_stride * 2};
        uint32_t rbox[3] = {64, TBM, 1};
        rc = encode_tmap(&maps.r, g.resid, 2, 3, rdims, rstr, rbox, 1);
        if (rc != A2F_OK) return rc;
        p.resid_tma = 1;
    }
    auto kern = gemm_tc2_kernel<BN, TC>;
    static bool attr_done = false;
    if (!attr_done) {
        A2F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
        attr_done = true;
    }
    const int total = p.tiles_m_per_batch * p.num_batches * p.tiles_n;
    const int max_pairs = sm_count() / 2;
    // wave quantisation: when the last wave is only partly filled, its tiles are cut into 2 or 4 column slices
    p.full_tiles = total;
    p.tail_split = 1;
    if (g_tail_split && total > max_pairs && g.N % BN == 0) {
        const int left = total % max_pairs;
        if (left != 0) {
            int best = 1;
            double best_cost = 1.0;
            for (int sp = 2; sp <= 4; sp *= 2) {
                const double cost = (double)((left * sp + max_pairs - 1) / max_pairs) / sp;
                if (cost < best_cost - 1e-9) { best_cost = cost; best = sp; }
            }
            if (best > 1) {
                p.full_tiles = total - left;
                p.tail_split = best;
                uint64_t bdims[2] = {(uint64_t)g.K, (uint64_t)g.N};
                uint64_t bstr[1] = {(uint64_t)g.ldw * 2};
                uint32_t bbox[2] = {TBK, (uint32_t)(BN / 2 / best)};
                rc = encode_tmap_bf16(&maps.bs, g.W, 2, bdims, bstr, bbox, 1);
                if (rc != A2F_OK) return rc;
            }
        }
    }
    const int vtotal = p.full_tiles + (total - p.full_tiles) * p.tail_split;
    const int pairs = vtotal < max_pairs ? vtotal : max_pairs;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * pairs, 1, 1);
    cfg.blockDim = dim3(TC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    A2F_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, p));
    count_launch();
    return A2F_OK;
}

template <int BN, typename TC, bool SCALAR, bool LOSS = false>
static int launch_tc(const TmapSet& maps, const TcParams& p, cudaStream_t s) {
    using Cfg = TcCfg<BN, TC, tc_wide_v<BN, TC, SCALAR>>;
    auto kern = gemm_tc_kernel<BN, TC, SCALAR, LOSS>;
    static bool attr_done = false;   // per instantiation
    if (!attr_done) {
        A2F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
        attr_done = true;
    }
    const int total = p.tiles_m_per_batch * p.num_batches * p.tiles_n;
    int grid = total < sm_count() ? total : sm_count();
    if (p.g.max_ctas > 0 && grid > p.g.max_ctas) grid = p.g.max_ctas;
    A2F_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, s, maps, p));
    count_launch();
    return A2F_OK;
}

template <int BN, typename TC>
static int encode_c_and_launch(TmapSet& maps, const TcParams& p, bool scalar, cudaStream_t s) {
    using Cfg = TcCfg<BN, TC>;
    if (scalar && p.g.loss_gt != nullptr) {
        if (BN == 256 && sizeof(TC) == 4) return launch_tc<256, float, true, true>(maps, p, s);
        return set_error(A2F_EINVAL, "gemm_tc: the fused-loss epilogue is the 256-wide fp32 vertex head only");
    }
    if (scalar) return launch_tc<BN, TC, true>(maps, p, s);
    const GemmParams& g = p.g;
    const int ncols_total = (p.mode == 2) ? 768 : g.N;
    uint64_t dims[3] = {(uint64_t)ncols_total, (uint64_t)g.rows_per_batch, (uint64_t)p.num_batches};
    uint64_t strides[2] = {(uint64_t)g.ldc * sizeof(TC), (uint64_t)g.c_batch_stride * sizeof(TC)};
    uint32_t box[3] = {(uint32_t)Cfg::SBW, TBM, 1};
    int rc = encode_tmap(&maps.c, g.C, (int)sizeof(TC), 3, dims, strides, box, Cfg::SWZ ? 1 : 0);
    if (rc != A2F_OK) return rc;
    return launch_tc<BN, TC, false>(maps, p, s);
}

template <int BN> static int dispatch_out(TmapSet& maps, const TcParams& p, int c_bf16, bool scalar, cudaStream_t s) {
    if (c_bf16) return encode_c_and_launch<BN, bf16>(maps, p, scalar, s);
    return encode_c_and_launch<BN, float>(maps, p, scalar, s);
}

static int g_force_bn = 0;   // debug: force a tile width (tests exercise every instantiation)
static int g_pair_mode = 1;  // debug: 0 = never use the CTA-pair kernel
static unsigned long long* g_timeline = nullptr;

int gemm_tc(const GemmParams& g_in, int c_bf16, int mode, cudaStream_t s) {
    if (g_in.M <= 0 || g_in.N <= 0) return A2F_OK;
    GemmParams g = g_in;
    normalize_gemm(g);
    TcParams p;
    p.g = g;
    p.mode = mode;
    p.lbo_enc = g_umma_fields[0];
    p.sbo_enc = g_umma_fields[1];
    p.desc_version = g_umma_fields[2];
    p.desc_layout = g_umma_fields[3];
    p.fast_gelu = c_bf16 ? 1 : 0;
    p.resid_tma = 0;
    p.full_tiles = 0;
    p.tail_split = 1;
    p.timeline = g_timeline;
    TmapSet maps;
    memset(&maps, 0, sizeof(maps));

    A2F_REQUIRE(g.rows_per_batch > 0 && g.M % g.rows_per_batch == 0, "gemm_tc: M must be a multiple of rows_per_batch");
    p.num_batches = g.M / g.rows_per_batch;
    p.tiles_m_per_batch = (g.rows_per_batch + TBM - 1) / TBM;

    int BN;
    bool use_pair = false;
    if (mode == 2) {
        // positional conv: g.N = 48 per group, 16 groups, K = 128 taps x 64 (48 live + 16 zero-weight) channels
        BN = 48;
        p.tiles_n = 16;
        p.num_k_blocks = 128;
        p.kb_per_seg = 128;
        uint64_t dims[3] = {768, (uint64_t)g.rows_per_batch, (uint64_t)p.num_batches};
        uint64_t strides[2] = {768 * 2, (uint64_t)g.rows_per_batch * 768 * 2};
        uint32_t box[3] = {TBK, TBM, 1};
        int rc = encode_tmap_bf16(&maps.a[0], g.A, 3, dims, strides, box, 1);
        if (rc != A2F_OK) return rc;
        uint64_t bdims[2] = {(uint64_t)128 * 64, (uint64_t)16 * 48};
        uint64_t bstr[1] = {(uint64_t)128 * 64 * 2};
        uint32_t bbox[2] = {TBK, 48};
        rc = encode_tmap_bf16(&maps.b, g.W, 2, bdims, bstr, bbox, 1);
        if (rc != A2F_OK) return rc;
    } else {
        A2F_REQUIRE(g.K % 8 == 0, "gemm_tc: K must be a multiple of 8 (16-byte TMA rows)");
        A2F_REQUIRE(g.a_row_stride % 8 == 0 && g.a_batch_stride % 8 == 0 && g.ldw % 8 == 0,
                    "gemm_tc: operand strides must be multiples of 8 elements (16 bytes)");
        if (g_force_bn) BN = g_force_bn;
        else if (g.N > 128) BN = 256;
        else if (g.N > 64) BN = 128;
        else BN = 64;
        {
            // CTA-pair kernel: 256-row tiles, TMA-store epilogue only
            const size_t csz0 = c_bf16 ? 2 : 4;
            const bool c_ok = (reinterpret_cast<uintptr_t>(g.C) % 16 == 0) && ((g.ldc * (long long)csz0) % 16 == 0) &&
                              ((g.c_batch_stride * (long long)csz0) % 16 == 0);
            use_pair = g_pair_mode && BN == 256 && g.tmpl == nullptr && g.loss_gt == nullptr && c_ok && g.rows_per_batch > TBM &&
                       sm_count() >= 2;
            if (use_pair) p.tiles_m_per_batch = (g.rows_per_batch + 2 * TBM - 1) / (2 * TBM);
        }
        p.tiles_n = (g.N + BN - 1) / BN;
        p.num_k_blocks = (g.K + TBK - 1) / TBK;
        if (g.n_seg > 0) {
            // explicit segments: one map over [a_rows x row_len], row_len = furthest column any segment touches
            const int kseg = g.K / g.n_seg;
            A2F_REQUIRE(kseg % TBK == 0, "gemm_tc: explicit K segments must be multiples of 64");
            int row_len = 0;
            for (int i = 0; i < g.n_seg; ++i) {
                A2F_REQUIRE(g.seg_col_off[i] >= 0 && g.seg_col_off[i] % 8 == 0, "gemm_tc: seg_col_off must be a multiple of 8");
                if (g.seg_col_off[i] + kseg > row_len) row_len = g.seg_col_off[i] + kseg;
            }
            A2F_REQUIRE(row_len <= g.a_row_stride, "gemm_tc: explicit segments must stay inside one A row");
            p.kb_per_seg = kseg / TBK;
            uint64_t dims[3] = {(uint64_t)row_len, (uint64_t)g.a_rows, (uint64_t)p.num_batches};
            uint64_t strides[2] = {(uint64_t)g.a_row_stride * 2,
                                   (uint64_t)(p.num_batches > 1 ? g.a_batch_stride : g.a_row_stride * g.a_rows) * 2};
            uint32_t box[3] = {TBK, TBM, 1};
            int rc = encode_tmap_bf16(&maps.a[0], g.A, 3, dims, strides, box, 1);
            if (rc != A2F_OK) return rc;
        }
        // K segments: every tensor map must have non-overlapping rows (segment length <= row stride)
        int nseg = 1;
        if (g.n_seg > 0) nseg = 0;
        else if (g.a_row_stride < g.K) nseg = (int)((g.K + g.a_row_stride - 1) / g.a_row_stride);
        A2F_REQUIRE(nseg <= MAX_SEGS, "gemm_tc: A rows overlap too much (more than 4 K segments)");
        A2F_REQUIRE(nseg == 0 || (g.K % nseg == 0 && (nseg == 1 || (g.K / nseg) % TBK == 0)),
                    "gemm_tc: K segment length must be a multiple of 64");
        const int kseg = nseg > 0 ? g.K / nseg : 0;
        if (nseg > 0) p.kb_per_seg = (nseg == 1) ? p.num_k_blocks : kseg / TBK;
        for (int sgi = 0; sgi < nseg; ++sgi) {
            uint64_t dims[3] = {(uint64_t)kseg, (uint64_t)g.a_rows, (uint64_t)p.num_batches};
            uint64_t strides[2] = {(uint64_t)g.a_row_stride * 2,
                                   (uint64_t)(p.num_batches > 1 ? g.a_batch_stride : g.a_row_stride * g.rows_per_batch) * 2};
            uint32_t box[3] = {TBK, TBM, 1};
            const bf16* base = static_cast<const bf16*>(g.A) + (size_t)sgi * kseg;
            int rc = encode_tmap_bf16(&maps.a[sgi], base, 3, dims, strides, box, 1);
            if (rc != A2F_OK) return rc;
        }
        uint64_t bdims[2] = {(uint64_t)g.K, (uint64_t)g.N};
        uint64_t bstr[1] = {(uint64_t)g.ldw * 2};
        uint32_t bbox[2] = {TBK, (uint32_t)(use_pair ? BN / 2 : BN)};
        int rc = encode_tmap_bf16(&maps.b, g.W, 2, bdims, bstr, bbox, 1);
        if (rc != A2F_OK) return rc;
    }

    const size_t csz = c_bf16 ? 2 : 4;
    const bool c_tma_ok = (reinterpret_cast<uintptr_t>(g.C) % 16 == 0) && ((g.ldc * (long long)csz) % 16 == 0) &&
                          ((g.c_batch_stride * (long long)csz) % 16 == 0);
    p.resid_vec_ok = 0;
    p.bias_vec_ok = (g.bias != nullptr) && (reinterpret_cast<uintptr_t>(g.bias) % 16 == 0);
    if (g.resid) {
        const size_t rsz = g.resid_bf16 ? 2 : 4;
        p.resid_vec_ok = (reinterpret_cast<uintptr_t>(g.resid) % 16 == 0) && ((g.ldr * (long long)rsz) % 16 == 0) &&
                         ((g.r_batch_stride * (long long)rsz) % 16 == 0);
    }
    // scalar (transposing) epilogue: outputs whose rows are not 16-byte aligned (the 15069-wide vertex head) and the
    // template-add epilogue.  fp32 output only.
    const bool scalar = (g.tmpl != nullptr) || !c_tma_ok || g.loss_gt != nullptr;
    if (scalar) {
        A2F_REQUIRE(g.resid_mode == A2F_RESID_ADD, "gemm_tc: the activation-backward epilogue needs 16-byte aligned outputs");
        A2F_REQUIRE(!c_bf16, "gemm_tc: template-add / unaligned outputs are fp32 only");
        A2F_REQUIRE(g.ldc < (1LL << 25) && g.N < (1 << 25), "gemm_tc: scalar epilogue needs ldc, N < 2^25");
        A2F_REQUIRE(mode != 2, "gemm_tc: posconv output must be 16-byte aligned");
    }

    if (use_pair) {
        if (c_bf16) return launch_tc2<256, bf16>(maps, p, s);
        return launch_tc2<256, float>(maps, p, s);
    }
    A2F_REQUIRE(g.C2 == nullptr, "gemm_tc: the two-output epilogue (C2) needs the CTA-pair kernel (N > 128, more than 128 rows "
                                 "per batch, 16-byte aligned outputs)");
    switch (BN) {
        case 256: return dispatch_out<256>(maps, p, c_bf16, scalar, s);
        case 128: return dispatch_out<128>(maps, p, c_bf16, scalar, s);
        case 64: return dispatch_out<64>(maps, p, c_bf16, scalar, s);
        case 48: return dispatch_out<48>(maps, p, c_bf16, scalar, s);
        default: return set_error(A2F_EINVAL, "gemm_tc: unsupported tile width");
    }
}

}  // namespace a2f

namespace a2f {
void set_mha_impl(int v);
void set_mha_tc_min_t(int v);
void set_mha_short_nqb(int v);
void set_dec_cluster(int v);
void set_posconv_impl(int v);
void set_posconv_swap(int v);
}
namespace a2f {
unsigned long long* debug_timeline() { return g_timeline; }
}
extern "C" int a2f_debug_set_timeline(void* dev_ptr) {
    a2f::g_timeline = static_cast<unsigned long long*>(dev_ptr);
    return A2F_OK;
}

extern "C" int a2f_debug_set_umma_field(int field, unsigned value) {
    if (field >= 0 && field < 4) {
        a2f::g_umma_fields[field] = value;
        return A2F_OK;
    }
    if (field == 5) {   // 0 = never use the CTA-pair (cta_group::2) kernel, 1 = automatic
        a2f::g_pair_mode = value ? 1 : 0;
        return A2F_OK;
    }
    if (field == 6) {   // encoder attention kernel: 0 = automatic, 1 = mma.sync, 2 = tcgen05
        if (value > 2) return a2f::set_error(A2F_EINVAL, "bad attention impl");
        a2f::set_mha_impl((int)value);
        return A2F_OK;
    }
    if (field == 8) {   // decoder rollout on long clips: 0 = automatic cluster size, 1 = single CTA, 2/4/8 = forced
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8) return a2f::set_error(A2F_EINVAL, "bad cluster size");
        a2f::set_dec_cluster((int)value);
        return A2F_OK;
    }
    if (field == 7) {   // automatic mode: shortest sequence that takes the tcgen05 attention kernel
        a2f::set_mha_tc_min_t((int)value);
        return A2F_OK;
    }
    if (field == 9) {   // positional conv of the tcgen05 backend: 0 = posconv_tc.cu (kpad 8 weights), 1 = legacy mode 2 (kpad 64)
        a2f::set_posconv_impl(value ? 1 : 0);
        return A2F_OK;
    }
    if (field == 10) {  // posconv_tc.cu: exchange the LBO / SBO descriptor fields (bring-up experiment)
        a2f::set_posconv_swap(value ? 1 : 0);
        return A2F_OK;
    }
    if (field == 12) {  // short-clip attention: 0 = automatic, 1 = one 80-query block per CTA (round-1 schedule)
        a2f::set_mha_short_nqb((int)value);
        return A2F_OK;
    }
    if (field == 11) {  // pair kernel: 0 = never cut the tiles of a partly filled last wave into column slices
        a2f::g_tail_split = value ? 1 : 0;
        return A2F_OK;
    }
    if (field == 4) {   // force tile width (0 = automatic)
        if (value != 0 && value != 64 && value != 128 && value != 256) return a2f::set_error(A2F_EINVAL, "bad BN");
        a2f::g_force_bn = (int)value;
        return A2F_OK;
    }
    return a2f::set_error(A2F_EINVAL, "unknown debug field");
}

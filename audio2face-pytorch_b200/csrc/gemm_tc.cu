// tcgen05 / TMEM / TMA GEMM for sm_100a: C = act(A W^T + bias) + resid + tmpl, bf16 operands, fp32 accumulate.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor into a STAGES-deep 128B-swizzled smem ring)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma, accumulators double-buffered in TMEM)
//   warps 2..5  epilogue       (tcgen05.ld -> bias/act/residual/template -> global), overlaps the next tile's mainloop
// A is addressed through <=3-D tensor maps so that the same mainloop serves
//   mode 0: plain [M,K] matrices and the stride-2 Conv1d stack as an implicit GEMM over channels-last activations
//           (rows of the im2col matrix are strided/overlapping views of the activation, split into K segments so that
//            every tensor map has non-overlapping rows),
//   mode 2: the k=128 grouped positional conv (K block = one tap, A box shifted by one time step per block).
#include "a2f_common.cuh"
#include "gemm_params.cuh"

namespace a2f {

constexpr int TBM = 128;          // tile rows (UMMA M)
constexpr int TBK = 64;           // K per stage: 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int TC_THREADS = 192;
constexpr int A_STAGE_BYTES = TBM * TBK * 2;
constexpr int MAX_SEGS = 4;

struct TmapSet {
    CUtensorMap a[MAX_SEGS];
    CUtensorMap b;
};

struct TcParams {
    GemmParams g;
    int mode;
    int tiles_m_per_batch, num_batches, tiles_n, num_k_blocks, kb_per_seg;
    int vec_ok;                     // 16-byte vector epilogue stores/loads are legal
    uint32_t lbo_enc, sbo_enc, desc_version, desc_layout;
};

// debug knobs (a2f_debug_set_umma_field): defaults are the CUTLASS-documented K-major SWIZZLE_128B values
static uint32_t g_umma_fields[4] = {1u, 64u, 1u, 2u};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, const TcParams& p) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(p.lbo_enc & 0x3FFFu) << 16) |
           ((uint64_t)(p.sbo_enc & 0x3FFFu) << 32) | ((uint64_t)(p.desc_version & 3u) << 46) |
           ((uint64_t)(p.desc_layout & 7u) << 61);
}

template <int BN> struct TcCfg {
    static constexpr int CH = (BN % 32 == 0) ? 32 : 16;             // epilogue column chunk
    static constexpr int ACC_COLS = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;
    static constexpr int TMEM_COLS = 2 * ACC_COLS;
    static constexpr int B_STAGE_BYTES = BN * TBK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128) ? 6 : 8;
    static constexpr int EPI_STAGE_FLOATS = 4 * 32 * 33;            // per-warp transpose buffers (staged epilogue)
    static constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES +
                                         EPI_STAGE_FLOATS * 4 + 2 * BN * 4 /*bias*/ + 256 /*barriers*/;
};

template <int BN, typename TC, bool STAGED>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ TmapSet maps, const TcParams p) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int CH = Cfg::CH;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)STAGES * A_STAGE_BYTES;
    float* sEpi = reinterpret_cast<float*>(smem + (size_t)STAGES * Cfg::STAGE_BYTES);
    float* sBias = sEpi + Cfg::EPI_STAGE_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 2 * BN);
    uint64_t* full_bar = bars;                 // [STAGES]
    uint64_t* empty_bar = bars + STAGES;       // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
    uint64_t* tempty_bar = tfull_bar + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GemmParams& g = p.g;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a[0]);
        tma_prefetch_desc(&maps.b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_m = p.tiles_m_per_batch * p.num_batches;
    const int total_tiles = tiles_m * p.tiles_n;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nb = tile % p.tiles_n, mb = tile / p.tiles_n;
                const int b = mb / p.tiles_m_per_batch, lt = mb % p.tiles_m_per_batch;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    uint8_t* dstA = sA + (size_t)stage * A_STAGE_BYTES;
                    uint8_t* dstB = sB + (size_t)stage * Cfg::B_STAGE_BYTES;
                    if (p.mode == 2) {
                        tma_load_3d(dstA, &maps.a[0], &full_bar[stage], nb * 48, lt * TBM + kb - 64, b);
                        tma_load_2d(dstB, &maps.b, &full_bar[stage], kb * TBK, nb * 48);
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
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (4 warps, TMEM lane quarter = warp % 4) =====================
        const int q = warp & 3;
        const int epi_tid = threadIdx.x - 64;     // 0..127
        float* stg = sEpi + (warp - 2) * (32 * 33);
        int acc = 0;
        uint32_t acc_phase = 0;
        TC* __restrict__ C = static_cast<TC*>(g.C);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nb = tile % p.tiles_n, mb = tile / p.tiles_n;
            const int b = mb / p.tiles_m_per_batch, lt = mb % p.tiles_m_per_batch;
            const int n_tile0 = (p.mode == 2) ? nb * 48 : nb * BN;          // first global column of this tile
            const int n_lim = (p.mode == 2) ? 48 : min(BN, g.N - n_tile0);   // live columns in this tile
            // stage the bias slice for this tile (double-buffered with the accumulator index)
            float* bias_s = sBias + acc * BN;
            for (int i = epi_tid; i < BN; i += 128)
                bias_s[i] = (g.bias != nullptr && i < n_lim) ? g.bias[n_tile0 + i] : 0.f;
            asm volatile("bar.sync 1, 128;" ::: "memory");

            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();

            const int r_in_batch = lt * TBM + q * 32 + lane;                 // this thread's row (direct mode)
            const bool row_ok = r_in_batch < g.rows_per_batch;
            const long long m = (long long)b * g.rows_per_batch + r_in_batch;
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::ACC_COLS);

#pragma unroll 1
            for (int c = 0; c < BN / CH; ++c) {
                if (c * CH >= n_lim) break;     // warp-uniform
                float v[CH];
                if (CH == 32) tmem_ld_32x32(t_row + c * CH, v);
                else tmem_ld_32x16(t_row + c * CH, v);
                tmem_ld_wait();
                const int ncol0 = n_tile0 + c * CH;   // global column of v[0]
                if (!STAGED) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) v[j] = apply_act_rt(v[j] + bias_s[c * CH + j], g.act);
                    if (row_ok) {
                        const bool full = (c * CH + CH <= n_lim) && p.vec_ok;
                        if (g.resid) {
                            if (g.resid_bf16) {
                                const bf16* rp = static_cast<const bf16*>(g.resid) + m * g.ldr + ncol0;
                                if (full) {
#pragma unroll
                                    for (int j = 0; j < CH; j += 8) {
                                        uint4 u = *reinterpret_cast<const uint4*>(rp + j);
                                        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            float2 f = __bfloat1622float2(h2[e]);
                                            v[j + 2 * e] += f.x;
                                            v[j + 2 * e + 1] += f.y;
                                        }
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < CH; ++j)
                                        if (c * CH + j < n_lim) v[j] += __bfloat162float(rp[j]);
                                }
                            } else {
                                const float* rp = static_cast<const float*>(g.resid) + m * g.ldr + ncol0;
                                if (full) {
#pragma unroll
                                    for (int j = 0; j < CH; j += 4) {
                                        float4 f = *reinterpret_cast<const float4*>(rp + j);
                                        v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < CH; ++j)
                                        if (c * CH + j < n_lim) v[j] += rp[j];
                                }
                            }
                        }
                        TC* cp = C + m * g.ldc + ncol0;
                        if (full) {
                            if (sizeof(TC) == 2) {
#pragma unroll
                                for (int j = 0; j < CH; j += 8) {
                                    uint4 u;
                                    u.x = pack_bf16x2(v[j], v[j + 1]);
                                    u.y = pack_bf16x2(v[j + 2], v[j + 3]);
                                    u.z = pack_bf16x2(v[j + 4], v[j + 5]);
                                    u.w = pack_bf16x2(v[j + 6], v[j + 7]);
                                    *reinterpret_cast<uint4*>(cp + j) = u;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < CH; j += 4)
                                    *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < CH; ++j)
                                if (c * CH + j < n_lim) st_from_float(cp + j, v[j]);
                        }
                    }
                } else {
                    // staged: transpose through smem so that one warp store covers 32 consecutive columns of a row
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < CH; ++j) stg[lane * 33 + j] = v[j];
                    __syncwarp();
                    const int ncol = ncol0 + (lane % CH);
                    const bool col_ok = (c * CH + (lane % CH)) < n_lim && lane < CH;
                    const float bj = bias_s[c * CH + (lane % CH)];
                    const int rows_left = g.rows_per_batch - (lt * TBM + q * 32);   // rows of this warp that exist
                    const long long m0w = (long long)b * g.rows_per_batch + lt * TBM + q * 32;
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) {
                        if (r >= rows_left) break;
                        if (col_ok) {
                            const long long mr = m0w + r;
                            float o = apply_act_rt(stg[r * 33 + (lane % CH)] + bj, g.act);
                            if (g.resid) {
                                const long long ri = mr * g.ldr + ncol;
                                o += g.resid_bf16 ? __bfloat162float(static_cast<const bf16*>(g.resid)[ri])
                                                  : static_cast<const float*>(g.resid)[ri];
                            }
                            if (g.tmpl) o += __ldg(g.tmpl + (mr / g.rows_per_tmpl) * (long long)g.N + ncol);
                            st_from_float(C + mr * g.ldc + ncol, o);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

template <int BN, typename TC, bool STAGED>
static int launch_tc(const TmapSet& maps, const TcParams& p, cudaStream_t s) {
    using Cfg = TcCfg<BN>;
    auto kern = gemm_tc_kernel<BN, TC, STAGED>;
    static bool attr_done = false;   // per instantiation
    if (!attr_done) {
        A2F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
        attr_done = true;
    }
    const int total = p.tiles_m_per_batch * p.num_batches * p.tiles_n;
    const int grid = total < sm_count() ? total : sm_count();
    kern<<<grid, TC_THREADS, Cfg::SMEM_BYTES, s>>>(maps, p);
    A2F_CHECK_LAUNCH("gemm_tc_kernel");
    count_launch();
    return A2F_OK;
}

template <int BN> static int dispatch_out(const TmapSet& maps, const TcParams& p, int c_bf16, bool staged, cudaStream_t s) {
    if (c_bf16) return staged ? launch_tc<BN, bf16, true>(maps, p, s) : launch_tc<BN, bf16, false>(maps, p, s);
    return staged ? launch_tc<BN, float, true>(maps, p, s) : launch_tc<BN, float, false>(maps, p, s);
}

static int g_force_bn = 0;   // debug: force a tile width (tests exercise every instantiation)

int gemm_tc(const GemmParams& g, int c_bf16, int mode, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0) return A2F_OK;
    TcParams p;
    p.g = g;
    p.mode = mode;
    p.lbo_enc = g_umma_fields[0];
    p.sbo_enc = g_umma_fields[1];
    p.desc_version = g_umma_fields[2];
    p.desc_layout = g_umma_fields[3];
    TmapSet maps;
    memset(&maps, 0, sizeof(maps));

    A2F_REQUIRE(g.rows_per_batch > 0 && g.M % g.rows_per_batch == 0, "gemm_tc: M must be a multiple of rows_per_batch");
    p.num_batches = g.M / g.rows_per_batch;
    p.tiles_m_per_batch = (g.rows_per_batch + TBM - 1) / TBM;

    int BN;
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
        p.tiles_n = (g.N + BN - 1) / BN;
        p.num_k_blocks = (g.K + TBK - 1) / TBK;
        // K segments: every tensor map must have non-overlapping rows (segment length <= row stride)
        int nseg = 1;
        if (g.a_row_stride < g.K) nseg = (int)((g.K + g.a_row_stride - 1) / g.a_row_stride);
        A2F_REQUIRE(nseg <= MAX_SEGS, "gemm_tc: A rows overlap too much (more than 4 K segments)");
        A2F_REQUIRE(g.K % nseg == 0 && (nseg == 1 || (g.K / nseg) % TBK == 0),
                    "gemm_tc: K segment length must be a multiple of 64");
        const int kseg = g.K / nseg;
        p.kb_per_seg = (nseg == 1) ? p.num_k_blocks : kseg / TBK;
        for (int sgi = 0; sgi < nseg; ++sgi) {
            uint64_t dims[3] = {(uint64_t)kseg, (uint64_t)g.rows_per_batch, (uint64_t)p.num_batches};
            uint64_t strides[2] = {(uint64_t)g.a_row_stride * 2,
                                   (uint64_t)(p.num_batches > 1 ? g.a_batch_stride : g.a_row_stride * g.rows_per_batch) * 2};
            uint32_t box[3] = {TBK, TBM, 1};
            const bf16* base = static_cast<const bf16*>(g.A) + (size_t)sgi * kseg;
            int rc = encode_tmap_bf16(&maps.a[sgi], base, 3, dims, strides, box, 1);
            if (rc != A2F_OK) return rc;
        }
        uint64_t bdims[2] = {(uint64_t)g.K, (uint64_t)g.N};
        uint64_t bstr[1] = {(uint64_t)g.ldw * 2};
        uint32_t bbox[2] = {TBK, (uint32_t)BN};
        int rc = encode_tmap_bf16(&maps.b, g.W, 2, bdims, bstr, bbox, 1);
        if (rc != A2F_OK) return rc;
    }

    const size_t csz = c_bf16 ? 2 : 4;
    const size_t vec_elems = 16 / csz;
    bool vec_ok = (reinterpret_cast<uintptr_t>(g.C) % 16 == 0) && (g.ldc % (long long)vec_elems == 0);
    if (g.resid) {
        const size_t rsz = g.resid_bf16 ? 2 : 4;
        vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(g.resid) % 16 == 0) && (g.ldr % (long long)(16 / rsz) == 0);
    }
    p.vec_ok = vec_ok ? 1 : 0;
    // staged (transposing) epilogue: coalesced 4-byte stores for outputs whose rows are not 16-byte aligned
    // (the 15069-wide vertex head) and for the template-add epilogue.
    const bool staged = (g.tmpl != nullptr) || !vec_ok;

    switch (BN) {
        case 256: return dispatch_out<256>(maps, p, c_bf16, staged, s);
        case 128: return dispatch_out<128>(maps, p, c_bf16, staged, s);
        case 64: return dispatch_out<64>(maps, p, c_bf16, staged, s);
        case 48: return dispatch_out<48>(maps, p, c_bf16, staged, s);
        default: return set_error(A2F_EINVAL, "gemm_tc: unsupported tile width");
    }
}

}  // namespace a2f

extern "C" int a2f_debug_set_umma_field(int field, unsigned value) {
    if (field >= 0 && field < 4) {
        a2f::g_umma_fields[field] = value;
        return A2F_OK;
    }
    if (field == 4) {   // force tile width (0 = automatic)
        if (value != 0 && value != 64 && value != 128 && value != 256) return a2f::set_error(A2F_EINVAL, "bad BN");
        a2f::g_force_bn = (int)value;
        return A2F_OK;
    }
    return a2f::set_error(A2F_EINVAL, "unknown debug field");
}

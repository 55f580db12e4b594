// Weight-gradient GEMM on tcgen05 for sm_100a:   dW[n, s*K + k] += sum_m dY[m, n] * X[(b, r + roff[s]), coff[s] + k]
//
// Both operands are read IN PLACE: dY [M,N] and X [M,K] are row-major with the reduction index m outermost, i.e.
// "MN-major" operands in UMMA terms.  TMA brings [64 reduction rows x 64 columns] boxes (128-byte rows, SWIZZLE_128B)
// and the shared-memory descriptors describe them as MN-major SW128 atoms (LBO = distance between 64-column atoms,
// SBO = distance between groups of 8 reduction rows); the instruction descriptor sets the a_major / b_major bits.
// No transposed copy of an activation or of a gradient is ever written to HBM.
//
// The reduction (M = B*T rows, up to 64k for the conv stack) is split over CTAs: work item = (output tile, batch,
// chunk of 64-row blocks); partial tiles leave through fp32 swizzled staging and cp.reduce.async.bulk.tensor (.add),
// so dW is accumulated in L2/HBM by the TMA unit -- the .grad accumulation semantics of autograd for free.
//
// Warp roles as in gemm_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue; accumulators are
// double-buffered in TMEM so the epilogue of item i overlaps the mainloop of item i+1.
#include "a2f_common.cuh"
#include "gemm_params.cuh"

namespace a2f {

constexpr int WBM = 128;          // dW rows per tile (= dY columns), UMMA M
constexpr int WBK = 64;           // reduction rows per stage
constexpr int W_THREADS = 320;
constexpr int W_ATOM_BYTES = 64 * WBK * 2;   // one [64 rows x 64 cols] bf16 box = 8 KB
constexpr int W_EPI_BYTES = 16384;           // 128 rows x 32 fp32 columns

struct WgradMaps {
    CUtensorMap dy;
    CUtensorMap x;
    CUtensorMap dw;
};

struct WgradTc {
    WgradParams g;
    int tiles_n, tiles_c, n_tiles;      // dW row tiles, column tiles
    int tiles_per_seg;                  // column tiles per segment (a tile never straddles two segments)
    int num_batches, kb_per_batch;      // 64-row reduction blocks per batch
    int chunks_per_batch, kb_per_chunk; // split of the reduction
    int total_items;
};

template <int BNW> struct WCfg {
    static constexpr int ACC_COLS = BNW;                    // 64 / 128 / 256 fp32 columns
    static constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
    static constexpr int A_STAGE = 2 * W_ATOM_BYTES;        // 128 dY columns
    static constexpr int B_STAGE = (BNW / 64) * W_ATOM_BYTES;
    static constexpr int STAGE = A_STAGE + B_STAGE;
    static constexpr int STAGES = (BNW >= 256) ? 4 : (BNW >= 128) ? 6 : 8;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE + 2 * W_EPI_BYTES + 256;
};

A2F_D void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}

// MN-major SWIZZLE_128B descriptor: LBO = 8192 B (next 64-column atom), SBO = 1024 B (next 8 reduction rows)
A2F_D uint64_t make_mn_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(W_ATOM_BYTES >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int BNW>
__global__ void __launch_bounds__(W_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradMaps maps, const WgradTc p) {
    using Cfg = WCfg<BNW>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)STAGES * Cfg::A_STAGE;
    uint8_t* sEpi = smem + (size_t)STAGES * Cfg::STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + 2 * W_EPI_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WgradParams& g = p.g;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.dy);
        tma_prefetch_desc(&maps.x);
        tma_prefetch_desc(&maps.dw);
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

    // item -> (tile, batch, chunk); items that share a reduction chunk are adjacent so their operand reads hit in L2
    auto decode = [&](int item, int& tn, int& tc, int& b, int& kb0, int& kb1) {
        const int tile = item % p.n_tiles, unit = item / p.n_tiles;
        tn = tile / p.tiles_c;
        tc = tile % p.tiles_c;
        b = unit / p.chunks_per_batch;
        const int ch = unit % p.chunks_per_batch;
        kb0 = ch * p.kb_per_chunk;
        kb1 = min(p.kb_per_batch, kb0 + p.kb_per_chunk);
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                int tn, tc, b, kb0, kb1;
                decode(item, tn, tc, b, kb0, kb1);
                const int sg = tc / p.tiles_per_seg;            // segment (tap) of the tile
                const int xcol = g.col_off(sg) + (tc - sg * p.tiles_per_seg) * BNW;
                const int xrow_off = g.row_off(sg);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], Cfg::STAGE);
                    uint8_t* dstA = sA + (size_t)stage * Cfg::A_STAGE;
                    uint8_t* dstB = sB + (size_t)stage * Cfg::B_STAGE;
#pragma unroll
                    for (int a = 0; a < 2; ++a)
                        tma_load_3d(dstA + a * W_ATOM_BYTES, &maps.dy, &full_bar[stage], tn * WBM + a * 64, kb * WBK, b);
#pragma unroll
                    for (int a = 0; a < BNW / 64; ++a)
                        tma_load_3d(dstB + a * W_ATOM_BYTES, &maps.x, &full_bar[stage], xcol + a * 64, kb * WBK + xrow_off, b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // D=f32, A=B=bf16, both MN-major (bits 15, 16), N=BNW, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(BNW >> 3) << 17) | ((uint32_t)(WBM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                int tn, tc, b, kb0, kb1;
                decode(item, tn, tc, b, kb0, kb1);
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = make_mn_desc(smem_u32(sA + (size_t)stage * Cfg::A_STAGE));
                    const uint64_t bdesc = make_mn_desc(smem_u32(sB + (size_t)stage * Cfg::B_STAGE));
#pragma unroll
                    for (int k = 0; k < WBK / 16; ++k) {
                        // 16 reduction rows = 2 groups of 8 rows x 128 B = 2048 B: +128 in the >>4 address field
                        umma_f16(d_tmem, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc,
                                 (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        __syncwarp();
    } else {
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        const bool leader = ((ew & 3) == 0) && lane == 0;
        const int bar_id = 1 + half;
        uint8_t* stage_buf = sEpi + half * W_EPI_BYTES;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int ktot = g.K * g.n_seg;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int tn, tc, b, kb0, kb1;
            decode(item, tn, tc, b, kb0, kb1);
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int r_tile = q * 32 + lane;
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::ACC_COLS);
            uint8_t* rowp = stage_buf + r_tile * 128;
            const int sg = tc / p.tiles_per_seg, within = tc - sg * p.tiles_per_seg;
            const int ocol = sg * g.K + within * BNW;         // first dW column of the tile
            const int n_lim = min(BNW, ktot - ocol);
#pragma unroll 1
            for (int blk = half; blk < BNW / 32; blk += 2) {
                const int col0 = blk * 32;
                if (col0 >= n_lim) break;
                float v[32];
                tmem_ld_32x32(t_row + col0, v);
                tmem_ld_wait();
                if (leader) tma_store_wait_read();
                named_bar_sync(bar_id, 128);
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const int pch = ch ^ (r_tile & 7);
                    uint4 u;
                    u.x = __float_as_uint(v[ch * 4 + 0]);
                    u.y = __float_as_uint(v[ch * 4 + 1]);
                    u.z = __float_as_uint(v[ch * 4 + 2]);
                    u.w = __float_as_uint(v[ch * 4 + 3]);
                    *reinterpret_cast<uint4*>(rowp + pch * 16) = u;
                }
                fence_proxy_async_smem();
                named_bar_sync(bar_id, 128);
                if (leader) {
                    tma_reduce_add_2d(&maps.dw, stage_buf, ocol + col0, tn * WBM);
                    tma_store_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (leader) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

template <int BNW> static int launch_wgrad(const WgradMaps& maps, const WgradTc& p, cudaStream_t s) {
    using Cfg = WCfg<BNW>;
    auto kern = wgrad_tc_kernel<BNW>;
    static bool attr_done = false;
    if (!attr_done) {
        A2F_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_done = true;
    }
    const int grid = p.total_items < sm_count() ? p.total_items : sm_count();
    kern<<<grid, W_THREADS, Cfg::SMEM, s>>>(maps, p);
    A2F_CHECK_LAUNCH("wgrad_tc_kernel");
    count_launch();
    return A2F_OK;
}

int wgrad_tc(const WgradParams& g, cudaStream_t s) {
    WgradTc p;
    p.g = g;
    const int ktot = g.K * g.n_seg;
    A2F_REQUIRE(g.N % 8 == 0 || g.dy_row_stride >= g.N, "wgrad_tc: bad dY shape");
    A2F_REQUIRE(g.dy_row_stride % 8 == 0 && g.dy_batch_stride % 8 == 0 && g.x_row_stride % 8 == 0 && g.x_batch_stride % 8 == 0,
                "wgrad_tc: operand strides must be multiples of 8 elements (16 bytes)");
    A2F_REQUIRE(g.ldw % 4 == 0 && reinterpret_cast<uintptr_t>(g.dW) % 16 == 0, "wgrad_tc: dW must be 16-byte aligned with ldw % 4 == 0");
    int BNW = g.K > 128 ? 256 : (g.K > 64 ? 128 : 64);
    if (g.n_seg > 1) {
        // a column tile must not straddle two segments: K is a multiple of the tile width, or one tile per segment
        // (columns K..BNW-1 of such a tile read zeros because the X map ends at x_col_off + K)
        while (BNW > 64 && g.K % BNW != 0) BNW >>= 1;
        A2F_REQUIRE(g.K % BNW == 0 || g.K < BNW, "wgrad_tc: with several segments K must be < 64 or a multiple of 64");
    }
    p.num_batches = g.M / g.rows_per_batch;
    p.tiles_n = (g.N + WBM - 1) / WBM;
    p.tiles_per_seg = (g.K + BNW - 1) / BNW;
    p.tiles_c = p.tiles_per_seg * g.n_seg;
    p.n_tiles = p.tiles_n * p.tiles_c;
    p.kb_per_batch = (g.rows_per_batch + WBK - 1) / WBK;
    // split the reduction until every SM has work (at least ~2 items per SM when the reduction is long enough)
    const int want_units = (2 * sm_count() + p.n_tiles - 1) / p.n_tiles;
    int chunks = (want_units + p.num_batches - 1) / p.num_batches;
    if (chunks < 1) chunks = 1;
    if (chunks > p.kb_per_batch) chunks = p.kb_per_batch;
    // never make chunks shorter than 4 blocks (256 rows): the fp32 reduce-add traffic would dominate
    const int max_chunks = (p.kb_per_batch + 3) / 4;
    if (chunks > max_chunks) chunks = max_chunks;
    p.kb_per_chunk = (p.kb_per_batch + chunks - 1) / chunks;
    p.chunks_per_batch = (p.kb_per_batch + p.kb_per_chunk - 1) / p.kb_per_chunk;
    p.total_items = p.n_tiles * p.num_batches * p.chunks_per_batch;

    WgradMaps maps;
    memset(&maps, 0, sizeof(maps));
    {
        uint64_t dims[3] = {(uint64_t)g.N, (uint64_t)g.rows_per_batch, (uint64_t)p.num_batches};
        uint64_t strides[2] = {(uint64_t)g.dy_row_stride * 2,
                               (uint64_t)(p.num_batches > 1 ? g.dy_batch_stride : g.dy_row_stride * g.rows_per_batch) * 2};
        uint32_t box[3] = {64, WBK, 1};
        int rc = encode_tmap_bf16(&maps.dy, g.dY, 3, dims, strides, box, 1);
        if (rc != A2F_OK) return rc;
    }
    {
        int row_len = 0;
        for (int i = 0; i < g.n_seg; ++i) {
            A2F_REQUIRE(g.col_off(i) >= 0 && g.col_off(i) % 8 == 0, "wgrad_tc: x_col_off must be a multiple of 8");
            if (g.col_off(i) + g.K > row_len) row_len = g.col_off(i) + g.K;
        }
        A2F_REQUIRE(row_len <= g.x_row_stride, "wgrad_tc: segments must stay inside one X row");
        uint64_t dims[3] = {(uint64_t)row_len, (uint64_t)g.x_rows, (uint64_t)p.num_batches};
        uint64_t strides[2] = {(uint64_t)g.x_row_stride * 2,
                               (uint64_t)(p.num_batches > 1 ? g.x_batch_stride : g.x_row_stride * g.x_rows) * 2};
        uint32_t box[3] = {64, WBK, 1};
        int rc = encode_tmap_bf16(&maps.x, g.X, 3, dims, strides, box, 1);
        if (rc != A2F_OK) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)g.N};
        uint64_t strides[1] = {(uint64_t)g.ldw * 4};
        uint32_t box[2] = {32, WBM};
        int rc = encode_tmap(&maps.dw, g.dW, 4, 2, dims, strides, box, 1);
        if (rc != A2F_OK) return rc;
    }
    switch (BNW) {
        case 256: return launch_wgrad<256>(maps, p, s);
        case 128: return launch_wgrad<128>(maps, p, s);
        default: return launch_wgrad<64>(maps, p, s);
    }
}

}  // namespace a2f

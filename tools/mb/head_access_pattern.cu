// microbenchmark: out[m,n] = tmpl[m,n] + 1 over [M,15069] fp32 with the access patterns of the vertex-head epilogue
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 15069, BN = 256, TBM = 128;
// CW: 32-column chunks handled per row before moving to the next row; RG: rows per load group (RG*CW loads in flight)
template <int CW, int RG, int PIPE>
__global__ void __launch_bounds__(256, 1) pat(const float* __restrict__ t, float* __restrict__ o, int M) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2;
    const int tiles_n = (N + BN - 1) / BN, tiles_m = M / TBM, total = tiles_n * tiles_m;
    constexpr int NCH = BN / 32 / 2;          // chunks per warp per tile (4)
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int nb = tile % tiles_n, mb = tile / tiles_n;
        const long long row0 = (long long)mb * TBM + q * 32;
#pragma unroll 1
        for (int cg = 0; cg < NCH / CW; ++cg) {
            // columns: CW == 1 -> interleaved chunks (c = half, half+2, ...) like the current kernel; else contiguous
            const int col0 = nb * BN + (CW == 1 ? (cg * 2 + half) * 32 : half * (BN / 2) + cg * CW * 32) + lane;
#pragma unroll 1
            for (int rg = 0; rg < 32; rg += RG) {
                float v[RG * CW];
#pragma unroll
                for (int r = 0; r < RG; ++r)
#pragma unroll
                    for (int c = 0; c < CW; ++c) {
                        const int col = col0 + c * 32;
                        v[r * CW + c] = col < N ? __ldg(t + (row0 + rg + r) * N + col) : 0.f;
                    }
#pragma unroll
                for (int r = 0; r < RG; ++r)
#pragma unroll
                    for (int c = 0; c < CW; ++c) {
                        const int col = col0 + c * 32;
                        if (col < N) o[(row0 + rg + r) * N + col] = v[r * CW + c] + 1.f;
                    }
            }
        }
    }
}
// rolling window like the WIDE epilogue: WR rows x 4 chunks per lane always in flight, every consumed slot refilled at once
template <int WR>
__global__ void __launch_bounds__(256, 1) roll(const float* __restrict__ t, float* __restrict__ o, int M) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2;
    const int tiles_n = (N + BN - 1) / BN, tiles_m = M / TBM, total = tiles_n * tiles_m;
    float v[WR * 4];
    auto base = [&](int tile, long long& row0, int& col0) {
        const int nb = tile % tiles_n, mb = tile / tiles_n;
        row0 = (long long)mb * TBM + q * 32;
        col0 = nb * BN + half * 128 + lane;
    };
    long long row0; int col0;
    base(blockIdx.x, row0, col0);
#pragma unroll
    for (int r = 0; r < WR; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) v[r * 4 + c] = (col0 + c * 32 < N) ? __ldg(t + (row0 + r) * N + col0 + c * 32) : 0.f;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        base(tile, row0, col0);
        long long rown = 0; int coln = 0;
        const bool hasn = tile + (int)gridDim.x < total;
        if (hasn) base(tile + gridDim.x, rown, coln);
#pragma unroll
        for (int ph = 0; ph < 32 / WR; ++ph) {
            const bool last = ph == 32 / WR - 1;
#pragma unroll
            for (int r = 0; r < WR; ++r) {
                const int rr = ph * WR + r;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float x = v[r * 4 + c];
                    if (!last) v[r * 4 + c] = (col0 + c * 32 < N) ? __ldg(t + (row0 + rr + WR) * N + col0 + c * 32) : 0.f;
                    else v[r * 4 + c] = (hasn && coln + c * 32 < N) ? __ldg(t + (rown + r) * N + coln + c * 32) : 0.f;
                    if (col0 + c * 32 < N) o[(row0 + rr) * N + col0 + c * 32] = x + 1.f;
                }
            }
        }
    }
}
__global__ void flat(const float* __restrict__ t, float* __restrict__ o, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) o[i] = t[i] + 1.f;
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e9;
    for (int i = 0; i < 5; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    const int M = 16384; const long long n = (long long)M * N;
    float *t, *o; cudaMalloc(&t, n * 4); cudaMalloc(&o, n * 4); cudaMemset(t, 0, n * 4);
    const double gb = 2.0 * n * 4 / 1e9;
    auto rep = [&](const char* name, float ms) { printf("%-40s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, gb / (ms * 1e-3)); };
    rep("flat 148x8x1024", timeit([&] { flat<<<148 * 8, 1024>>>(t, o, n); }));
    rep("rows32 x 128B (current), 8 warps", timeit([&] { pat<1, 32, 0><<<148, 256>>>(t, o, M); }));
    rep("rows8 x 512B, 8 warps", timeit([&] { pat<4, 8, 0><<<148, 256>>>(t, o, M); }));
    rep("rows16 x 512B (64 in flight), 8 warps", timeit([&] { pat<4, 16, 0><<<148, 256>>>(t, o, M); }));
    rep("rows16 x 256B, 8 warps", timeit([&] { pat<2, 16, 0><<<148, 256>>>(t, o, M); }));
    rep("rows32 x 128B, 2 CTAs/SM", timeit([&] { pat<1, 32, 0><<<296, 256>>>(t, o, M); }));
    rep("rows8 x 512B, 2 CTAs/SM", timeit([&] { pat<4, 8, 0><<<296, 256>>>(t, o, M); }));
    rep("rows32 x 128B, 4 CTAs/SM", timeit([&] { pat<1, 32, 0><<<592, 256>>>(t, o, M); }));
    rep("rows8 x 512B, 4 CTAs/SM", timeit([&] { pat<4, 8, 0><<<592, 256>>>(t, o, M); }));
    rep("rolling window 8 rows x 512B, 8 warps", timeit([&] { roll<8><<<148, 256>>>(t, o, M); }));
    rep("rolling window 16 rows x 512B, 8 warps", timeit([&] { roll<16><<<148, 256>>>(t, o, M); }));
    rep("rolling window 4 rows x 512B, 8 warps", timeit([&] { roll<4><<<148, 256>>>(t, o, M); }));
    // the same kernel with the GEMM's shared-memory carve-out (224 KB dynamic smem -> ~28 KB of L1 left)
    cudaFuncSetAttribute(roll<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    rep("rolling 8 rows x 512B, 224 KB smem", timeit([&] { roll<8><<<148, 256, 224 * 1024>>>(t, o, M); }));
    cudaFuncSetAttribute(roll<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    rep("rolling 8 rows x 512B, 160 KB smem", timeit([&] { roll<8><<<148, 256, 160 * 1024>>>(t, o, M); }));
    cudaFuncSetAttribute(roll<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    rep("rolling 4 rows x 512B, 224 KB smem", timeit([&] { roll<4><<<148, 256, 224 * 1024>>>(t, o, M); }));
    cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
    return 0;
}

// conv0_apply variants: 0 = full, 1 = stores only (no math), 2 = math only (one store per CTA), 3 = full with st.global.cs
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
    const float2 u = fmul2(x, ffma2(make_float2(0.0356774081f, 0.0356774081f), fmul2(x, x), make_float2(0.7978845608f, 0.7978845608f)));
    float2 t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
    const float2 hx = fmul2(x, make_float2(0.5f, 0.5f));
    return ffma2(hx, t, hx);
}
__device__ __forceinline__ unsigned pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<unsigned*>(&v);
}
constexpr int TCH = 64;
template <int MODE>
__global__ void __launch_bounds__(256) k(const float* __restrict__ audio, const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                         long long N, int L0) {
    const int b = blockIdx.y, t0 = blockIdx.x * TCH;
    __shared__ __align__(16) float xs[5 * TCH + 8];
    const float* x = audio + (long long)b * N;
    for (int i = threadIdx.x; i < 5 * TCH + 8; i += blockDim.x) {
        const long long gi = 5LL * t0 + i;
        xs[i] = (gi < N && i < 5 * TCH + 5) ? x[gi] : 0.f;
    }
    const int c = threadIdx.x * 2;
    float2 wp[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) wp[k] = make_float2(w[c * 10 + k], w[(c + 1) * 10 + k]);
    const float2 a_init = make_float2(w[c], w[c + 1]);
    __syncthreads();
    __nv_bfloat16* o = out + (long long)b * L0 * 512;
    const int tn = min(TCH, L0 - t0);
    float2 sink = make_float2(0.f, 0.f);
    for (int t4 = 0; t4 + 4 <= tn; t4 += 4) {
        float xw[28];
#pragma unroll
        for (int qd = 0; qd < 7; ++qd) {
            const float4 f = *reinterpret_cast<const float4*>(xs + 5 * t4 + 4 * qd);
            xw[4 * qd] = f.x; xw[4 * qd + 1] = f.y; xw[4 * qd + 2] = f.z; xw[4 * qd + 3] = f.w;
        }
        float2 a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = a_init;
        if (MODE != 1) {
#pragma unroll
            for (int k = 0; k < 10; ++k)
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = ffma2(wp[k], make_float2(xw[5 * u + k], xw[5 * u + k]), a[u]);
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u].x += xw[5 * u];
        }
        __nv_bfloat16* p = o + (long long)(t0 + t4) * 512 + c;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 y = MODE != 1 ? gelu_fast2(a[u]) : a[u];
            if (MODE == 2) { sink.x += y.x; sink.y += y.y; }
            else if (MODE == 3) asm volatile("st.global.cs.b32 [%0], %1;" ::"l"(p + u * 512), "r"(pack(y.x, y.y)) : "memory");
            else *reinterpret_cast<unsigned*>(p + u * 512) = pack(y.x, y.y);
        }
    }
    if (MODE == 2 && sink.x == 1234.5f) o[c] = __float2bfloat16(sink.y);
}
template <int MODE> void run(const char* name, const float* a, const float* w, __nv_bfloat16* o, long long N, int L0, int B, void* fl, size_t flb) {
    dim3 grid((L0 + TCH - 1) / TCH, B);
    float best = 1e9f;
    for (int r = 0; r < 6; ++r) {
        cudaMemsetAsync(fl, r, flb);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<MODE><<<grid, 256>>>(a, w, o, N, L0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    printf("%-34s %7.1f us   (%.2f TB/s of output)\n", name, best * 1e3, (double)B * L0 * 1024 / (best * 1e-3) / 1e12);
}
int main() {
    const int B = 32; const long long N = 80000; const int L0 = (int)((N - 10) / 5 + 1);
    float *a, *w; __nv_bfloat16* o; void* fl; const size_t flb = 256u << 20;
    cudaMalloc(&a, B * N * 4); cudaMalloc(&w, 512 * 10 * 4); cudaMalloc(&o, (size_t)B * L0 * 1024); cudaMalloc(&fl, flb);
    cudaMemset(a, 0, B * N * 4); cudaMemset(w, 0, 5120 * 4);
    run<0>("full", a, w, o, N, L0, B, fl, flb);
    run<1>("stores only", a, w, o, N, L0, B, fl, flb);
    run<2>("math only", a, w, o, N, L0, B, fl, flb);
    run<3>("full, st.global.cs", a, w, o, N, L0, B, fl, flb);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

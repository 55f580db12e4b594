// VocaLoss / FaceFormerLoss (ref:src/loss/loss.py:4-55) as one pass over pred and gt:
//   rec = mean_{row,v} sum_xyz (p-g)^2            (ref loss.py:29-30)
//   vel = mean_{pair,v} sum_xyz ((p1-p0)-(g1-g0))^2 over non-overlapping row pairs (2k,2k+1)   (ref loss.py:32-40)
//   loss = k_rec*rec + k_vel*vel                  (ref loss.py:51-55)
// Deterministic: per-CTA fp64 partials in the workspace, summed by a second single-CTA kernel (no atomics).
#include "a2f_common.cuh"
#include "gemm_params.cuh"

namespace a2f {

constexpr int LOSS_THREADS = 256;

__global__ void __launch_bounds__(LOSS_THREADS) loss_partial_kernel(const float* __restrict__ pred,
                                                                    const float* __restrict__ gt, long long pairs,
                                                                    int V3, double* __restrict__ partial) {
    const long long total = pairs * V3;
    double rec = 0.0, vel = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long k = i / V3;
        const int e = (int)(i - k * V3);
        const long long i0 = 2 * k * V3 + e, i1 = i0 + V3;
        const float p0 = __ldg(pred + i0), p1 = __ldg(pred + i1);
        const float g0 = __ldg(gt + i0), g1 = __ldg(gt + i1);
        const float d0 = p0 - g0, d1 = p1 - g1;
        const float dv = (p1 - p0) - (g1 - g0);
        rec += (double)(d0 * d0) + (double)(d1 * d1);
        vel += (double)(dv * dv);
    }
    __shared__ double sh[2][LOSS_THREADS / 32];
    rec = warp_sum_d(rec);
    vel = warp_sum_d(vel);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = rec; sh[1][warp] = vel; }
    __syncthreads();
    if (warp == 0) {
        double r = lane < LOSS_THREADS / 32 ? sh[0][lane] : 0.0;
        double v = lane < LOSS_THREADS / 32 ? sh[1][lane] : 0.0;
        r = warp_sum_d(r);
        v = warp_sum_d(v);
        if (lane == 0) { partial[2 * blockIdx.x] = r; partial[2 * blockIdx.x + 1] = v; }
    }
}

__global__ void loss_final_kernel(const double* __restrict__ partial, int nblocks, double inv_rec, double inv_vel,
                                  float k_rec, float k_vel, float* __restrict__ out3) {
    double r = 0.0, v = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) { r += partial[2 * i]; v += partial[2 * i + 1]; }
    r = warp_sum_d(r);
    v = warp_sum_d(v);
    if (threadIdx.x == 0) {
        const float rec = (float)(r * inv_rec), vel = (float)(v * inv_vel);
        out3[0] = rec * k_rec + vel * k_vel;
        out3[1] = rec;
        out3[2] = vel;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_bwd_kernel(const float* __restrict__ pred,
                                                                const float* __restrict__ gt, long long pairs, int V3,
                                                                float c_rec, float c_vel,
                                                                const float* __restrict__ gscale,
                                                                float* __restrict__ dpred) {
    const long long total = pairs * V3;
    const float gs = gscale ? __ldg(gscale) : 1.f;
    const float cr = c_rec * gs, cv = c_vel * gs;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long k = i / V3;
        const int e = (int)(i - k * V3);
        const long long i0 = 2 * k * V3 + e, i1 = i0 + V3;
        const float p0 = __ldg(pred + i0), p1 = __ldg(pred + i1);
        const float g0 = __ldg(gt + i0), g1 = __ldg(gt + i1);
        const float d0 = p0 - g0, d1 = p1 - g1;
        const float dv = (p1 - p0) - (g1 - g0);
        dpred[i0] = cr * d0 - cv * dv;
        dpred[i1] = cr * d1 + cv * dv;
    }
}

static int loss_grid() { return 4 * sm_count(); }

}  // namespace a2f

using namespace a2f;

extern "C" {

size_t a2f_voca_loss_workspace_bytes(void) { return (size_t)loss_grid() * 2 * sizeof(double); }

int a2f_voca_loss_fwd(const float* pred, const float* gt, long long rows, int V3, float k_rec, float k_vel,
                      float* out3, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(pred && gt && out3 && workspace, "a2f_voca_loss_fwd: NULL argument");
    A2F_REQUIRE(rows > 0 && rows % 2 == 0, "a2f_voca_loss_fwd: rows must be positive and even (ref loss.py:34 view(-1,2,..))");
    A2F_REQUIRE(V3 > 0 && V3 % 3 == 0, "a2f_voca_loss_fwd: V3 must be a positive multiple of 3");
    const int grid = loss_grid();
    A2F_REQUIRE(workspace_bytes >= (size_t)grid * 2 * sizeof(double), "a2f_voca_loss_fwd: workspace too small");
    A2F_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 8 == 0, "a2f_voca_loss_fwd: workspace must be 8-byte aligned");
    double* partial = static_cast<double*>(workspace);
    loss_partial_kernel<<<grid, LOSS_THREADS, 0, as_stream(stream)>>>(pred, gt, rows / 2, V3, partial);
    A2F_CHECK_LAUNCH("loss_partial_kernel");
    const double nv = (double)(V3 / 3);
    loss_final_kernel<<<1, 32, 0, as_stream(stream)>>>(partial, grid, 1.0 / ((double)rows * nv),
                                                      1.0 / ((double)(rows / 2) * nv), k_rec, k_vel, out3);
    A2F_CHECK_LAUNCH("loss_final_kernel");
    count_launch(2);
    return A2F_OK;
}

int a2f_voca_loss_bwd(const float* pred, const float* gt, long long rows, int V3, float k_rec, float k_vel,
                      const float* gscale, float* dpred, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(pred && gt && dpred, "a2f_voca_loss_bwd: NULL argument");
    A2F_REQUIRE(rows > 0 && rows % 2 == 0, "a2f_voca_loss_bwd: rows must be positive and even");
    A2F_REQUIRE(V3 > 0 && V3 % 3 == 0, "a2f_voca_loss_bwd: V3 must be a positive multiple of 3");
    const double nv = (double)(V3 / 3);
    const float c_rec = (float)(2.0 * k_rec / ((double)rows * nv));
    const float c_vel = (float)(2.0 * k_vel / ((double)(rows / 2) * nv));
    loss_bwd_kernel<<<loss_grid(), LOSS_THREADS, 0, as_stream(stream)>>>(pred, gt, rows / 2, V3, c_rec, c_vel, gscale,
                                                                         dpred);
    A2F_CHECK_LAUNCH("loss_bwd_kernel");
    count_launch();
    return A2F_OK;
}

/* ---- vertex head with the loss fused into its epilogue (tcgen05 path) ---- */
size_t a2f_vertex_head_loss_workspace_bytes(void) { return (size_t)sm_count() * 8 * 2 * sizeof(double); }

int a2f_vertex_head_loss(const void* z3, const void* w3, int K3, const float* bias, const float* tmpl, int rows_per_tmpl,
                         const float* gt, long long rows, int V3, float k_rec, float k_vel, float* pred, void* dy_bf16,
                         long long ld_dy, float* out3, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(z3 && w3 && gt && dy_bf16 && out3 && workspace, "a2f_vertex_head_loss: NULL argument");
    A2F_REQUIRE(rows > 0 && rows % 2 == 0 && rows < (1LL << 30), "a2f_vertex_head_loss: rows must be positive and even");
    A2F_REQUIRE(V3 > 0 && V3 % 3 == 0 && K3 > 0 && K3 % 8 == 0, "a2f_vertex_head_loss: bad V3 / K");
    A2F_REQUIRE(ld_dy >= V3, "a2f_vertex_head_loss: ld_dy must cover V3 columns");
    A2F_REQUIRE(tmpl == nullptr || rows_per_tmpl > 0, "a2f_vertex_head_loss: rows_per_tmpl must be positive");
    const int n_part = sm_count() * 8;
    A2F_REQUIRE(workspace_bytes >= (size_t)n_part * 2 * sizeof(double), "a2f_vertex_head_loss: workspace too small");
    A2F_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 8 == 0, "a2f_vertex_head_loss: workspace must be 8-byte aligned");
    cudaStream_t s = as_stream(stream);
    A2F_CHECK_CUDA(cudaMemsetAsync(workspace, 0, (size_t)n_part * 2 * sizeof(double), s));
    const double nv = (double)(V3 / 3);
    GemmParams p;
    p.M = (int)rows; p.N = V3; p.K = K3;
    p.A = z3; p.a_row_stride = K3; p.a_batch_stride = 0; p.rows_per_batch = (int)rows;
    p.W = w3; p.ldw = K3;
    p.bias = bias; p.act = A2F_ACT_NONE;
    p.resid = nullptr; p.resid_bf16 = 0; p.ldr = 0;
    p.tmpl = tmpl; p.rows_per_tmpl = tmpl ? rows_per_tmpl : 1;
    p.C = pred; p.ldc = V3; p.c_batch_stride = 0;
    p.loss_gt = gt;
    p.loss_dy = dy_bf16;
    p.ld_dy = ld_dy;
    p.c_rec = (float)(2.0 * k_rec / ((double)rows * nv));
    p.c_vel = (float)(2.0 * k_vel / ((double)(rows / 2) * nv));
    p.loss_partial = static_cast<double*>(workspace);
    rc = gemm_tc(p, 0, 0, s);
    if (rc != A2F_OK) return rc;
    loss_final_kernel<<<1, 32, 0, s>>>(static_cast<const double*>(workspace), n_part, 1.0 / ((double)rows * nv),
                                       1.0 / ((double)(rows / 2) * nv), k_rec, k_vel, out3);
    A2F_CHECK_LAUNCH("loss_final_kernel");
    count_launch();
    return A2F_OK;
}

/* ---- vertex head that follows the decoder rollout frame group by frame group (inference, tcgen05 path) ---- */
int a2f_vertex_head_stream_rows(int B, int T) {
    if (B <= 0 || T <= 0 || 128 % B != 0 || B % 32 != 0) return 0;
    const int fg = 128 / B;
    return (T + fg - 1) / fg * 128;
}

int a2f_vertex_head_stream(const void* z3_frame_major, const void* w3, int K3, const float* bias, const float* tmpl, int B, int T,
                           int V3, float* out, const unsigned* frames_done, int reserve_sms, void* stream) {
    int rc = require_sm100();
    if (rc != A2F_OK) return rc;
    A2F_REQUIRE(z3_frame_major && w3 && out && frames_done, "a2f_vertex_head_stream: NULL argument");
    A2F_REQUIRE(B > 0 && B % 32 == 0 && 128 % B == 0, "a2f_vertex_head_stream: B must be 32, 64 or 128");
    A2F_REQUIRE(T > 0 && V3 > 0 && K3 > 0 && K3 % 8 == 0, "a2f_vertex_head_stream: bad T / V3 / K");
    A2F_REQUIRE((long long)T * V3 < (1LL << 25), "a2f_vertex_head_stream: T * V3 must stay below 2^25 (row stride of the epilogue)");
    A2F_REQUIRE(reserve_sms >= 0 && reserve_sms < sm_count(), "a2f_vertex_head_stream: reserve_sms out of range");
    const int fg = 128 / B;                                   // frames per 128-row group
    GemmParams p;
    p.M = a2f_vertex_head_stream_rows(B, T); p.N = V3; p.K = K3;
    p.A = z3_frame_major; p.a_row_stride = K3; p.a_batch_stride = 128LL * K3; p.rows_per_batch = 128;
    p.W = w3; p.ldw = K3;
    p.bias = bias; p.act = A2F_ACT_NONE;
    p.resid = nullptr; p.resid_bf16 = 0; p.ldr = 0;
    p.tmpl = tmpl; p.rows_per_tmpl = 1;
    p.C = out; p.ldc = (long long)T * V3;                     // utterance u -> out + u * T * V3
    p.c_batch_stride = (long long)fg * V3;                    // frame group g -> frames g * fg ...
    p.perm_rows = B;
    p.perm_stride = V3;                                       // ... frame f of the group -> + f * V3
    p.live_rows = (long long)T * B;
    p.wait_counters = frames_done;
    p.wait_target = B;
    p.wait_per_batch = fg;
    p.wait_n = T;
    p.max_ctas = sm_count() - reserve_sms;
    return gemm_tc(p, 0, 0, as_stream(stream));
}

}  // extern "C"

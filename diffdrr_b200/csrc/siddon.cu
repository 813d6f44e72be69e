// siddon.cu -- Siddon exact-path DRR kernels for sm_100a (forward, backward, visit counter).
// One thread walks one ray (ray_math.cuh); blockIdx.y is the pose, blockIdx.x tiles the rays of that pose.
#include "kernels.h"
#include "ray_math.cuh"

namespace b200drr {

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads) siddon_fwd_general_kernel(const float* __restrict__ vol, VolDims dims,
                                                                      const float* __restrict__ src,
                                                                      const float* __restrict__ tgt,
                                                                      const float* __restrict__ raylen,
                                                                      float* __restrict__ out, int64_t N, float shift,
                                                                      float eps, int reduce, int align_corners)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int b = blockIdx.y;
    const int64_t r = (int64_t)b * N + n;
    const Ray ray = load_ray(src, tgt, b, r, eps);
    out[r] = siddon_ray_general(vol, dims, ray, __ldg(raylen + r), shift, reduce, align_corners);
}

__global__ void __launch_bounds__(kThreads) siddon_fwd_fast_kernel(const float* __restrict__ vol, VolDims dims,
                                                                   const float* __restrict__ src,
                                                                   const float* __restrict__ tgt,
                                                                   const float* __restrict__ raylen,
                                                                   float* __restrict__ out, int64_t N, float shift,
                                                                   float eps)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int b = blockIdx.y;
    const int64_t r = (int64_t)b * N + n;
    const Ray ray = load_ray(src, tgt, b, r, eps);
    out[r] = __ldg(raylen + r) * siddon_ray_fast<false>(vol, dims, ray, shift, nullptr);
}

__global__ void __launch_bounds__(kThreads) siddon_visits_kernel(VolDims dims, const float* __restrict__ src,
                                                                 const float* __restrict__ tgt,
                                                                 int32_t* __restrict__ visits, int64_t N, float shift,
                                                                 float eps)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int b = blockIdx.y;
    const int64_t r = (int64_t)b * N + n;
    const Ray ray = load_ray(src, tgt, b, r, eps);
    int count = 0;
    siddon_ray_fast<true>(nullptr, dims, ray, shift, &count);
    visits[r] = count;
}

__global__ void __launch_bounds__(kThreads) siddon_bwd_kernel(const float* __restrict__ vol, VolDims dims,
                                                              const float* __restrict__ src,
                                                              const float* __restrict__ tgt,
                                                              const float* __restrict__ raylen,
                                                              const float* __restrict__ gout,
                                                              float* __restrict__ g_src, float* __restrict__ g_tgt,
                                                              float* __restrict__ g_raylen, float* __restrict__ g_vol,
                                                              int64_t N, float shift, float eps, int stop_grad)
{
    __shared__ float red[32];
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const bool active = n < N;
    float gs[3] = {0.0f, 0.0f, 0.0f};
    if (active) {
        const int64_t r = (int64_t)b * N + n;
        const Ray ray = load_ray(src, tgt, b, r, eps);
        const float L = __ldg(raylen + r), g = __ldg(gout + r);
        float gt[3];
        const float acc = siddon_ray_bwd(vol, dims, ray, shift, g * L, stop_grad ? nullptr : g_vol, gs, gt);
        if (g_tgt) {
#pragma unroll
            for (int a = 0; a < 3; ++a) g_tgt[r * 3 + a] = gt[a];
        }
        if (g_raylen) g_raylen[r] = stop_grad ? 0.0f : g * acc;
    }
    if (g_src) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float tot = block_sum(gs[a], red);
            if (threadIdx.x == 0) atomicAdd(g_src + b * 3 + a, tot);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Detector-grid kernel: the rays of a pose are the full H x W detector (n = h*W + w).  A CTA owns a
// TW x TH pixel tile and every warp an 8 x 4 sub-tile, so the 32 lanes of a warp walk a compact ray
// bundle (~16 x 8 voxels across) and their gathers share 128-byte lines / 32-byte sectors in L1.
// U = voxel loads kept in flight per thread.
// ---------------------------------------------------------------------------------------------------
template <int TW, int TH, int U>
__global__ void __launch_bounds__(TW* TH) siddon_fwd_grid_kernel(const float* __restrict__ vol, VolDims dims,
                                                                 const float* __restrict__ src,
                                                                 const float* __restrict__ tgt,
                                                                 const float* __restrict__ raylen,
                                                                 float* __restrict__ out, int H, int W, float shift,
                                                                 float eps)
{
    constexpr int WX = TW / 8;  // warps per tile row
    const int tiles_x = (W + TW - 1) / TW;
    const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int px = tile_x * TW + (warp % WX) * 8 + (lane & 7);
    const int py = tile_y * TH + (warp / WX) * 4 + (lane >> 3);
    if (px >= W || py >= H) return;
    const int b = blockIdx.y;
    const int64_t r = ((int64_t)b * H + py) * W + px;
    const Ray ray = load_ray(src, tgt, b, r, eps);
    out[r] = __ldg(raylen + r) * siddon_ray_fast_ilp<U>(vol, dims, ray, shift);
}

template <int TW, int TH, int U>
static cudaError_t launch_grid_variant(const float* vol, VolDims dims, const float* src, const float* tgt,
                                       const float* raylen, float* out, int B, int H, int W, float shift, float eps,
                                       cudaStream_t stream)
{
    const dim3 grid((unsigned)(((W + TW - 1) / TW) * ((H + TH - 1) / TH)), (unsigned)B, 1);
    siddon_fwd_grid_kernel<TW, TH, U><<<grid, TW * TH, 0, stream>>>(vol, dims, src, tgt, raylen, out, H, W, shift, eps);
    return cudaGetLastError();
}

cudaError_t launch_siddon_fwd_grid(const float* vol, VolDims dims, const float* src, const float* tgt,
                                   const float* raylen, float* out, int B, int H, int W, float shift, float eps,
                                   int variant, cudaStream_t stream)
{
#define V(id, TW, TH, U) \
    case id: return launch_grid_variant<TW, TH, U>(vol, dims, src, tgt, raylen, out, B, H, W, shift, eps, stream);
    switch (variant) {
        V(0, 16, 8, 4)
        V(1, 16, 8, 1)
        V(2, 16, 8, 2)
        V(3, 16, 8, 8)
        V(4, 16, 16, 4)
        V(5, 32, 8, 4)
        V(6, 8, 8, 4)
        V(7, 8, 16, 4)
        V(8, 32, 16, 4)
        V(9, 8, 4, 4)
        default: return cudaErrorInvalidValue;
    }
#undef V
}

static inline dim3 ray_grid(int B, int64_t N) { return dim3((unsigned)((N + kThreads - 1) / kThreads), (unsigned)B, 1); }

cudaError_t launch_siddon_fwd(const float* vol, VolDims dims, const float* src, const float* tgt,
                              const float* raylen, float* out, int B, int64_t N, float shift, float eps, int reduce,
                              int align_corners, cudaStream_t stream)
{
    if (reduce == 0 && !align_corners)
        siddon_fwd_fast_kernel<<<ray_grid(B, N), kThreads, 0, stream>>>(vol, dims, src, tgt, raylen, out, N, shift, eps);
    else
        siddon_fwd_general_kernel<<<ray_grid(B, N), kThreads, 0, stream>>>(vol, dims, src, tgt, raylen, out, N, shift,
                                                                          eps, reduce, align_corners);
    return cudaGetLastError();
}

cudaError_t launch_siddon_visits(VolDims dims, const float* src, const float* tgt, int32_t* visits, int B, int64_t N,
                                 float shift, float eps, cudaStream_t stream)
{
    siddon_visits_kernel<<<ray_grid(B, N), kThreads, 0, stream>>>(dims, src, tgt, visits, N, shift, eps);
    return cudaGetLastError();
}

cudaError_t launch_siddon_bwd(const float* vol, VolDims dims, const float* src, const float* tgt,
                              const float* raylen, const float* gout, float* g_src, float* g_tgt, float* g_raylen,
                              float* g_vol, int B, int64_t N, float shift, float eps, int stop_grad,
                              cudaStream_t stream)
{
    if (g_src) {
        cudaError_t e = cudaMemsetAsync(g_src, 0, sizeof(float) * 3 * (size_t)B, stream);
        if (e != cudaSuccess) return e;
    }
    siddon_bwd_kernel<<<ray_grid(B, N), kThreads, 0, stream>>>(vol, dims, src, tgt, raylen, gout, g_src, g_tgt,
                                                               g_raylen, g_vol, N, shift, eps, stop_grad);
    return cudaGetLastError();
}

}  // namespace b200drr

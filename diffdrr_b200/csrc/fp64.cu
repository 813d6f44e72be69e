// fp64.cu -- double-precision Siddon / trilinear renderers for sm_100a.
//
// The reference reaches fp64 through `drr.to(torch.float64)` (drr.py:75 comment, utils.py:110): the same tensor algebra
// runs in double.  These kernels restate that algebra literally, one thread per ray (no tiling, no packed copies: fp64 is
// the accuracy path, the B200's fp64 rate is a small fraction of its fp32 rate), with the same closed-form backward as the
// fp32 kernels:
//   Siddon    renderers.py:34-76 + 94-113 + 143-169: every plane alpha of the three axes, merged ascending, midpoint ->
//             nearest voxel (half-to-even, zero padding), L * v * (alpha_{j+1} - alpha_j), sum or max;
//   Trilinear renderers.py:205-240: alpha_m = linspace(0,1,P)[m] * (amax - amin) + amin (the linspace is built in fp32 by
//             the reference even in the fp64 run, renderers.py:224), 8-corner interpolation with per-corner zero padding,
//             L * value * step, sum or max.
#include <math.h>

#include "kernels.h"

namespace b200drr {

namespace {

struct Ray64 {
    double s[3], d[3];
};

__device__ __forceinline__ Ray64 load_ray64(const double* src, const double* tgt, int b, int64_t r, double eps)
{
    Ray64 ray;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        ray.s[a] = src[b * 3 + a];
        ray.d[a] = (tgt[r * 3 + a] - ray.s[a]) + eps;  // renderers.py:104-106: eps is added to the direction
    }
    return ray;
}

// continuous voxel coordinate grid_sample sees for the point s + alpha * d (renderers.py:148-152 + ATen un-normalisation)
__device__ __forceinline__ double pix_at(double alpha, double s, double d, double shift, int D, int align_corners)
{
    const double g = 2.0 * ((s + alpha * d) + shift) / (double)D - 1.0;
    return align_corners ? (g + 1.0) * 0.5 * (double)(D - 1) : ((g + 1.0) * (double)D - 1.0) * 0.5;
}

// ascending merge of the three per-axis plane sequences alpha_a(i) = ((i - shift) - s_a) / d_a, i = 0..D_a
struct PlaneMerge {
    double cur[3];
    int i[3], step[3], left[3];
    double shift;
    __device__ void init(const Ray64& ray, const VolDims& dims, double sh)
    {
        shift = sh;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const bool up = ray.d[a] > 0.0;
            i[a] = up ? 0 : dims.d[a];
            step[a] = up ? 1 : -1;
            left[a] = dims.d[a] + 1;
            cur[a] = (((double)i[a] - shift) - ray.s[a]) / ray.d[a];
        }
    }
    __device__ bool pop(const Ray64& ray, double& alpha, int& axis)
    {
        axis = -1;
        double best = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (left[a] > 0 && (axis < 0 || cur[a] < best)) {
                best = cur[a];
                axis = a;
            }
        if (axis < 0) return false;
        alpha = best;
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (a == axis) {
                i[a] += step[a];
                --left[a];
                cur[a] = (((double)i[a] - shift) - ray.s[a]) / ray.d[a];
            }
        return true;
    }
};

__device__ __forceinline__ int64_t nearest_voxel(const Ray64& ray, const VolDims& dims, double mid, double shift, int align_corners)
{
    int64_t flat = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double r = nearbyint(pix_at(mid, ray.s[a], ray.d[a], shift, dims.d[a], align_corners));  // half to even
        if (!(r >= 0.0 && r < (double)dims.d[a])) return -1;                                             // zero padding
        flat = flat * dims.d[a] + (int64_t)r;
    }
    return flat;
}

__global__ void __launch_bounds__(128) siddon_fwd_f64_kernel(const double* __restrict__ vol, VolDims dims,
                                                             const double* __restrict__ src, const double* __restrict__ tgt,
                                                             const double* __restrict__ raylen, double* __restrict__ out, int64_t N,
                                                             double shift, double eps, int reduce, int align_corners)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int b = blockIdx.y;
    const int64_t r = (int64_t)b * N + n;
    const Ray64 ray = load_ray64(src, tgt, b, r, eps);
    const double L = raylen[r];
    PlaneMerge pm;
    pm.init(ray, dims, shift);
    double prev, alpha, acc = 0.0;
    int axis;
    bool first = true;
    pm.pop(ray, prev, axis);
    while (pm.pop(ray, alpha, axis)) {
        const int64_t v = nearest_voxel(ray, dims, 0.5 * (prev + alpha), shift, align_corners);
        const double term = (L * (v >= 0 ? vol[v] : 0.0)) * (alpha - prev);
        if (reduce == 0) acc += term;
        else if (first || term > acc) acc = term;
        first = false;
        prev = alpha;
    }
    out[r] = acc;
}

// closed form: dI/dalpha_m = L (v_{m-1} - v_m), dalpha/ds_a = (alpha - 1)/d_a, dalpha/dt_a = -alpha/d_a
__global__ void __launch_bounds__(128) siddon_bwd_f64_kernel(const double* __restrict__ vol, VolDims dims,
                                                             const double* __restrict__ src, const double* __restrict__ tgt,
                                                             const double* __restrict__ raylen, const double* __restrict__ gout,
                                                             double* __restrict__ g_src, double* __restrict__ g_tgt,
                                                             double* __restrict__ g_raylen, double* __restrict__ g_vol, int64_t N,
                                                             double shift, double eps, int stop_grad, int align_corners)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int b = blockIdx.y;
    const int64_t r = (int64_t)b * N + n;
    const Ray64 ray = load_ray64(src, tgt, b, r, eps);
    const double L = raylen[r], g = gout[r], gL = g * L;
    PlaneMerge pm;
    pm.init(ray, dims, shift);
    double A[3] = {0.0, 0.0, 0.0}, C[3] = {0.0, 0.0, 0.0};
    double prev, alpha, vprev = 0.0, acc = 0.0;
    int axprev, axis;
    pm.pop(ray, prev, axprev);
    while (pm.pop(ray, alpha, axis)) {
        const int64_t v = nearest_voxel(ray, dims, 0.5 * (prev + alpha), shift, align_corners);
        const double val = v >= 0 ? vol[v] : 0.0;
        const double len = alpha - prev;
        acc += val * len;
        if (g_vol && !stop_grad && v >= 0) atomicAdd(g_vol + v, gL * len);
        // the crossing at `prev` (axis axprev) separates the previous segment (vprev) from this one (val)
        const double coef = gL * (vprev - val);
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (a == axprev) {
                A[a] += coef * prev;
                C[a] += coef;
            }
        vprev = val;
        prev = alpha;
        axprev = axis;
    }
    {  // the last plane: beyond it the volume is zero padding
        const double coef = gL * vprev;
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (a == axprev) {
                A[a] += coef * prev;
                C[a] += coef;
            }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (g_tgt) g_tgt[r * 3 + a] = -A[a] / ray.d[a];
        if (g_src) atomicAdd(g_src + b * 3 + a, (A[a] - C[a]) / ray.d[a]);
    }
    if (g_raylen) g_raylen[r] = stop_grad ? 0.0 : g * acc;
}

// ATen's fp32 linspace(0, 1, P)[m] (symmetric about the midpoint, the second half as end - step * k), then cast to double
__device__ __forceinline__ double linspace01_f32(int m, int P)
{
    const float step = 1.0f / (float)(P - 1);
    if (m < P / 2) return (double)__fmul_rn(step, (float)m);
    return (double)fmaf(-step, (float)(P - 1 - m), 1.0f);
}

// 8-corner interpolation with per-corner zero padding; optionally the analytic gradient w.r.t. pix and a volume scatter
__device__ __forceinline__ double trilerp64(const double* __restrict__ vol, const VolDims& dims, const double pix[3], double grad[3],
                                            double* g_vol, double g_scale)
{
    double f[3];
    int64_t i0[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double fl = floor(pix[a]);
        i0[a] = (int64_t)fl;
        f[a] = pix[a] - fl;
    }
    double val = 0.0;
    if (grad) grad[0] = grad[1] = grad[2] = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int o[3] = {c & 1, (c >> 1) & 1, (c >> 2) & 1};
        double w[3];
        int64_t id[3];
        bool inb = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            id[a] = i0[a] + o[a];
            w[a] = o[a] ? f[a] : 1.0 - f[a];
            inb = inb && id[a] >= 0 && id[a] < dims.d[a];
        }
        if (!inb) continue;
        const int64_t flat = (id[0] * dims.d[1] + id[1]) * dims.d[2] + id[2];
        const double v = vol[flat];
        val += v * w[0] * w[1] * w[2];
        if (grad) {
            grad[0] += v * (o[0] ? 1.0 : -1.0) * w[1] * w[2];
            grad[1] += v * (o[1] ? 1.0 : -1.0) * w[0] * w[2];
            grad[2] += v * (o[2] ? 1.0 : -1.0) * w[0] * w[1];
        }
        if (g_vol) atomicAdd(g_vol + flat, g_scale * w[0] * w[1] * w[2]);
    }
    return val;
}

__global__ void __launch_bounds__(128) trilinear_fwd_f64_kernel(const double* __restrict__ vol, VolDims dims,
                                                                const double* __restrict__ src, const double* __restrict__ tgt,
                                                                const double* __restrict__ raylen, double* __restrict__ out, int64_t N,
                                                                double shift, double eps, int P,
                                                                const double* __restrict__ alpha_range, int reduce, int align_corners)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int b = blockIdx.y;
    const int64_t r = (int64_t)b * N + n;
    const Ray64 ray = load_ray64(src, tgt, b, r, eps);
    const double L = raylen[r], amin = alpha_range[0], amax = alpha_range[1];
    const double step = (amax - amin) / (double)(P - 1);
    double acc = 0.0;
    for (int m = 0; m < P; ++m) {
        const double alpha = linspace01_f32(m, P) * (amax - amin) + amin;
        double pix[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) pix[a] = pix_at(alpha, ray.s[a], ray.d[a], shift, dims.d[a], align_corners);
        const double term = (L * trilerp64(vol, dims, pix, nullptr, nullptr, 0.0)) * step;
        if (reduce == 0) acc += term;
        else if (m == 0 || term > acc) acc = term;
    }
    out[r] = acc;
}

__global__ void __launch_bounds__(128) trilinear_bwd_f64_kernel(const double* __restrict__ vol, VolDims dims,
                                                                const double* __restrict__ src, const double* __restrict__ tgt,
                                                                const double* __restrict__ raylen, const double* __restrict__ gout,
                                                                double* __restrict__ g_src, double* __restrict__ g_tgt,
                                                                double* __restrict__ g_raylen, double* __restrict__ g_vol,
                                                                double* __restrict__ g_alpha_range, int64_t N, double shift,
                                                                double eps, int P, const double* __restrict__ alpha_range,
                                                                int align_corners)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int b = blockIdx.y;
    const int64_t r = (int64_t)b * N + n;
    const Ray64 ray = load_ray64(src, tgt, b, r, eps);
    const double L = raylen[r], g = gout[r], amin = alpha_range[0], amax = alpha_range[1];
    const double range = amax - amin, step = range / (double)(P - 1);
    double scale[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) scale[a] = align_corners ? (double)(dims.d[a] - 1) / (double)dims.d[a] : 1.0;  // dpix/dx
    double gs[3] = {0.0, 0.0, 0.0}, gt[3] = {0.0, 0.0, 0.0}, sumv = 0.0, d_amin = 0.0, d_amax = 0.0;
    for (int m = 0; m < P; ++m) {
        const double lin = linspace01_f32(m, P), alpha = lin * range + amin;
        double pix[3], G[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) pix[a] = pix_at(alpha, ray.s[a], ray.d[a], shift, dims.d[a], align_corners);
        sumv += trilerp64(vol, dims, pix, G, g_vol, g * L * step);
        double gd = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double ga = G[a] * scale[a];
            gs[a] += (1.0 - alpha) * ga;
            gt[a] += alpha * ga;
            gd += ga * ray.d[a];
        }
        d_amin += (1.0 - lin) * gd;
        d_amax += lin * gd;
    }
    const double k = g * L * step;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (g_tgt) g_tgt[r * 3 + a] = k * gt[a];
        if (g_src) atomicAdd(g_src + b * 3 + a, k * gs[a]);
    }
    if (g_raylen) g_raylen[r] = g * step * sumv;
    if (g_alpha_range) {
        atomicAdd(g_alpha_range + 0, g * L * (-sumv / (double)(P - 1) + step * d_amin));
        atomicAdd(g_alpha_range + 1, g * L * (sumv / (double)(P - 1) + step * d_amax));
    }
}

inline dim3 grid64(int B, int64_t N) { return dim3((unsigned)((N + 127) / 128), (unsigned)B, 1); }

}  // namespace

cudaError_t launch_siddon_fwd_f64(const double* vol, VolDims dims, const double* src, const double* tgt, const double* raylen,
                                  double* out, int B, int64_t N, double shift, double eps, int reduce, int align_corners,
                                  cudaStream_t stream)
{
    siddon_fwd_f64_kernel<<<grid64(B, N), 128, 0, stream>>>(vol, dims, src, tgt, raylen, out, N, shift, eps, reduce, align_corners);
    return cudaGetLastError();
}

cudaError_t launch_siddon_bwd_f64(const double* vol, VolDims dims, const double* src, const double* tgt, const double* raylen,
                                  const double* gout, double* g_src, double* g_tgt, double* g_raylen, double* g_vol, int B,
                                  int64_t N, double shift, double eps, int stop_grad, int align_corners, cudaStream_t stream)
{
    if (g_src) {
        cudaError_t e = cudaMemsetAsync(g_src, 0, sizeof(double) * 3 * (size_t)B, stream);
        if (e != cudaSuccess) return e;
    }
    siddon_bwd_f64_kernel<<<grid64(B, N), 128, 0, stream>>>(vol, dims, src, tgt, raylen, gout, g_src, g_tgt, g_raylen, g_vol, N,
                                                           shift, eps, stop_grad, align_corners);
    return cudaGetLastError();
}

cudaError_t launch_trilinear_fwd_f64(const double* vol, VolDims dims, const double* src, const double* tgt, const double* raylen,
                                     double* out, int B, int64_t N, double shift, double eps, int n_points,
                                     const double* alpha_range, int reduce, int align_corners, cudaStream_t stream)
{
    trilinear_fwd_f64_kernel<<<grid64(B, N), 128, 0, stream>>>(vol, dims, src, tgt, raylen, out, N, shift, eps, n_points, alpha_range,
                                                              reduce, align_corners);
    return cudaGetLastError();
}

cudaError_t launch_trilinear_bwd_f64(const double* vol, VolDims dims, const double* src, const double* tgt, const double* raylen,
                                     const double* gout, double* g_src, double* g_tgt, double* g_raylen, double* g_vol,
                                     double* g_alpha_range, int B, int64_t N, double shift, double eps, int n_points,
                                     const double* alpha_range, int align_corners, cudaStream_t stream)
{
    if (g_src) {
        cudaError_t e = cudaMemsetAsync(g_src, 0, sizeof(double) * 3 * (size_t)B, stream);
        if (e != cudaSuccess) return e;
    }
    trilinear_bwd_f64_kernel<<<grid64(B, N), 128, 0, stream>>>(vol, dims, src, tgt, raylen, gout, g_src, g_tgt, g_raylen, g_vol,
                                                              g_alpha_range, N, shift, eps, n_points, alpha_range, align_corners);
    return cudaGetLastError();
}

}  // namespace b200drr

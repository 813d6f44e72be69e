// hostemu.cpp -- TEST INFRASTRUCTURE: compiles the product's per-ray device math
// (diffdrr_b200/csrc/ray_math.cuh) with g++ and loops over rays on the CPU, so that the CPU-only build
// container can check the kernel logic against the oracle and the goldens before any GPU time is spent.
// Never linked into libb200drr.so and never imported by the product package.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../diffdrr_b200/csrc/brick.cuh"
#include "../../diffdrr_b200/csrc/ncc_math.cuh"
#include "../../diffdrr_b200/csrc/psync.cuh"

using namespace b200drr;

static VolDims mk(int D0, int D1, int D2)
{
    VolDims d;
    d.d[0] = D0;
    d.d[1] = D1;
    d.d[2] = D2;
    return d;
}

extern "C" {

void emu_siddon_fwd(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt, const float* raylen,
                    float* out, int B, long N, float shift, float eps, int reduce, int align_corners)
{
    const VolDims dims = mk(D0, D1, D2);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            if (reduce == 0 && !align_corners) out[r] = raylen[r] * siddon_ray_fast<false>(vol, dims, ray, shift, nullptr);
            else out[r] = siddon_ray_general(vol, dims, ray, raylen[r], shift, reduce, align_corners);
        }
}

void emu_siddon_fwd_ilp(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt, const float* raylen,
                        float* out, int B, long N, float shift, float eps, int unroll)
{
    const VolDims dims = mk(D0, D1, D2);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            out[r] = raylen[r] * (unroll == 4    ? siddon_ray_fast_ilp<4>(vol, dims, ray, shift)
                                  : unroll == 3  ? siddon_ray_fast_ilp<3>(vol, dims, ray, shift)
                                  : unroll == -4 ? siddon_ray_lean<4>(vol, dims, ray, shift)
                                                 : siddon_ray_lean<1>(vol, dims, ray, shift));
        }
}

void emu_siddon_fwd_psync(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt,
                          const float* raylen, float* out, int B, long N, float shift, float eps, int slab, int unroll)
{
    const VolDims dims = mk(D0, D1, D2);
    const int n_slabs = slab > 0 ? (D0 + slab - 1) / slab : 1;
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            float acc = 0.0f;
            for (int sl = 0; sl < n_slabs; ++sl) {
                const int lo_v[3] = {slab > 0 ? sl * slab : 0, 0, 0};
                const int hi_v[3] = {slab > 0 ? std::min(D0, (sl + 1) * slab) : D0, D1, D2};
                const unsigned nvox = (unsigned)D0 * D1 * D2;
                acc += unroll == 2 ? siddon_ray_psync<2>(vol, nvox, lo_v, hi_v, D1 * D2, D2, 1, ray, shift, NoAlign())
                                   : siddon_ray_psync<1>(vol, nvox, lo_v, hi_v, D1 * D2, D2, 1, ray, shift, NoAlign());
            }
            out[r] = raylen[r] * acc;
        }
}

void emu_siddon_general(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt,
                        const float* raylen, float* out, int B, long N, float shift, float eps, int reduce,
                        int align_corners)
{
    const VolDims dims = mk(D0, D1, D2);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            out[r] = siddon_ray_general(vol, dims, ray, raylen[r], shift, reduce, align_corners);
        }
}

void emu_siddon_visits(int D0, int D1, int D2, const float* src, const float* tgt, int* visits, int B, long N,
                       float shift, float eps)
{
    const VolDims dims = mk(D0, D1, D2);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            siddon_ray_fast<true>(nullptr, dims, ray, shift, &visits[r]);
        }
}

void emu_siddon_bwd(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt, const float* raylen,
                    const float* gout, float* g_src, float* g_tgt, float* g_raylen, float* g_vol, int B, long N,
                    float shift, float eps, int stop_grad)
{
    const VolDims dims = mk(D0, D1, D2);
    std::memset(g_src, 0, sizeof(float) * 3 * B);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            float gs[3], gt[3];
            const float acc =
                siddon_ray_bwd(vol, dims, ray, shift, gout[r] * raylen[r], stop_grad ? nullptr : g_vol, gs, gt);
            for (int a = 0; a < 3; ++a) {
                g_tgt[r * 3 + a] = gt[a];
                g_src[b * 3 + a] += gs[a];
            }
            g_raylen[r] = stop_grad ? 0.0f : gout[r] * acc;
        }
}

void emu_siddon_bwd_lean(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt,
                         const float* raylen, const float* gout, float* g_src, float* g_tgt, float* g_raylen, float* g_vol,
                         int B, long N, float shift, float eps, int stop_grad, int slab)
{
    const VolDims dims = mk(D0, D1, D2);
    const int n_slabs = slab > 0 ? (D0 + slab - 1) / slab : 1;
    std::memset(g_src, 0, sizeof(float) * 3 * B);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            const float gL = gout[r] * raylen[r];
            float A[3] = {0, 0, 0}, C[3] = {0, 0, 0}, acc = 0;
            for (int sl = 0; sl < n_slabs; ++sl) {
                const int lo_v[3] = {slab > 0 ? sl * slab : 0, 0, 0};
                const int hi_v[3] = {slab > 0 ? std::min(D0, (sl + 1) * slab) : D0, D1, D2};
                acc += stop_grad ? siddon_ray_bwd_lean_box<4, false>(vol, dims, lo_v, hi_v, D1 * D2, D2, 1, ray, shift, gL, nullptr, A, C)
                                 : siddon_ray_bwd_lean_box<4, true>(vol, dims, lo_v, hi_v, D1 * D2, D2, 1, ray, shift, gL, g_vol, A, C);
            }
            for (int a = 0; a < 3; ++a) {
                g_tgt[r * 3 + a] = -gL * A[a] * ray.inv[a];
                g_src[b * 3 + a] += gL * (A[a] - C[a]) * ray.inv[a];
            }
            g_raylen[r] = stop_grad ? 0.0f : gout[r] * acc;
        }
}

// The training-step fast path: per-ray sensitivities from the two-axis walk (ray_math.cuh: siddon_ray_sens_box), then the
// elementwise backward exactly as sens_bwd_kernel forms it.  slab <= 0: whole volume.
void emu_siddon_sens(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt, const float* raylen,
                     const float* gout, float* g_src, float* g_tgt, float* g_raylen, float* out, int B, long N, float shift,
                     float eps, int stop_grad, int slab)
{
    const VolDims dims = mk(D0, D1, D2);
    // slab > 0: slabs of that many planes along axis 0; slab < 0: -slab pieces along each ray's own major axis (MAJ kernels)
    const int n_slabs = slab > 0 ? (D0 + slab - 1) / slab : slab < 0 ? -slab : 1;
    std::memset(g_src, 0, sizeof(float) * 3 * B);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            const float L = raylen[r];
            float sens[8] = {0, 0, 0, 0, 0, 0, 0, 0}, img = 0;
            for (int sl = 0; sl < n_slabs; ++sl) {
                int lo_v[3] = {slab > 0 ? sl * slab : 0, 0, 0};
                int hi_v[3] = {slab > 0 ? std::min(D0, (sl + 1) * slab) : D0, D1, D2};
                if (slab < 0 && !major_axis_piece(ray, dims, sl, -slab, lo_v, hi_v)) continue;
                if (slab != 0 && box_surely_missed(ray, lo_v, hi_v, shift)) continue;  // as the slab kernels do
                float A[3] = {0, 0, 0}, C[3] = {0, 0, 0};
                const float S = slab < 0 ? siddon_ray_sens_box<4, LoadPlain, true>(vol, dims, lo_v, hi_v, D1 * D2, D2, 1, ray, shift, A, C)
                                         : siddon_ray_sens_box<4>(vol, dims, lo_v, hi_v, D1 * D2, D2, 1, ray, shift, A, C);
                for (int a = 0; a < 3; ++a) {
                    const float k = L * ray.inv[a];
                    sens[a] += -k * A[a];
                    sens[4 + a] += k * (A[a] - C[a]);
                }
                sens[3] += S;
                img += L * S;
            }
            out[r] = img;
            for (int a = 0; a < 3; ++a) {
                g_tgt[r * 3 + a] = gout[r] * sens[a];
                g_src[b * 3 + a] += gout[r] * sens[4 + a];
            }
            g_raylen[r] = stop_grad ? 0.0f : gout[r] * sens[3];
        }
}

void emu_siddon_fwd_mask(const float* vol, const float* mask, int D0, int D1, int D2, const float* src, const float* tgt,
                         const float* raylen, float* out, int B, long N, int C, float shift, float eps)
{
    const VolDims dims = mk(D0, D1, D2);
    std::memset(out, 0, sizeof(float) * (size_t)B * C * N);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            siddon_ray_lean_mask<4>(vol, mask, dims, ray, shift, raylen[r], out + (long)b * C * N + n, N, C);
        }
}

void emu_trilinear_fwd_mask(const float* vol, const float* mask, int D0, int D1, int D2, const float* src,
                            const float* tgt, const float* raylen, float* out, int B, long N, int C, float shift, float eps,
                            int P, float amin, float amax, int align_corners)
{
    const VolDims dims = mk(D0, D1, D2);
    const float step = (amax - amin) / (float)(P - 1);
    std::memset(out, 0, sizeof(float) * (size_t)B * C * N);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            trilinear_ray_fwd_mask(vol, mask, dims, ray, shift, P, amin, amax, align_corners, raylen[r] * step,
                                   out + (long)b * C * N + n, N, C);
        }
}

// packed-corner path: host replica of pack_corners_kernel + the packed marchers
static std::vector<float4> pack_host(const float* vol, const VolDims& dims)
{
    const long n1 = dims.d[1] + 1, n2 = dims.d[2] + 1;
    std::vector<float4> packed(2 * (size_t)(dims.d[0] + 1) * n1 * n2);
    for (int i0 = -1; i0 < dims.d[0]; ++i0)
        for (int i1 = -1; i1 < dims.d[1]; ++i1)
            for (int i2 = -1; i2 < dims.d[2]; ++i2) {
                float v[8];
                for (int k = 0; k < 8; ++k) {
                    const int x = i0 + (k & 1), y = i1 + ((k >> 1) & 1), z = i2 + ((k >> 2) & 1);
                    const bool inb = x >= 0 && x < dims.d[0] && y >= 0 && y < dims.d[1] && z >= 0 && z < dims.d[2];
                    v[k] = inb ? vol[((long)x * dims.d[1] + y) * dims.d[2] + z] : 0.0f;
                }
                const size_t c = ((size_t)(i0 + 1) * n1 + (i1 + 1)) * n2 + (i2 + 1);
                packed[2 * c] = float4{v[0], v[1], v[2], v[3]};
                packed[2 * c + 1] = float4{v[4], v[5], v[6], v[7]};
            }
    return packed;
}

void emu_trilinear_packed(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt, const float* raylen,
                          const float* gout, float* out, float* g_src, float* g_tgt, float* g_raylen, float* g_alpha_range,
                          int B, long N, float shift, float eps, int P, float amin, float amax, int slab)
{
    const VolDims dims = mk(D0, D1, D2);
    const std::vector<float4> packed = pack_host(vol, dims);
    const float step = (amax - amin) / (float)(P - 1);
    std::memset(g_src, 0, sizeof(float) * 3 * B);
    const int n_slabs = slab > 0 ? (D0 + 1 + slab - 1) / slab : 1;
    double ga0 = 0, ga1 = 0;
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            out[r] = 0.0f;
            g_raylen[r] = 0.0f;
            for (int a = 0; a < 3; ++a) g_tgt[r * 3 + a] = 0.0f;
            for (int sl = 0; sl < n_slabs; ++sl) {
                const float s_lo = slab > 0 ? (float)(-1 + sl * slab) : -INFINITY;
                const float s_hi = slab > 0 ? (float)(-1 + (sl + 1) * slab) : INFINITY;
                out[r] += trilinear_ray_fwd_packed(packed.data(), dims, ray, shift, P, amin, amax, s_lo, s_hi) * (raylen[r] * step);
                const TriGrad tg =
                    trilinear_ray_bwd_packed(packed.data(), dims, ray, shift, P, amin, amax, gout[r], raylen[r], s_lo, s_hi);
                for (int a = 0; a < 3; ++a) {
                    g_tgt[r * 3 + a] += tg.gt[a];
                    g_src[b * 3 + a] += tg.gs[a];
                }
                g_raylen[r] += gout[r] * step * tg.sumV;
                ga0 += tg.ga0;
                ga1 += tg.ga1;
            }
        }
    g_alpha_range[0] = (float)ga0;
    g_alpha_range[1] = (float)ga1;
}

void emu_trilinear_fwd(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt,
                       const float* raylen, float* out, int B, long N, float shift, float eps, int P, float amin,
                       float amax, int reduce, int align_corners)
{
    const VolDims dims = mk(D0, D1, D2);
    const float step = (amax - amin) / (float)(P - 1);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            out[r] = trilinear_ray_fwd(vol, dims, ray, shift, P, amin, amax, reduce, align_corners) * (raylen[r] * step);
        }
}

void emu_trilinear_bwd(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt,
                       const float* raylen, const float* gout, float* g_src, float* g_tgt, float* g_raylen,
                       float* g_vol, float* g_alpha_range, int B, long N, float shift, float eps, int P, float amin,
                       float amax, int align_corners)
{
    const VolDims dims = mk(D0, D1, D2);
    const float step = (amax - amin) / (float)(P - 1);
    std::memset(g_src, 0, sizeof(float) * 3 * B);
    double ga0 = 0, ga1 = 0;
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            const TriGrad tg =
                trilinear_ray_bwd(vol, dims, ray, shift, P, amin, amax, align_corners, gout[r], raylen[r], g_vol);
            for (int a = 0; a < 3; ++a) {
                g_tgt[r * 3 + a] = tg.gt[a];
                g_src[b * 3 + a] += tg.gs[a];
            }
            g_raylen[r] = gout[r] * step * tg.sumV;
            ga0 += tg.ga0;
            ga1 += tg.ga1;
        }
    g_alpha_range[0] = (float)ga0;
    g_alpha_range[1] = (float)ga1;
}

// box_surely_missed (pre-test of the slab-major kernels) against the exact set-ups it short-cuts:
// out = {(ray, slab) pairs, pairs skipped, pairs skipped although an exact set-up reports a hit (must be 0), exact hits}
void emu_box_pretest(int D0, int D1, int D2, const float* src, const float* tgt, int B, long N, int slab, float shift,
                     float eps, double* out)
{
    const VolDims dims = mk(D0, D1, D2);
    const int n_slabs = (D0 + slab - 1) / slab;
    double pairs = 0, skipped = 0, wrong = 0, hits = 0;
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const Ray ray = load_ray(src, tgt, b, (long)b * N + n, eps);
            for (int sl = 0; sl < n_slabs; ++sl) {
                const int lo_v[3] = {sl * slab, 0, 0};
                const int hi_v[3] = {std::min(D0, (sl + 1) * slab), D1, D2};
                const bool hit = start_walk_box(ray, lo_v, hi_v, shift).hit || start_walk_frame(ray, dims, lo_v, hi_v, shift).hit;
                const bool skip = box_surely_missed(ray, lo_v, hi_v, shift);
                pairs += 1;
                skipped += skip;
                hits += hit;
                wrong += (skip && hit);
            }
        }
    out[0] = pairs; out[1] = skipped; out[2] = wrong; out[3] = hits;
}

// Backward of the general walk (reduce "sum"/"max", align_corners) and trilinear reduce="max" through the device routines
void emu_siddon_bwd_general(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt, const float* raylen,
                            const float* gout, float* g_src, float* g_tgt, float* g_raylen, float* g_vol, int B, long N,
                            float shift, float eps, int stop_grad, int reduce, int align_corners)
{
    const VolDims dims = mk(D0, D1, D2);
    std::memset(g_src, 0, sizeof(float) * 3 * B);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            float gs[3], gt[3];
            const float acc = siddon_ray_general_bwd(vol, dims, ray, raylen[r], gout[r] * raylen[r], shift, reduce,
                                                     align_corners, stop_grad ? nullptr : g_vol, gs, gt);
            for (int a = 0; a < 3; ++a) {
                g_tgt[r * 3 + a] = gt[a];
                g_src[b * 3 + a] += gs[a];
            }
            g_raylen[r] = stop_grad ? 0.0f : gout[r] * acc;
        }
}

void emu_trilinear_bwd_max(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt, const float* raylen,
                           const float* gout, float* g_src, float* g_tgt, float* g_raylen, float* g_vol,
                           float* g_alpha_range, int B, long N, float shift, float eps, int P, float amin, float amax,
                           int align_corners)
{
    const VolDims dims = mk(D0, D1, D2);
    const float step = (amax - amin) / (float)(P - 1);
    std::memset(g_src, 0, sizeof(float) * 3 * B);
    double ga0 = 0, ga1 = 0;
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            const TriGrad tg =
                trilinear_ray_bwd_max(vol, dims, ray, shift, P, amin, amax, align_corners, gout[r], raylen[r], g_vol);
            for (int a = 0; a < 3; ++a) {
                g_tgt[r * 3 + a] = tg.gt[a];
                g_src[b * 3 + a] += tg.gs[a];
            }
            g_raylen[r] = gout[r] * step * tg.sumV;
            ga0 += tg.ga0;
            ga1 += tg.ga1;
        }
    g_alpha_range[0] = (float)ga0;
    g_alpha_range[1] = (float)ga1;
}

// Siddon mode="bilinear" through the device routine: image and (grads != 0) gradients in one call
void emu_siddon_bilinear(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt, const float* raylen,
                         const float* gout, float* out, float* g_src, float* g_tgt, float* g_raylen, float* g_vol, int B,
                         long N, float shift, float eps, int stop_grad, int reduce, int align_corners, int grads)
{
    const VolDims dims = mk(D0, D1, D2);
    std::memset(g_src, 0, sizeof(float) * 3 * B);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            float gs[3] = {0, 0, 0}, gt[3] = {0, 0, 0}, sum_tl = 0;
            if (grads) {
                out[r] = siddon_ray_bilinear<true>(vol, dims, ray, raylen[r], shift, 0, align_corners, gout[r], stop_grad != 0,
                                                   stop_grad ? nullptr : g_vol, gs, gt, sum_tl);
                for (int a = 0; a < 3; ++a) {
                    g_tgt[r * 3 + a] = gt[a];
                    g_src[b * 3 + a] += gs[a];
                }
                g_raylen[r] = stop_grad ? 0.0f : gout[r] * sum_tl;
            } else {
                out[r] = siddon_ray_bilinear<false>(vol, dims, ray, raylen[r], shift, reduce, align_corners, 0.0f, true, nullptr,
                                                    gs, gt, sum_tl);
            }
        }
}

// EXPERIMENT: chunk-reuse lean walk over a transposed (major-axis-fastest), padded copy; slab > 0 cuts axis 0.
// `axis` = which volume axis is fastest in volT: volT[i_p][i_q][i_axis] with (p, q) the other two axes in ascending order.
void emu_siddon_fwd_chunk(const float* volT, int D0, int D1, int D2, int axis, int width, const float* src, const float* tgt,
                          const float* raylen, float* out, int B, long N, float shift, float eps, int slab)
{
    const int D[3] = {D0, D1, D2};
    const int p = axis == 0 ? 1 : 0, q = axis == 2 ? 1 : 2;
    int st[3];
    st[axis] = 1;
    st[q] = D[axis];
    st[p] = D[axis] * D[q];
    const int n_slabs = slab > 0 ? (D0 + slab - 1) / slab : 1;
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            float acc = 0.0f;
            for (int sl = 0; sl < n_slabs; ++sl) {
                const int lo_v[3] = {slab > 0 ? sl * slab : 0, 0, 0};
                const int hi_v[3] = {slab > 0 ? std::min(D0, (sl + 1) * slab) : D0, D1, D2};
                acc += width == 4 ? siddon_ray_lean_box_chunk<4, 4>(volT, lo_v, hi_v, st[0], st[1], st[2], ray, shift)
                                  : siddon_ray_lean_box_chunk<4, 2>(volT, lo_v, hi_v, st[0], st[1], st[2], ray, shift);
            }
            out[r] = raylen[r] * acc;
        }
}

// EXPERIMENT: the sensitivities walk through the chunk loader; writes sens [B][N][8] like the kernel
void emu_siddon_sens_chunk(const float* volT, int D0, int D1, int D2, int axis, int width, const float* src, const float* tgt,
                           const float* raylen, float* out, float* sens, int B, long N, float shift, float eps, int slab,
                           int chunked)
{
    const VolDims dims = mk(D0, D1, D2);
    const int D[3] = {D0, D1, D2};
    int st[3] = {D1 * D2, D2, 1};
    if (chunked) {
        const int p = axis == 0 ? 1 : 0, q = axis == 2 ? 1 : 2;
        st[axis] = 1;
        st[q] = D[axis];
        st[p] = D[axis] * D[q];
    }
    const int n_slabs = slab > 0 ? (D0 + slab - 1) / slab : 1;
    std::memset(sens, 0, sizeof(float) * 8 * (size_t)B * N);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            const float L = raylen[r];
            float img = 0;
            for (int sl = 0; sl < n_slabs; ++sl) {
                const int lo_v[3] = {slab > 0 ? sl * slab : 0, 0, 0};
                const int hi_v[3] = {slab > 0 ? std::min(D0, (sl + 1) * slab) : D0, D1, D2};
                float A[3] = {0, 0, 0}, C[3] = {0, 0, 0};
                const float S = !chunked ? siddon_ray_sens_box<4>(volT, dims, lo_v, hi_v, st[0], st[1], st[2], ray, shift, A, C)
                                : width == 4
                                    ? siddon_ray_sens_box<4, LoadChunk<4>>(volT, dims, lo_v, hi_v, st[0], st[1], st[2], ray, shift, A, C)
                                    : siddon_ray_sens_box<4, LoadChunk<2>>(volT, dims, lo_v, hi_v, st[0], st[1], st[2], ray, shift, A, C);
                for (int a = 0; a < 3; ++a) {
                    const float k = L * ray.inv[a];
                    sens[r * 8 + a] += -k * A[a];
                    sens[r * 8 + 4 + a] += k * (A[a] - C[a]);
                }
                sens[r * 8 + 3] += S;
                img += L * S;
            }
            out[r] = img;
        }
}

// the reference for it: the plain lean walk on the ORIGINAL layout with the same slab cuts
void emu_siddon_fwd_lean_slab(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt,
                              const float* raylen, float* out, int B, long N, float shift, float eps, int slab)
{
    const VolDims dims = mk(D0, D1, D2);
    const int n_slabs = slab > 0 ? (D0 + slab - 1) / slab : slab < 0 ? -slab : 1;  // slab < 0: major-axis pieces (MAJ kernels)
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            float acc = 0.0f;
            for (int sl = 0; sl < n_slabs; ++sl) {
                int lo_v[3] = {slab > 0 ? sl * slab : 0, 0, 0};
                int hi_v[3] = {slab > 0 ? std::min(D0, (sl + 1) * slab) : D0, D1, D2};
                if (slab < 0 && (!major_axis_piece(ray, dims, sl, -slab, lo_v, hi_v) || box_surely_missed(ray, lo_v, hi_v, shift)))
                    continue;
                acc += siddon_ray_lean_box<4>(vol, lo_v, hi_v, D1 * D2, D2, 1, ray, shift);
            }
            out[r] = raylen[r] * acc;
        }
}

// mask_to_channels backward through the device routines (FetchMasked / SampleGradMasked): gout is [B][C][N]
void emu_siddon_bwd_mask(const float* vol, const float* mask, int D0, int D1, int D2, const float* src, const float* tgt,
                         const float* raylen, const float* gout, float* g_src, float* g_tgt, float* g_raylen, float* g_vol,
                         int B, long N, int C, float shift, float eps, int stop_grad)
{
    const VolDims dims = mk(D0, D1, D2);
    std::memset(g_src, 0, sizeof(float) * 3 * B);
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            float gs[3], gt[3];
            const FetchMasked fetch{vol, mask, gout + (long)b * C * N + n, N, C};
            const float acc = siddon_ray_bwd_f(fetch, dims, ray, shift, raylen[r], stop_grad ? nullptr : g_vol, gs, gt);
            for (int a = 0; a < 3; ++a) {
                g_tgt[r * 3 + a] = gt[a];
                g_src[b * 3 + a] += gs[a];
            }
            g_raylen[r] = stop_grad ? 0.0f : acc;
        }
}

void emu_trilinear_bwd_mask(const float* vol, const float* mask, int D0, int D1, int D2, const float* src, const float* tgt,
                            const float* raylen, const float* gout, float* g_src, float* g_tgt, float* g_raylen,
                            float* g_vol, float* g_alpha_range, int B, long N, int C, float shift, float eps, int P,
                            float amin, float amax, int align_corners)
{
    const VolDims dims = mk(D0, D1, D2);
    const float step = (amax - amin) / (float)(P - 1);
    std::memset(g_src, 0, sizeof(float) * 3 * B);
    double ga0 = 0, ga1 = 0;
    for (int b = 0; b < B; ++b)
        for (long n = 0; n < N; ++n) {
            const long r = (long)b * N + n;
            const Ray ray = load_ray(src, tgt, b, r, eps);
            const SampleGradMasked sg{mask, gout + (long)b * C * N + n, N, C};
            const TriGrad tg = trilinear_ray_bwd_g(GatherPlain{vol}, dims, ray, shift, P, amin, amax, align_corners, 1.0f,
                                                   raylen[r], g_vol, -INFINITY, INFINITY, sg);
            for (int a = 0; a < 3; ++a) {
                g_tgt[r * 3 + a] = tg.gt[a];
                g_src[b * 3 + a] += tg.gs[a];
            }
            g_raylen[r] = step * tg.sumV;
            ga0 += tg.ga0;
            ga1 += tg.ga1;
        }
    g_alpha_range[0] = (float)ga0;
    g_alpha_range[1] = (float)ga1;
}


// Brick-major forward (siddon_brick.cu) as the kernel does it: ray table, per-pose detector geometry from three corner
// rays, per (brick, pose) pixel rectangle from the projected corners, 8x4 pixel tiles, conservative hit test, exact
// per-pair walk on a zero-filled copy of the brick.  Returns the number of RECTANGLE VIOLATIONS: pixels outside a
// brick's rectangle whose exact walk set-up hits the brick (must be 0; counted only when check != 0 -- it is O(all pairs)).
long emu_siddon_fwd_brick(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt, const float* raylen,
                          float* out, int B, int H, int W, float shift, float eps, int BX, int BY, int BZ, int check,
                          long* stats /* [4]: candidates, hits (test), hits (exact), pairs walked */)
{
    const VolDims dims = mk(D0, D1, D2);
    const long N = (long)H * W;
    std::vector<Ray> rays((size_t)B * N);
    std::vector<PoseGeo> geo(B);
    for (int b = 0; b < B; ++b) {
        for (long n = 0; n < N; ++n) rays[(size_t)b * N + n] = load_ray(src, tgt, b, (long)b * N + n, eps);
        const Ray& r00 = rays[(size_t)b * N];
        const Ray& r0w = rays[(size_t)b * N + (W - 1)];
        const Ray& rh0 = rays[(size_t)b * N + (long)(H - 1) * W];
        float t00[3], t0w[3], th0[3];
        for (int a = 0; a < 3; ++a) {
            t00[a] = r00.s[a] + r00.d[a];
            t0w[a] = r00.s[a] + r0w.d[a];
            th0[a] = r00.s[a] + rh0.d[a];
        }
        geo[b] = make_pose_geo(r00.s, t00, t0w, th0, H, W);
    }
    std::fill(out, out + (size_t)B * N, 0.0f);
    long violations = 0;
    long st[4] = {0, 0, 0, 0};
    std::vector<float> brick((size_t)BX * BY * BZ);
    const int nb0 = (D0 + BX - 1) / BX, nb1 = (D1 + BY - 1) / BY, nb2 = (D2 + BZ - 1) / BZ;
    for (int i0 = 0; i0 < nb0; ++i0)
        for (int i1 = 0; i1 < nb1; ++i1)
            for (int i2 = 0; i2 < nb2; ++i2) {
                const int org[3] = {i0 * BX, i1 * BY, i2 * BZ};
                const int lo_v[3] = {org[0], org[1], org[2]};
                const int hi_v[3] = {std::min(org[0] + BX, D0), std::min(org[1] + BY, D1), std::min(org[2] + BZ, D2)};
                for (int x = 0; x < BX; ++x)  // what the TMA box copy leaves in shared memory (zero fill past the volume)
                    for (int y = 0; y < BY; ++y)
                        for (int z = 0; z < BZ; ++z) {
                            const int g0 = org[0] + x, g1 = org[1] + y, g2 = org[2] + z;
                            brick[((size_t)x * BY + y) * BZ + z] =
                                (g0 < D0 && g1 < D1 && g2 < D2) ? vol[((size_t)g0 * D1 + g1) * D2 + g2] : 0.0f;
                        }
                LdHost ld{brick.data()};
                for (int b = 0; b < B; ++b) {
                    const PixRect rc = brick_rect(geo[b], lo_v, hi_v, shift, H, W);
                    float clo[3], chi[3];
                    for (int a = 0; a < 3; ++a) {
                        clo[a] = ((float)lo_v[a] - shift) - geo[b].S[a];
                        chi[a] = ((float)hi_v[a] - shift) - geo[b].S[a];
                    }
                    if (check) {
                        for (int py = 0; py < H; ++py)
                            for (int px = 0; px < W; ++px) {
                                const bool inside = rc.x0 <= rc.x1 && px >= rc.x0 && px <= rc.x1 && py >= rc.y0 && py <= rc.y1;
                                const Ray& ray = rays[(size_t)b * N + (long)py * W + px];
                                const bool exact = start_walk_box(ray, lo_v, hi_v, shift).hit;
                                float a_in, a_out;
                                const bool maybe = brick_maybe_hit(ray.inv, clo, chi, a_in, a_out);
                                if (exact && (!inside || !maybe)) ++violations;
                            }
                    }
                    if (rc.x0 > rc.x1 || rc.y0 > rc.y1) continue;
                    const int tw = (rc.x1 - rc.x0) / 8 + 1, th = (rc.y1 - rc.y0) / 4 + 1;
                    for (int t = 0; t < tw * th; ++t)
                        for (int lane = 0; lane < 32; ++lane) {
                            const int ty = t / tw, tx = t - ty * tw;
                            const int px = rc.x0 + tx * 8 + (lane & 7), py = rc.y0 + ty * 4 + (lane >> 3);
                            if (px > rc.x1 || py > rc.y1) continue;
                            ++st[0];
                            const long r = (long)b * N + (long)py * W + px;
                            const Ray& ray = rays[r];
                            float a_in, a_out;
                            if (!brick_maybe_hit(ray.inv, clo, chi, a_in, a_out)) continue;
                            ++st[1];
                            const unsigned it = pack_item(step_bin(a_in, a_out, fabsf(ray.d[0]) + fabsf(ray.d[1]) + fabsf(ray.d[2]),
                                                                   (float)kBrickBins / (float)(BX + BY + BZ + 8)),
                                                          b % kBrickPoseChunk, py * W + px);
                            if (item_ray(it) != py * W + px || item_pose(it) != b % kBrickPoseChunk) ++violations;
                            if (start_walk_box(ray, lo_v, hi_v, shift).hit) ++st[2];
                            const float part = brick_pair_fwd<4>(ld, ray, lo_v, hi_v, org, BY * BZ, BZ, 1, shift);
                            if (part != 0.0f) {
                                ++st[3];
                                out[r] += raylen[r] * part;
                            }
                        }
                }
            }
    if (stats)
        for (int i = 0; i < 4; ++i) stats[i] = st[i];
    return violations;
}

void emu_set_rcp_perturb(float p) { g_emu_rcp_perturb = p; }

// Production decomposition of siddon_brick.cu (second generation): per (brick, pose) the 4-row tile bands are clipped to
// the projected outline (row_span), the per-pair walk is brick_pair_fwd_lean (no entry fix-ups, accumulated alphas).
// Same return value / stats as emu_siddon_fwd_brick.
long emu_siddon_fwd_brick2(const float* vol, int D0, int D1, int D2, const float* src, const float* tgt, const float* raylen,
                           float* out, int B, int H, int W, float shift, float eps, int BX, int BY, int BZ, int check, int lean,
                           long* stats)
{
    const long N = (long)H * W;
    std::vector<Ray> rays((size_t)B * N);
    std::vector<PoseGeo> geo(B);
    for (int b = 0; b < B; ++b) {
        for (long n = 0; n < N; ++n) rays[(size_t)b * N + n] = load_ray(src, tgt, b, (long)b * N + n, eps);
        const Ray& r00 = rays[(size_t)b * N];
        const Ray& r0w = rays[(size_t)b * N + (W - 1)];
        const Ray& rh0 = rays[(size_t)b * N + (long)(H - 1) * W];
        float t00[3], t0w[3], th0[3];
        for (int a = 0; a < 3; ++a) {
            t00[a] = r00.s[a] + r00.d[a];
            t0w[a] = r00.s[a] + r0w.d[a];
            th0[a] = r00.s[a] + rh0.d[a];
        }
        geo[b] = make_pose_geo(r00.s, t00, t0w, th0, H, W);
    }
    std::fill(out, out + (size_t)B * N, 0.0f);
    long violations = 0;
    long st[4] = {0, 0, 0, 0};
    std::vector<float> brick((size_t)BX * BY * BZ);
    std::vector<char> cand((size_t)N);
    const int nb0 = (D0 + BX - 1) / BX, nb1 = (D1 + BY - 1) / BY, nb2 = (D2 + BZ - 1) / BZ;
    for (int i0 = 0; i0 < nb0; ++i0)
        for (int i1 = 0; i1 < nb1; ++i1)
            for (int i2 = 0; i2 < nb2; ++i2) {
                const int org[3] = {i0 * BX, i1 * BY, i2 * BZ};
                const int lo_v[3] = {org[0], org[1], org[2]};
                const int hi_v[3] = {std::min(org[0] + BX, D0), std::min(org[1] + BY, D1), std::min(org[2] + BZ, D2)};
                for (int x = 0; x < BX; ++x)
                    for (int y = 0; y < BY; ++y)
                        for (int z = 0; z < BZ; ++z) {
                            const int g0 = org[0] + x, g1 = org[1] + y, g2 = org[2] + z;
                            brick[((size_t)x * BY + y) * BZ + z] =
                                (g0 < D0 && g1 < D1 && g2 < D2) ? vol[((size_t)g0 * D1 + g1) * D2 + g2] : 0.0f;
                        }
                LdHost ld{brick.data()};
                for (int b = 0; b < B; ++b) {
                    float uv[16], umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY, dmin = INFINITY,
                                  dmax = -INFINITY;
                    bool bad = false;
                    for (int c = 0; c < 8; ++c) {
                        const float X[3] = {(float)((c & 1) ? hi_v[0] : lo_v[0]) - shift, (float)((c & 2) ? hi_v[1] : lo_v[1]) - shift,
                                            (float)((c & 4) ? hi_v[2] : lo_v[2]) - shift};
                        float u, v, den;
                        project_corner(geo[b], X, u, v, den);
                        bad = bad || u != u || v != v || den != den;
                        uv[2 * c] = u;
                        uv[2 * c + 1] = v;
                        umin = fminf(umin, u); umax = fmaxf(umax, u);
                        vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
                        dmin = fminf(dmin, den); dmax = fmaxf(dmax, den);
                    }
                    if (bad) dmin = NAN;
                    const PixRect rc = rect_from_extents(umin, umax, vmin, vmax, dmin, dmax, H, W);
                    const bool ok = outline_valid(umin, umax, vmin, vmax, dmin, dmax);
                    float clo[3], chi[3];
                    for (int a = 0; a < 3; ++a) {
                        clo[a] = ((float)lo_v[a] - shift) - geo[b].S[a];
                        chi[a] = ((float)hi_v[a] - shift) - geo[b].S[a];
                    }
                    std::fill(cand.begin(), cand.end(), 0);
                    long pend_r = -1;
                    if (rc.x0 <= rc.x1 && rc.y0 <= rc.y1) {
                        const int th = (rc.y1 - rc.y0) / 4 + 1;
                        for (int ty = 0; ty < th; ++ty) {
                            const int py0 = rc.y0 + 4 * ty;
                            int px_lo, px_hi;
                            if (!row_span(uv, ok, rc, py0, px_lo, px_hi)) continue;
                            const int cnt = (px_hi - px_lo) / 8 + 1;
                            for (int t = 0; t < cnt; ++t)
                                for (int lane = 0; lane < 32; ++lane) {
                                    const int px = px_lo + 8 * t + (lane & 7), py = py0 + (lane >> 3);
                                    if (px > px_hi || py > rc.y1) continue;
                                    ++st[0];
                                    cand[(size_t)py * W + px] = 1;
                                    const long r = (long)b * N + (long)py * W + px;
                                    const Ray& ray = rays[r];
                                    float a_in, a_out;
                                    if (!brick_maybe_hit(ray.inv, clo, chi, a_in, a_out)) continue;
                                    ++st[1];
                                    if (start_walk_box(ray, lo_v, hi_v, shift).hit) ++st[2];
                                    if (lean == 3) {  // two-rays-per-thread path: pair this hit with the previous pending one
                                        if (pend_r < 0) {
                                            pend_r = r;
                                            continue;
                                        }
                                        const Ray& rb = rays[pend_r];
                                        AccState wa, wb;
                                        AccConst ka, kb;
                                        brick_pair_setup_acc(ld, true, ray.s, ray.inv, clo, chi, lo_v, hi_v, org, BY * BZ, BZ, 1, shift, wa, ka);
                                        brick_pair_setup_acc(ld, true, rb.s, rb.inv, clo, chi, lo_v, hi_v, org, BY * BZ, BZ, 1, shift, wb, kb);
                                        float pa, pb;
                                        brick_pair2_walk<2>(ld, wa, ka, wb, kb, pa, pb);
                                        if (pa != 0.0f) { ++st[3]; out[r] += raylen[r] * pa; }
                                        if (pb != 0.0f) { ++st[3]; out[pend_r] += raylen[pend_r] * pb; }
                                        pend_r = -1;
                                        continue;
                                    }
                                    const float part =
                                        lean == 1 ? brick_pair_fwd_lean<4>(ld, ray.s, ray.inv, clo, chi, lo_v, hi_v, org, BY * BZ, BZ, 1, shift)
                                        : lean == 2 ? brick_pair_fwd_lean<4, LdHost, false>(ld, ray.s, ray.inv, clo, chi, lo_v, hi_v, org, BY * BZ, BZ, 1, shift)
                                             : brick_pair_fwd<4>(ld, ray, lo_v, hi_v, org, BY * BZ, BZ, 1, shift);
                                    if (part != 0.0f) {
                                        ++st[3];
                                        out[r] += raylen[r] * part;
                                    }
                                }
                        }
                    }
                    if (pend_r >= 0) {  // odd hit left over: paired with an empty slot
                        const Ray& rb = rays[pend_r];
                        AccState wa, wb;
                        AccConst ka, kb;
                        brick_pair_setup_acc(ld, false, rb.s, rb.inv, clo, chi, lo_v, hi_v, org, BY * BZ, BZ, 1, shift, wa, ka);
                        brick_pair_setup_acc(ld, true, rb.s, rb.inv, clo, chi, lo_v, hi_v, org, BY * BZ, BZ, 1, shift, wb, kb);
                        float pa, pb;
                        brick_pair2_walk<2>(ld, wa, ka, wb, kb, pa, pb);
                        if (pa != 0.0f) ++violations;  // an empty slot must contribute nothing
                        if (pb != 0.0f) { ++st[3]; out[pend_r] += raylen[pend_r] * pb; }
                    }
                    if (check)
                        for (int py = 0; py < H; ++py)
                            for (int px = 0; px < W; ++px) {
                                const Ray& ray = rays[(size_t)b * N + (long)py * W + px];
                                float a_in, a_out;
                                const bool maybe = brick_maybe_hit(ray.inv, clo, chi, a_in, a_out);
                                if (start_walk_box(ray, lo_v, hi_v, shift).hit && (!cand[(size_t)py * W + px] || !maybe)) ++violations;
                            }
                }
            }
    if (stats)
        for (int i = 0; i < 4; ++i) stats[i] = st[i];
    return violations;
}

// Volume gradient through the brick decomposition (siddon_brick.cu, BWD mode): per brick a zeroed accumulator brick, every
// candidate pair of every pose scatters (gout * raylen) * chord length into it (brick_pair_bwd_lean), then the brick is stored
// (clipped to the volume) -- each voxel belongs to exactly one brick, so g_vol is OVERWRITTEN.
void emu_siddon_bwd_vol_brick(int D0, int D1, int D2, const float* src, const float* tgt, const float* raylen, const float* gout,
                              float* g_vol, int B, int H, int W, float shift, float eps, int BX, int BY, int BZ)
{
    const long N = (long)H * W;
    std::vector<Ray> rays((size_t)B * N);
    std::vector<PoseGeo> geo(B);
    for (int b = 0; b < B; ++b) {
        for (long n = 0; n < N; ++n) rays[(size_t)b * N + n] = load_ray(src, tgt, b, (long)b * N + n, eps);
        const Ray& r00 = rays[(size_t)b * N];
        const Ray& r0w = rays[(size_t)b * N + (W - 1)];
        const Ray& rh0 = rays[(size_t)b * N + (long)(H - 1) * W];
        float t00[3], t0w[3], th0[3];
        for (int a = 0; a < 3; ++a) {
            t00[a] = r00.s[a] + r00.d[a];
            t0w[a] = r00.s[a] + r0w.d[a];
            th0[a] = r00.s[a] + rh0.d[a];
        }
        geo[b] = make_pose_geo(r00.s, t00, t0w, th0, H, W);
    }
    std::vector<float> brick((size_t)BX * BY * BZ);
    const int nb0 = (D0 + BX - 1) / BX, nb1 = (D1 + BY - 1) / BY, nb2 = (D2 + BZ - 1) / BZ;
    for (int i0 = 0; i0 < nb0; ++i0)
        for (int i1 = 0; i1 < nb1; ++i1)
            for (int i2 = 0; i2 < nb2; ++i2) {
                const int org[3] = {i0 * BX, i1 * BY, i2 * BZ};
                const int lo_v[3] = {org[0], org[1], org[2]};
                const int hi_v[3] = {std::min(org[0] + BX, D0), std::min(org[1] + BY, D1), std::min(org[2] + BZ, D2)};
                std::fill(brick.begin(), brick.end(), 0.0f);
                StHost st{brick.data()};
                for (int b = 0; b < B; ++b) {
                    float uv[16], umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY, dmin = INFINITY,
                                  dmax = -INFINITY;
                    bool bad = false;
                    for (int c = 0; c < 8; ++c) {
                        const float X[3] = {(float)((c & 1) ? hi_v[0] : lo_v[0]) - shift, (float)((c & 2) ? hi_v[1] : lo_v[1]) - shift,
                                            (float)((c & 4) ? hi_v[2] : lo_v[2]) - shift};
                        float u, v, den;
                        project_corner(geo[b], X, u, v, den);
                        bad = bad || u != u || v != v || den != den;
                        uv[2 * c] = u;
                        uv[2 * c + 1] = v;
                        umin = fminf(umin, u); umax = fmaxf(umax, u);
                        vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
                        dmin = fminf(dmin, den); dmax = fmaxf(dmax, den);
                    }
                    if (bad) dmin = NAN;
                    const PixRect rc = rect_from_extents(umin, umax, vmin, vmax, dmin, dmax, H, W);
                    const bool ok = outline_valid(umin, umax, vmin, vmax, dmin, dmax);
                    float clo[3], chi[3];
                    for (int a = 0; a < 3; ++a) {
                        clo[a] = ((float)lo_v[a] - shift) - geo[b].S[a];
                        chi[a] = ((float)hi_v[a] - shift) - geo[b].S[a];
                    }
                    if (!(rc.x0 <= rc.x1 && rc.y0 <= rc.y1)) continue;
                    const int th = (rc.y1 - rc.y0) / 4 + 1;
                    for (int ty = 0; ty < th; ++ty) {
                        const int py0 = rc.y0 + 4 * ty;
                        int px_lo, px_hi;
                        if (!row_span(uv, ok, rc, py0, px_lo, px_hi)) continue;
                        const int cnt = (px_hi - px_lo) / 8 + 1;
                        for (int t = 0; t < cnt; ++t)
                            for (int lane = 0; lane < 32; ++lane) {
                                const int px = px_lo + 8 * t + (lane & 7), py = py0 + (lane >> 3);
                                if (px > px_hi || py > rc.y1) continue;
                                const long r = (long)b * N + (long)py * W + px;
                                const Ray& ray = rays[r];
                                float a_in, a_out;
                                if (!brick_maybe_hit(ray.inv, clo, chi, a_in, a_out)) continue;
                                brick_pair_bwd_lean<2>(st, gout[r] * raylen[r], ray.s, ray.inv, clo, chi, lo_v, hi_v, org, BY * BZ, BZ,
                                                       1, shift);
                            }
                    }
                }
                for (int x = 0; x < hi_v[0] - org[0]; ++x)
                    for (int y = 0; y < hi_v[1] - org[1]; ++y)
                        for (int z = 0; z < hi_v[2] - org[2]; ++z)
                            g_vol[((size_t)(org[0] + x) * D1 + (org[1] + y)) * D2 + (org[2] + z)] = brick[((size_t)x * BY + y) * BZ + z];
            }
}

// Debug: per-brick partial sums of ONE ray through the exact and the lean pair paths (bricks in linear order).
long emu_brick_ray_debug(const float* vol, int D0, int D1, int D2, const float* src3, const float* tgt3, float shift, float eps,
                         int BX, int BY, int BZ, float* part_full, float* part_lean, int* brick_ids, int max_out)
{
    Ray ray;
    for (int a = 0; a < 3; ++a) {
        ray.s[a] = src3[a];
        ray.d[a] = (tgt3[a] - src3[a]) + eps;
        ray.inv[a] = 1.0f / ray.d[a];
    }
    std::vector<float> brick((size_t)BX * BY * BZ);
    const int nb0 = (D0 + BX - 1) / BX, nb1 = (D1 + BY - 1) / BY, nb2 = (D2 + BZ - 1) / BZ;
    long n = 0;
    for (int i0 = 0; i0 < nb0; ++i0)
        for (int i1 = 0; i1 < nb1; ++i1)
            for (int i2 = 0; i2 < nb2; ++i2) {
                const int org[3] = {i0 * BX, i1 * BY, i2 * BZ};
                const int lo_v[3] = {org[0], org[1], org[2]};
                const int hi_v[3] = {std::min(org[0] + BX, D0), std::min(org[1] + BY, D1), std::min(org[2] + BZ, D2)};
                float clo[3], chi[3];
                for (int a = 0; a < 3; ++a) {
                    clo[a] = ((float)lo_v[a] - shift) - ray.s[a];
                    chi[a] = ((float)hi_v[a] - shift) - ray.s[a];
                }
                float a_in, a_out;
                if (!brick_maybe_hit(ray.inv, clo, chi, a_in, a_out)) continue;
                for (int x = 0; x < BX; ++x)
                    for (int y = 0; y < BY; ++y)
                        for (int z = 0; z < BZ; ++z) {
                            const int g0 = org[0] + x, g1 = org[1] + y, g2 = org[2] + z;
                            brick[((size_t)x * BY + y) * BZ + z] =
                                (g0 < D0 && g1 < D1 && g2 < D2) ? vol[((size_t)g0 * D1 + g1) * D2 + g2] : 0.0f;
                        }
                LdHost ld{brick.data()};
                if (n < max_out) {
                    part_full[n] = brick_pair_fwd<4>(ld, ray, lo_v, hi_v, org, BY * BZ, BZ, 1, shift);
                    part_lean[n] = brick_pair_fwd_lean<4, LdHost, false>(ld, ray.s, ray.inv, clo, chi, lo_v, hi_v, org, BY * BZ, BZ, 1, shift);
                    brick_ids[n] = (i0 * nb1 + i1) * nb2 + i2;
                }
                ++n;
            }
    return n;
}
}  // extern "C"

// ---- access-pattern analysis (tuning aid): distinct 32-byte sectors / 128-byte lines per warp-wide gather ----------
// Emulates the lock-step lean walk of one warp = WX x WY pixel patch and counts, per walk step, how many distinct
// sectors and lines the 32 lanes touch.  strides (st0,st1,st2) describe the volume layout being evaluated.
#include <set>
// layout < 0: linear with strides (st0,st1,st2).  layout = 1: bricked, 128-B line = 4x4x2 voxels (axes 0,1,2), 32-B
// sector = 2x2x2.  layout = 2: line = 2x4x4, sector 2x2x2.  layout = 3: line = 4x2x4.  out[4] = distinct sectors
// touched by the warp over the whole item (a lower bound on its L1 fills).
static long brick_addr(int layout, int D1, int D2, int i0, int i1, int i2)
{
    int b0, b1, b2;  // log2 of the line extent per axis
    if (layout == 1) { b0 = 2; b1 = 2; b2 = 1; } else if (layout == 2) { b0 = 1; b1 = 2; b2 = 2; } else { b0 = 2; b1 = 1; b2 = 2; }
    const long n1 = (D1 + (1 << b1) - 1) >> b1, n2 = (D2 + (1 << b2) - 1) >> b2;
    const long line = (((long)(i0 >> b0)) * n1 + (i1 >> b1)) * n2 + (i2 >> b2);
    // sector = which 2x2x2 sub-brick inside the line; word = position inside the sector
    const int s0 = (i0 & ((1 << b0) - 1)) >> 1, s1 = (i1 & ((1 << b1) - 1)) >> 1, s2 = (i2 & ((1 << b2) - 1)) >> 1;
    const int sector = (s0 * (b1 > 1 ? 2 : 1) + s1) * (b2 > 1 ? 2 : 1) + s2;
    const int word = (i0 & 1) * 4 + (i1 & 1) * 2 + (i2 & 1);
    return line * 32 + sector * 8 + word;
}

extern "C" void emu_warp_sectors(int D0, int D1, int D2, const float* src, const float* tgt, int B, int H, int W,
                                 int WX, int WY, int slab, long st0, long st1, long st2, float shift, float eps,
                                 int sample_every, double* out /* [5]: steps, lane-visits, sectors, lines, item sectors */)
{
    const int layout = st2 < 0 ? (int)(-st2) : -1;
    double item_sectors = 0;
    const VolDims dims = mk(D0, D1, D2);
    double steps = 0, visits = 0, sectors = 0, lines = 0;
    const int n_slabs = slab > 0 ? (D0 + slab - 1) / slab : 1;
    long warp_id = 0;
    for (int b = 0; b < B; ++b)
        for (int ty = 0; ty + WY <= H; ty += WY)
            for (int tx = 0; tx + WX <= W; tx += WX, ++warp_id) {
                if (warp_id % sample_every) continue;
                for (int sl = 0; sl < n_slabs; ++sl) {
                    const int lo_v[3] = {slab > 0 ? sl * slab : 0, 0, 0};
                    const int hi_v[3] = {slab > 0 ? std::min(D0, (sl + 1) * slab) : D0, D1, D2};
                    const int n = WX * WY;
                    std::vector<Walk> w(n);
                    std::vector<float> acur(n);
                    std::vector<bool> live(n);
                    int alive = 0;
                    for (int l = 0; l < n; ++l) {
                        const long r = ((long)b * H + ty + l / WX) * W + tx + l % WX;
                        const Ray ray = load_ray(src, tgt, b, r, eps);
                        w[l] = start_walk_box(ray, lo_v, hi_v, shift);
                        live[l] = w[l].hit;
                        acur[l] = w[l].a_in;
                        alive += live[l];
                    }
                    std::set<long> item_sec;
                    while (alive > 0) {
                        std::set<long> sec, lin;
                        for (int l = 0; l < n; ++l) {
                            if (!live[l]) continue;
                            const long off = layout < 0 ? w[l].idx[0] * st0 + w[l].idx[1] * st1 + w[l].idx[2] * st2
                                                        : brick_addr(layout, D1, D2, w[l].idx[0], w[l].idx[1], w[l].idx[2]);
                            sec.insert(off >> 3);
                            item_sec.insert(off >> 3);
                            lin.insert(off >> 5);
                            visits += 1;
                            const float anext = fminf(fminf(w[l].an[0], w[l].an[1]), w[l].an[2]);
                            if (!(anext < w[l].a_out)) { live[l] = false; --alive; continue; }
                            for (int a = 0; a < 3; ++a)
                                if (w[l].an[a] == anext) {
                                    w[l].idx[a] += w[l].sti[a];
                                    w[l].nf[a] += 1.0f;
                                    w[l].an[a] = fmaf(w[l].nf[a], w[l].da[a], w[l].a0[a]);
                                }
                        }
                        steps += 1;
                        sectors += sec.size();
                        lines += lin.size();
                    }
                    item_sectors += item_sec.size();
                }
            }
    out[0] = steps; out[1] = visits; out[2] = sectors; out[3] = lines; out[4] = item_sectors;
}


// ---- tuning aid: the lean walk with the lanes of a warp re-synchronised every K planes of the (common) major axis ----
// K = 1 reproduces the request pattern of the plane-synchronous walk (1..3 requests per plane: all lanes, then the
// lanes with one minor crossing, then those with two).  out: [steps, lane-visits, sectors, lines]
extern "C" void emu_warp_sectors_sync(int D0, int D1, int D2, const float* src, const float* tgt, int B, int H, int W,
                                      int WX, int WY, int slab, int K, float shift, float eps, int sample_every, double* out)
{
    double steps = 0, visits = 0, sectors = 0, lines = 0;
    const int n_slabs = slab > 0 ? (D0 + slab - 1) / slab : 1;
    const long st0 = (long)D1 * D2, st1 = D2;
    long warp_id = 0;
    for (int b = 0; b < B; ++b)
        for (int ty = 0; ty + WY <= H; ty += WY)
            for (int tx = 0; tx + WX <= W; tx += WX, ++warp_id) {
                if (warp_id % sample_every) continue;
                for (int sl = 0; sl < n_slabs; ++sl) {
                    const int lo_v[3] = {slab > 0 ? sl * slab : 0, 0, 0};
                    const int hi_v[3] = {slab > 0 ? std::min(D0, (sl + 1) * slab) : D0, D1, D2};
                    const int n = WX * WY;
                    std::vector<Walk> w(n);
                    std::vector<bool> live(n);
                    std::vector<int> major(n);
                    int alive = 0;
                    for (int l = 0; l < n; ++l) {
                        const long r = ((long)b * H + ty + l / WX) * W + tx + l % WX;
                        const Ray ray = load_ray(src, tgt, b, r, eps);
                        w[l] = start_walk_box(ray, lo_v, hi_v, shift);
                        live[l] = w[l].hit;
                        alive += live[l];
                        const float a0 = fabsf(ray.d[0]), a1 = fabsf(ray.d[1]), a2 = fabsf(ray.d[2]);
                        major[l] = (a0 >= a1 && a0 >= a2) ? 0 : (a1 >= a2 ? 1 : 2);
                    }
                    while (alive > 0) {
                        // the chunk the warp is working on = that of the most lagging live lane
                        int m = -1, dir = 0;
                        bool uniform = K > 0;
                        for (int l = 0; l < n && uniform; ++l) {
                            if (!live[l]) continue;
                            if (m < 0) { m = major[l]; dir = w[l].sti[m]; }
                            else if (major[l] != m || w[l].sti[m] != dir) uniform = false;
                        }
                        int target = 0;
                        if (uniform) {
                            bool first = true;
                            for (int l = 0; l < n; ++l) {
                                if (!live[l]) continue;
                                const int c = w[l].idx[m] / K;
                                if (first || (dir > 0 ? c < target : c > target)) { target = c; first = false; }
                            }
                        }
                        std::set<long> sec, lin;
                        for (int l = 0; l < n; ++l) {
                            if (!live[l]) continue;
                            if (uniform && w[l].idx[m] / K != target) continue;  // waits for the others
                            const long off = w[l].idx[0] * st0 + w[l].idx[1] * st1 + w[l].idx[2];
                            sec.insert(off >> 3);
                            lin.insert(off >> 5);
                            visits += 1;
                            const float anext = fminf(fminf(w[l].an[0], w[l].an[1]), w[l].an[2]);
                            if (!(anext < w[l].a_out)) { live[l] = false; --alive; continue; }
                            for (int a = 0; a < 3; ++a)
                                if (w[l].an[a] == anext) {
                                    w[l].idx[a] += w[l].sti[a];
                                    w[l].nf[a] += 1.0f;
                                    w[l].an[a] = fmaf(w[l].nf[a], w[l].da[a], w[l].a0[a]);
                                }
                        }
                        steps += 1;
                        sectors += sec.size();
                        lines += lin.size();
                    }
                }
            }
    out[0] = steps; out[1] = visits; out[2] = sectors; out[3] = lines;
}

// ---- tuning aid: per-lane 16-byte chunk reuse along the MAJOR axis (transposed copy with the major axis fastest) ------
// For every warp-step counts the lanes that must load a new aligned 4-voxel chunk and the distinct 128-byte lines among
// those loads.  out: [steps, lane-visits, chunk loads, sum over steps of distinct lines among loading lanes]
extern "C" void emu_chunk_reuse(int D0, int D1, int D2, const float* src, const float* tgt, int B, int H, int W, int WX, int WY,
                                int slab, float shift, float eps, int sample_every, double* out)
{
    double steps = 0, visits = 0, loads = 0, lines = 0;
    const int n_slabs = slab > 0 ? (D0 + slab - 1) / slab : 1;
    long warp_id = 0;
    for (int b = 0; b < B; ++b)
        for (int ty = 0; ty + WY <= H; ty += WY)
            for (int tx = 0; tx + WX <= W; tx += WX, ++warp_id) {
                if (warp_id % sample_every) continue;
                for (int sl = 0; sl < n_slabs; ++sl) {
                    const int lo_v[3] = {slab > 0 ? sl * slab : 0, 0, 0};
                    const int hi_v[3] = {slab > 0 ? std::min(D0, (sl + 1) * slab) : D0, D1, D2};
                    const int n = WX * WY;
                    std::vector<Walk> w(n);
                    std::vector<bool> live(n);
                    std::vector<long> chunk(n, -1);
                    std::vector<int> major(n);
                    int alive = 0;
                    for (int l = 0; l < n; ++l) {
                        const long r = ((long)b * H + ty + l / WX) * W + tx + l % WX;
                        const Ray ray = load_ray(src, tgt, b, r, eps);
                        w[l] = start_walk_box(ray, lo_v, hi_v, shift);
                        live[l] = w[l].hit;
                        alive += live[l];
                        const float a0 = fabsf(ray.d[0]), a1 = fabsf(ray.d[1]), a2 = fabsf(ray.d[2]);
                        major[l] = (a0 >= a1 && a0 >= a2) ? 0 : (a1 >= a2 ? 1 : 2);
                    }
                    const int dims[3] = {D0, D1, D2};
                    while (alive > 0) {
                        std::set<long> lin;
                        for (int l = 0; l < n; ++l) {
                            if (!live[l]) continue;
                            const int m = major[l], u = (m + 1) % 3, v = (m + 2) % 3;
                            // transposed copy: major axis fastest, then v, then u
                            const long row = (long)w[l].idx[u] * dims[v] + w[l].idx[v];
                            const long off = row * dims[m] + w[l].idx[m];
                            visits += 1;
                            if ((off >> 2) != chunk[l]) { chunk[l] = off >> 2; loads += 1; lin.insert(off >> 5); }
                            const float anext = fminf(fminf(w[l].an[0], w[l].an[1]), w[l].an[2]);
                            if (!(anext < w[l].a_out)) { live[l] = false; --alive; continue; }
                            for (int a = 0; a < 3; ++a)
                                if (w[l].an[a] == anext) {
                                    w[l].idx[a] += w[l].sti[a];
                                    w[l].nf[a] += 1.0f;
                                    w[l].an[a] = fmaf(w[l].nf[a], w[l].da[a], w[l].a0[a]);
                                }
                        }
                        steps += 1;
                        lines += lin.size();
                    }
                }
            }
    out[0] = steps; out[1] = visits; out[2] = loads; out[3] = lines;
}

extern "C" {
// ncc.cu on the CPU: the same chunked moments (ncc_partial_kernel -> ncc_finalize_kernel) and closed-form gradient (ncc_bwd_kernel).
void emu_ncc_fwd(const float* x1, const float* x2, int B, int C, long N, float eps, float* stats, float* score)
{
    for (int b = 0; b < B; ++b) {
        double total = 0.0;
        for (int c = 0; c < C; ++c) {
            const long img = (long)b * C + c;
            NccSums acc = {0.0, 0.0, 0.0, 0.0, 0.0};
            for (long lo = 0; lo < N; lo += kNccChunk) {
                NccSums part = {0.0, 0.0, 0.0, 0.0, 0.0};
                for (long i = lo; i < std::min<long>(N, lo + kNccChunk); ++i) ncc_accumulate(part, x1[img * N + i], x2[img * N + i]);
                ncc_merge(acc, part);
            }
            const NccStats s = ncc_finalize(acc, N, eps);
            std::memcpy(stats + img * 8, &s, sizeof(s));
            total += (double)s.score;
        }
        score[b] = (float)(total / (double)C);
    }
}

void emu_ncc_bwd(const float* x1, const float* x2, const float* stats, const float* gscore, float* g_x1, float* g_x2, int B, int C,
                 long N)
{
    for (long img = 0; img < (long)B * C; ++img) {
        NccStats s;
        std::memcpy(&s, stats + img * 8, sizeof(s));
        const float k = gscore[img / C] / ((float)C * (float)N);
        for (long i = 0; i < N; ++i) ncc_grad(s, k, x1[img * N + i], x2[img * N + i], g_x1[img * N + i], g_x2[img * N + i]);
    }
}
}

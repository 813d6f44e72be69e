// hostemu.cpp -- TEST INFRASTRUCTURE: compiles the product's per-ray device math
// (diffdrr_b200/csrc/ray_math.cuh) with g++ and loops over rays on the CPU, so that the CPU-only build
// container can check the kernel logic against the oracle and the goldens before any GPU time is spent.
// Never linked into libb200drr.so and never imported by the product package.
#include <cstring>

#include "../../diffdrr_b200/csrc/ray_math.cuh"

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

}  // extern "C"

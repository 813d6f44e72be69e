// ncc_math.cuh -- per-image statistics and closed-form gradient of the zero-normalised cross correlation that drives the
// 2D/3D registration loop (host+device; the kernels of ncc.cu are thin wrappers, tests/hostemu compiles the same source).
//
// Replaces reference diffdrr/metrics.py:21-44 (NormalizedCrossCorrelation2d.forward / .norm with patch_size = None):
//   norm(x) = (x - mean_hw x) / sqrt(var_hw x + eps)   (population variance),   score[b] = mean_{c,h,w} norm(x1) norm(x2).
// The reference runs ~15 elementwise / reduction launches forward and ~25 backward through autograd.  Here the five moments
// { sum x1, sum x2, sum x1^2, sum x2^2, sum x1 x2 } of one image pair are accumulated in ONE pass (in double: the one-pass
// variance q/N - mu^2 cancels, fp64 keeps 16 digits of it), and the gradient is the closed form below -- one elementwise pass.
//
// With s = sqrt(var + eps), n = (x - mu)/s and L = mean(n1 n2):   dL/dx2_j = (n1_j - n2_j L) / (N s2)
// (d n2_i/d x2_j = (delta_ij - 1/N)/s2 - n2_i n2_j/(N s2);  the -1/N term multiplies sum_i n1_i = 0).  Symmetric for x1.
#pragma once

#include "common.cuh"

namespace b200drr {

constexpr int kNccChunk = 2048;  // elements of one (image pair, chunk) partial: 256 threads x 2 float4

struct NccSums {
    double s1, s2, q1, q2, p;
};

B200_HD void ncc_accumulate(NccSums& a, float x1, float x2)
{
    const double u = (double)x1, v = (double)x2;
    a.s1 += u;
    a.s2 += v;
    a.q1 += u * u;
    a.q2 += v * v;
    a.p += u * v;
}

B200_HD void ncc_merge(NccSums& a, const NccSums& b)
{
    a.s1 += b.s1;
    a.s2 += b.s2;
    a.q1 += b.q1;
    a.q2 += b.q2;
    a.p += b.p;
}

// stats of one image pair, as the backward pass reads them: { mu1, 1/s1, mu2, 1/s2, L = mean(n1 n2), 0, 0, 0 }
struct NccStats {
    float mu1, rs1, mu2, rs2, score, pad0, pad1, pad2;
};

B200_HD NccStats ncc_finalize(const NccSums& a, int64_t N, float eps)
{
    const double inv_n = 1.0 / (double)N;
    const double mu1 = a.s1 * inv_n, mu2 = a.s2 * inv_n;
    double v1 = a.q1 * inv_n - mu1 * mu1, v2 = a.q2 * inv_n - mu2 * mu2;
    v1 = v1 > 0.0 ? v1 : 0.0;  // a constant image: the cancellation may leave -1e-17
    v2 = v2 > 0.0 ? v2 : 0.0;
    const double rs1 = 1.0 / sqrt(v1 + (double)eps), rs2 = 1.0 / sqrt(v2 + (double)eps);
    NccStats s;
    s.mu1 = (float)mu1;
    s.rs1 = (float)rs1;
    s.mu2 = (float)mu2;
    s.rs2 = (float)rs2;
    s.score = (float)((a.p * inv_n - mu1 * mu2) * rs1 * rs2);
    s.pad0 = s.pad1 = s.pad2 = 0.0f;
    return s;
}

// k = gscore[b] / (C N).  Gradient of score[b] with respect to x2 (and x1) at one pixel.
B200_HD void ncc_grad(const NccStats& s, float k, float x1, float x2, float& g1, float& g2)
{
    const float n1 = (x1 - s.mu1) * s.rs1, n2 = (x2 - s.mu2) * s.rs2;
    g1 = k * s.rs1 * (n2 - n1 * s.score);
    g2 = k * s.rs2 * (n1 - n2 * s.score);
}

}  // namespace b200drr

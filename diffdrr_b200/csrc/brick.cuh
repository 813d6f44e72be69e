// brick.cuh -- per-(ray, brick) logic of the brick-major Siddon kernels (siddon_brick.cu).
//
// Brick-major scheduling: a CTA owns one BX x BY x BZ voxel brick, staged in shared memory by ONE TMA box copy
// (cp.async.bulk.tensor.3d, zero fill outside the volume), and integrates over it every ray of every pose of the batch
// whose line crosses it; the partial line integrals are combined with red.global.add.  A voxel fetched once from L2
// then serves ~0.4 visits per pose x all poses of the batch from shared memory, instead of one 32-byte sector per
// visit from L1/L2 (the limit of the slab-major kernels).
//
// Everything in this header is host+device code so tests/hostemu can run the same decomposition (brick loop, detector
// rectangle of a brick, hit test, per-pair walk) on the CPU against the oracle; the TMA / mbarrier / sorting machinery
// is device-only and lives in siddon_brick.cu.  Replaces reference renderers.py:94-113 (+156-169) on this path.
#pragma once

#include "ray_math.cuh"

namespace b200drr {

// Per-pose detector geometry in voxel-index space, derived from three corner rays of the full H x W grid (the grid is
// affine in the pixel indices: detector.py:105-132 builds it as a cartesian product).  For a point X, the line
// source->X meets the detector plane at pixel (u, v):  t = num / (n . (X - S)),  rel = (S - O) + t (X - S),
// u = rel . A,  v = rel . Bv.
struct PoseGeo {
    float S[3];   // source
    float n[3];   // eu x ev (plane normal, not normalised)
    float num;    // n . (O - S);  NaN marks a degenerate grid (H or W == 1): no rectangle can be derived
    float A[3];   // (ev x n) / |n|^2
    float Bv[3];  // (n x eu) / |n|^2
    float O[3];   // target of pixel (0, 0)
};

B200_HD void cross3(const float a[3], const float b[3], float c[3])
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// t00 / t0w / th0: targets of pixels (row 0, col 0), (row 0, col W-1), (row H-1, col 0)
B200_HD PoseGeo make_pose_geo(const float s[3], const float t00[3], const float t0w[3], const float th0[3], int H, int W)
{
    PoseGeo g;
    float eu[3], ev[3], os[3];
    const float iu = W > 1 ? 1.0f / (float)(W - 1) : 0.0f, iv = H > 1 ? 1.0f / (float)(H - 1) : 0.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        eu[a] = (t0w[a] - t00[a]) * iu;
        ev[a] = (th0[a] - t00[a]) * iv;
        g.S[a] = s[a];
        g.O[a] = t00[a];
        os[a] = t00[a] - s[a];
    }
    cross3(eu, ev, g.n);
    const float nn = g.n[0] * g.n[0] + g.n[1] * g.n[1] + g.n[2] * g.n[2];
    g.num = g.n[0] * os[0] + g.n[1] * os[1] + g.n[2] * os[2];
    float en[3], ne[3];
    cross3(ev, g.n, en);
    cross3(g.n, eu, ne);
    const float inn = 1.0f / nn;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        g.A[a] = en[a] * inn;
        g.Bv[a] = ne[a] * inn;
    }
    if (!(nn > 0.0f) || H < 2 || W < 2) g.num = NAN;
    return g;
}

// Projection of one brick corner: pixel coordinates (u, v) and the denominator n . (X - S).
B200_HD void project_corner(const PoseGeo& g, const float X[3], float& u, float& v, float& den)
{
    float w[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) w[a] = X[a] - g.S[a];
    den = g.n[0] * w[0] + g.n[1] * w[1] + g.n[2] * w[2];
    const float t = g.num / den;
    float rel[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) rel[a] = (g.S[a] - g.O[a]) + t * w[a];
    u = rel[0] * g.A[0] + rel[1] * g.A[1] + rel[2] * g.A[2];
    v = rel[0] * g.Bv[0] + rel[1] * g.Bv[1] + rel[2] * g.Bv[2];
}

struct PixRect {
    int x0, y0, x1, y1;  // inclusive; x0 > x1 = empty
};

// Pixel rectangle that certainly contains every pixel whose LINE (quirk Q1: Siddon integrates the infinite line,
// renderers.py:97-113) meets the box: bounding box of the 8 projected corners +- 1 pixel when the box stays on one
// side of the plane through the source parallel to the detector (a projective map keeps a convex set convex there);
// otherwise -- or when a corner comes close to that plane -- the whole detector.
B200_HD PixRect rect_from_extents(float umin, float umax, float vmin, float vmax, float dmin, float dmax, int H, int W)
{
    PixRect r;
    const float lo = fminf(fabsf(dmin), fabsf(dmax)), hi = fmaxf(fabsf(dmin), fabsf(dmax));
    const bool one_side = (dmin > 0.0f) || (dmax < 0.0f);
    const bool ok = one_side && (lo >= 1e-3f * hi) && (umin <= umax) && (vmin <= vmax);  // false for NaN as well
    if (!ok) {
        r.x0 = 0; r.y0 = 0; r.x1 = W - 1; r.y1 = H - 1;
        return r;
    }
    const float fx0 = floorf(umin) - 1.0f, fx1 = ceilf(umax) + 1.0f, fy0 = floorf(vmin) - 1.0f, fy1 = ceilf(vmax) + 1.0f;
    if (fx1 < 0.0f || fy1 < 0.0f || fx0 > (float)(W - 1) || fy0 > (float)(H - 1)) {
        r.x0 = 1; r.x1 = 0; r.y0 = 1; r.y1 = 0;  // projects off the detector
        return r;
    }
    r.x0 = (int)fmaxf(fx0, 0.0f);
    r.y0 = (int)fmaxf(fy0, 0.0f);
    r.x1 = (int)fminf(fx1, (float)(W - 1));
    r.y1 = (int)fminf(fy1, (float)(H - 1));
    return r;
}

B200_HD PixRect brick_rect(const PoseGeo& g, const int lo_v[3], const int hi_v[3], float shift, int H, int W)
{
    float umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY, dmin = INFINITY, dmax = -INFINITY;
    for (int c = 0; c < 8; ++c) {
        const float X[3] = {(float)((c & 1) ? hi_v[0] : lo_v[0]) - shift, (float)((c & 2) ? hi_v[1] : lo_v[1]) - shift,
                            (float)((c & 4) ? hi_v[2] : lo_v[2]) - shift};
        float u, v, den;
        project_corner(g, X, u, v, den);
        // NaN-propagating min / max: a NaN corner must invalidate the rectangle (fminf would drop it)
        umin = (u < umin || u != u) ? u : umin;
        umax = (u > umax || u != u) ? u : umax;
        vmin = (v < vmin || v != v) ? v : vmin;
        vmax = (v > vmax || v != v) ? v : vmax;
        dmin = (den < dmin || den != den) ? den : dmin;
        dmax = (den > dmax || den != den) ? den : dmax;
    }
    return rect_from_extents(umin, umax, vmin, vmax, dmin, dmax, H, W);
}

// Conservative (ray, box) test from the ray-table entry: the same expressions as box_surely_missed with the
// per-(pose, box) constants clo = (lo - shift) - s, chi = (hi - shift) - s hoisted.  Returns false only when the line
// certainly misses; a_in / a_out feed the step estimate the work list is sorted by.
B200_HD bool brick_maybe_hit(const float inv[3], const float clo[3], const float chi[3], float& a_in, float& a_out)
{
    a_in = -INFINITY;
    a_out = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float a0 = clo[a] * inv[a], a1 = chi[a] * inv[a];
        a_in = fmaxf(a_in, fminf(a0, a1));
        a_out = fminf(a_out, fmaxf(a0, a1));
    }
    const float margin = 1e-4f * (1.0f + fabsf(a_in) + fabsf(a_out));
    return !(a_in > a_out + margin);
}

// Work item: bin (5 bits) | pose within the chunk (5 bits) | row (11 bits) | column (11 bits)
constexpr int kBrickPoseChunk = 32;
constexpr int kBrickMaxSide = 2048;
constexpr int kBrickBins = 32;
B200_HD unsigned pack_item(int bin, int bl, int py, int px) { return ((unsigned)bin << 27) | ((unsigned)bl << 22) | ((unsigned)py << 11) | (unsigned)px; }
B200_HD int item_bin(unsigned it) { return (int)(it >> 27); }
B200_HD int item_pose(unsigned it) { return (int)((it >> 22) & 31u); }
B200_HD int item_row(unsigned it) { return (int)((it >> 11) & 2047u); }
B200_HD int item_col(unsigned it) { return (int)(it & 2047u); }

// Estimated number of voxels a hit visits inside the box -> sort bin (longest first keeps the lanes of a warp alike)
B200_HD int step_bin(float a_in, float a_out, float sumabs_d, float inv_width)
{
    float est = (a_out - a_in) * sumabs_d;      // L1 length of the chord in voxels
    est = fminf(fmaxf(est, 0.0f), 4096.0f);     // NaN -> 0
    const int b = (int)(est * inv_width);
    return b < kBrickBins - 1 ? b : kBrickBins - 1;
}

// Host emulation loader: the brick is a plain array, offsets are element offsets.
struct LdHost {
    const float* brick;
    static constexpr int kScale = 1;
    B200_HD int base() const { return 0; }
    B200_HD float operator()(int off) const { return brick[off]; }
};

// Line integral (sum of v * dalpha) of one ray over the voxels [lo_v, hi_v) of a brick whose voxel (org) sits at element
// 0 of the staged copy with element strides (st0, st1, st2).  Same walk, same order, same rounding as
// siddon_ray_lean_box; only the loads differ (shared memory instead of global).
template <int U, class Ld>
B200_HD float brick_pair_fwd(const Ld& ld, const Ray& ray, const int lo_v[3], const int hi_v[3], const int org[3], int st0,
                             int st1, int st2, float shift)
{
    const Walk w = start_walk_box(ray, lo_v, hi_v, shift);
    if (!w.hit) return 0.0f;
    LeanConst k;
    LeanState s;
    lean_init(w, st0 * Ld::kScale, st1 * Ld::kScale, st2 * Ld::kScale, s, k);
    s.off += ld.base() - (org[0] * st0 + org[1] * st1 + org[2] * st2) * Ld::kScale;
    float acc = 0.0f;
    while (s.acur < k.a_out) {
        float len[U], v[U];
        int offs[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            offs[j] = s.off;
            len[j] = lean_step(s, k);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) v[j] = ld(offs[j]);
#pragma unroll
        for (int j = 0; j < U; ++j) acc = fmaf(len[j], v[j], acc);
    }
    return acc;
}

}  // namespace b200drr

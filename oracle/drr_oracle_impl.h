/*
 * drr_oracle_impl.h -- CPU restatement of DiffDRR's renderer hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Included twice by drr_oracle.c, once with REAL=float (SUF=f32) and once with REAL=double (SUF=f64).
 * Each function cites the reference lines it restates (paths relative to /root/reference/diffdrr/).
 * The algorithm of the gather itself lives in a third-party dependency that is not vendored in the
 * reference: torch.nn.functional.grid_sample (ATen grid_sampler_3d, installed torch 2.11.0; the
 * reference pins only `torch`, pyproject.toml:15).  Its published semantics are restated here:
 *   un-normalise (align_corners=False):  pix = ((g + 1) * size - 1) / 2
 *   un-normalise (align_corners=True):   pix = (g + 1) / 2 * (size - 1)
 *   mode="nearest":  index = nearbyint(pix)  (round-half-to-even), zero outside [0, size)
 *   mode="bilinear": 8-corner lerp from floor(pix), every corner zero-padded independently
 *   grid component 0 addresses the LAST dim of the sampled tensor; the reference samples
 *   volume.permute(2,1,0) (renderers.py:160), so component a addresses volume axis a.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call
 * this code, and only as the checker / the reported CPU baseline -- never as the product path.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* ---- ATen grid_sampler un-normalisation ------------------------------------------------------- */
static inline REAL FN(unnormalize)(REAL g, int size, int align_corners)
{
    if (align_corners) return ((g + (REAL)1) / (REAL)2) * (REAL)(size - 1);
    return ((g + (REAL)1) * (REAL)size - (REAL)1) / (REAL)2;
}

/* renderers.py:143-153 (_get_xyzs): x = s + alpha * (t - s + eps), then 2*(x + shift)/dims - 1,
 * followed by ATen's un-normalisation: returns the continuous voxel coordinate `pix` on axis a. */
static inline REAL FN(pix_at)(REAL alpha, REAL s, REAL d, REAL shift, int size, int align_corners)
{
    REAL x = s + alpha * d;
    REAL g = (REAL)2 * (x + shift) / (REAL)size - (REAL)1;
    return FN(unnormalize)(g, size, align_corners);
}

/* renderers.py:94-113 (_get_alphas): alpha_a[i] = ((i - shift) - s_a) / (t_a - s_a + eps), i = 0..D_a,
 * concatenated over the three axes and sorted ascending.  Each per-axis sequence is monotone in i, so
 * the sort is done as a 3-way merge; `axis`/`plane` (optional) record where every alpha came from. */
static void FN(sorted_alphas)(const int dims[3], const REAL s[3], const REAL d[3], REAL shift,
                              REAL *alpha, int *axis, int *plane)
{
    int pos[3], step[3], left[3];
    for (int a = 0; a < 3; ++a) {
        /* ascending in i when d > 0, descending when d < 0: walk the planes in ascending-alpha order */
        if (d[a] > 0) { pos[a] = 0; step[a] = 1; } else { pos[a] = dims[a]; step[a] = -1; }
        left[a] = dims[a] + 1;
    }
    int M = dims[0] + dims[1] + dims[2] + 3;
    REAL head[3];
    for (int a = 0; a < 3; ++a) head[a] = (((REAL)pos[a] - shift) - s[a]) / d[a];
    for (int m = 0; m < M; ++m) {
        int best = -1;
        for (int a = 0; a < 3; ++a)
            if (left[a] > 0 && (best < 0 || head[a] < head[best])) best = a;
        alpha[m] = head[best];
        if (axis) axis[m] = best;
        if (plane) plane[m] = pos[best];
        pos[best] += step[best];
        if (--left[best] > 0) head[best] = (((REAL)pos[best] - shift) - s[best]) / d[best];
    }
}

/* Nearest-voxel lookup at parameter `amid` (renderers.py:156-164 with mode="nearest").
 * Returns the flat index into vol (axis 2 fastest) or -1 for zero padding. */
static inline long FN(nearest_index)(REAL amid, const REAL s[3], const REAL d[3], REAL shift,
                                     const int dims[3], int align_corners)
{
    long idx[3];
    for (int a = 0; a < 3; ++a) {
        REAL pix = FN(pix_at)(amid, s[a], d[a], shift, dims[a], align_corners);
        REAL r = (REAL)nearbyint((double)pix);
        if (!(r >= 0 && r < (REAL)dims[a])) return -1;
        idx[a] = (long)r;
    }
    return (idx[0] * dims[1] + idx[1]) * (long)dims[2] + idx[2];
}

/* renderers.py:34-76 (Siddon.forward, mask=None): out[b,n] = reduce_j (L * v_j) * (alpha_{j+1} - alpha_j).
 * reduce: 0 = sum, 1 = max (renderers.py:175-183). */
void FN(oracle_siddon_fwd)(const REAL *vol, int D0, int D1, int D2, const REAL *src, const REAL *tgt,
                           const REAL *raylen, REAL *out, int B, long N, REAL shift, REAL eps, int reduce,
                           int align_corners)
{
    const int dims[3] = {D0, D1, D2};
    const int M = D0 + D1 + D2 + 3;
#pragma omp parallel
    {
        REAL *alpha = (REAL *)malloc(sizeof(REAL) * (size_t)M);
#pragma omp for schedule(dynamic, 64)
        for (long r = 0; r < (long)B * N; ++r) {
            const int b = (int)(r / N);
            REAL s[3], d[3];
            for (int a = 0; a < 3; ++a) {
                s[a] = src[b * 3 + a];
                d[a] = (tgt[r * 3 + a] - s[a]) + eps; /* renderers.py:104-106 */
            }
            FN(sorted_alphas)(dims, s, d, shift, alpha, NULL, NULL);
            const REAL L = raylen[r];
            REAL acc = 0;
            int first = 1;
            for (int j = 0; j + 1 < M; ++j) {
                REAL amid = (alpha[j] + alpha[j + 1]) / (REAL)2; /* renderers.py:57 */
                long idx = FN(nearest_index)(amid, s, d, shift, dims, align_corners);
                REAL v = idx < 0 ? (REAL)0 : vol[idx];
                REAL term = (L * v) * (alpha[j + 1] - alpha[j]); /* renderers.py:166,70-71 */
                if (reduce == 0) acc += term;
                else if (first || term > acc) acc = term;
                first = 0;
            }
            out[r] = acc;
        }
        free(alpha);
    }
}

/* renderers.py:77-89 (Siddon.forward with a label mask, "mask_to_channels"): every segment's term is routed to the
 * channel given by the nearest-sampled label at the same midpoint (zero padding -> label 0): out [B][C][N]. */
void FN(oracle_siddon_fwd_mask)(const REAL *vol, const REAL *mask, int D0, int D1, int D2, const REAL *src,
                                const REAL *tgt, const REAL *raylen, REAL *out, int B, long N, int C, REAL shift,
                                REAL eps, int align_corners)
{
    const int dims[3] = {D0, D1, D2};
    const int M = D0 + D1 + D2 + 3;
    memset(out, 0, sizeof(REAL) * (size_t)B * C * N);
#pragma omp parallel
    {
        REAL *alpha = (REAL *)malloc(sizeof(REAL) * (size_t)M);
#pragma omp for schedule(dynamic, 64)
        for (long r = 0; r < (long)B * N; ++r) {
            const int b = (int)(r / N);
            const long n = r % N;
            REAL s[3], d[3];
            for (int a = 0; a < 3; ++a) {
                s[a] = src[b * 3 + a];
                d[a] = (tgt[r * 3 + a] - s[a]) + eps;
            }
            FN(sorted_alphas)(dims, s, d, shift, alpha, NULL, NULL);
            const REAL L = raylen[r];
            for (int j = 0; j + 1 < M; ++j) {
                REAL amid = (alpha[j] + alpha[j + 1]) / (REAL)2;
                long idx = FN(nearest_index)(amid, s, d, shift, dims, align_corners);
                if (idx < 0) continue; /* density 0, label 0: adds 0 to channel 0 */
                long c = (long)mask[idx];
                if (c < 0 || c >= C) continue;
                out[((long)b * C + c) * N + n] += (L * vol[idx]) * (alpha[j + 1] - alpha[j]);
            }
        }
        free(alpha);
    }
}

/* Autograd of Siddon.forward restated in closed form (checked against the
 * reference's own autograd through the tests/golden siddon fixtures):
 *   dI/dalpha_m = L * (v_{m-1} - v_m)            (v_{-1} = v_{M-1} = 0; nearest sampling has no
 *                                                 gradient w.r.t. the sample position)
 *   alpha from axis a: dalpha/ds_a = (alpha - 1)/d_a,  dalpha/dt_a = -alpha/d_a
 *   dI/dL = sum_j v_j * len_j,   dI/dV[voxel_j] += L * len_j
 * stop_grad != 0 restates stop_gradients_through_grid_sample=True (renderers.py:63-65): the gather and
 * the "* L" einsum are constants, so g_vol and g_raylen receive nothing.
 * g_src [B,3], g_tgt [B,N,3], g_raylen [B,N] are overwritten; g_vol [D0,D1,D2] is accumulated into
 * (caller zero-fills); any of them may be NULL. reduce must be 0 (sum). */
void FN(oracle_siddon_bwd)(const REAL *vol, int D0, int D1, int D2, const REAL *src, const REAL *tgt,
                           const REAL *raylen, const REAL *gout, REAL *g_src, REAL *g_tgt, REAL *g_raylen,
                           REAL *g_vol, int B, long N, REAL shift, REAL eps, int stop_grad, int align_corners)
{
    const int dims[3] = {D0, D1, D2};
    const int M = D0 + D1 + D2 + 3;
    if (g_src) memset(g_src, 0, sizeof(REAL) * (size_t)B * 3);
#pragma omp parallel
    {
        REAL *alpha = (REAL *)malloc(sizeof(REAL) * (size_t)M);
        int *axis = (int *)malloc(sizeof(int) * (size_t)M);
        REAL *v = (REAL *)malloc(sizeof(REAL) * (size_t)(M + 1));
#pragma omp for schedule(dynamic, 64)
        for (long r = 0; r < (long)B * N; ++r) {
            const int b = (int)(r / N);
            REAL s[3], d[3];
            for (int a = 0; a < 3; ++a) {
                s[a] = src[b * 3 + a];
                d[a] = (tgt[r * 3 + a] - s[a]) + eps;
            }
            FN(sorted_alphas)(dims, s, d, shift, alpha, axis, NULL);
            const REAL L = raylen[r], g = gout[r];
            REAL gL = 0;
            for (int j = 0; j + 1 < M; ++j) {
                REAL amid = (alpha[j] + alpha[j + 1]) / (REAL)2;
                long idx = FN(nearest_index)(amid, s, d, shift, dims, align_corners);
                v[j] = idx < 0 ? (REAL)0 : vol[idx];
                REAL len = alpha[j + 1] - alpha[j];
                gL += v[j] * len;
                if (g_vol && !stop_grad && idx >= 0) {
                    REAL add = g * L * len;
#pragma omp atomic
                    g_vol[idx] += add;
                }
            }
            REAL gs[3] = {0, 0, 0}, gt[3] = {0, 0, 0};
            for (int m = 0; m < M; ++m) {
                REAL vm1 = m > 0 ? v[m - 1] : (REAL)0;
                REAL vm = m + 1 < M ? v[m] : (REAL)0;
                REAL c = g * L * (vm1 - vm);
                int a = axis[m];
                gs[a] += c * (alpha[m] - (REAL)1) / d[a];
                gt[a] += c * (-alpha[m]) / d[a];
            }
            for (int a = 0; a < 3; ++a) {
                if (g_tgt) g_tgt[r * 3 + a] = gt[a];
                if (g_src) {
#pragma omp atomic
                    g_src[b * 3 + a] += gs[a];
                }
            }
            if (g_raylen) g_raylen[r] = stop_grad ? (REAL)0 : g * gL;
        }
        free(alpha);
        free(axis);
        free(v);
    }
}

/* Autograd of Siddon.forward WITH a label mask (renderers.py:77-89): the scatter_add_ routes segment j to channel
 * c_j = label at the segment's voxel, so the upstream gradient of segment j is g_j = gout[b][c_j][n] and everything
 * above holds with v_j replaced by g_j v_j and g = 1:
 *   dLoss/dalpha_m = L (g_{m-1} v_{m-1} - g_m v_m),  dLoss/dL = sum_j g_j v_j len_j,  dLoss/dV[voxel_j] += g_j L len_j.
 * gout is [B][C][N]; labels outside [0, C) carry no gradient (the forward drops them). */
void FN(oracle_siddon_bwd_mask)(const REAL *vol, const REAL *mask, int D0, int D1, int D2, const REAL *src, const REAL *tgt,
                                const REAL *raylen, const REAL *gout, REAL *g_src, REAL *g_tgt, REAL *g_raylen,
                                REAL *g_vol, int B, long N, int C, REAL shift, REAL eps, int stop_grad, int align_corners)
{
    const int dims[3] = {D0, D1, D2};
    const int M = D0 + D1 + D2 + 3;
    if (g_src) memset(g_src, 0, sizeof(REAL) * (size_t)B * 3);
#pragma omp parallel
    {
        REAL *alpha = (REAL *)malloc(sizeof(REAL) * (size_t)M);
        int *axis = (int *)malloc(sizeof(int) * (size_t)M);
        REAL *v = (REAL *)malloc(sizeof(REAL) * (size_t)(M + 1));
#pragma omp for schedule(dynamic, 64)
        for (long r = 0; r < (long)B * N; ++r) {
            const int b = (int)(r / N);
            const long n = r % N;
            REAL s[3], d[3];
            for (int a = 0; a < 3; ++a) {
                s[a] = src[b * 3 + a];
                d[a] = (tgt[r * 3 + a] - s[a]) + eps;
            }
            FN(sorted_alphas)(dims, s, d, shift, alpha, axis, NULL);
            const REAL L = raylen[r];
            REAL gL = 0;
            for (int j = 0; j + 1 < M; ++j) {
                REAL amid = (alpha[j] + alpha[j + 1]) / (REAL)2;
                long idx = FN(nearest_index)(amid, s, d, shift, dims, align_corners);
                REAL gj = 0;
                if (idx >= 0) {
                    long c = (long)mask[idx];
                    if (c >= 0 && c < C) gj = gout[((long)b * C + c) * N + n];
                }
                v[j] = idx < 0 ? (REAL)0 : gj * vol[idx];
                REAL len = alpha[j + 1] - alpha[j];
                gL += v[j] * len;
                if (g_vol && !stop_grad && idx >= 0) {
                    REAL add = gj * L * len;
#pragma omp atomic
                    g_vol[idx] += add;
                }
            }
            REAL gs[3] = {0, 0, 0}, gt[3] = {0, 0, 0};
            for (int m = 0; m < M; ++m) {
                REAL vm1 = m > 0 ? v[m - 1] : (REAL)0;
                REAL vm = m + 1 < M ? v[m] : (REAL)0;
                REAL c = L * (vm1 - vm);
                int a = axis[m];
                gs[a] += c * (alpha[m] - (REAL)1) / d[a];
                gt[a] += c * (-alpha[m]) / d[a];
            }
            for (int a = 0; a < 3; ++a) {
                if (g_tgt) g_tgt[r * 3 + a] = gt[a];
                if (g_src) {
#pragma omp atomic
                    g_src[b * 3 + a] += gs[a];
                }
            }
            if (g_raylen) g_raylen[r] = stop_grad ? (REAL)0 : gL;
        }
        free(alpha);
        free(axis);
        free(v);
    }
}

/* Autograd of Siddon.forward with reducefn="max" (renderers.py:175-183: img.max(dim=-1).values): only the FIRST maximal
 * segment j* carries gradient (torch.max returns the first maximal index):
 *   dI/dalpha_{j*+1} = +L v,  dI/dalpha_{j*} = -L v,  dI/dL = v len,  dI/dV[voxel] = L len   (v, len of segment j*). */
void FN(oracle_siddon_bwd_max)(const REAL *vol, int D0, int D1, int D2, const REAL *src, const REAL *tgt,
                               const REAL *raylen, const REAL *gout, REAL *g_src, REAL *g_tgt, REAL *g_raylen,
                               REAL *g_vol, int B, long N, REAL shift, REAL eps, int stop_grad, int align_corners)
{
    const int dims[3] = {D0, D1, D2};
    const int M = D0 + D1 + D2 + 3;
    if (g_src) memset(g_src, 0, sizeof(REAL) * (size_t)B * 3);
#pragma omp parallel
    {
        REAL *alpha = (REAL *)malloc(sizeof(REAL) * (size_t)M);
        int *axis = (int *)malloc(sizeof(int) * (size_t)M);
#pragma omp for schedule(dynamic, 64)
        for (long r = 0; r < (long)B * N; ++r) {
            const int b = (int)(r / N);
            REAL s[3], d[3];
            for (int a = 0; a < 3; ++a) {
                s[a] = src[b * 3 + a];
                d[a] = (tgt[r * 3 + a] - s[a]) + eps;
            }
            FN(sorted_alphas)(dims, s, d, shift, alpha, axis, NULL);
            const REAL L = raylen[r], g = gout[r];
            int jbest = 0;
            long ibest = -1;
            REAL tbest = 0, vbest = 0;
            for (int j = 0; j + 1 < M; ++j) {
                REAL amid = (alpha[j] + alpha[j + 1]) / (REAL)2;
                long idx = FN(nearest_index)(amid, s, d, shift, dims, align_corners);
                REAL v = idx < 0 ? (REAL)0 : vol[idx];
                REAL term = (L * v) * (alpha[j + 1] - alpha[j]);
                if (j == 0 || term > tbest) {
                    tbest = term;
                    jbest = j;
                    vbest = v;
                    ibest = idx;
                }
            }
            const REAL len = alpha[jbest + 1] - alpha[jbest];
            REAL gs[3] = {0, 0, 0}, gt[3] = {0, 0, 0};
            for (int e = 0; e < 2; ++e) {
                const int m = jbest + e;
                const REAL c = (e ? (REAL)1 : (REAL)-1) * g * L * vbest;
                const int a = axis[m];
                gs[a] += c * (alpha[m] - (REAL)1) / d[a];
                gt[a] += c * (-alpha[m]) / d[a];
            }
            for (int a = 0; a < 3; ++a) {
                if (g_tgt) g_tgt[r * 3 + a] = gt[a];
                if (g_src) {
#pragma omp atomic
                    g_src[b * 3 + a] += gs[a];
                }
            }
            if (g_raylen) g_raylen[r] = stop_grad ? (REAL)0 : g * vbest * len;
            if (g_vol && !stop_grad && ibest >= 0) {
                REAL add = g * L * len;
#pragma omp atomic
                g_vol[ibest] += add;
            }
        }
        free(alpha);
        free(axis);
    }
}

/* renderers.py:124-140 (_get_alpha_minmax) followed by the batch-global .min()/.max() of
 * renderers.py:221-223.  Far plane is dims + 1 - shift (quirk Q4). */
void FN(oracle_alpha_minmax)(const REAL *src, const REAL *tgt, int D0, int D1, int D2, int B, long N,
                             REAL shift, REAL eps, REAL *amin_out, REAL *amax_out)
{
    const int dims[3] = {D0, D1, D2};
    REAL gmin = (REAL)INFINITY, gmax = -(REAL)INFINITY;
    for (long r = 0; r < (long)B * N; ++r) {
        const int b = (int)(r / N);
        REAL amin = -(REAL)INFINITY, amax = (REAL)INFINITY;
        for (int a = 0; a < 3; ++a) {
            REAL s = src[b * 3 + a];
            REAL d = (tgt[r * 3 + a] - s) + eps;
            REAL a0 = (((REAL)0 - shift) - s) / d;
            REAL a1 = (((REAL)(dims[a] + 1) - shift) - s) / d;
            REAL lo = a0 < a1 ? a0 : a1, hi = a0 < a1 ? a1 : a0;
            if (lo > amin) amin = lo;
            if (hi < amax) amax = hi;
        }
        if (amin < 0) amin = 0;
        if (amax > 1) amax = 1;
        if (amin < gmin) gmin = amin;
        if (amax > gmax) gmax = amax;
    }
    *amin_out = gmin;
    *amax_out = gmax;
}

/* torch.linspace(0, 1, P)[m] as ATen computes it (symmetric about the midpoint).  renderers.py:224 builds it
 * in the default dtype (fp32) and only then casts `.to(volume)`, so the fp64 run sees fp32-rounded values. */
static inline REAL FN(linspace01)(int m, int P)
{
    float step = 1.0f / (float)(P - 1);
    if (m < P / 2) return (REAL)(step * (float)m);
    return (REAL)fmaf(-step, (float)(P - 1 - m), 1.0f); /* ATen's vectorised kernel fuses end - step*k */
}

/* ATen grid_sampler_3d, mode="bilinear", padding_mode="zeros": value and (optionally) the analytic
 * gradient w.r.t. the continuous voxel coordinate; every corner is zero-padded on its own. */
static inline REAL FN(trilerp)(const REAL *vol, const int dims[3], const REAL pix[3], REAL grad[3],
                               long corner_idx[8], REAL corner_w[8])
{
    REAL f[3];
    long i0[3];
    for (int a = 0; a < 3; ++a) {
        REAL fl = (REAL)floor((double)pix[a]);
        i0[a] = (long)fl;
        f[a] = pix[a] - fl;
    }
    REAL val = 0;
    if (grad) grad[0] = grad[1] = grad[2] = 0;
    for (int c = 0; c < 8; ++c) {
        int o[3] = {c & 1, (c >> 1) & 1, (c >> 2) & 1};
        REAL w[3];
        long id[3];
        int inb = 1;
        for (int a = 0; a < 3; ++a) {
            w[a] = o[a] ? f[a] : (REAL)1 - f[a];
            id[a] = i0[a] + o[a];
            if (id[a] < 0 || id[a] >= dims[a]) inb = 0;
        }
        long flat = inb ? (id[0] * dims[1] + id[1]) * (long)dims[2] + id[2] : -1;
        REAL v = inb ? vol[flat] : (REAL)0;
        val += v * w[0] * w[1] * w[2];
        if (grad) {
            grad[0] += v * (o[0] ? (REAL)1 : (REAL)-1) * w[1] * w[2];
            grad[1] += v * w[0] * (o[1] ? (REAL)1 : (REAL)-1) * w[2];
            grad[2] += v * w[0] * w[1] * (o[2] ? (REAL)1 : (REAL)-1);
        }
        if (corner_idx) { corner_idx[c] = flat; corner_w[c] = w[0] * w[1] * w[2]; }
    }
    return val;
}

/* renderers.py:205-240 (Trilinear.forward, mask=None): alpha_m = lin_m*(amax-amin)+amin,
 * out[b,n] = reduce_m (L * tri(V, s + alpha_m d)) * step, step = (amax-amin)/(P-1). */
void FN(oracle_trilinear_fwd)(const REAL *vol, int D0, int D1, int D2, const REAL *src, const REAL *tgt,
                              const REAL *raylen, REAL *out, int B, long N, REAL shift, REAL eps, int n_points,
                              REAL alphamin, REAL alphamax, int reduce, int align_corners)
{
    const int dims[3] = {D0, D1, D2};
    const REAL step = (alphamax - alphamin) / (REAL)(n_points - 1);
#pragma omp parallel for schedule(dynamic, 64)
    for (long r = 0; r < (long)B * N; ++r) {
        const int b = (int)(r / N);
        REAL s[3], d[3];
        for (int a = 0; a < 3; ++a) {
            s[a] = src[b * 3 + a];
            d[a] = (tgt[r * 3 + a] - s[a]) + eps;
        }
        const REAL L = raylen[r];
        REAL acc = 0;
        for (int m = 0; m < n_points; ++m) {
            REAL alpha = FN(linspace01)(m, n_points) * (alphamax - alphamin) + alphamin;
            REAL pix[3];
            for (int a = 0; a < 3; ++a) pix[a] = FN(pix_at)(alpha, s[a], d[a], shift, dims[a], align_corners);
            REAL term = (L * FN(trilerp)(vol, dims, pix, NULL, NULL, NULL)) * step;
            if (reduce == 0) acc += term;
            else if (m == 0 || term > acc) acc = term;
        }
        out[r] = acc;
    }
}

/* renderers.py:242-252 (Trilinear.forward with a label mask): the label is sampled with mode="nearest" at the SAME
 * point as the trilinear density sample. */
void FN(oracle_trilinear_fwd_mask)(const REAL *vol, const REAL *mask, int D0, int D1, int D2, const REAL *src,
                                   const REAL *tgt, const REAL *raylen, REAL *out, int B, long N, int C, REAL shift,
                                   REAL eps, int n_points, REAL alphamin, REAL alphamax, int align_corners)
{
    const int dims[3] = {D0, D1, D2};
    const REAL step = (alphamax - alphamin) / (REAL)(n_points - 1);
    memset(out, 0, sizeof(REAL) * (size_t)B * C * N);
#pragma omp parallel for schedule(dynamic, 64)
    for (long r = 0; r < (long)B * N; ++r) {
        const int b = (int)(r / N);
        const long n = r % N;
        REAL s[3], d[3];
        for (int a = 0; a < 3; ++a) {
            s[a] = src[b * 3 + a];
            d[a] = (tgt[r * 3 + a] - s[a]) + eps;
        }
        const REAL L = raylen[r];
        for (int m = 0; m < n_points; ++m) {
            REAL alpha = FN(linspace01)(m, n_points) * (alphamax - alphamin) + alphamin;
            REAL pix[3];
            long c = 0, flat = 0;
            int inb = 1;
            for (int a = 0; a < 3; ++a) {
                pix[a] = FN(pix_at)(alpha, s[a], d[a], shift, dims[a], align_corners);
                REAL rr = (REAL)nearbyint((double)pix[a]);
                if (!(rr >= 0 && rr < (REAL)dims[a])) inb = 0;
                flat = flat * dims[a] + (long)(inb ? rr : 0);
            }
            if (inb) c = (long)mask[flat];
            if (c < 0 || c >= C) continue;
            out[((long)b * C + c) * N + n] += (L * FN(trilerp)(vol, dims, pix, NULL, NULL, NULL)) * step;
        }
    }
}

/* Autograd of Trilinear.forward in closed form, checked against the
 * reference's autograd through the tests/golden trilinear fixtures.  With x_m = s + alpha_m d (voxel units;
 * dpix/dx = 1 for align_corners=False and (D-1)/D for True), G_m = grad tri(V)(x_m):
 *   g_s  = g L step sum_m (1 - alpha_m) G_m        g_t = g L step sum_m alpha_m G_m
 *   g_L  = g step sum_m V_m
 *   g_amin = g L [ -sum V_m/(P-1) + step sum (1 - lin_m) G_m . d ]
 *   g_amax = g L [ +sum V_m/(P-1) + step sum      lin_m  G_m . d ]
 *   g_V[corner] += g L step w_corner
 * g_amin/g_amax are summed over all rays into two scalars. */
void FN(oracle_trilinear_bwd)(const REAL *vol, int D0, int D1, int D2, const REAL *src, const REAL *tgt,
                              const REAL *raylen, const REAL *gout, REAL *g_src, REAL *g_tgt, REAL *g_raylen,
                              REAL *g_vol, REAL *g_amin, REAL *g_amax, int B, long N, REAL shift, REAL eps,
                              int n_points, REAL alphamin, REAL alphamax, int align_corners)
{
    const int dims[3] = {D0, D1, D2};
    const REAL step = (alphamax - alphamin) / (REAL)(n_points - 1);
    if (g_src) memset(g_src, 0, sizeof(REAL) * (size_t)B * 3);
    REAL tot_amin = 0, tot_amax = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : tot_amin, tot_amax)
    for (long r = 0; r < (long)B * N; ++r) {
        const int b = (int)(r / N);
        REAL s[3], d[3], scale[3];
        for (int a = 0; a < 3; ++a) {
            s[a] = src[b * 3 + a];
            d[a] = (tgt[r * 3 + a] - s[a]) + eps;
            scale[a] = align_corners ? (REAL)(dims[a] - 1) / (REAL)dims[a] : (REAL)1;
        }
        const REAL L = raylen[r], g = gout[r];
        REAL sumV = 0, gs[3] = {0, 0, 0}, gt[3] = {0, 0, 0}, ga0 = 0, ga1 = 0;
        for (int m = 0; m < n_points; ++m) {
            REAL lin = FN(linspace01)(m, n_points);
            REAL alpha = lin * (alphamax - alphamin) + alphamin;
            REAL pix[3], G[3], cw[8];
            long ci[8];
            for (int a = 0; a < 3; ++a) pix[a] = FN(pix_at)(alpha, s[a], d[a], shift, dims[a], align_corners);
            REAL v = FN(trilerp)(vol, dims, pix, G, ci, cw);
            sumV += v;
            REAL Gd = 0;
            for (int a = 0; a < 3; ++a) {
                G[a] *= scale[a];
                gs[a] += ((REAL)1 - alpha) * G[a];
                gt[a] += alpha * G[a];
                Gd += G[a] * d[a];
            }
            ga0 += ((REAL)1 - lin) * Gd;
            ga1 += lin * Gd;
            if (g_vol)
                for (int c = 0; c < 8; ++c)
                    if (ci[c] >= 0) {
                        REAL add = g * L * step * cw[c];
#pragma omp atomic
                        g_vol[ci[c]] += add;
                    }
        }
        for (int a = 0; a < 3; ++a) {
            if (g_tgt) g_tgt[r * 3 + a] = g * L * step * gt[a];
            if (g_src) {
                REAL add = g * L * step * gs[a];
#pragma omp atomic
                g_src[b * 3 + a] += add;
            }
        }
        if (g_raylen) g_raylen[r] = g * step * sumV;
        tot_amin += g * L * (-sumV / (REAL)(n_points - 1) + step * ga0);
        tot_amax += g * L * (sumV / (REAL)(n_points - 1) + step * ga1);
    }
    if (g_amin) *g_amin = tot_amin;
    if (g_amax) *g_amax = tot_amax;
}

/* renderers.py:34-76 with mode="bilinear" (renderers.py:18,66 -> grid_sample(mode="bilinear")): the density of segment j
 * is the trilinear interpolant T at the segment MIDPOINT x_j = s + abar_j d, abar_j = (alpha_j + alpha_{j+1})/2:
 *   I = L sum_j len_j T(x_j).
 * Backward (closed form; stop_grad restates stop_gradients_through_grid_sample=True, i.e. T is a constant):
 *   dI/dalpha_m = L [ (T_{m-1} - T_m) + (len_{m-1} G_{m-1}.d + len_m G_m.d)/2 ],  G = grad T w.r.t. x (x dpix/dx)
 *   dI/ds = sum_m dI/dalpha_m dalpha_m/ds + L sum_j len_j (1 - abar_j) G_j,   dI/dt likewise with abar_j
 *   dI/dL = sum_j len_j T_j,   dI/dV[corner] += L len_j w_corner.
 * g_* == NULL everywhere: forward only (out [B][N]).  reduce must be 0 (sum). */
void FN(oracle_siddon_bilinear)(const REAL *vol, int D0, int D1, int D2, const REAL *src, const REAL *tgt,
                                const REAL *raylen, const REAL *gout, REAL *out, REAL *g_src, REAL *g_tgt, REAL *g_raylen,
                                REAL *g_vol, int B, long N, REAL shift, REAL eps, int stop_grad, int align_corners)
{
    const int dims[3] = {D0, D1, D2};
    const int M = D0 + D1 + D2 + 3;
    if (g_src) memset(g_src, 0, sizeof(REAL) * (size_t)B * 3);
#pragma omp parallel
    {
        REAL *alpha = (REAL *)malloc(sizeof(REAL) * (size_t)M);
        int *axis = (int *)malloc(sizeof(int) * (size_t)M);
        REAL *T = (REAL *)malloc(sizeof(REAL) * (size_t)(M + 1));
        REAL *Gd = (REAL *)malloc(sizeof(REAL) * (size_t)(M + 1));
#pragma omp for schedule(dynamic, 64)
        for (long r = 0; r < (long)B * N; ++r) {
            const int b = (int)(r / N);
            REAL s[3], d[3], scale[3];
            for (int a = 0; a < 3; ++a) {
                s[a] = src[b * 3 + a];
                d[a] = (tgt[r * 3 + a] - s[a]) + eps;
                scale[a] = align_corners ? (REAL)(dims[a] - 1) / (REAL)dims[a] : (REAL)1;
            }
            FN(sorted_alphas)(dims, s, d, shift, alpha, axis, NULL);
            const REAL L = raylen[r], g = gout ? gout[r] : (REAL)0;
            REAL acc = 0, sumTL = 0, gs[3] = {0, 0, 0}, gt[3] = {0, 0, 0};
            for (int j = 0; j + 1 < M; ++j) {
                const REAL amid = (alpha[j] + alpha[j + 1]) / (REAL)2, len = alpha[j + 1] - alpha[j];
                REAL pix[3], G[3], cw[8];
                long ci[8];
                for (int a = 0; a < 3; ++a) pix[a] = FN(pix_at)(amid, s[a], d[a], shift, dims[a], align_corners);
                T[j] = FN(trilerp)(vol, dims, pix, G, ci, cw);
                acc += (L * T[j]) * len;
                sumTL += T[j] * len;
                Gd[j] = 0;
                for (int a = 0; a < 3; ++a) {
                    G[a] *= scale[a];
                    Gd[j] += G[a] * d[a];
                    if (!stop_grad) { /* direct dependence of the midpoint on s and t */
                        gs[a] += g * L * len * ((REAL)1 - amid) * G[a];
                        gt[a] += g * L * len * amid * G[a];
                    }
                }
                if (g_vol && !stop_grad)
                    for (int k = 0; k < 8; ++k)
                        if (ci[k] >= 0) {
                            REAL add = g * L * len * cw[k];
#pragma omp atomic
                            g_vol[ci[k]] += add;
                        }
            }
            if (out) out[r] = acc;
            for (int m = 0; m < M; ++m) {
                const REAL Tm1 = m > 0 ? T[m - 1] : (REAL)0, Tm = m + 1 < M ? T[m] : (REAL)0;
                REAL c = Tm1 - Tm;
                if (!stop_grad) {
                    if (m > 0) c += (alpha[m] - alpha[m - 1]) * Gd[m - 1] / (REAL)2;
                    if (m + 1 < M) c += (alpha[m + 1] - alpha[m]) * Gd[m] / (REAL)2;
                }
                c *= g * L;
                const int a = axis[m];
                gs[a] += c * (alpha[m] - (REAL)1) / d[a];
                gt[a] += c * (-alpha[m]) / d[a];
            }
            for (int a = 0; a < 3; ++a) {
                if (g_tgt) g_tgt[r * 3 + a] = gt[a];
                if (g_src) {
#pragma omp atomic
                    g_src[b * 3 + a] += gs[a];
                }
            }
            if (g_raylen) g_raylen[r] = stop_grad ? (REAL)0 : g * sumTL;
        }
        free(alpha);
        free(axis);
        free(T);
        free(Gd);
    }
}

/* Autograd of Trilinear.forward with reducefn="max": I = max_m (L V_m) step, gradient through the FIRST maximal sample m*
 * only; the closed forms of oracle_trilinear_bwd with every sum over m replaced by its m* term. */
void FN(oracle_trilinear_bwd_max)(const REAL *vol, int D0, int D1, int D2, const REAL *src, const REAL *tgt,
                                  const REAL *raylen, const REAL *gout, REAL *g_src, REAL *g_tgt, REAL *g_raylen,
                                  REAL *g_vol, REAL *g_amin, REAL *g_amax, int B, long N, REAL shift, REAL eps,
                                  int n_points, REAL alphamin, REAL alphamax, int align_corners)
{
    const int dims[3] = {D0, D1, D2};
    const REAL step = (alphamax - alphamin) / (REAL)(n_points - 1);
    if (g_src) memset(g_src, 0, sizeof(REAL) * (size_t)B * 3);
    REAL tot_amin = 0, tot_amax = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : tot_amin, tot_amax)
    for (long r = 0; r < (long)B * N; ++r) {
        const int b = (int)(r / N);
        REAL s[3], d[3], scale[3];
        for (int a = 0; a < 3; ++a) {
            s[a] = src[b * 3 + a];
            d[a] = (tgt[r * 3 + a] - s[a]) + eps;
            scale[a] = align_corners ? (REAL)(dims[a] - 1) / (REAL)dims[a] : (REAL)1;
        }
        const REAL L = raylen[r], g = gout[r];
        int mbest = 0;
        REAL tbest = 0;
        for (int m = 0; m < n_points; ++m) {
            REAL alpha = FN(linspace01)(m, n_points) * (alphamax - alphamin) + alphamin;
            REAL pix[3];
            for (int a = 0; a < 3; ++a) pix[a] = FN(pix_at)(alpha, s[a], d[a], shift, dims[a], align_corners);
            REAL term = (L * FN(trilerp)(vol, dims, pix, NULL, NULL, NULL)) * step;
            if (m == 0 || term > tbest) {
                tbest = term;
                mbest = m;
            }
        }
        const REAL lin = FN(linspace01)(mbest, n_points);
        const REAL alpha = lin * (alphamax - alphamin) + alphamin;
        REAL pix[3], G[3], cw[8];
        long ci[8];
        for (int a = 0; a < 3; ++a) pix[a] = FN(pix_at)(alpha, s[a], d[a], shift, dims[a], align_corners);
        const REAL v = FN(trilerp)(vol, dims, pix, G, ci, cw);
        REAL Gd = 0;
        for (int a = 0; a < 3; ++a) {
            G[a] *= scale[a];
            Gd += G[a] * d[a];
            if (g_tgt) g_tgt[r * 3 + a] = g * L * step * alpha * G[a];
            if (g_src) {
                REAL add = g * L * step * ((REAL)1 - alpha) * G[a];
#pragma omp atomic
                g_src[b * 3 + a] += add;
            }
        }
        if (g_vol)
            for (int k = 0; k < 8; ++k)
                if (ci[k] >= 0) {
                    REAL add = g * L * step * cw[k];
#pragma omp atomic
                    g_vol[ci[k]] += add;
                }
        if (g_raylen) g_raylen[r] = g * step * v;
        tot_amin += g * L * (-v / (REAL)(n_points - 1) + step * ((REAL)1 - lin) * Gd);
        tot_amax += g * L * (v / (REAL)(n_points - 1) + step * lin * Gd);
    }
    if (g_amin) *g_amin = tot_amin;
    if (g_amax) *g_amax = tot_amax;
}

/* Autograd of Trilinear.forward WITH a label mask (renderers.py:242-252): sample m goes to channel c_m = nearest label
 * at the sample point (zero padding -> label 0), so its upstream gradient is g_m = gout[b][c_m][n]; the closed forms
 * above hold with every per-sample term weighted by g_m (and g = 1). */
void FN(oracle_trilinear_bwd_mask)(const REAL *vol, const REAL *mask, int D0, int D1, int D2, const REAL *src,
                                   const REAL *tgt, const REAL *raylen, const REAL *gout, REAL *g_src, REAL *g_tgt,
                                   REAL *g_raylen, REAL *g_vol, REAL *g_amin, REAL *g_amax, int B, long N, int C, REAL shift,
                                   REAL eps, int n_points, REAL alphamin, REAL alphamax, int align_corners)
{
    const int dims[3] = {D0, D1, D2};
    const REAL step = (alphamax - alphamin) / (REAL)(n_points - 1);
    if (g_src) memset(g_src, 0, sizeof(REAL) * (size_t)B * 3);
    REAL tot_amin = 0, tot_amax = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : tot_amin, tot_amax)
    for (long r = 0; r < (long)B * N; ++r) {
        const int b = (int)(r / N);
        const long n = r % N;
        REAL s[3], d[3], scale[3];
        for (int a = 0; a < 3; ++a) {
            s[a] = src[b * 3 + a];
            d[a] = (tgt[r * 3 + a] - s[a]) + eps;
            scale[a] = align_corners ? (REAL)(dims[a] - 1) / (REAL)dims[a] : (REAL)1;
        }
        const REAL L = raylen[r];
        REAL sumV = 0, gs[3] = {0, 0, 0}, gt[3] = {0, 0, 0}, ga0 = 0, ga1 = 0;
        for (int m = 0; m < n_points; ++m) {
            REAL lin = FN(linspace01)(m, n_points);
            REAL alpha = lin * (alphamax - alphamin) + alphamin;
            REAL pix[3], G[3], cw[8];
            long ci[8], c = 0, flat = 0;
            int inb = 1;
            for (int a = 0; a < 3; ++a) {
                pix[a] = FN(pix_at)(alpha, s[a], d[a], shift, dims[a], align_corners);
                REAL rr = (REAL)nearbyint((double)pix[a]);
                if (!(rr >= 0 && rr < (REAL)dims[a])) inb = 0;
                flat = flat * dims[a] + (long)(inb ? rr : 0);
            }
            if (inb) c = (long)mask[flat];
            const REAL g = (c >= 0 && c < C) ? gout[((long)b * C + c) * N + n] : (REAL)0;
            REAL v = FN(trilerp)(vol, dims, pix, G, ci, cw);
            sumV += g * v;
            REAL Gd = 0;
            for (int a = 0; a < 3; ++a) {
                G[a] *= scale[a] * g;
                gs[a] += ((REAL)1 - alpha) * G[a];
                gt[a] += alpha * G[a];
                Gd += G[a] * d[a];
            }
            ga0 += ((REAL)1 - lin) * Gd;
            ga1 += lin * Gd;
            if (g_vol)
                for (int k = 0; k < 8; ++k)
                    if (ci[k] >= 0) {
                        REAL add = g * L * step * cw[k];
#pragma omp atomic
                        g_vol[ci[k]] += add;
                    }
        }
        for (int a = 0; a < 3; ++a) {
            if (g_tgt) g_tgt[r * 3 + a] = L * step * gt[a];
            if (g_src) {
                REAL add = L * step * gs[a];
#pragma omp atomic
                g_src[b * 3 + a] += add;
            }
        }
        if (g_raylen) g_raylen[r] = step * sumV;
        tot_amin += L * (-sumV / (REAL)(n_points - 1) + step * ga0);
        tot_amax += L * (sumV / (REAL)(n_points - 1) + step * ga1);
    }
    if (g_amin) *g_amin = tot_amin;
    if (g_amax) *g_amax = tot_amax;
}

#undef FN
#undef CAT
#undef CAT_

// softras_math.cuh -- per-(pixel, face) arithmetic of the SoftRas rasterizer.
//
// Semantics follow the reference device functions (jrender/renderer/dr/softras/cuda/
// soft_rasterize.py:20-173 forward, :1118-1174 backward); the evaluation ORDER of every
// fp32 expression is kept so that results are bit-identical to oracle/softras_oracle.c
// (see common.cuh for the arithmetic contract).  Where the reference promotes to double
// through a literal (`1. / (1. + exp(..))`, `alpha *= 1. - D`) the same promotion is
// done here; where the promotion provably cannot change the fp32 result (clamps against
// representable constants, `1./x` of a float: 53 >= 2*24+2 so double rounding is
// innocuous) plain fp32 is used.  Divisions go through exact_math.cuh's fast_div, which
// returns the correctly rounded quotient without a MUFU when the reciprocal is known.
#pragma once
#include "common.cuh"
#include "exact_math.cuh"

namespace b200r {

// Reciprocals of the per-launch constant denominators, computed once per thread.
struct DivConst {
    float sigma, r_sigma, gamma, r_gamma, span, r_span;  // span = far - near
    bool s_sigma, s_gamma, s_span;
    __device__ __forceinline__ void init(const SoftRasParams& P) {
        sigma = P.sigma; gamma = P.gamma; span = P.far_ - P.near_;
        r_sigma = rcp_refined(sigma); r_gamma = rcp_refined(gamma); r_span = rcp_refined(span);
        s_sigma = midrange(sigma); s_gamma = midrange(gamma); s_span = midrange(span);
    }
    __device__ __forceinline__ float by_sigma(float a) const { return fast_div(a, sigma, r_sigma, s_sigma); }
    __device__ __forceinline__ float by_gamma(float a) const { return fast_div(a, gamma, r_gamma, s_gamma); }
    __device__ __forceinline__ float by_span(float a) const { return fast_div(a, span, r_span, s_span); }
    // EXACT = false: relaxed_div (arguments of exp() and gradient terms only)
    template <bool EXACT> __device__ __forceinline__ float by_sigma_t(float a) const { return div_t<EXACT>(a, sigma, r_sigma, s_sigma); }
    template <bool EXACT> __device__ __forceinline__ float by_gamma_t(float a) const { return div_t<EXACT>(a, gamma, r_gamma, s_gamma); }
    template <bool EXACT> __device__ __forceinline__ float by_span_t(float a) const { return div_t<EXACT>(a, span, r_span, s_span); }
    // optimistic (branch-free) flavours, see DivGuard in exact_math.cuh.  The three denominators are launch constants:
    // a pair's guard STARTS as consts_ok() (so a launch with an out-of-range sigma / gamma / far - near simply runs
    // every pair through the guarded functions) and the per-division test covers the numerator only.
    __device__ __forceinline__ bool consts_ok() const { return s_sigma && s_gamma && s_span; }
    __device__ __forceinline__ float by_sigma_o(float a, DivGuard& g) const { return optimistic_div_okden(a, sigma, r_sigma, g); }
    __device__ __forceinline__ float by_gamma_o(float a, DivGuard& g) const { return optimistic_div_okden(a, gamma, r_gamma, g); }
    __device__ __forceinline__ float by_span_o(float a, DivGuard& g) const { return optimistic_div_okden(a, span, r_span, g); }
};

// :20-25
__device__ __forceinline__ void barycentric_coordinate(float w[3], float x, float y, const float* inv) {
    w[0] = inv[0] * x + inv[1] * y + inv[2];
    w[1] = inv[3] * x + inv[4] * y + inv[5];
    w[2] = inv[6] * x + inv[7] * y + inv[8];
}

// :43-46
__device__ __forceinline__ bool check_pixel_inside(const float w[3]) {
    return w[0] <= 1.f && w[0] >= 0.f && w[1] <= 1.f && w[1] >= 0.f && w[2] <= 1.f && w[2] >= 0.f;
}

// Unchecked Markstein step: correctly rounded a / b when the caller has established that
// b is midrange and a is zero or midrange (exact_math.cuh explains the sequence).
__device__ __forceinline__ float div_nocheck(float a, float b, float r) {
    const float q = a * r;
    const float rem = __fmaf_rn(q, -b, a);
    return __fmaf_rn(r, rem, q);
}

// barycentric_clip (:49-54) followed by zp = 1. / (w0/z0 + w1/z1 + w2/z2) (:364, :1296).
// max(min(w,1.),0.) and max(sum,1e-5) are exact in fp32 (see header comment).  The six divisions
// share ONE range check: the clamped weights lie in [0, 1], w_sum in [1e-5, 3], so if every
// weight is 0 or >= 2^-58 (and the face's z are midrange, flag bits 4-6) all numerators stay
// inside fast_div's safe range before and after the normalisation; otherwise the individually
// checked path runs.  Zero numerators are exact in the unchecked sequence because they are +0.
// STRICT = false (backward, where zp only scales gradient terms): weights below 2^-58 are not
// diverted to the checked path (relaxed_div's argument applies).
template <bool STRICT = true>
__device__ __forceinline__ float clip_and_z(float w[3], const FaceRec* rec) {
#pragma unroll
    for (int k = 0; k < 3; k++) w[k] = fmaxf(fminf(w[k], 1.f), 0.f);
    const float w_sum = fmaxf(w[0] + w[1] + w[2], 1e-5f);
    const uint32_t fl = rec->flags;
    const float tiny = 3.4694469519536142e-18f;  // 2^-58
    const bool ok = ((fl & 0x70u) == 0x70u) &&
                    (!STRICT || ((w[0] == 0.f || w[0] >= tiny) && (w[1] == 0.f || w[1] >= tiny) && (w[2] == 0.f || w[2] >= tiny)));
    if (ok) {
        if (w_sum != 1.f) {  // x / 1 == x exactly
            const float r = rcp_refined(w_sum);
#pragma unroll
            for (int k = 0; k < 3; k++) w[k] = div_nocheck(w[k], w_sum, r);
        }
        const float a = div_nocheck(w[0], rec->v[2], rec->rz[0]);
        const float b = div_nocheck(w[1], rec->v[5], rec->rz[1]);
        const float c = div_nocheck(w[2], rec->v[8], rec->rz[2]);
        return 1.f / (a + b + c);
    }
    if (w_sum != 1.f) {
        const float r = rcp_refined(w_sum);
        const bool safe = midrange(w_sum);
#pragma unroll
        for (int k = 0; k < 3; k++) w[k] = fast_div(w[k], w_sum, r, safe);
    }
    const float a = fast_div(w[0], rec->v[2], rec->rz[0], (fl & 16u) != 0);
    const float b = fast_div(w[1], rec->v[5], rec->rz[1], (fl & 32u) != 0);
    const float c = fast_div(w[2], rec->v[8], rec->rz[2], (fl & 64u) != 0);
    return 1.f / (a + b + c);
}

// Optimistic flavour of clip_and_z: always the unchecked sequence; the per-face flag test, (STRICT: the forward,
// where zp orders the top-K list) the "weights are 0 or >= 2^-58" condition and the final reciprocal's range go
// into the guard.  With the guard up the result equals clip_and_z<STRICT>'s bit for bit.
template <bool STRICT = false>
__device__ __forceinline__ float clip_and_z_opt(float w[3], const FaceRec* rec, DivGuard& g) {
#pragma unroll
    for (int k = 0; k < 3; k++) w[k] = fmaxf(fminf(w[k], 1.f), 0.f);
    const float w_sum = fmaxf(w[0] + w[1] + w[2], 1e-5f);
    g.ok = g.ok && ((rec->flags & 0x70u) == 0x70u);
    if (STRICT) {
        // each clamped weight (in [+0, 1]) must be 0 or >= 2^-58: on the bit patterns, (bits - 1) >= (bits(2^-58) - 1)
        // unsigned (0 wraps to 0xffffffff) -- one add + one compare-and-AND per weight
        const uint32_t lim = 0x22800000u - 1u;  // bits(2^-58) - 1
#pragma unroll
        for (int k = 0; k < 3; k++) g.ok = g.ok && (__float_as_uint(w[k]) - 1u) >= lim;
    }
    if (w_sum != 1.f) {  // x / 1 == x exactly
        const float r = rcp_refined(w_sum);
#pragma unroll
        for (int k = 0; k < 3; k++) w[k] = div_nocheck(w[k], w_sum, r);
    }
    const float a = div_nocheck(w[0], rec->v[2], rec->rz[0]);
    const float b = div_nocheck(w[1], rec->v[5], rec->rz[1]);
    const float c = div_nocheck(w[2], rec->v[8], rec->rz[2]);
    return optimistic_rcp(a + b + c, g);
}

__device__ __forceinline__ float sel3(int i, float a, float b, float c) { return i == 0 ? a : (i == 1 ? b : c); }

// :57-147.  rec->a0 holds the pre-subtracted Gram-matrix rows.  Returns sign; writes
// dis_x, dis_y and (for the backward) t[3].
template <bool STRICT = true>
__device__ __forceinline__ float euclidean_p2f_distance(float& dis_x, float& dis_y, float t[3],
                                                        const float w[3], const FaceRec* rec,
                                                        float xp, float yp, DivGuard* guard = nullptr) {
    const float* f = rec->v;
    const uint32_t fl = rec->flags;
    // guard != nullptr (a compile-time fact after inlining): branch-free division, range folded into the guard
    auto edge_div = [&](float a, float b, float r, bool safe) -> float {
        if (guard != nullptr) return STRICT ? optimistic_div_exact(a, b, r, safe, *guard) : optimistic_div(a, b, r, safe, *guard);
        return div_t<STRICT>(a, b, r, safe);
    };
    if (w[0] > 0.f && w[1] > 0.f && w[2] > 0.f && w[0] < 1.f && w[1] < 1.f && w[2] < 1.f) {
        float dis_min = 100000000.f;
        float dis_x_min = 0.f, dis_y_min = 0.f;
        t[0] = t[1] = t[2] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int v0 = k, v1 = (k + 1) % 3, v2 = (k + 2) % 3;
            const float* a = rec->a0 + 3 * k;
            float t0[3];
            t0[v0] = edge_div(w[0] * a[0] + w[1] * a[1] + w[2] * a[2] - a[v1], a[v0] - a[v1], rec->rden[k],
                              (fl & (128u << k)) != 0);
            t0[v1] = 1.f - t0[v0];
            t0[v2] = 0.f;
            t0[0] -= w[0];
            t0[1] -= w[1];
            t0[2] -= w[2];
            const float dx = t0[0] * f[0] + t0[1] * f[3] + t0[2] * f[6];
            const float dy = t0[0] * f[1] + t0[1] * f[4] + t0[2] * f[7];
            const float dis = dx * dx + dy * dy;
            if (dis < dis_min) {
                dis_min = dis;
                dis_x_min = dx;
                dis_y_min = dy;
                t[0] = t0[0];
                t[1] = t0[1];
                t[2] = t0[2];
            }
        }
        dis_x = dis_x_min;
        dis_y = dis_y_min;
        return 1.f;
    } else {
        int v0 = 0;  // the reference leaves v0 = -1 (UB) when no branch fires; oracle uses 0 too
        if (w[1] <= 0.f && w[2] <= 0.f) {
            v0 = 0;
            if ((fl & 1u) && (xp - f[0]) * (f[6] - f[0]) + (yp - f[1]) * (f[7] - f[1]) > 0.f) v0 = 2;
        } else if (w[2] <= 0.f && w[0] <= 0.f) {
            v0 = 1;
            if ((fl & 2u) && (xp - f[3]) * (f[0] - f[3]) + (yp - f[4]) * (f[1] - f[4]) > 0.f) v0 = 0;
        } else if (w[0] <= 0.f && w[1] <= 0.f) {
            v0 = 2;
            if ((fl & 4u) && (xp - f[6]) * (f[3] - f[6]) + (yp - f[7]) * (f[4] - f[7]) > 0.f) v0 = 1;
        } else if (w[0] <= 0.f) v0 = 1;
        else if (w[1] <= 0.f) v0 = 2;
        else if (w[2] <= 0.f) v0 = 0;

        const int v1 = v0 == 2 ? 0 : v0 + 1;
        const float* a = rec->a0 + 3 * v0;
        const float a_0 = a[0], a_1 = a[1], a_2 = a[2];
        const float a_v0 = a[v0];   // the record lives in shared memory (forward) / L1 (backward): two indexed loads
        const float a_v1 = a[v1];   // are cheaper than two compare-select chains
        const float tv0 = edge_div(w[0] * a_0 + w[1] * a_1 + w[2] * a_2 - a_v1, a_v0 - a_v1, rec->rden[v0],
                                   (fl & (128u << v0)) != 0);
        const float tv1 = 1.f - tv0;
        const float c0 = fminf(fmaxf(tv0, 0.f), 1.f);
        const float c1 = fminf(fmaxf(tv1, 0.f), 1.f);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float tk = (k == v0) ? c0 : ((k == v1) ? c1 : 0.f);
            t[k] = tk - w[k];
        }
        dis_x = t[0] * f[0] + t[1] * f[3] + t[2] * f[6];
        dis_y = t[0] * f[1] + t[1] * f[4] + t[2] * f[7];
        return -1.f;
    }
}

// :150-154
__device__ __forceinline__ float barycentric_p2f_distance(const float w[3]) {
    float dis = w[0] > w[1] ? (w[1] > w[2] ? w[2] : w[1]) : (w[0] > w[2] ? w[2] : w[0]);
    dis = dis > 0.f ? dis * dis : -dis * dis;
    return dis;
}

// `1. / (1. + exp(x))` with x already = -sign*dis/sigma  (:338, :344).
// EXACT: the reference's double-precision tail, bit for bit.  !EXACT: the same expression in fp32
// (two roundings instead of one: <= 1 ulp from the exact tail, 6e-8 relative) -- about 20 fewer
// instructions per (pixel, face) pair and no FP64 pipe.  D feeds no discrete decision except the
// `D > 0.5` of hard alpha, so every index / depth output stays bit-identical either way.
template <bool EXACT>
__device__ __forceinline__ float sigmoid_from_negarg(float x) {
    if (EXACT) return sigmoid_tail(expf(x));
    return 1.f / (1.f + expf(x));
}

// optimistic flavour of sigmoid_from_negarg<false>: same two roundings, reciprocal without the IEEE slow path
__device__ __forceinline__ float sigmoid_from_negarg_opt(float x, DivGuard& g) {
    return optimistic_rcp(1.f + expf(x), g);
}

// alpha "prod" aggregation  alpha *= 1. - D  (:357); see sigmoid_from_negarg for EXACT.
template <bool EXACT>
__device__ __forceinline__ float alpha_prod_t(float alpha, float D) {
    if (EXACT) return alpha_prod(alpha, D);
    return alpha * (1.f - D);
}

// surface texel index of forward_sample_texture / backward_sample_texture (:159-166, :1157-1168).
// R == 1: both branches of the reference give texel 0 (w is clipped to >= 0).
__device__ __forceinline__ int surface_texel(const float w[3], int R) {
    if (R == 1) return 0;
    const int w_x = (int)fminf(w[0] * R, (float)(R - 1));
    const int w_y = (int)fminf(w[1] * R, (float)(R - 1));
    if ((w[0] + w[1]) * R - w_x - w_y <= 1.f) return w_y * R + w_x;
    return (R - 1 - w_y) * R + (R - 1 - w_x);
}

// :156-173 forward flavour (vertex mode perspective-correct); tex points at this face's texels
__device__ __forceinline__ void sample_texture_fwd(float col[3], const float* __restrict__ btex, int fn, int T,
                                                   const float w[3], int R, int tex_type,
                                                   const FaceRec* rec, float z) {
    if (tex_type == 0) {
        if (R == 1) {   // the face colour travels in the record: no texture address arithmetic at all
            col[0] = rec->col[0]; col[1] = rec->col[1]; col[2] = rec->col[2];
            return;
        }
    }
    const float* __restrict__ tex = btex + (size_t)fn * T * 3;   // this face's texels
    if (tex_type == 0) {
        {
            const int j = surface_texel(w, R);
#pragma unroll
            for (int k = 0; k < 3; k++) col[k] = __ldg(tex + j * 3 + k);
        }
    } else {
        const uint32_t fl = rec->flags;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float c = fast_div(w[0] * __ldg(tex + k), rec->v[2], rec->rz[0], (fl & 16u) != 0) +
                            fast_div(w[1] * __ldg(tex + 3 + k), rec->v[5], rec->rz[1], (fl & 32u) != 0) +
                            fast_div(w[2] * __ldg(tex + 6 + k), rec->v[8], rec->rz[2], (fl & 64u) != 0);
            col[k] = c * z;
        }
    }
}

}  // namespace b200r

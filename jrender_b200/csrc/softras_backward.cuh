// softras_backward.cuh -- SoftRas top-K backward (replaces K6 of the reference:
// backward_soft_rasterize_cuda_kernel, cuda/soft_rasterize.py:1177-1360).
//
// The reference issues 9 + 3T global atomicAdd per (pixel, face) pair.  Here a warp owns
// an 8x4 pixel block; every lane sorts its <=K saved face ids ascending, and the warp then
// walks the UNION of its lanes' ids in ascending order (redux.sync min picks the next
// face).  Lanes that hold the face compute its gradient contribution, the 12 values are
// reduced across the warp with shuffles, and one lane issues the atomics: global atomics
// drop from one set per (pixel, face) to one set per (warp, face).
#pragma once
#include "softras_math.cuh"

namespace b200r {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

template <int DIST, int RGB>
__global__ void __launch_bounds__(B200R_TILE_THREADS, 2)
k_softras_backward(const SoftRasParams P, const FaceRec* __restrict__ recs, const float* __restrict__ textures,
                   const float* __restrict__ soft_colors, const float* __restrict__ aggrs_info,
                   const int* __restrict__ ids_in, const float* __restrict__ grad_soft_colors,
                   float* __restrict__ grad_faces, float* __restrict__ grad_textures) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* s_id = reinterpret_cast<int*>(smem_raw);                                     // [K][256]
    float* s_wrec = reinterpret_cast<float*>(s_id + (size_t)P.K * B200R_TILE_THREADS);  // [8][40]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int is = P.is, nf = P.nf, K = P.K, T = P.T;
    const int b = blockIdx.y;
    const int tx = blockIdx.x % P.ntx, ty = blockIdx.x / P.ntx;
    const int px = tx * B200R_TILE + (warp & 1) * 8 + (lane & 7);
    const int row = ty * B200R_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool valid = px < is && row < is;
    const float xp = b200r_pix_coord(px, is);
    const float yp = b200r_pix_coord(is - 1 - row, is);
    const size_t npix = (size_t)is * is;
    const size_t pn = (size_t)row * is + px;
    DivConst dc;
    dc.init(P);

    // ---- load + insertion-sort this pixel's ids (ascending; list ends at the first -1, :1236)
    int n = 0;
    if (valid) {
        const int* src = ids_in + (size_t)b * K * npix + pn;
        for (int k = 0; k < K; k++) {
            const int v = __ldg(src + (size_t)k * npix);
            if (v < 0) break;
            int j = n;
            while (j > 0) {
                const int u = s_id[(j - 1) * B200R_TILE_THREADS + tid];
                if (u <= v) break;
                s_id[j * B200R_TILE_THREADS + tid] = u;
                --j;
            }
            s_id[j * B200R_TILE_THREADS + tid] = v;
            ++n;
        }
    }
    // warp-uniform early out for empty blocks
    if (__ballot_sync(0xffffffffu, n > 0) == 0u) return;

    float g[4] = {0.f, 0.f, 0.f, 0.f}, oc[4] = {0.f, 0.f, 0.f, 0.f};
    float softmax_sum = 1.f, softmax_max = 0.f;
    if (valid) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            g[k] = __ldg(grad_soft_colors + ((size_t)b * 4 + k) * npix + pn);
            oc[k] = __ldg(soft_colors + ((size_t)b * 4 + k) * npix + pn);
        }
        softmax_sum = __ldg(aggrs_info + ((size_t)b * 2 + 0) * npix + pn);
        softmax_max = __ldg(aggrs_info + ((size_t)b * 2 + 1) * npix + pn);
    }

    // per-pixel / per-launch denominators of the backward
    const float r_ssum = rcp_refined(softmax_sum);
    const bool s_ssum = midrange(softmax_sum);
    const double d_galpha = (double)g[3];
    const double d_one_minus_alpha = (double)(1.f - oc[3]);
    const float nmf = P.near_ - P.far_;
    const float r_nmf = rcp_refined(nmf);
    const bool s_nmf = midrange(nmf);

    const FaceRec* brecs = recs + (size_t)b * nf;
    const float* btex = textures + (size_t)b * nf * T * 3;
    float* bgf = grad_faces + (size_t)b * nf * 9;
    float* bgt = grad_textures + (size_t)b * nf * T * 3;
    FaceRec* wrec = reinterpret_cast<FaceRec*>(s_wrec + warp * 40);

    int p = 0;
    int cur = (p < n) ? s_id[tid] : 0x7fffffff;
    while (true) {
        const int fn = __reduce_min_sync(0xffffffffu, cur);
        if (fn == 0x7fffffff) break;
        // stage the record for the warp: one 4-byte word per lane
        __syncwarp();
        reinterpret_cast<uint32_t*>(wrec)[lane] = __ldg(reinterpret_cast<const uint32_t*>(brecs + fn) + lane);
        if (lane < 8) reinterpret_cast<uint32_t*>(wrec)[32 + lane] = __ldg(reinterpret_cast<const uint32_t*>(brecs + fn) + 32 + lane);
        __syncwarp();
        const bool mine = (cur == fn);

        float gv[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // [k*3 + l]
        float gt[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // T==1: [k]; vertex: [j*3+k]
        if (mine) {
            const FaceRec* rec = wrec;
            const float* f = rec->v;
            float w[3], t[3] = {0.f, 0.f, 0.f};
            float dis = 0.f, dis_x = 0.f, dis_y = 0.f, sign = 0.f, soft_fragment;
            barycentric_coordinate(w, xp, yp, rec->inv);
            if (DIST == 0) {
                soft_fragment = 1.f;  // :1259
            } else if (DIST == 1) {
                dis = barycentric_p2f_distance(w);
                t[0] = w[0]; t[1] = w[1]; t[2] = w[2];
                soft_fragment = sigmoid_from_negarg(dc.by_sigma(-dis));
            } else {
                sign = euclidean_p2f_distance(dis_x, dis_y, t, w, rec, xp, yp);
                dis = dis_x * dis_x + dis_y * dis_y;
                soft_fragment = sigmoid_from_negarg(dc.by_sigma(-sign * dis));
            }

            float C_grad_xy = 0.f;
            float C_grad_xy_alpha = g[3];
            if (P.alpha_func == 1) {
                C_grad_xy_alpha = C_grad_xy_alpha / (float)nf;
            } else if (P.alpha_func == 2) {
                // (float)((double)g_a * ((double)(1 - alpha_out) / max((double)(1 - D), 1e-6)))  (:1289)
                const float omd = 1.f - soft_fragment;
                const double den = fmax(midrange(omd) ? f2d_mid(omd) : (double)omd, 1e-6);
                const double prod = d_galpha * (d_one_minus_alpha / den);
                C_grad_xy_alpha = d_midrange(prod) ? d2f_mid(prod) : (float)prod;
            }
            C_grad_xy += C_grad_xy_alpha;

            const float w0[3] = {w[0], w[1], w[2]};
            barycentric_clip(w);
            const float zp = interp_z(w, rec);

            const float* tex = btex + (size_t)fn * T * 3;
            if (RGB == 0) {
                if ((float)fn == softmax_max) {  // :1300 (int vs float compare, Q10)
                    if (P.tex_type == 0) {
                        const int j = surface_texel(w, P.R);
                        if (T == 1) { gt[0] = g[0]; gt[1] = g[1]; gt[2] = g[2]; }
                        else {
#pragma unroll
                            for (int k = 0; k < 3; k++) atomicAdd(bgt + ((size_t)fn * T + j) * 3 + k, g[k]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 3; j++)
#pragma unroll
                            for (int k = 0; k < 3; k++) gt[j * 3 + k] = w[j] * g[k];
                    }
                }
            } else if (RGB == 1) {
                float C_grad_xyz_rgb = 0.f;
                const float zp_norm = dc.by_span(P.far_ - zp);
                const float zp_softmax = fast_div(soft_fragment * expf(dc.by_gamma(zp_norm - softmax_max)), softmax_sum, r_ssum, s_ssum);
                float col[3];
                if (P.tex_type == 0) {
                    const int j = surface_texel(w, P.R);
                    if (T == 1) { col[0] = rec->col[0]; col[1] = rec->col[1]; col[2] = rec->col[2]; }
                    else {
#pragma unroll
                        for (int k = 0; k < 3; k++) col[k] = __ldg(tex + j * 3 + k);
                    }
                    if (T == 1) {
#pragma unroll
                        for (int k = 0; k < 3; k++) gt[k] = zp_softmax * g[k];
                    } else {
#pragma unroll
                        for (int k = 0; k < 3; k++) atomicAdd(bgt + ((size_t)fn * T + j) * 3 + k, zp_softmax * g[k]);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        col[k] = w[0] * __ldg(tex + k) + w[1] * __ldg(tex + 3 + k) + w[2] * __ldg(tex + 6 + k);
#pragma unroll
                    for (int j = 0; j < 3; j++)
#pragma unroll
                        for (int k = 0; k < 3; k++) gt[j * 3 + k] = zp_softmax * (w[j] * g[k]);
                }
#pragma unroll
                for (int k = 0; k < 3; k++) C_grad_xyz_rgb += g[k] * (col[k] - oc[k]);
                C_grad_xyz_rgb *= zp_softmax;
                C_grad_xy += C_grad_xyz_rgb / soft_fragment;

                const float C_grad_z_rgb = fast_div(dc.by_gamma(C_grad_xyz_rgb), nmf, r_nmf, s_nmf) * zp * zp;
                const uint32_t fl = rec->flags;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const bool sz = (fl & (16u << k)) != 0;
                    const float z = f[3 * k + 2], rz = rec->rz[k];
                    gv[k * 3 + 2] = fast_div(fast_div(C_grad_z_rgb * w[k], z, rz, sz), z, rz, sz);
                }
            }

            C_grad_xy *= dc.by_sigma(soft_fragment * (1.f - soft_fragment));  // :1336
            if (DIST == 1) {  // backward_barycentric_p2f_distance (:1118-1132), w := t (unclipped)
                const int pm = t[0] > t[1] ? (t[1] > t[2] ? 2 : 1) : (t[0] > t[2] ? 2 : 0);
                const float* inv = rec->inv;
                const float scale2 = dis > 0.f ? sqrtf(dis) : sqrtf(-dis);
#pragma unroll
                for (int l = 0; l < 2; l++)
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        float grad_kl = 0.f;
                        const float a = -inv[3 * pm + l];
                        grad_kl += a * inv[3 * k + 0] * xp;
                        grad_kl += a * inv[3 * k + 1] * yp;
                        grad_kl += a * inv[3 * k + 2] * 1.f;
                        float v = grad_kl * C_grad_xy;
                        v = (float)((double)v * (2.0 * (double)scale2));
                        gv[k * 3 + l] = v;
                    }
            } else if (DIST == 2) {  // :1341-1347
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    gv[k * 3 + 0] = 2.f * sign * C_grad_xy * (t[k] + w0[k]) * dis_x;
                    gv[k * 3 + 1] = 2.f * sign * C_grad_xy * (t[k] + w0[k]) * dis_y;
                }
            }
            ++p;
            cur = (p < n) ? s_id[p * B200R_TILE_THREADS + tid] : 0x7fffffff;
        }

        // ---- warp reduction + one set of atomics per (warp, face)
#pragma unroll
        for (int c = 0; c < 9; c++) gv[c] = warp_sum(gv[c]);
        const int ngt = (P.tex_type == 1) ? 9 : (T == 1 ? 3 : 0);
        if (RGB != 2) {
#pragma unroll
            for (int c = 0; c < 9; c++)
                if (c < ngt) gt[c] = warp_sum(gt[c]);
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 9; c++) atomicAdd(bgf + (size_t)fn * 9 + c, gv[c]);
            if (RGB != 2) {
#pragma unroll
                for (int c = 0; c < 9; c++)
                    if (c < ngt) atomicAdd(bgt + (size_t)fn * T * 3 + c, gt[c]);
            }
        }
    }
}

}  // namespace b200r

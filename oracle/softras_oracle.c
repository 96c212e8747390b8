/*
 * softras_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, IEEE fp32/fp64, no FMA contraction) of the
 * reference's SoftRas rasterizer kernels.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Follows, function by function (all paths relative to /root/reference):
 *   jrender/renderer/dr/softras/cuda/soft_rasterize.py
 *     :20-25    barycentric_coordinate
 *     :28-34    check_border
 *     :37-40    check_face_frontside
 *     :43-46    check_pixel_inside
 *     :49-54    barycentric_clip
 *     :57-147   euclidean_p2f_distance
 *     :150-154  forward_barycentric_p2f_distance
 *     :156-173  forward_sample_texture (forward flavour, perspective-correct)
 *     :176-236  forward_soft_rasterize_inv_cuda_kernel        (K1)
 *     :243-456  forward_soft_rasterize_cuda_kernel            (K2)
 *     :1118-1132 backward_barycentric_p2f_distance
 *     :1135-1151 forward_sample_texture (backward flavour, NOT perspective)
 *     :1154-1174 backward_sample_texture
 *     :1177-1360 backward_soft_rasterize_cuda_kernel (top-K)   (K6)
 *   host semantics: jrender/renderer/dr/softras/soft_rasterize.py:25,39-42,59-71,108
 *
 * Arithmetic contract.  scalar_t == float in the reference.  Every expression is
 * evaluated with the C/CUDA usual arithmetic conversions of the reference source:
 * a double literal (`1.`, `0.`, `1e-5`, `2.`) promotes that sub-expression to
 * double, and the result is rounded to float on assignment.  Build with
 * -ffp-contract=off so the compiler never fuses a*b+c (the product CUDA kernels
 * are built with -fmad=false for the same reason): +,-,*,/ and sqrt are then
 * bit-identical between this oracle and the GPU, and only expf() (libm vs CUDA
 * libdevice, both <= 2 ulp) differs.
 *
 * Documented deviations from the reference (both are UB in the reference):
 *   - backward_sample_texture returns an UNINITIALISED value for non-hit texels
 *     (:1155-1173); this oracle defines it as 0.
 *   - euclidean_p2f_distance with v0 == -1 (all w > 0 but some w >= 1, :107-121)
 *     indexes arrays at -1; this oracle skips the face's distance update by
 *     treating it as v0 = 0 (never hit by any test mesh; flagged in DESIGN.md).
 *
 * Parity status: pinned against outputs of the reference's own kernel strings
 * compiled for sm_100a and run on a B200 (tests/golden/ref_gpu_*.npz, generated
 * by oracle/make_ref_golden.py).  See DESIGN.md "Oracle".
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define K_MAX_POINTS_PER_PIXEL 64 /* :16 */

/* :20-25 */
static inline void barycentric_coordinate(float *w, float x, float y, const float *fi) {
    w[0] = fi[0] * x + fi[1] * y + fi[2];
    w[1] = fi[3] * x + fi[4] * y + fi[5];
    w[2] = fi[6] * x + fi[7] * y + fi[8];
}

/* :28-34 */
static inline int check_border(float x, float y, const float *face, float threshold) {
    return (x > fmaxf(fmaxf(face[0], face[3]), face[6]) + threshold ||
            x < fminf(fminf(face[0], face[3]), face[6]) - threshold ||
            y > fmaxf(fmaxf(face[1], face[4]), face[7]) + threshold ||
            y < fminf(fminf(face[1], face[4]), face[7]) - threshold);
}

/* :37-40 */
static inline int check_face_frontside(const float *face) {
    return (face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0]);
}

/* :43-46 */
static inline int check_pixel_inside(const float *w) {
    return w[0] <= 1 && w[0] >= 0 && w[1] <= 1 && w[1] >= 0 && w[2] <= 1 && w[2] >= 0;
}

/* :49-54 -- min/max against double literals, sum clamp against 1e-5 (double) */
static inline void barycentric_clip(float *w) {
    for (int k = 0; k < 3; k++) w[k] = (float)fmax(fmin((double)w[k], 1.), 0.);
    const float w_sum = (float)fmax((double)(w[0] + w[1] + w[2]), 1e-5);
    for (int k = 0; k < 3; k++) w[k] /= w_sum;
}

/* :57-147 */
static inline void euclidean_p2f_distance(float *sign, float *dis_x, float *dis_y,
                                          const float *w, float *t,
                                          const float *face, const float *face_info,
                                          float xp, float yp) {
    const float *face_sym = face_info + 9;
    const float *face_obt = face_info + 18;

    if (w[0] > 0 && w[1] > 0 && w[2] > 0 && w[0] < 1 && w[1] < 1 && w[2] < 1) {
        float dis_min = 100000000;
        float dis_x_min = 0;
        float dis_y_min = 0;
        float a0[3];
        float t0[3];
        for (int k = 0; k < 3; k++) {
            int v0 = k;
            int v1 = (k + 1) % 3;
            int v2 = (k + 2) % 3;
            a0[0] = face_sym[3 * v0 + 0] - face_sym[3 * v1 + 0];
            a0[1] = face_sym[3 * v0 + 1] - face_sym[3 * v1 + 1];
            a0[2] = face_sym[3 * v0 + 2] - face_sym[3 * v1 + 2];

            t0[v0] = (w[0] * a0[0] + w[1] * a0[1] + w[2] * a0[2] - a0[v1]) / (a0[v0] - a0[v1]);
            t0[v1] = 1 - t0[v0];
            t0[v2] = 0;

            t0[0] -= w[0];
            t0[1] -= w[1];
            t0[2] -= w[2];

            float dx = t0[0] * face[0] + t0[1] * face[3] + t0[2] * face[6];
            float dy = t0[0] * face[1] + t0[1] * face[4] + t0[2] * face[7];
            float dis = dx * dx + dy * dy;

            if (dis < dis_min) {
                dis_min = dis;
                dis_x_min = dx;
                dis_y_min = dy;
                t[0] = t0[0];
                t[1] = t0[1];
                t[2] = t0[2];
            }
        }
        *dis_x = dis_x_min;
        *dis_y = dis_y_min;
        *sign = 1;
    } else {
        int v0 = -1;

        if (w[1] <= 0 && w[2] <= 0) {
            v0 = 0;
            if (face_obt[0] == 1 && (xp - face[0]) * (face[6] - face[0]) + (yp - face[1]) * (face[7] - face[1]) > 0) v0 = 2;
        } else if (w[2] <= 0 && w[0] <= 0) {
            v0 = 1;
            if (face_obt[1] == 1 && (xp - face[3]) * (face[0] - face[3]) + (yp - face[4]) * (face[1] - face[4]) > 0) v0 = 0;
        } else if (w[0] <= 0 && w[1] <= 0) {
            v0 = 2;
            if (face_obt[2] == 1 && (xp - face[6]) * (face[3] - face[6]) + (yp - face[7]) * (face[4] - face[7]) > 0) v0 = 1;
        } else if (w[0] <= 0) v0 = 1;
        else if (w[1] <= 0) v0 = 2;
        else if (w[2] <= 0) v0 = 0;

        if (v0 < 0) v0 = 0; /* reference UB (index -1); see header */

        const int v1 = (v0 + 1) % 3;
        const int v2 = (v0 + 2) % 3;

        float a0[3];
        a0[0] = face_sym[3 * v0 + 0] - face_sym[3 * v1 + 0];
        a0[1] = face_sym[3 * v0 + 1] - face_sym[3 * v1 + 1];
        a0[2] = face_sym[3 * v0 + 2] - face_sym[3 * v1 + 2];

        t[v0] = (w[0] * a0[0] + w[1] * a0[1] + w[2] * a0[2] - a0[v1]) / (a0[v0] - a0[v1]);
        t[v1] = 1 - t[v0];
        t[v2] = 0;

        for (int k = 0; k < 3; k++) {
            t[k] = (float)fmin(fmax((double)t[k], 0.), 1.);
            t[k] -= w[k];
        }

        *dis_x = t[0] * face[0] + t[1] * face[3] + t[2] * face[6];
        *dis_y = t[0] * face[1] + t[1] * face[4] + t[2] * face[7];
        *sign = -1;
    }
}

/* :150-154 */
static inline float forward_barycentric_p2f_distance(const float *w) {
    float dis = w[0] > w[1] ? (w[1] > w[2] ? w[2] : w[1]) : (w[0] > w[2] ? w[2] : w[0]);
    dis = dis > 0 ? dis * dis : -dis * dis;
    return dis;
}

/* :156-173 (forward flavour) */
static inline float forward_sample_texture_fwd(const float *texture, const float *w, int R, int k,
                                               int texture_sample_type, const float *face, float z) {
    float texture_k = 0;
    if (texture_sample_type == 0) {
        const int w_x = (int)fminf(w[0] * R, (float)(R - 1));
        const int w_y = (int)fminf(w[1] * R, (float)(R - 1));
        if ((w[0] + w[1]) * R - w_x - w_y <= 1) {
            texture_k = texture[(w_y * R + w_x) * 3 + k];
        } else {
            texture_k = texture[((R - 1 - w_y) * R + (R - 1 - w_x)) * 3 + k];
        }
    } else if (texture_sample_type == 1) {
        texture_k = w[0] * texture[k] / face[2] + w[1] * texture[3 + k] / face[5] + w[2] * texture[6 + k] / face[8];
        texture_k *= z;
    }
    return texture_k;
}

/* :1135-1151 (backward flavour: vertex mode is NOT perspective-correct, Q8) */
static inline float forward_sample_texture_bwd(const float *texture, const float *w, int R, int k,
                                               int texture_sample_type) {
    float texture_k = 0;
    if (texture_sample_type == 0) {
        const int w_x = (int)fminf(w[0] * R, (float)(R - 1));
        const int w_y = (int)fminf(w[1] * R, (float)(R - 1));
        if ((w[0] + w[1]) * R - w_x - w_y <= 1) {
            texture_k = texture[(w_y * R + w_x) * 3 + k];
        } else {
            texture_k = texture[((R - 1 - w_y) * R + (R - 1 - w_x)) * 3 + k];
        }
    } else if (texture_sample_type == 1) {
        texture_k = w[0] * texture[k] + w[1] * texture[3 + k] + w[2] * texture[6 + k];
    }
    return texture_k;
}

/* :1154-1174; non-hit texels are UB in the reference, defined as 0 here (Q9) */
static inline float backward_sample_texture(float grad_color, const float *w, int R, int k,
                                            int texture_sample_type) {
    float grad_texture_k = 0;
    if (texture_sample_type == 0) {
        const int w_x = (int)fminf(w[0] * R, (float)(R - 1));
        const int w_y = (int)fminf(w[1] * R, (float)(R - 1));
        if ((w[0] + w[1]) * R - w_x - w_y <= 1) {
            if (k == w_y * R + w_x) grad_texture_k = grad_color;
        } else {
            if (k == (R - 1 - w_y) * R + (R - 1 - w_x)) grad_texture_k = grad_color;
        }
    } else if (texture_sample_type == 1) {
        grad_texture_k = w[k] * grad_color;
    }
    return grad_texture_k;
}

/* :1118-1132 */
static inline void backward_barycentric_p2f_distance(float grad_v[3][3], const float *w,
                                                     const float *face_info, float xp, float yp,
                                                     float dis, float C) {
    const int p = w[0] > w[1] ? (w[1] > w[2] ? 2 : 1) : (w[0] > w[2] ? 2 : 0);
    const float *face_inv = face_info;
    for (int l = 0; l < 2; l++) {
        for (int k = 0; k < 3; k++) {
            float grad_kl = 0;
            for (int q = 0; q < 3; q++) {
                grad_kl += -face_inv[3 * p + l] * face_inv[3 * k + q] * (q == 0 ? xp : (q == 1 ? yp : 1));
            }
            grad_v[k][l] = grad_kl * C;
            grad_v[k][l] = (float)((double)grad_v[k][l] *
                                   (dis > 0 ? (2. * (double)sqrtf(dis)) : (2. * (double)sqrtf(-dis))));
        }
    }
}

/* ------------------------------------------------------------------ K1 :176-236 */
void softras_oracle_face_info(const float *faces, float *faces_info, int batch_size, int num_faces) {
    const long n = (long)batch_size * num_faces;
    memset(faces_info, 0, sizeof(float) * 27 * (size_t)n); /* cudaMemsetAsync(out0_p, 0, ...) :467 */
    for (long i = 0; i < n; i++) {
        const float *face = &faces[i * 9];
        float *face_inv = &faces_info[i * 27];
        float *face_sym = &faces_info[i * 27 + 9];
        float *face_obt = &faces_info[i * 27 + 18];
        float p[3][2];
        for (int num = 0; num < 3; num++)
            for (int dim = 0; dim < 2; dim++) p[num][dim] = face[3 * num + dim];
        float face_inv_star[9] = {
            p[1][1] - p[2][1], p[2][0] - p[1][0], p[1][0] * p[2][1] - p[2][0] * p[1][1],
            p[2][1] - p[0][1], p[0][0] - p[2][0], p[2][0] * p[0][1] - p[0][0] * p[2][1],
            p[0][1] - p[1][1], p[1][0] - p[0][0], p[0][0] * p[1][1] - p[1][0] * p[0][1]};
        float det = (p[2][0] * (p[0][1] - p[1][1]) + p[0][0] * (p[1][1] - p[2][1]) + p[1][0] * (p[2][1] - p[0][1]));
        det = det > 0 ? (float)fmax((double)det, 1e-10) : (float)fmin((double)det, -1e-10);
        for (int k = 0; k < 9; k++) face_inv[k] = face_inv_star[k] / det;
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++)
                face_sym[j * 3 + k] = face[j * 3 + 0] * face[k * 3 + 0] + face[j * 3 + 1] * face[k * 3 + 1] + 1;
        for (int k = 0; k < 3; k++) {
            const int k0 = k, k1 = (k + 1) % 3, k2 = (k + 2) % 3;
            if ((p[k1][0] - p[k0][0]) * (p[k2][0] - p[k0][0]) + (p[k1][1] - p[k0][1]) * (p[k2][1] - p[k0][1]) < 0) {
                face_obt[k0] = 1;
                break;
            }
        }
    }
}

typedef struct {
    int id;
    float z;
} Pixel; /* :238-241 */

/* ------------------------------------------------------------------ K2 :243-456
 * Processes output rows row_begin, row_begin+row_stride, ... < row_end of every batch
 * element (row 0 = top); the full op is row_begin=0, row_end=image_size, row_stride=1.  Buffers must be pre-initialised
 * by softras_oracle_forward below (memsets of :467-470). */
static void forward_rows(const float *faces, const float *textures, const float *faces_info,
                         float *aggrs_info, float *soft_colors, int32_t *faces_id_buffers,
                         int batch_size, int num_faces, int image_size, int max_faces_id,
                         int texture_size, int texture_res, float near, float far, float eps,
                         float sigma_val, int func_id_dist, float dist_eps, float gamma_val,
                         int func_id_rgb, int func_id_alpha, int texture_sample_type,
                         int double_side, int row_begin, int row_end, int row_stride, int nthreads) {
    const int is = image_size;
    const int nf = num_faces;
    const long npix = (long)is * is;
    const float threshold = dist_eps * sigma_val; /* :289 */
    const float border = sqrtf(threshold);        /* :316 */

    /* Hoisted check_border operands (same float adds as :30-33, computed once per face). */
    float *bb = (float *)malloc(sizeof(float) * 4 * (size_t)batch_size * nf);
    for (long f = 0; f < (long)batch_size * nf; f++) {
        const float *face = &faces[f * 9];
        bb[4 * f + 0] = fmaxf(fmaxf(face[0], face[3]), face[6]) + border;
        bb[4 * f + 1] = fminf(fminf(face[0], face[3]), face[6]) - border;
        bb[4 * f + 2] = fmaxf(fmaxf(face[1], face[4]), face[7]) + border;
        bb[4 * f + 3] = fminf(fminf(face[1], face[4]), face[7]) - border;
    }

    if (row_stride < 1) row_stride = 1;
    const long rows_per_image = (row_end - row_begin + row_stride - 1) / row_stride;
    const long rows = (long)batch_size * rows_per_image;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (long r = 0; r < rows; r++) {
        const int bn = (int)(r / rows_per_image);
        const int row = row_begin + (int)(r % rows_per_image) * row_stride;
        for (int col = 0; col < is; col++) {
            const long pn = (long)row * is + col;
            const int yi = is - 1 - row; /* :280 */
            const int xi = col;
            const float yp = (float)((2. * yi + 1. - is) / is);
            const float xp = (float)((2. * xi + 1. - is) / is);

            const float *face = &faces[(long)bn * nf * 9] - 9;
            const float *texture = &textures[(long)bn * nf * texture_size * 3] - texture_size * 3;
            const float *face_info = &faces_info[(long)bn * nf * 27] - 27;
            const float *fbb = &bb[(long)bn * nf * 4] - 4;

            float soft_color[4] = {1.f, 1.f, 1.f, 0.f};
            if (func_id_alpha == 2) soft_color[3] = 1.f;
            float softmax_sum = expf(eps / gamma_val);
            float softmax_max = eps;
            for (int k = 0; k < 3; k++) {
                if (func_id_rgb == 0) soft_color[k] = soft_colors[((long)bn * 4 + k) * npix + pn];
                else if (func_id_rgb == 1) soft_color[k] = soft_colors[((long)bn * 4 + k) * npix + pn] * softmax_sum;
            }
            float depth_min = 10000000;
            int face_index_min = -1;
            Pixel q[K_MAX_POINTS_PER_PIXEL];
            int q_size = 0;
            float q_max_z = -1;
            int q_max_id = -1;

            for (int fn = 0; fn < nf; fn++) {
                face += 9;
                texture += texture_size * 3;
                face_info += 27;
                fbb += 4;

                if (xp > fbb[0] || xp < fbb[1] || yp > fbb[2] || yp < fbb[3]) continue; /* :316 */

                float dis, dis_x, dis_y, t[3], w[3], w_clip[3], sign, soft_fragment;

                barycentric_coordinate(w, xp, yp, face_info);

                if (func_id_dist == 0) {
                    soft_fragment = check_pixel_inside(w) ? 1.f : 0.f;
                    if (soft_fragment == 0.f) continue;
                } else if (func_id_dist == 1) {
                    dis = forward_barycentric_p2f_distance(w);
                    if (-dis >= threshold) continue;
                    soft_fragment = (float)(1. / (1. + (double)expf(-dis / sigma_val)));
                } else {
                    euclidean_p2f_distance(&sign, &dis_x, &dis_y, w, t, face, face_info, xp, yp);
                    dis = dis_x * dis_x + dis_y * dis_y;
                    if (sign < 0 && dis >= threshold) continue;
                    soft_fragment = (float)(1. / (1. + (double)expf(-sign * dis / sigma_val)));
                }

                if (func_id_alpha == 0) {
                    if (soft_fragment > 0.5) soft_color[3] = 1.f;
                } else if (func_id_alpha == 1) {
                    soft_color[3] += soft_fragment;
                } else if (func_id_alpha == 2) {
                    soft_color[3] = (float)((double)soft_color[3] * (1. - (double)soft_fragment));
                }

                for (int k = 0; k < 3; k++) w_clip[k] = w[k];
                barycentric_clip(w_clip);
                const float zp = (float)(1. / (double)(w_clip[0] / face[2] + w_clip[1] / face[5] + w_clip[2] / face[8]));
                if (zp < near || zp > far) continue;

                /* top-K by z, "replace the current max" policy :367-385 */
                if (q_size < max_faces_id) {
                    q[q_size].id = fn;
                    q[q_size].z = zp;
                    if (zp > q_max_z) {
                        q_max_z = zp;
                        q_max_id = q_size;
                    }
                    q_size++;
                } else if (zp < q_max_z) {
                    q[q_max_id].id = fn;
                    q[q_max_id].z = zp;
                    q_max_z = -1;
                    for (int k = 0; k < q_size; k++) {
                        if (q[k].z > q_max_z) {
                            q_max_z = q[k].z;
                            q_max_id = k;
                        }
                    }
                }

                if (func_id_rgb == 0) {
                    if (zp < depth_min && check_pixel_inside(w) && (double_side || check_face_frontside(face))) {
                        depth_min = zp;
                        face_index_min = fn;
                        for (int k = 0; k < 3; k++)
                            soft_color[k] = forward_sample_texture_fwd(texture, w_clip, texture_res, k, texture_sample_type, face, zp);
                    }
                } else if (func_id_rgb == 1) {
                    if (check_face_frontside(face) || double_side) {
                        const float zp_norm = (far - zp) / (far - near);
                        float exp_delta_zp = 1.f;
                        if (zp_norm > softmax_max) {
                            exp_delta_zp = expf((softmax_max - zp_norm) / gamma_val);
                            softmax_max = zp_norm;
                        }
                        const float exp_z = expf((zp_norm - softmax_max) / gamma_val);
                        softmax_sum = exp_delta_zp * softmax_sum + exp_z * soft_fragment;
                        for (int k = 0; k < 3; k++) {
                            const float color_k = forward_sample_texture_fwd(texture, w_clip, texture_res, k, texture_sample_type, face, zp);
                            soft_color[k] = exp_delta_zp * soft_color[k] + exp_z * soft_fragment * color_k;
                        }
                    }
                }
            }

            /* finalise :425-455 */
            if (func_id_alpha == 0) soft_colors[((long)bn * 4 + 3) * npix + pn] = soft_color[3];
            else if (func_id_alpha == 1) soft_colors[((long)bn * 4 + 3) * npix + pn] = soft_color[3] / nf;
            else if (func_id_alpha == 2) soft_colors[((long)bn * 4 + 3) * npix + pn] = (float)(1. - (double)soft_color[3]);

            if (func_id_rgb == 0) {
                if (face_index_min != -1)
                    for (int k = 0; k < 3; k++) soft_colors[((long)bn * 4 + k) * npix + pn] = soft_color[k];
                aggrs_info[((long)bn * 2 + 0) * npix + pn] = depth_min;
                aggrs_info[((long)bn * 2 + 1) * npix + pn] = (float)face_index_min;
            } else if (func_id_rgb == 1) {
                for (int k = 0; k < 3; k++) soft_colors[((long)bn * 4 + k) * npix + pn] = soft_color[k] / softmax_sum;
                aggrs_info[((long)bn * 2 + 0) * npix + pn] = softmax_sum;
                aggrs_info[((long)bn * 2 + 1) * npix + pn] = softmax_max;
            }
            for (int k = 0; k < q_size; k++)
                faces_id_buffers[((long)bn * max_faces_id + k) * npix + pn] = q[k].id;
        }
    }
    free(bb);
}

/* Full forward op = the reference's cuda_src :467-516: memsets, K1, K2.
 * faces_id_buffer layout [B, K, H, W] (as the forward writes it, :454). */
void softras_oracle_forward(const float *faces, const float *textures, float *faces_info,
                            float *aggrs_info, float *soft_colors, int32_t *faces_id_buffer,
                            int batch_size, int num_faces, int texture_size, int image_size,
                            int max_faces_id, float near, float far, float eps, float sigma_val,
                            int func_id_dist, float dist_eps, float gamma_val, int func_id_rgb,
                            int func_id_alpha, int texture_sample_type, int double_side,
                            int row_begin, int row_end, int row_stride, int nthreads) {
    const size_t npix = (size_t)image_size * image_size;
    memset(aggrs_info, 0, sizeof(float) * 2 * npix * batch_size);
    memset(soft_colors, 0, sizeof(float) * 4 * npix * batch_size); /* background is always 0 (Q1) */
    memset(faces_id_buffer, 0xFF, sizeof(int32_t) * (size_t)max_faces_id * npix * batch_size);
    softras_oracle_face_info(faces, faces_info, batch_size, num_faces);
    const int texture_res = (int)sqrt((double)texture_size); /* :475 */
    forward_rows(faces, textures, faces_info, aggrs_info, soft_colors, faces_id_buffer, batch_size,
                 num_faces, image_size, max_faces_id, texture_size, texture_res, near, far, eps,
                 sigma_val, func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha,
                 texture_sample_type, double_side, row_begin, row_end, row_stride, nthreads);
}

/* ------------------------------------------------------------------ K6 :1177-1360
 * Sequential restatement: pixels in ascending linear index, slots m ascending,
 * `+=` in place of atomicAdd.  accumulate_double != 0 accumulates in double and
 * rounds once at the end (order-free centre for tolerance tests); 0 accumulates
 * in float exactly like a serialised execution of the reference.
 * faces_id_buffer layout [B, K, H, W] (the reference transposes to [B,H,W,K] on
 * the host first, soft_rasterize.py:108; the values read are the same). */
void softras_oracle_backward(const float *faces, const float *textures, const float *soft_colors,
                             const float *faces_info, const float *aggrs_info,
                             const int32_t *faces_id_buffer, const float *grad_soft_colors,
                             float *grad_faces, float *grad_textures, int batch_size,
                             int num_faces, int texture_size, int image_size, int max_faces_id,
                             float near, float far, float eps, float sigma_val, int func_id_dist,
                             float dist_eps, float gamma_val, int func_id_rgb, int func_id_alpha,
                             int texture_sample_type, int double_side, int accumulate_double,
                             int row_begin, int row_end, int row_stride) {
    (void)eps;
    (void)double_side;
    const int is = image_size;
    const int nf = num_faces;
    const int mf = max_faces_id;
    const long npix = (long)is * is;
    const int texture_res = (int)sqrt((double)texture_size);
    const float threshold = dist_eps * sigma_val;
    const size_t ngf = (size_t)batch_size * nf * 9;
    const size_t ngt = (size_t)batch_size * nf * texture_size * 3;
    double *gf = (double *)calloc(ngf, sizeof(double));
    double *gt = (double *)calloc(ngt, sizeof(double));
#define ACC(buf, idx, val)                                                   \
    do {                                                                     \
        if (accumulate_double) buf[idx] += (double)(val);                    \
        else buf[idx] = (double)((float)buf[idx] + (float)(val));            \
    } while (0)

    for (int bn = 0; bn < batch_size; bn++)
        for (int row = row_begin; row < row_end; row += (row_stride < 1 ? 1 : row_stride))
            for (int col = 0; col < is; col++) {
                const long pn = (long)row * is + col;
                const int yi = is - 1 - row;
                const int xi = col;
                const float yp = (float)((2. * yi + 1 - is) / is);
                const float xp = (float)((2. * xi + 1 - is) / is);

                const float softmax_sum = aggrs_info[((long)bn * 2 + 0) * npix + pn];
                const float softmax_max = aggrs_info[((long)bn * 2 + 1) * npix + pn];

                for (int m = 0; m < mf; m++) {
                    const int fn = faces_id_buffer[((long)bn * mf + m) * npix + pn];
                    if (fn == -1) break;

                    const float *face = &faces[((long)bn * nf + fn) * 9];
                    const float *texture = &textures[((long)bn * nf + fn) * texture_size * 3];
                    const float *face_info = &faces_info[((long)bn * nf + fn) * 27];

                    if (check_border(xp, yp, face, sqrtf(threshold))) continue;

                    float dis = 0, dis_x = 0, dis_y = 0, t[3] = {0, 0, 0}, w[3], w0[3], sign = 0, soft_fragment;

                    barycentric_coordinate(w, xp, yp, face_info);

                    if (func_id_dist == 0) {
                        soft_fragment = 1;
                    } else if (func_id_dist == 1) {
                        dis = forward_barycentric_p2f_distance(w);
                        for (int k = 0; k < 3; k++) t[k] = w[k];
                        soft_fragment = (float)(1. / (1. + (double)expf(-dis / sigma_val)));
                    } else {
                        euclidean_p2f_distance(&sign, &dis_x, &dis_y, w, t, face, face_info, xp, yp);
                        dis = dis_x * dis_x + dis_y * dis_y;
                        soft_fragment = (float)(1. / (1. + (double)expf(-sign * dis / sigma_val)));
                    }

                    const size_t gfo = ((size_t)bn * nf + fn) * 9;
                    const size_t gto = ((size_t)bn * nf + fn) * texture_size * 3;
                    float grad_v[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
                    float C_grad_xy = 0;

                    float C_grad_xy_alpha = grad_soft_colors[((long)bn * 4 + 3) * npix + pn];
                    if (func_id_alpha == 0) {
                    } else if (func_id_alpha == 1) {
                        C_grad_xy_alpha /= nf;
                    } else if (func_id_alpha == 2) {
                        C_grad_xy_alpha = (float)((double)C_grad_xy_alpha *
                                                  ((double)(1 - soft_colors[((long)bn * 4 + 3) * npix + pn]) /
                                                   fmax((double)(1 - soft_fragment), 1e-6)));
                    }
                    C_grad_xy += C_grad_xy_alpha;

                    for (int k = 0; k < 3; k++) w0[k] = w[k];
                    barycentric_clip(w);
                    const float zp = (float)(1. / (double)(w[0] / face[2] + w[1] / face[5] + w[2] / face[8]));

                    if (func_id_rgb == 0) {
                        if ((float)fn == softmax_max) { /* int vs float compare, Q10 */
                            for (int k = 0; k < 3; k++)
                                for (int j = 0; j < texture_size; j++)
                                    ACC(gt, gto + 3 * j + k,
                                        backward_sample_texture(grad_soft_colors[((long)bn * 4 + k) * npix + pn], w, texture_res, j, texture_sample_type));
                        }
                    } else if (func_id_rgb == 1) {
                        float C_grad_xyz_rgb = 0.f;
                        const float zp_norm = (far - zp) / (far - near);
                        const float zp_softmax = soft_fragment * expf((zp_norm - softmax_max) / gamma_val) / softmax_sum;

                        for (int k = 0; k < 3; k++) {
                            const float grad_soft_color_k = grad_soft_colors[((long)bn * 4 + k) * npix + pn];
                            for (int j = 0; j < texture_size; j++) {
                                const float grad_t = backward_sample_texture(grad_soft_color_k, w, texture_res, j, texture_sample_type);
                                ACC(gt, gto + 3 * j + k, zp_softmax * grad_t);
                            }
                            const float color_k = forward_sample_texture_bwd(texture, w, texture_res, k, texture_sample_type);
                            C_grad_xyz_rgb += grad_soft_color_k * (color_k - soft_colors[((long)bn * 4 + k) * npix + pn]);
                        }
                        C_grad_xyz_rgb *= zp_softmax;
                        C_grad_xy += C_grad_xyz_rgb / soft_fragment;

                        const float C_grad_z_rgb = C_grad_xyz_rgb / gamma_val / (near - far) * zp * zp;
                        grad_v[0][2] = C_grad_z_rgb * w[0] / face[2] / face[2];
                        grad_v[1][2] = C_grad_z_rgb * w[1] / face[5] / face[5];
                        grad_v[2][2] = C_grad_z_rgb * w[2] / face[8] / face[8];
                    }

                    C_grad_xy *= soft_fragment * (1 - soft_fragment) / sigma_val;
                    if (func_id_dist == 1) {
                        backward_barycentric_p2f_distance(grad_v, t, face_info, xp, yp, dis, C_grad_xy);
                    } else if (func_id_dist == 2) {
                        for (int k = 0; k < 3; k++)
                            for (int l = 0; l < 2; l++)
                                grad_v[k][l] = 2 * sign * C_grad_xy * (t[k] + w0[k]) * (l == 0 ? dis_x : dis_y);
                    }

                    ACC(gf, gfo + 0, grad_v[0][0]);
                    ACC(gf, gfo + 1, grad_v[0][1]);
                    ACC(gf, gfo + 3, grad_v[1][0]);
                    ACC(gf, gfo + 4, grad_v[1][1]);
                    ACC(gf, gfo + 6, grad_v[2][0]);
                    ACC(gf, gfo + 7, grad_v[2][1]);
                    ACC(gf, gfo + 2, grad_v[0][2]);
                    ACC(gf, gfo + 5, grad_v[1][2]);
                    ACC(gf, gfo + 8, grad_v[2][2]);
                }
            }
#undef ACC
    for (size_t i = 0; i < ngf; i++) grad_faces[i] = (float)gf[i];
    for (size_t i = 0; i < ngt; i++) grad_textures[i] = (float)gt[i];
    free(gf);
    free(gt);
}

/* Multi-threaded timing variant for the CPU baseline (bench.py): thread t runs the sequential
 * restatement above on rows row_begin + t*stride, + nthreads*stride, ... into private
 * gradient buffers (double accumulation), which are then summed -- SURVEY.md 8(d)'s
 * "per-thread private grad buffers reduced at the end".  Not used by the parity tests. */
void softras_oracle_backward_mt(const float *faces, const float *textures, const float *soft_colors,
                                const float *faces_info, const float *aggrs_info,
                                const int32_t *faces_id_buffer, const float *grad_soft_colors,
                                float *grad_faces, float *grad_textures, int batch_size,
                                int num_faces, int texture_size, int image_size, int max_faces_id,
                                float near, float far, float eps, float sigma_val, int func_id_dist,
                                float dist_eps, float gamma_val, int func_id_rgb, int func_id_alpha,
                                int texture_sample_type, int double_side,
                                int row_begin, int row_end, int row_stride, int nthreads) {
    const size_t ngf = (size_t)batch_size * num_faces * 9;
    const size_t ngt = (size_t)batch_size * num_faces * texture_size * 3;
    if (row_stride < 1) row_stride = 1;
    if (nthreads < 1) nthreads = 1;
    float *pf = (float *)calloc(ngf * nthreads, sizeof(float));
    float *pt = (float *)calloc(ngt * nthreads, sizeof(float));
#pragma omp parallel for schedule(static, 1) num_threads(nthreads)
    for (int t = 0; t < nthreads; t++)
        softras_oracle_backward(faces, textures, soft_colors, faces_info, aggrs_info, faces_id_buffer,
                                grad_soft_colors, pf + ngf * t, pt + ngt * t, batch_size, num_faces,
                                texture_size, image_size, max_faces_id, near, far, eps, sigma_val,
                                func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha,
                                texture_sample_type, double_side, 1, row_begin + t * row_stride, row_end,
                                row_stride * nthreads);
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (long i = 0; i < (long)ngf; i++) {
        double a = 0;
        for (int t = 0; t < nthreads; t++) a += pf[ngf * t + i];
        grad_faces[i] = (float)a;
    }
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (long i = 0; i < (long)ngt; i++) {
        double a = 0;
        for (int t = 0; t < nthreads; t++) a += pt[ngt * t + i];
        grad_textures[i] = (float)a;
    }
    free(pf);
    free(pt);
}

int softras_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// softras_forward2.cuh -- two-phase SoftRas forward (replaces K2 / K5 of the reference:
// forward_soft_rasterize_cuda_kernel, cuda/soft_rasterize.py:243-456 and
// cuda/soft_rasterize_coarse_to_fine.py:513-761).
//
// One warp = one CTA = one 8x4 pixel block at a time, pulled from the cost-ordered tile queue (k_tile_order).  The
// reference's per-pixel loop `for (fn = 0; fn < nf; fn++)` (:311) has an order-INDEPENDENT part -- barycentric
// coordinates, point-to-face distance, sigmoid, clipped-barycentric 1/z (:318-364; ~250 instructions) -- and a cheap
// order-DEPENDENT part -- alpha, top-K insertion, online softmax (:349-419).  The first version ran both per pixel
// with the lanes walking private face lists in lock-step: 19 of 32 lanes busy (uneven list lengths, and the rare
// "pixel strictly inside the face" path of euclidean_p2f_distance executed by ~3 lanes while 29 waited).  Here:
//
//   stage   the next <= R face records of the block's list: one cp.async.bulk (TMA) per record into shared memory,
//           completion on an mbarrier (no register round trip, no per-record index arithmetic);
//   list    every lane turns its pixel's face mask into a run of the block's PAIR LIST (pixel-major, so a pixel's
//           pairs are contiguous and in ascending face id);
//   phase A walks that list 32 pairs at a time, all lanes busy whatever the per-pixel list lengths: geometry of
//           (pixel, face) -> (D, zp) in shared memory.  Pairs whose pixel is strictly inside the face are set aside
//           and evaluated together afterwards (uniform branch: no 3-lane detours);
//   phase B each lane folds its own pixel's results in ascending face id: alpha, near/far, top-K, softmax.
//
// Results are bit-identical to the first version's (same device functions, same order per pixel).
#pragma once
#include "softras_forward.cuh"

namespace b200r {

#ifndef B200R_F2_R
#define B200R_F2_R 16        // face records staged per round
#endif
#ifndef B200R_F2_PMAX
#define B200R_F2_PMAX 256    // (pixel, face) pairs per round (a round is cut short when its faces cover more)
#endif
#ifndef B200R_F2_MINB
#define B200R_F2_MINB 20     // resident one-warp CTAs per SM the register allocation must allow
#endif
#ifndef B200R_F2_DEFER_INSIDE
#define B200R_F2_DEFER_INSIDE 1
#endif

struct __align__(16) Fwd2Smem {
    FaceRecS rec[B200R_F2_R];                    // staged records (16-byte aligned: bulk-copy destination); reused as the output staging area
    __align__(16) uint32_t lmask[B200R_F2_R];    // per staged record: the lanes (pixels) inside its rectangle (read as uint4)
    __align__(16) float2 res[B200R_F2_PMAX];     // pair -> (D, zp);  D == -1: the pair contributes nothing (:333/:337/:343)
    unsigned long long mbar;
    int ids[B200R_F2_R + 4 * 32];                // pending block-face ids, ascending
    unsigned short plist[B200R_F2_PMAX];         // pair -> (record slot << 5) | lane, pixel-major
    unsigned short ilist[B200R_F2_PMAX];         // pairs set aside for the "strictly inside" distance path
};
static_assert(sizeof(Fwd2Smem) % 16 == 0, "the top-K depth lists behind Fwd2Smem are read with 16-byte loads");

static inline size_t fwd2_smem_bytes(int K) {
    const size_t b = sizeof(Fwd2Smem) + (size_t)32 * (fwd_qz_stride(K) + K) * 4;
    return (b + 15) & ~(size_t)15;
}

// Geometry of one (pixel, face) pair, order-independent part of the reference's loop body (:328-364): soft fragment D
// (distance -> sigmoid) and the clipped-barycentric depth zp.  Returns false when the pair is skipped by the inside
// test (hard, :332), or the distance threshold (:337 / :343).  Guarded divisions (EXACT as in shade_face).
template <int DIST, bool EXACT>
__device__ __forceinline__ bool pair_geometry(const FaceRec* rec, const SoftRasParams& P, const DivConst& dc, float xp, float yp,
                                              const float w[3], float threshold, float& D, float& zp) {
    if (DIST == 0) {
        if (!check_pixel_inside(w)) return false;
        D = 1.f;
    } else if (DIST == 1) {
        const float dis = barycentric_p2f_distance(w);
        if (-dis >= threshold) return false;
        D = sigmoid_from_negarg<EXACT>(dc.template by_sigma_t<EXACT>(-dis));
    } else {
        float dis_x, dis_y, t[3];
        const float sign = euclidean_p2f_distance(dis_x, dis_y, t, w, rec, xp, yp);
        const float dis = dis_x * dis_x + dis_y * dis_y;
        if (sign < 0.f && dis >= threshold) return false;
        D = sigmoid_from_negarg<EXACT>(dc.template by_sigma_t<EXACT>(-sign * dis));
    }
    float wc[3] = {w[0], w[1], w[2]};
    zp = clip_and_z(wc, rec);
    return true;
}

// The same with branch-free optimistic divisions (exact_math.cuh, DivGuard); when the guard drops the caller re-runs
// the pair with pair_geometry<DIST, false>, whose intermediates have the same bits whenever the guard holds.
template <int DIST>
__device__ __forceinline__ bool pair_geometry_opt(const FaceRec* rec, const DivConst& dc, float xp, float yp, const float w[3],
                                                  float threshold, float& D, float& zp, DivGuard& g) {
    D = 1.f;
    if (DIST == 0) {
        if (!check_pixel_inside(w)) return false;
    } else if (DIST == 1) {
        const float dis = barycentric_p2f_distance(w);
        if (-dis >= threshold) return false;
        D = sigmoid_from_negarg_opt(dc.by_sigma_o(-dis, g), g);
    } else {
        float dis_x, dis_y, t[3];
        const float sign = euclidean_p2f_distance<true>(dis_x, dis_y, t, w, rec, xp, yp, &g);
        const float dis = dis_x * dis_x + dis_y * dis_y;
        if (sign < 0.f && dis >= threshold && g.ok) return false;   // an out-of-range pair falls through to the re-run
        D = sigmoid_from_negarg_opt(dc.by_sigma_o(-sign * dis, g), g);
    }
    float wc[3] = {w[0], w[1], w[2]};
    zp = clip_and_z_opt<true>(wc, rec, g);
    return true;
}

// Order-dependent part of the loop body for one pair (:349-419): alpha, near/far, top-K, colour aggregation.
template <int RGB, bool EXACT>
__device__ __forceinline__ void fold_pair(const FaceRec* rec, PixState& st, const SoftRasParams& P, const DivConst& dc, bool consts_ok,
                                          float D, float zp, float xp, float yp, float* my_qz, int* s_qid, int tpix,
                                          const float* __restrict__ btex) {
    // alpha aggregation, before any z test (:349-358, Q2); the default ('prod') is tested first
    if (P.alpha_func == 2) st.alpha = alpha_prod_t<EXACT>(st.alpha, D);
    else if (P.alpha_func == 1) st.alpha += D;
    else if (D > 0.5f) st.alpha = 1.f;
    if (zp < P.near_ || zp > P.far_) return;   // :365
    const int fn = (int)rec->face_id;
    topk_insert<32>(st, zp, fn, P.K, my_qz, s_qid, tpix);
    if (RGB == 2) return;
    const bool front = (rec->flags & 8u) != 0;
    if (RGB == 0) {   // :390-397
        if (zp < st.depth_min && (P.double_side || front)) {
            float w[3];
            barycentric_coordinate(w, xp, yp, rec->inv);
            if (check_pixel_inside(w)) {
                st.depth_min = zp;
                st.face_index_min = fn;
                float col[3];
                if (P.tex_type == 0 && P.R == 1) {
                    col[0] = rec->col[0]; col[1] = rec->col[1]; col[2] = rec->col[2];
                } else {
                    clip_and_z(w, rec);   // w := barycentric_clip(w) (:363), the same bits phase A used
                    sample_texture_fwd(col, btex, fn, P.T, w, P.R, P.tex_type, rec, zp);
                }
                st.sc0 = col[0]; st.sc1 = col[1]; st.sc2 = col[2];
            }
        }
        return;
    }
    if (!(front || P.double_side)) return;   // :400
    float col[3];
    if (P.tex_type == 0 && P.R == 1) {
        col[0] = rec->col[0]; col[1] = rec->col[1]; col[2] = rec->col[2];
    } else {
        float w[3];
        barycentric_coordinate(w, xp, yp, rec->inv);
        clip_and_z(w, rec);
        sample_texture_fwd(col, btex, fn, P.T, w, P.R, P.tex_type, rec, zp);
    }
    if constexpr (!EXACT) {
        DivGuard g;
        g.ok = consts_ok;
        const float zp_norm = dc.by_span_o(P.far_ - zp, g);
        const float dz = zp_norm - st.softmax_max;
        const float ex = expf(dc.by_gamma_o(-fabsf(dz), g));   // = exp_delta_zp when dz > 0, exp_z otherwise (:402-407)
        if (g.ok) {
            const bool up = dz > 0.f;
            const float exp_delta_zp = up ? ex : 1.f;
            const float exp_z = up ? 1.f : ex;   // exp(0 / gamma) == 1 exactly
            if (up) st.softmax_max = zp_norm;
            st.softmax_sum = exp_delta_zp * st.softmax_sum + exp_z * D;
            st.sc0 = exp_delta_zp * st.sc0 + exp_z * D * col[0];
            st.sc1 = exp_delta_zp * st.sc1 + exp_z * D * col[1];
            st.sc2 = exp_delta_zp * st.sc2 + exp_z * D * col[2];
            return;
        }
    }
    const float zp_norm = dc.template by_span_t<EXACT>(P.far_ - zp);
    float exp_delta_zp = 1.f;
    if (zp_norm > st.softmax_max) {
        exp_delta_zp = expf(dc.template by_gamma_t<EXACT>(st.softmax_max - zp_norm));
        st.softmax_max = zp_norm;
    }
    const float exp_z = expf(dc.template by_gamma_t<EXACT>(zp_norm - st.softmax_max));
    st.softmax_sum = exp_delta_zp * st.softmax_sum + exp_z * D;
    st.sc0 = exp_delta_zp * st.sc0 + exp_z * D * col[0];
    st.sc1 = exp_delta_zp * st.sc1 + exp_z * D * col[1];
    st.sc2 = exp_delta_zp * st.sc2 + exp_z * D * col[2];
}

template <int DIST, int RGB, bool EXACT>
__global__ void __launch_bounds__(32, B200R_F2_MINB)
k_softras_forward2(const SoftRasParams P, const FaceRec* __restrict__ recs, const uint2* __restrict__ rects,
                   const int* __restrict__ coarse_cnt, const int* __restrict__ coarse_ids,
                   const float* __restrict__ textures, float* __restrict__ soft_colors,
                   float* __restrict__ aggrs_info, int* __restrict__ ids_out, int* tile_counter,
                   const int* __restrict__ tile_order, float* __restrict__ pooled) {
    constexpr int R = B200R_F2_R, PMAX = B200R_F2_PMAX, UNR = 4, TW = 8, TH = 4, NT = 32;
    static_assert(R <= 32 && PMAX >= 32 * 8 && PMAX <= 32 * R, "round geometry");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Fwd2Smem& S = *reinterpret_cast<Fwd2Smem*>(smem_raw);
    const int qzs = fwd_qz_stride(P.K);
    float* s_qz_all = reinterpret_cast<float*>(smem_raw + sizeof(Fwd2Smem));   // [32][qzs]
    int* s_qid = reinterpret_cast<int*>(s_qz_all + (size_t)NT * qzs);          // [K][32]
    const int lane = threadIdx.x;
    float* my_qz = s_qz_all + (size_t)lane * qzs;
    const int is = P.is, nf = P.nf, K = P.K;
    const int tiles_per_image = P.fntx * P.fnty;
    const int lx = lane & 7, ly = lane >> 3;
    const float threshold = P.dist_eps * P.sigma;  // :289
    const size_t npix = (size_t)is * is;
    DivConst dc;
    dc.init(P);
    const bool consts_ok = dc.consts_ok();
    const float softmax_sum0 = expf(P.eps / P.gamma);
    const uint32_t lt_mask = (1u << lane) - 1u;
    if (lane == 0) f2_mbar_init(&S.mbar, 1);
    __syncwarp();
    uint32_t parity = 0;

    for (int titer = 0;; titer++) {
        // ---- which tile
        int t;
        if (tile_counter == nullptr) {
            if (titer > 0) break;
            t = blockIdx.y * tiles_per_image + blockIdx.x;
        } else {
            __syncwarp();  // previous tile fully written
            int q = 0;
            if (lane == 0) q = atomicAdd(tile_counter, 1);
            q = __shfl_sync(0xffffffffu, q, 0);
            if (q >= P.queue_len) break;
            t = __ldg(tile_order + q);  // most expensive tiles first (k_tile_order)
            if (t < 0) continue;
        }
        const int b = t / tiles_per_image;
        const int tt = t - b * tiles_per_image;
        const int tx = tt % P.fntx, ty = tt / P.fntx;
        const int tx0 = tx * TW, tr0 = ty * TH, tx1 = tx0 + TW - 1, tr1 = tr0 + TH - 1;
        const int px = tx0 + lx, row = tr0 + ly;
        const float xp = b200r_pix_coord(px, is);
        const float yp = b200r_pix_coord(is - 1 - row, is);

        // ---- per-pixel state, initialised as :291-309 (background buffer is all zero, Q1)
        PixState st;
        st.softmax_sum = softmax_sum0;
        st.softmax_max = P.eps;
        if (RGB == 0) { st.sc0 = st.sc1 = st.sc2 = 0.f; }
        else if (RGB == 1) { st.sc0 = st.sc1 = st.sc2 = 0.f * softmax_sum0; }
        else { st.sc0 = st.sc1 = st.sc2 = 1.f; }
        st.alpha = (P.alpha_func == 2) ? 1.f : 0.f;
        st.depth_min = 10000000.f;
        st.face_index_min = -1;
        st.q_size = 0;
        st.q_max_z = -1.f;
        st.q_max_id = -1;
        const bool vec_ids = ((is & 3) == 0) && ((K & 3) == 0);
        if (vec_ids) {
            for (int j = lane; j < K * NT / 4; j += NT) reinterpret_cast<int4*>(s_qid)[j] = make_int4(-1, -1, -1, -1);
        }

        const int cbin = (tr0 / P.coarse_px) * P.ncs + (tx0 / P.coarse_px);
        const int n_coarse = coarse_cnt[b * P.ncs * P.ncs + cbin];
        const int* clist = coarse_ids + ((size_t)b * P.ncs * P.ncs + cbin) * nf;
        const uint2* brects = rects + (size_t)b * nf;
        const FaceRec* brecs = recs + (size_t)b * nf;
        const float* btex = textures + (size_t)b * nf * P.T * 3;

        int n_pending = 0, head = 0;  // pending ids live in S.ids[head, head + n_pending); warp-uniform
        for (int base = 0; base < n_coarse; base += NT * UNR) {
            // ---- leftover of the previous pass (< R entries) back to the front
            if (head > 0) {
                int keep = 0;
                if (lane < n_pending) keep = S.ids[head + lane];
                __syncwarp();
                if (lane < n_pending) S.ids[lane] = keep;
                head = 0;
            }
            // ---- fine filter: next UNR*32 coarse entries -> S.ids (ordered).  All id loads are issued before the
            // dependent rectangle gathers so one pass costs two memory latencies, not 2*UNR.
            {
                int id[UNR];
                uint2 rc[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    const int i = base + u * NT + lane;
                    id[u] = (i < n_coarse) ? __ldg(clist + i) : -1;
                }
#pragma unroll
                for (int u = 0; u < UNR; u++) rc[u] = (id[u] >= 0) ? __ldg(brects + id[u]) : make_uint2(1u, 1u);
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    const bool pass = id[u] >= 0 && rect_overlaps(rc[u], tx0, tx1, tr0, tr1);
                    const unsigned bal = __ballot_sync(0xffffffffu, pass);
                    if (pass) S.ids[n_pending + __popc(bal & lt_mask)] = id[u];
                    n_pending += __popc(bal);
                }
            }
            const bool last = base + NT * UNR >= n_coarse;

            while (n_pending >= R || (last && n_pending > 0)) {
                int m = min(n_pending, R);
                __syncwarp();  // S.ids complete; previous round's readers of S.rec / S.res done
                // ---- stage m records: one bulk copy each, completion counted in bytes on the mbarrier
                f2_fence_async_smem();
                if (lane == 0) f2_mbar_expect_tx(&S.mbar, (uint32_t)m * (uint32_t)sizeof(FaceRec));
                __syncwarp();
                if (lane < m) f2_bulk_g2s(&S.rec[lane].r, brecs + S.ids[head + lane], (uint32_t)sizeof(FaceRec), &S.mbar);
                f2_mbar_wait(&S.mbar, parity);
                parity ^= 1u;

                // ---- lane j: record j's rectangle, clipped to the 8x4 block, as the set of covered lanes
                uint32_t lm = 0u;
                if (lane < m) {
                    const uint32_t rx = S.rec[lane].r.rect_x, rr = S.rec[lane].r.rect_r;
                    const int cx0 = max((int)(rx & 0xffffu) - tx0, 0), cx1 = min((int)(rx >> 16) - tx0, 7);
                    const int ry0 = max((int)(rr & 0xffffu) - tr0, 0), ry1 = min((int)(rr >> 16) - tr0, 3);
                    if (cx0 <= cx1 && ry0 <= ry1) {
                        const uint32_t cols = (2u << cx1) - (1u << cx0);
                        const uint32_t rows = (0xffffffffu >> (24 - 8 * ry1)) & (0xffffffffu << (8 * ry0));
                        lm = (cols * 0x01010101u) & rows;
                    }
                }
                // a round evaluates at most PMAX pairs: keep the longest prefix of faces that fits (>= 8: 8 x 32 <= PMAX)
                int total = __reduce_add_sync(0xffffffffu, __popc(lm));
                if (total > PMAX) {
                    int incl = __popc(lm);
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int n = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += n;
                    }
                    m = __popc(__ballot_sync(0xffffffffu, lane < m && incl <= PMAX));
                    if (lane >= m) lm = 0u;
                    total = __reduce_add_sync(0xffffffffu, __popc(lm));
                }
                if (lane < R) S.lmask[lane] = lm;
                __syncwarp();
                // ---- every lane: its pixel's faces as a bit mask over the staged records (ascending record = ascending id)
                uint32_t mask = 0u;
#pragma unroll
                for (int q = 0; q < R / 4; q++) {
                    const uint4 m4 = reinterpret_cast<const uint4*>(S.lmask)[q];
                    mask |= ((m4.x >> lane) & 1u) << (4 * q + 0);
                    mask |= ((m4.y >> lane) & 1u) << (4 * q + 1);
                    mask |= ((m4.z >> lane) & 1u) << (4 * q + 2);
                    mask |= ((m4.w >> lane) & 1u) << (4 * q + 3);
                }
                const int cnt = __popc(mask);
                int off = cnt;   // inclusive scan -> this pixel's run of the pair list is [off - cnt, off)
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int n = __shfl_up_sync(0xffffffffu, off, d);
                    if (lane >= d) off += n;
                }
                off -= cnt;
                {
                    uint32_t mm = mask;
                    int o = off;
                    while (mm != 0u) {
                        const int j = __ffs(mm) - 1;
                        mm &= mm - 1u;
                        S.plist[o++] = (unsigned short)((j << 5) | lane);
                    }
                }
                __syncwarp();

                // ---- phase A: all pairs, 32 at a time.  mode 0: the pair list (pairs strictly inside their face are set
                // aside when DIST is euclidean); mode 1: the pairs set aside.
                int n_inside = 0;
#pragma unroll 1
                for (int mode = 0; mode < 2; mode++) {
                    const int n_items = (mode == 0) ? total : n_inside;
#pragma unroll 1
                    for (int q0 = 0; q0 < n_items; q0 += 32) {
                        const int q = q0 + lane;
                        bool valid = q < n_items;
                        int p = 0;
                        if (valid) p = (mode == 0) ? q : (int)S.ilist[q];
                        const int e = S.plist[p];
                        const int l = e & 31;
                        const FaceRec* rec = &S.rec[e >> 5].r;
                        const float pxp = __shfl_sync(0xffffffffu, xp, l), pyp = __shfl_sync(0xffffffffu, yp, l);
                        float w[3];
                        barycentric_coordinate(w, pxp, pyp, rec->inv);
                        if (DIST == 2 && B200R_F2_DEFER_INSIDE && mode == 0) {
                            const bool inside = valid && w[0] > 0.f && w[1] > 0.f && w[2] > 0.f && w[0] < 1.f && w[1] < 1.f && w[2] < 1.f;
                            const unsigned bal = __ballot_sync(0xffffffffu, inside);
                            if (inside) S.ilist[n_inside + __popc(bal & lt_mask)] = (unsigned short)p;
                            n_inside += __popc(bal);
                            valid = valid && !inside;
                        }
                        if (valid) {
                            float D, zp = 0.f;
                            bool keep;
                            if constexpr (!EXACT) {
                                DivGuard g;
                                g.ok = consts_ok;
                                keep = pair_geometry_opt<DIST>(rec, dc, pxp, pyp, w, threshold, D, zp, g);
                                if (!g.ok) keep = pair_geometry<DIST, false>(rec, P, dc, pxp, pyp, w, threshold, D, zp);   // cold
                            } else {
                                keep = pair_geometry<DIST, true>(rec, P, dc, pxp, pyp, w, threshold, D, zp);
                            }
                            S.res[p] = make_float2(keep ? D : -1.f, zp);
                        }
                    }
                    __syncwarp();
                    if (!(DIST == 2 && B200R_F2_DEFER_INSIDE)) break;
                }

                // ---- phase B: every lane folds its pixel's pairs in ascending face id
                {
                    const int maxcnt = __reduce_max_sync(0xffffffffu, cnt);
                    uint32_t mm = mask;
                    const float2* my_res = S.res + off;
#pragma unroll 1
                    for (int i = 0; i < maxcnt; i++) {
                        if (mm != 0u) {
                            const FaceRec* rec = &S.rec[__ffs(mm) - 1].r;
                            mm &= mm - 1u;
                            const float2 r = my_res[i];
                            if (r.x != -1.f)
                                fold_pair<RGB, EXACT>(rec, st, P, dc, consts_ok, r.x, r.y, xp, yp, my_qz, s_qid, lane, btex);
                        }
                    }
                }
                head += m;
                n_pending -= m;
            }
        }

        // ---- finalise (:425-455)
        float out_a;
        if (P.alpha_func == 0) out_a = st.alpha;
        else if (P.alpha_func == 1) out_a = st.alpha / (float)nf;
        else out_a = (float)(1.0 - (double)st.alpha);
        float o0, o1, o2, g0, g1;
        if (RGB == 0) {
            o0 = st.sc0; o1 = st.sc1; o2 = st.sc2;  // stays at the (zero) background when no face was hit
            g0 = st.depth_min; g1 = (float)st.face_index_min;
        } else if (RGB == 1) {
            o0 = st.sc0 / st.softmax_sum; o1 = st.sc1 / st.softmax_sum; o2 = st.sc2 / st.softmax_sum;
            g0 = st.softmax_sum; g1 = st.softmax_max;
        } else {
            o0 = o1 = o2 = 0.f; g0 = g1 = 0.f;
        }

        // Stage the 6 output planes of the block in shared memory and write each plane row with 16-byte stores.
        __syncwarp();
        float* s_out = reinterpret_cast<float*>(S.rec);  // 6 * 32 floats
        s_out[0 * NT + lane] = o0;
        s_out[1 * NT + lane] = o1;
        s_out[2 * NT + lane] = o2;
        s_out[3 * NT + lane] = out_a;
        s_out[4 * NT + lane] = g0;
        s_out[5 * NT + lane] = g1;
        __syncwarp();
        if (pooled != nullptr) store_pooled_8x4(s_out, pooled, b, tx0, tr0, is, lane);   // anti-aliasing epilogue
        if ((is & 3) == 0) {
            constexpr int QPR = TW / 4;  // float4 per block row
            for (int j = lane; j < 6 * TH * QPR; j += NT) {
                const int ch = j / (TH * QPR), r = (j % (TH * QPR)) / QPR, q = j % QPR;
                const int orow = tr0 + r, ocol = tx0 + q * 4;
                if (orow < is && ocol < is) {
                    const float4 v = *reinterpret_cast<const float4*>(&s_out[ch * NT + r * TW + q * 4]);
                    float* dst = (ch < 4) ? soft_colors + ((size_t)b * 4 + ch) * npix
                                          : aggrs_info + ((size_t)b * 2 + (ch - 4)) * npix;
                    *reinterpret_cast<float4*>(dst + (size_t)orow * is + ocol) = v;
                }
            }
        } else {
            for (int j = lane; j < 6 * NT; j += NT) {
                const int ch = j / NT, r = (j % NT) / TW, c = j % TW;
                const int orow = tr0 + r, ocol = tx0 + c;
                if (orow < is && ocol < is) {
                    float* dst = (ch < 4) ? soft_colors + ((size_t)b * 4 + ch) * npix
                                          : aggrs_info + ((size_t)b * 2 + (ch - 4)) * npix;
                    dst[(size_t)orow * is + ocol] = s_out[j];
                }
            }
        }
        // top-K ids, slot order, -1 padded (replaces cudaMemsetAsync(out3_p, -1, ...) :470 + :453-455)
        if (vec_ids) {
            constexpr int QPR = TW / 4, QPP = NT / 4;
            int* bids = ids_out + (size_t)b * K * npix;
            for (int u = lane; u < K * QPP; u += NT) {
                const int k = u / QPP, pq = u - k * QPP;
                const int orow = tr0 + pq / QPR, ocol = tx0 + (pq % QPR) * 4;
                if (orow < is && ocol < is)
                    *reinterpret_cast<int4*>(bids + (size_t)k * npix + (size_t)orow * is + ocol) = reinterpret_cast<const int4*>(s_qid)[u];
            }
        } else if (px < is && row < is) {
            int* dst = ids_out + (size_t)b * K * npix + (size_t)row * is + px;
            for (int k = 0; k < K; k++)
                dst[(size_t)k * npix] = (k < st.q_size) ? s_qid[k * NT + lane] : -1;
        }
    }
}

}  // namespace b200r

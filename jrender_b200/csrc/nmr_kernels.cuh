// nmr_kernels.cuh -- Neural Mesh Renderer (dr_type='n3mr') hard rasterizer, forward + backward.
//
// Replaces the reference kernels of jrender/renderer/dr/n3mr/cuda/rasterize.py:
//   K7  forward_face_index_map_cuda_kernel   :30-164   thread per FACE scanning its bounding box,
//                                                       per-pixel atomicCAS spin lock for the z test
//   K8  forward_texture_sampling_cuda_kernel :227-298
//   K9  backward_pixel_map_cuda_kernel       :351-610  thread per face, serial scans
//   K10 backward_textures_cuda_kernel        :659-694
//   K11 backward_depth_map_cuda_kernel       :738-788
//
// B200 design: the forward is tile-centric like the SoftRas forward (same exact binning
// infrastructure): a thread owns a pixel and keeps its z-buffer entry in registers while the
// CTA streams the tile's faces through shared memory in ascending id -- no lock, no atomics,
// and a deterministic tie rule (lowest face id wins equal depth; the reference's winner is
// race-dependent).  K7 and K8 are fused: the pixel samples its winning face's texture and
// writes every map once.  K9 keeps the reference's per-face edge walk but gives a WARP to each
// face so that the long "out" scans (up to image_size pixels each) run 32 pixels at a time on
// coalesced rows.  Maps keep the kernels' orientation [B, yi, xi] with yi up.
#pragma once
#include "common.cuh"
#include "exact_math.cuh"
#include "softras_setup.cuh"

namespace b200r {

struct __align__(16) NmrRec {   // 128 bytes
    uint32_t rect_x;   // ix_min | ix_max << 16   (:102-103)
    uint32_t rect_r;   // iy_min | iy_max << 16   (:104-105), yi up
    uint32_t flags;    // bit 4+k: midrange(z_k)
    uint32_t face_id;
    float inv[9];      // pixel-space face_inv (:75-87), unclamped determinant
    float v[9];        // NDC x, y and z per vertex
    float rz[3];       // rcp_refined(z_k)
    float pad[7];
};
static_assert(sizeof(NmrRec) == 128, "NmrRec must be 128 bytes");

struct NmrParams {
    int B, nf, ts, is;
    float near_, far_, eps;
    float bg[3];
    int return_rgb, return_alpha, return_depth;
    int ntx, coarse_px, ncs;
};

// K7 prologue per face (:59-105): back-face cull, pixel-space inverse, integer bounding box.
__global__ void __launch_bounds__(256) k_nmr_setup(const float* __restrict__ faces, NmrRec* __restrict__ recs,
                                                   uint2* __restrict__ rects, int total_faces, int nf, int is) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_faces) return;
    float f[9];
#pragma unroll
    for (int k = 0; k < 9; k++) f[k] = __ldg(faces + (size_t)i * 9 + k);
    NmrRec r;
    r.face_id = (uint32_t)(i % nf);
    uint32_t flags = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) r.v[k] = f[k];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        r.rz[k] = rcp_refined(f[3 * k + 2]);
        if (midrange(f[3 * k + 2])) flags |= 16u << k;
    }
#pragma unroll
    for (int k = 0; k < 7; k++) r.pad[k] = 0.f;
    float p[3][2];
#pragma unroll
    for (int num = 0; num < 3; num++)
#pragma unroll
        for (int dim = 0; dim < 2; dim++) p[num][dim] = 0.5f * (f[3 * num + dim] * is + is - 1);  // :70 (0.5*x exact)
    float fi[9] = {p[1][1] - p[2][1], p[2][0] - p[1][0], p[1][0] * p[2][1] - p[2][0] * p[1][1],
                   p[2][1] - p[0][1], p[0][0] - p[2][0], p[2][0] * p[0][1] - p[0][0] * p[2][1],
                   p[0][1] - p[1][1], p[1][0] - p[0][0], p[0][0] * p[1][1] - p[1][0] * p[0][1]};
    const float den = (p[2][0] * (p[0][1] - p[1][1]) + p[0][0] * (p[1][1] - p[2][1]) + p[1][0] * (p[2][1] - p[0][1]));
#pragma unroll
    for (int k = 0; k < 9; k++) r.inv[k] = fi[k] / den;
    float x_min = (float)is, y_min = (float)is, x_max = 0.f, y_max = 0.f;
#pragma unroll
    for (int num = 0; num < 3; num++) {
        if (p[num][0] < x_min) x_min = p[num][0];
        if (p[num][0] > x_max) x_max = p[num][0];
        if (p[num][1] < y_min) y_min = p[num][1];
        if (p[num][1] > y_max) y_max = p[num][1];
    }
    int ix_min = max(0, (int)x_min), ix_max = min(is - 1, (int)x_max);
    int iy_min = max(0, (int)y_min), iy_max = min(is - 1, (int)y_max);
    const bool back = (f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0]);  // :63
    if (back || ix_min > ix_max || iy_min > iy_max) { ix_min = 1; ix_max = 0; iy_min = 1; iy_max = 0; }
    r.rect_x = (uint32_t)ix_min | ((uint32_t)ix_max << 16);
    r.rect_r = (uint32_t)iy_min | ((uint32_t)iy_max << 16);
    r.flags = flags;
    const uint4* src = reinterpret_cast<const uint4*>(&r);
    uint4* dst = reinterpret_cast<uint4*>(recs + i);
#pragma unroll
    for (int k = 0; k < 8; k++) dst[k] = src[k];
    rects[i] = make_uint2(r.rect_x, r.rect_r);
}

#define B200R_NMR_CHUNK 128

struct NmrFwdSmem {
    NmrRec rec[B200R_NMR_CHUNK];         // 16 KB
    int ids[B200R_NMR_CHUNK + 256];
    unsigned char wlist[8][B200R_NMR_CHUNK];
    int s_warp[8];
    int s_tile;
};

// K7 + K8 + background / alpha (n3mr.py:135-148), persistent over the cost-ordered tile queue.
__global__ void __launch_bounds__(B200R_TILE_THREADS, 3)
k_nmr_forward(const NmrParams P, const NmrRec* __restrict__ recs, const uint2* __restrict__ rects,
              const int* __restrict__ coarse_cnt, const int* __restrict__ coarse_ids,
              const float* __restrict__ faces, const float* __restrict__ textures,
              int* __restrict__ face_index_map, float* __restrict__ weight_map, float* __restrict__ depth_map,
              float* __restrict__ rgb_map, float* __restrict__ alpha_map, int* __restrict__ sampling_index_map,
              float* __restrict__ sampling_weight_map, float* __restrict__ face_inv_map,
              int* tile_counter, const int* __restrict__ tile_order) {
    __shared__ NmrFwdSmem S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int is = P.is, nf = P.nf;
    const int tiles_per_image = P.ntx * P.ntx;
    const int total_tiles = tiles_per_image * P.B;
    const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);

    while (true) {
        __syncthreads();
        if (tid == 0) S.s_tile = atomicAdd(tile_counter, 1);
        __syncthreads();
        const int q = S.s_tile;
        if (q >= total_tiles) break;
        const int t = __ldg(tile_order + q);
        const int b = t / tiles_per_image;
        const int tt = t - b * tiles_per_image;
        const int tx = tt % P.ntx, ty = tt / P.ntx;
        const int xi = tx * B200R_TILE + lx, yi = ty * B200R_TILE + ly;
        const float xp = b200r_pix_coord(xi, is);  // :110-111
        const float yp = b200r_pix_coord(yi, is);
        const float fxi = (float)xi, fyi = (float)yi;
        const int tx0 = tx * B200R_TILE, tx1 = tx0 + B200R_TILE - 1;
        const int tr0 = ty * B200R_TILE, tr1 = tr0 + B200R_TILE - 1;
        const int wx0 = tx0 + (warp & 1) * 8, wx1 = wx0 + 7;
        const int wr0 = tr0 + (warp >> 1) * 4, wr1 = wr0 + 3;

        float depth = P.far_;   // thrust::fill(depth_map, far) :181-182
        int best = -1;          // thrust::fill(face_index_map, -1) :176-177
        float bw0 = 0.f, bw1 = 0.f, bw2 = 0.f;

        const int cbin = (tr0 / P.coarse_px) * P.ncs + (tx0 / P.coarse_px);
        const int n_coarse = coarse_cnt[b * P.ncs * P.ncs + cbin];
        const int* clist = coarse_ids + ((size_t)b * P.ncs * P.ncs + cbin) * nf;
        const uint2* brects = rects + (size_t)b * nf;
        const NmrRec* brecs = recs + (size_t)b * nf;

        int n_pending = 0;
        for (int base = 0; base < n_coarse; base += B200R_TILE_THREADS) {
            {
                const int i = base + tid;
                int id = -1;
                bool pass = false;
                if (i < n_coarse) {
                    id = __ldg(clist + i);
                    pass = rect_overlaps(__ldg(brects + id), tx0, tx1, tr0, tr1);
                }
                int total;
                const int off = n_pending + block_excl_scan_256(pass ? 1 : 0, S.s_warp, total);
                if (pass) S.ids[off] = id;
                n_pending += total;
            }
            const bool last = base + B200R_TILE_THREADS >= n_coarse;
            if (n_pending < B200R_NMR_CHUNK && !last) continue;
            while (n_pending >= B200R_NMR_CHUNK || (last && n_pending > 0)) {
                const int m = min(n_pending, B200R_NMR_CHUNK);
                __syncthreads();
                for (int j = tid; j < m * 8; j += B200R_TILE_THREADS) {
                    const int f = j >> 3, qq = j & 7;
                    reinterpret_cast<uint4*>(&S.rec[f])[qq] = __ldg(reinterpret_cast<const uint4*>(brecs + S.ids[f]) + qq);
                }
                __syncthreads();
                const int rest = n_pending - m;
                int keep0 = 0;
                if (tid < rest) keep0 = S.ids[m + tid];
                int wcnt = 0;
                for (int j0 = 0; j0 < m; j0 += 32) {
                    const int j = j0 + lane;
                    bool pass = false;
                    if (j < m) pass = rect_overlaps(make_uint2(S.rec[j].rect_x, S.rec[j].rect_r), wx0, wx1, wr0, wr1);
                    const unsigned bal = __ballot_sync(0xffffffffu, pass);
                    if (pass) S.wlist[warp][wcnt + __popc(bal & ((1u << lane) - 1u))] = (unsigned char)j;
                    wcnt += __popc(bal);
                }
                __syncwarp();
                for (int it = 0; it < wcnt; it++) {
                    const NmrRec* rec = &S.rec[S.wlist[warp][it]];
                    {
                        const uint32_t rx = rec->rect_x, rr = rec->rect_r;
                        const uint32_t x0 = rx & 0xffffu, r0 = rr & 0xffffu;
                        if ((uint32_t)(xi - (int)x0) > (rx >> 16) - x0) continue;
                        if ((uint32_t)(yi - (int)r0) > (rr >> 16) - r0) continue;
                    }
                    const float* f = rec->v;
                    // inside test in NDC (:113-116)
                    if (((yp - f[1]) * (f[3] - f[0]) < (xp - f[0]) * (f[4] - f[1])) ||
                        ((yp - f[4]) * (f[6] - f[3]) < (xp - f[3]) * (f[7] - f[4])) ||
                        ((yp - f[7]) * (f[0] - f[6]) < (xp - f[6]) * (f[1] - f[7])))
                        continue;
                    const float* inv = rec->inv;
                    float w0 = inv[0] * fxi + inv[1] * fyi + inv[2];   // :121-123
                    float w1 = inv[3] * fxi + inv[4] * fyi + inv[5];
                    float w2 = inv[6] * fxi + inv[7] * fyi + inv[8];
                    w0 = fminf(fmaxf(w0, 0.f), 1.f);                   // :128
                    w1 = fminf(fmaxf(w1, 0.f), 1.f);
                    w2 = fminf(fmaxf(w2, 0.f), 1.f);
                    const float w_sum = ((0.f + w0) + w1) + w2;       // :126-129
                    if (w_sum != 1.f) {
                        const float r = rcp_refined(w_sum);
                        const bool safe = midrange(w_sum);
                        w0 = fast_div(w0, w_sum, r, safe);            // :132
                        w1 = fast_div(w1, w_sum, r, safe);
                        w2 = fast_div(w2, w_sum, r, safe);
                    }
                    const uint32_t fl = rec->flags;
                    const float zp = 1.f / (fast_div(w0, f[2], rec->rz[0], (fl & 16u) != 0) +
                                            fast_div(w1, f[5], rec->rz[1], (fl & 32u) != 0) +
                                            fast_div(w2, f[8], rec->rz[2], (fl & 64u) != 0));   // :135
                    if (zp <= P.near_ || P.far_ <= zp) continue;       // :136
                    if (zp < depth) {                                   // :146, ascending id => lowest id wins ties
                        depth = zp;
                        best = (int)rec->face_id;
                        bw0 = w0; bw1 = w1; bw2 = w2;
                    }
                }
                __syncthreads();
                if (tid < rest) S.ids[tid] = keep0;
                n_pending = rest;
            }
        }

        // ---- write the pixel's maps once (replaces the fills/memsets :176-184, :311-313)
        if (xi < is && yi < is) {
            const size_t i1 = ((size_t)b * is + yi) * is + xi;
            face_index_map[i1] = best;
            weight_map[i1 * 3 + 0] = bw0;
            weight_map[i1 * 3 + 1] = bw1;
            weight_map[i1 * 3 + 2] = bw2;
            depth_map[i1] = depth;
            if (P.return_alpha) alpha_map[i1] = best >= 0 ? 1.f : 0.f;   // n3mr.py:145-148
            if (P.return_depth) {
#pragma unroll
                for (int k = 0; k < 9; k++)
                    face_inv_map[i1 * 9 + k] = best >= 0 ? __ldg(&brecs[best].inv[k]) : 0.f;   // :152-156
            }
            if (P.return_rgb) {
                float px0 = P.bg[0], px1 = P.bg[1], px2 = P.bg[2];   // forward_background, n3mr.py:135-143
                int sidx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                float sw[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (best >= 0) {  // K8 :248-297
                    const int ts = P.ts;
                    const float* face = faces + ((size_t)b * nf + best) * 9;
                    const float* texture = textures + ((size_t)b * nf + best) * ts * ts * ts * 3;
                    const float wk[3] = {bw0, bw1, bw2};
                    float tif[3];
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        float v = wk[k] * (ts - 1) * (depth / __ldg(face + 3 * k + 2));
                        v = fmaxf(v, 0.f);
                        v = fminf(v, ts - 1 - P.eps);
                        tif[k] = v;
                    }
                    float np0 = 0.f, np1 = 0.f, np2 = 0.f;
#pragma unroll
                    for (int pn = 0; pn < 8; pn++) {
                        float w = 1.f;
                        int tii[3];
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const int base_i = (int)tif[k];
                            if (((pn >> k) & 1) == 0) { w *= 1 - (tif[k] - base_i); tii[k] = base_i; }
                            else { w *= tif[k] - base_i; tii[k] = base_i + 1; }
                        }
                        const int isc = tii[0] * ts * ts + tii[1] * ts + tii[2];
                        // texture_size 1: ts - 1 - eps < 0 and the "+1" taps leave the face's block; the
                        // reference reads whatever follows (next faces' texels; past the tensor for the
                        // last faces).  Taps inside the tensor are read likewise, taps past it give 0.
                        const size_t tap = ((size_t)b * nf + best) * ts * ts * ts + (size_t)isc;
                        if (isc >= 0 && tap < (size_t)P.B * nf * ts * ts * ts) {
                            np0 += w * __ldg(texture + isc * 3 + 0);
                            np1 += w * __ldg(texture + isc * 3 + 1);
                            np2 += w * __ldg(texture + isc * 3 + 2);
                        }
                        sidx[pn] = isc;
                        sw[pn] = w;
                    }
                    // rgb * mask + (1 - mask) * background with mask == 1
                    px0 = np0 * 1.f + 0.f * P.bg[0];
                    px1 = np1 * 1.f + 0.f * P.bg[1];
                    px2 = np2 * 1.f + 0.f * P.bg[2];
                } else {
                    px0 = 0.f * 0.f + 1.f * P.bg[0];
                    px1 = 0.f * 0.f + 1.f * P.bg[1];
                    px2 = 0.f * 0.f + 1.f * P.bg[2];
                }
                rgb_map[i1 * 3 + 0] = px0;
                rgb_map[i1 * 3 + 1] = px1;
                rgb_map[i1 * 3 + 2] = px2;
                int4* si = reinterpret_cast<int4*>(sampling_index_map + i1 * 8);
                si[0] = make_int4(sidx[0], sidx[1], sidx[2], sidx[3]);
                si[1] = make_int4(sidx[4], sidx[5], sidx[6], sidx[7]);
                float4* sv = reinterpret_cast<float4*>(sampling_weight_map + i1 * 8);
                sv[0] = make_float4(sw[0], sw[1], sw[2], sw[3]);
                sv[1] = make_float4(sw[4], sw[5], sw[6], sw[7]);
            }
        }
    }
}

// ---------------------------------------------------------------- K9 prologue: packed pixel records
// K9's scans evaluate, for long pixel runs (along x for axis 1, along y for axis 0),
//     diff_grad = sum_k (map_k - ref_k) * grad_k            (k over r, g, b and/or alpha; ref = the sample's in / out pixel)
// which is  A - sum_k ref_k * grad_k  with the per-pixel moment  A = sum_k map_k * grad_k.  This kernel stores (A, grad)
// per pixel -- 16 bytes for rgb (A, gr, gg, gb), 8 for alpha (A, ga), 32 for both -- twice: row-major (ph[b][y][x]) and
// column-major (pv[b][x][y]), so both scan directions read consecutive records.  (Reading the four maps separately, and
// columns with a stride of a whole image row, made the first version of the backward 20x slower than its instruction
// count warrants; 32-byte (rgb, grad, alpha, grad) records were the second version: the scans then ran at the L1's
// bandwidth, which this halves.)  32x32 tile transpose through shared memory.
// MODE 1: rgb, 2: alpha, 3: both.  FL = floats per record (4, 2, 8).
template <int MODE>
struct K9Rec {
    static constexpr int FL = (MODE == 1) ? 4 : ((MODE == 2) ? 2 : 8);
};

template <int MODE>
__global__ void __launch_bounds__(256)
k_nmr_pack(const float* __restrict__ rgb_map, const float* __restrict__ alpha_map, const float* __restrict__ grad_rgb_map,
           const float* __restrict__ grad_alpha_map, float* __restrict__ ph, float* __restrict__ pv, int is) {
    constexpr int FL = K9Rec<MODE>::FL;
    __shared__ float s[32][33][FL];
    const int b = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t img = (size_t)b * is * is;
    for (int r = ty; r < 32; r += 8) {
        const int x = x0 + tx, y = y0 + r;
        float v[FL];
#pragma unroll
        for (int k = 0; k < FL; k++) v[k] = 0.f;
        if (x < is && y < is) {
            const size_t i = img + (size_t)y * is + x;
            float A = 0.f;
            if (MODE & 1) {
                const float g0 = __ldg(grad_rgb_map + i * 3 + 0), g1 = __ldg(grad_rgb_map + i * 3 + 1), g2 = __ldg(grad_rgb_map + i * 3 + 2);
                A = __fmaf_rn(__ldg(rgb_map + i * 3 + 0), g0, A);
                A = __fmaf_rn(__ldg(rgb_map + i * 3 + 1), g1, A);
                A = __fmaf_rn(__ldg(rgb_map + i * 3 + 2), g2, A);
                v[1] = g0; v[2] = g1; v[3] = g2;
            }
            if (MODE & 2) {
                const float ga = __ldg(grad_alpha_map + i);
                A = __fmaf_rn(__ldg(alpha_map + i), ga, A);
                v[(MODE == 2) ? 1 : 4] = ga;
            }
            v[0] = A;
#pragma unroll
            for (int k = 0; k < FL; k++) ph[i * FL + k] = v[k];
        }
#pragma unroll
        for (int k = 0; k < FL; k++) s[r][tx][k] = v[k];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int x = x0 + r, y = y0 + tx;   // pv[b][x][y] <- pixel (y, x) = s[y - y0][x - x0]
        if (x < is && y < is) {
            const size_t j = img + (size_t)x * is + y;
#pragma unroll
            for (int k = 0; k < FL; k++) pv[j * FL + k] = s[tx][r][k];
        }
    }
}

// ---------------------------------------------------------------- K9: warp per (batch, face)
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// 1 / x to 1 ulp (MUFU.RCP).  The edge-scan sums below run in a different order from the
// reference's serial per-face loop anyway (lanes, then a warp reduction), so the terms
// diff_grad / dist are formed with this reciprocal and one FMA instead of an IEEE division:
// <= 1.5 ulp per term, against the 2e-5 relative tolerance of the gradient parity tests.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // |x| >= eps: never denormal
    return r;
}

__device__ __forceinline__ float k9_sel3(int i, float a, float b, float c) { return i == 0 ? a : (i == 1 ? b : c); }

// One line scan of K9 (:486-520 outwards / :556-601 inwards): lanes take consecutive pixels d1_from + lane, + 32, ...
// of the packed row `prow` (k_nmr_pack) and add their  -diff_grad / dist  terms to acc0 / acc1.
//   OWNER  inward scan: only pixels whose face_index_map entry is this face count (:573)
//   BOTH   both vertex terms present (p[1][0] != d0 and p[0][0] != d0, :508/:513): the common, branch-free flavour
// The loop is the whole cost of the kernel (the outward scan runs to the image border for every edge sample), so it
// is kept to one record load and ~20 arithmetic instructions per 32 pixels: pointers and the pixel coordinate
// advance by constants, the `diff_grad <= 0` skip is a select (a NaN still propagates like the reference's
// `continue` would let it), products are fused (sums of ~10^2..10^3 terms, compared at 2e-5 relative).
// ref = (r, g, b, alpha) of the sample's in-pixel (outward scan) / out-pixel (inward scan).
template <int MODE, bool OWNER, bool BOTH>
__device__ __forceinline__ void k9_scan(const float* __restrict__ prow, const int* __restrict__ fim, long fim_stride,
                                        int d1_from, int d1_to, int lane, int fn, float r0, float r1, float r2, float ra,
                                        float d1_cross, float k0, float k1, bool has0, bool has1, float eps,
                                        float& acc0, float& acc1) {
    constexpr int FL = K9Rec<MODE>::FL;
    const float* __restrict__ p = prow + (long)(d1_from + lane) * FL;
    const int* __restrict__ o = OWNER ? fim + (long)(d1_from + lane) * fim_stride : nullptr;
    float d1f = (float)(d1_from + lane);   // exact below 2^24
    for (int left = d1_to - d1_from - lane; left >= 0; left -= 32) {
        float diff;
        if (MODE == 1) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(p));
            diff = q.x - __fmaf_rn(r2, q.w, __fmaf_rn(r1, q.z, r0 * q.y));
        } else if (MODE == 2) {
            const float2 q = __ldg(reinterpret_cast<const float2*>(p));
            diff = q.x - ra * q.y;
        } else {
            const float4 q = __ldg(reinterpret_cast<const float4*>(p));
            const float ga = __ldg(p + 4);
            diff = q.x - __fmaf_rn(ra, ga, __fmaf_rn(r2, q.w, __fmaf_rn(r1, q.z, r0 * q.y)));
        }
        bool mine = true;
        if (OWNER) { mine = __ldg(o) == fn; o += 32 * fim_stride; }
        p += 32 * FL;
        const float d = (mine && !(diff <= 0.f)) ? diff : 0.f;   // :503 / :587 `if (diff_grad <= 0) continue`
        const float delta = d1f - d1_cross;   // (d1 - d1_cross), one rounding like the reference's
        d1f += 32.f;
        if (BOTH || has0) {
            float dist = delta * k0;
            dist = (0.f < dist) ? dist + eps : dist - eps;      // :510
            acc0 = __fmaf_rn(-d, rcp_approx(dist), acc0);
        }
        if (BOTH || has1) {
            float dist = delta * k1;
            dist = (0.f < dist) ? dist + eps : dist - eps;      // :515
            acc1 = __fmaf_rn(-d, rcp_approx(dist), acc1);
        }
    }
}

// K9 (backward_pixel_map_cuda_kernel, n3mr/cuda/rasterize.py:351-610): warp per (batch, face).  The six (edge, axis)
// passes of a face run as a real loop (six inlined copies thrash the instruction cache).  Per pass the three slopes
// of :423 / :529-533 are divided once (the reference re-divides per d0 with the same operands: same bits), the
// per-sample quotients that only scale `dist` use the approximate reciprocal, and the two scans are the tight loops
// of k9_scan.  d1_cross / d0_cross2 -- which pick pixels through floor / ceil -- keep the reference's exact,
// unfused arithmetic.
#ifndef B200R_K9_MINB
#define B200R_K9_MINB 6   // resident 256-thread CTAs per SM (40 registers; with the 32-byte records: 7.10 ms at C4 against 7.33 at 5 and 7.68 at 4 -- the scans are latency-bound)
#endif
template <int MODE>
__global__ void __launch_bounds__(256, B200R_K9_MINB)
k_nmr_backward_pixel_map(const float* __restrict__ faces, const int* __restrict__ face_index_map,
                         const float* __restrict__ rgb_map, const float* __restrict__ alpha_map,
                         const float* __restrict__ ph, const float* __restrict__ pv,
                         float* __restrict__ grad_faces, int batch_size, int num_faces, int is, float eps) {
    constexpr int FL = K9Rec<MODE>::FL;
    const int lane = threadIdx.x & 31;
    const long i = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= (long)batch_size * num_faces) return;
    const int bn = (int)(i / num_faces);
    const int fn = (int)(i % num_faces);
    const float* __restrict__ face = faces + i * 9;
    float fx[3], fy[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { fx[k] = __ldg(face + 3 * k); fy[k] = __ldg(face + 3 * k + 1); }
    if ((fy[2] - fy[0]) * (fx[1] - fx[0]) < (fy[1] - fy[0]) * (fx[2] - fx[0])) return;  // :377 (zeros stay)
    // pixel-space vertices, pp[num][dim] = 0.5 * (face * is + is - 1)  (:389-393)
    float px[3], py[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { px[k] = 0.5f * (fx[k] * is + is - 1); py[k] = 0.5f * (fy[k] * is + is - 1); }

    float gacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // [vertex*2 + (0:x, 1:y)]
    const float ndc_scale = 2.f / (float)is;
    const long img = (long)bn * is * is;

#pragma unroll 1
    for (int pass = 0; pass < 6; pass++) {
        const int edge_num = pass >> 1, axis = pass & 1;
        const int pi0 = edge_num, pi1 = (edge_num == 2) ? 0 : edge_num + 1, pi2 = (edge_num == 0) ? 2 : edge_num - 1;
        // p[num][dim] = pp[num][(dim + axis) % 2]   (:396-401)
        const float ax0 = k9_sel3(pi0, px[0], px[1], px[2]), ay0 = k9_sel3(pi0, py[0], py[1], py[2]);
        const float ax1 = k9_sel3(pi1, px[0], px[1], px[2]), ay1 = k9_sel3(pi1, py[0], py[1], py[2]);
        const float ax2 = k9_sel3(pi2, px[0], px[1], px[2]), ay2 = k9_sel3(pi2, py[0], py[1], py[2]);
        const float p00 = axis ? ay0 : ax0, p01 = axis ? ax0 : ay0;
        const float p10 = axis ? ay1 : ax1, p11 = axis ? ax1 : ay1;
        const float p20 = axis ? ay2 : ax2, p21 = axis ? ax2 : ay2;
        int direction;
        if (axis == 0) direction = (p00 < p10) ? -1 : 1;
        else direction = (p00 < p10) ? 1 : -1;
        const int d0_from = (int)fmax((double)ceilf(fminf(p00, p10)), 0.);
        const int d0_to = (int)fmin((double)fmaxf(p00, p10), is - 1.);
        if (d0_from > d0_to) continue;
        const float slope01 = (p11 - p01) / (p10 - p00);   // :423
        const float slope02 = (p21 - p01) / (p20 - p00);   // :530
        const float slope21 = (p11 - p21) / (p10 - p20);   // :533
        const float span = p10 - p00;
        // packed records: axis 1 scans rows of ph, axis 0 scans rows of pv; in both the record of (d0, d1) sits at
        // (img + d0 * is + d1) * FL
        const float* __restrict__ pk = ((axis == 0) ? pv : ph) + img * FL;
        const long fim_stride = (axis == 0) ? is : 1;
        float acc0 = 0.f, acc1 = 0.f;  // contributions to vertex pi0 / pi1, coordinate (1 - axis)
        for (int d0 = d0_from; d0 <= d0_to; d0++) {
            const float d0f = (float)d0;
            const float d1_cross = slope01 * (d0f - p00) + p01;
            const int d1_in = (0 < direction) ? (int)floorf(d1_cross) : (int)ceilf(d1_cross);
            const int d1_out = d1_in + direction;
            if (d1_in < 0 || is <= d1_in) continue;
            if (d1_out < 0 || is <= d1_out) continue;
            const float* __restrict__ prow = pk + (long)d0 * is * FL;
            const bool has0 = p10 != d0f, has1 = p00 != d0f;
            // pixel -> NDC: the reference's (float)((double)(q * delta) * 2. / is) is formed as delta * (q * (2 / is)); q by
            // approximate reciprocal -- a few ulp on a quantity that is then offset by eps and inverted approximately
            const float k0 = span * rcp_approx(p10 - d0f) * ndc_scale, k1 = span * rcp_approx(d0f - p00) * ndc_scale;
            const long fim_base = (axis == 0) ? img + d0 : img + (long)d0 * is;
            const long idx_in = fim_base + (long)d1_in * fim_stride, idx_out = fim_base + (long)d1_out * fim_stride;
            const bool visible = __ldg(face_index_map + idx_in) == fn;
            if (visible) {   // outwards from the edge to the image border (:461-521); reference colour = the in-pixel's
                float r0 = 0.f, r1 = 0.f, r2 = 0.f, ra = 0.f;
                if (MODE & 1) { r0 = __ldg(rgb_map + idx_in * 3); r1 = __ldg(rgb_map + idx_in * 3 + 1); r2 = __ldg(rgb_map + idx_in * 3 + 2); }
                if (MODE & 2) ra = __ldg(alpha_map + idx_in);
                const int lim = (0 < direction) ? is - 1 : 0;
                const int d1_from = max(min(d1_out, lim), 0), d1_to = min(max(d1_out, lim), is - 1);
                if (has0 && has1) k9_scan<MODE, false, true>(prow, nullptr, 0, d1_from, d1_to, lane, fn, r0, r1, r2, ra, d1_cross, k0, k1, true, true, eps, acc0, acc1);
                else k9_scan<MODE, false, false>(prow, nullptr, 0, d1_from, d1_to, lane, fn, r0, r1, r2, ra, d1_cross, k0, k1, has0, has1, eps, acc0, acc1);
            }
            {   // inwards up to the opposite edge, pixels owned by the face (:523-602); reference colour = the out-pixel's
                float d0_cross2;
                if ((d0f - p00) * (d0f - p20) < 0.f) d0_cross2 = slope02 * (d0f - p00) + p01;
                else d0_cross2 = slope21 * (d0f - p20) + p21;
                const int lim = (0 < direction) ? (int)ceilf(d0_cross2) : (int)floorf(d0_cross2);
                const int d1_from = max(min(d1_in, lim), 0), d1_to = min(max(d1_in, lim), is - 1);
                if (d1_from <= d1_to) {
                    float r0 = 0.f, r1 = 0.f, r2 = 0.f, ra = 0.f;
                    if (MODE & 1) { r0 = __ldg(rgb_map + idx_out * 3); r1 = __ldg(rgb_map + idx_out * 3 + 1); r2 = __ldg(rgb_map + idx_out * 3 + 2); }
                    if (MODE & 2) ra = __ldg(alpha_map + idx_out);
                    k9_scan<MODE, true, false>(prow, face_index_map + fim_base, fim_stride, d1_from, d1_to, lane, fn, r0, r1, r2, ra,
                                               d1_cross, k0, k1, has0, has1, eps, acc0, acc1);
                }
            }
        }
        const int g0 = pi0 * 2 + (1 - axis), g1 = pi1 * 2 + (1 - axis);
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if (k == g0) gacc[k] += acc0;
            if (k == g1) gacc[k] += acc1;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) gacc[k] = warp_sum_f(gacc[k]);
    if (lane == 0) {
        float* g = grad_faces + i * 9;
        g[0] = gacc[0]; g[1] = gacc[1]; g[2] = 0.f;
        g[3] = gacc[2]; g[4] = gacc[3]; g[5] = 0.f;
        g[6] = gacc[4]; g[7] = gacc[5]; g[8] = 0.f;
    }
}

// ---------------------------------------------------------------- K10 + K11 fused: thread per pixel
__global__ void __launch_bounds__(256)
k_nmr_backward_maps(const float* __restrict__ faces, const int* __restrict__ face_index_map,
                    const float* __restrict__ weight_map, const float* __restrict__ depth_map,
                    const float* __restrict__ face_inv_map, const int* __restrict__ sampling_index_map,
                    const float* __restrict__ sampling_weight_map, const float* __restrict__ grad_rgb_map,
                    const float* __restrict__ grad_depth_map, float* __restrict__ grad_faces,
                    float* __restrict__ grad_textures, int batch_size, int nf, int is, int ts,
                    int return_rgb, int return_depth) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)batch_size * is * is) return;
    const int fn = __ldg(face_index_map + i);
    if (fn < 0) return;
    const int bn = (int)(i / ((size_t)is * is));
    if (return_rgb) {  // K10 :675-692
        float* gt = grad_textures + ((size_t)bn * nf + fn) * ts * ts * ts * 3;
        const float g0 = __ldg(grad_rgb_map + i * 3 + 0), g1 = __ldg(grad_rgb_map + i * 3 + 1), g2 = __ldg(grad_rgb_map + i * 3 + 2);
#pragma unroll
        for (int pn = 0; pn < 8; pn++) {
            const float w = __ldg(sampling_weight_map + i * 8 + pn);
            const int isc = __ldg(sampling_index_map + i * 8 + pn);
            const size_t tap = ((size_t)bn * nf + fn) * ts * ts * ts + (size_t)isc;
            if (isc < 0 || tap >= (size_t)batch_size * nf * ts * ts * ts) continue;  // tap past the tensor (ts == 1, see K8)
            atomicAdd(gt + isc * 3 + 0, w * g0);
            atomicAdd(gt + isc * 3 + 1, w * g1);
            atomicAdd(gt + isc * 3 + 2, w * g2);
        }
    }
    if (return_depth) {  // K11 :755-786
        const float* face = faces + ((size_t)bn * nf + fn) * 9;
        const float depth = __ldg(depth_map + i);
        const float depth2 = depth * depth;
        const float grad_depth = __ldg(grad_depth_map + i);
        float* g = grad_faces + ((size_t)bn * nf + fn) * 9;
        float w[3], z[3], fi[9];
#pragma unroll
        for (int k = 0; k < 3; k++) { w[k] = __ldg(weight_map + i * 3 + k); z[k] = __ldg(face + 3 * k + 2); }
#pragma unroll
        for (int k = 0; k < 9; k++) fi[k] = __ldg(face_inv_map + i * 9 + k);
#pragma unroll
        for (int k = 0; k < 3; k++) atomicAdd(g + 3 * k + 2, grad_depth * w[k] * depth2 / (z[k] * z[k]));
        float tmp[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int l = 0; l < 3; l++) tmp[k] += -fi[3 * l + k] / z[l];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int l = 0; l < 2; l++) atomicAdd(g + 3 * k + l, -grad_depth * tmp[l] * w[k] * depth2 * is / 2);
    }
}

}  // namespace b200r

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
// K9's scans read (rgb, grad_rgb, alpha, grad_alpha) of long pixel runs: along x for axis 1 and
// along y for axis 0.  Reading four separate maps, and columns with a stride of a whole image
// row, made the first version of the backward 20x slower than its instruction count warrants.
// This kernel gathers the 8 floats of every pixel into one 32-byte record and writes them twice:
// row-major (ph[b][y][x]) and column-major (pv[b][x][y]), so that both scan directions read
// consecutive 32-byte sectors.  32x32 tile transpose through shared memory.
__global__ void __launch_bounds__(256)
k_nmr_pack(const float* __restrict__ rgb_map, const float* __restrict__ alpha_map, const float* __restrict__ grad_rgb_map,
           const float* __restrict__ grad_alpha_map, float4* __restrict__ ph, float4* __restrict__ pv, int is,
           int return_rgb, int return_alpha) {
    __shared__ float4 s[32][33][2];
    const int b = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t img = (size_t)b * is * is;
    for (int r = ty; r < 32; r += 8) {
        const int x = x0 + tx, y = y0 + r;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x < is && y < is) {
            const size_t i = img + (size_t)y * is + x;
            if (return_rgb) {
                a.x = __ldg(rgb_map + i * 3 + 0); a.y = __ldg(rgb_map + i * 3 + 1); a.z = __ldg(rgb_map + i * 3 + 2);
                a.w = __ldg(grad_rgb_map + i * 3 + 0); c.x = __ldg(grad_rgb_map + i * 3 + 1); c.y = __ldg(grad_rgb_map + i * 3 + 2);
            }
            if (return_alpha) { c.z = __ldg(alpha_map + i); c.w = __ldg(grad_alpha_map + i); }
            ph[i * 2] = a;
            ph[i * 2 + 1] = c;
        }
        s[r][tx][0] = a;
        s[r][tx][1] = c;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int x = x0 + r, y = y0 + tx;   // pv[b][x][y] <- pixel (y, x) = s[y - y0][x - x0]
        if (x < is && y < is) {
            const size_t j = img + (size_t)x * is + y;
            pv[j * 2] = s[tx][r][0];
            pv[j * 2 + 1] = s[tx][r][1];
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

// One (edge, axis) pass of the reference's per-face loop (:386-608).  The six passes of a face
// run as a real loop (not unrolled): the body is long and six copies of it thrash the
// instruction cache.
#ifndef B200R_K9_MINB
#define B200R_K9_MINB 6   // resident 256-thread CTAs per SM for the single-pixel-per-trip scan (40 registers; measured best of 4/5/6/8)
#endif
template <int U>
__global__ void __launch_bounds__(256, U == 1 ? B200R_K9_MINB : 1)
k_nmr_backward_pixel_map(const float* __restrict__ faces, const int* __restrict__ face_index_map,
                         const float* __restrict__ rgb_map, const float* __restrict__ alpha_map,
                         const float* __restrict__ grad_rgb_map, const float* __restrict__ grad_alpha_map,
                         const float4* __restrict__ ph, const float4* __restrict__ pv,
                         float* __restrict__ grad_faces, int batch_size, int num_faces, int is, float eps,
                         int return_rgb, int return_alpha) {
    const int lane = threadIdx.x & 31;
    const long i = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= (long)batch_size * num_faces) return;
    const int bn = (int)(i / num_faces);
    const int fn = (int)(i % num_faces);
    const float* __restrict__ face = faces + i * 9;
    {
        const float x0 = __ldg(face + 0), y0 = __ldg(face + 1), x1 = __ldg(face + 3), y1 = __ldg(face + 4);
        const float x2 = __ldg(face + 6), y2 = __ldg(face + 7);
        if ((y2 - y0) * (x1 - x0) < (y1 - y0) * (x2 - x0)) return;  // :377 (zeros stay)
    }

    float gacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // [vertex*2 + (0:x, 1:y)]
    const float ndc_scale = 2.f / (float)is;
    const long img = (long)bn * is * is;

#pragma unroll 1
    for (int pass = 0; pass < 6; pass++) {
        const int edge_num = pass >> 1, axis = pass & 1;
        const int pi0 = edge_num, pi1 = (edge_num == 2) ? 0 : edge_num + 1, pi2 = (edge_num == 0) ? 2 : edge_num - 1;
        // p[num][dim] = 0.5 * (face[3 * pi_num + (dim + axis) % 2] * is + is - 1)   (:389-399)
        const float p00 = 0.5f * (__ldg(face + 3 * pi0 + axis) * is + is - 1), p01 = 0.5f * (__ldg(face + 3 * pi0 + (axis ^ 1)) * is + is - 1);
        const float p10 = 0.5f * (__ldg(face + 3 * pi1 + axis) * is + is - 1), p11 = 0.5f * (__ldg(face + 3 * pi1 + (axis ^ 1)) * is + is - 1);
        const float p20 = 0.5f * (__ldg(face + 3 * pi2 + axis) * is + is - 1), p21 = 0.5f * (__ldg(face + 3 * pi2 + (axis ^ 1)) * is + is - 1);
        int direction;
        if (axis == 0) direction = (p00 < p10) ? -1 : 1;
        else direction = (p00 < p10) ? 1 : -1;
        const int d0_from = (int)fmax((double)ceilf(fminf(p00, p10)), 0.);
        const int d0_to = (int)fmin((double)fmaxf(p00, p10), is - 1.);
        // packed records: axis 1 scans rows of ph, axis 0 scans rows of pv; in both the record
        // of (d0, d1) sits at img + d0 * is + d1
        const float4* __restrict__ pk = ((axis == 0) ? pv : ph) + img * 2;
        float acc0 = 0.f, acc1 = 0.f;  // contributions to vertex pi0 / pi1, coordinate (1 - axis)
        for (int d0 = d0_from; d0 <= d0_to; d0++) {
            const float d1_cross = (p11 - p01) / (p10 - p00) * (d0 - p00) + p01;
            const int d1_in = (0 < direction) ? (int)floorf(d1_cross) : (int)ceilf(d1_cross);
            const int d1_out = d1_in + direction;
            if (d1_in < 0 || is <= d1_in) continue;
            if (d1_out < 0 || is <= d1_out) continue;
            const float4* __restrict__ prow = pk + (long)d0 * is * 2;
            const float4 ia = __ldg(prow + d1_in * 2), ic = __ldg(prow + d1_in * 2 + 1);
            const float4 oa = __ldg(prow + d1_out * 2), oc = __ldg(prow + d1_out * 2 + 1);
            const bool has0 = p10 != d0, has1 = p00 != d0;
            const float q0 = (p10 - p00) / (p10 - d0);
            const float q1 = (p10 - p00) / (d0 - p00);
            // pixel -> NDC: the reference's (float)((double)(q * delta) * 2. / is) is formed as
            // delta * (q * (2 / is)) -- a few ulp on a quantity that is then offset by eps and
            // inverted approximately (see rcp_approx); the sign test `0 < dist` is unaffected
            const float k0 = q0 * ndc_scale, k1 = q1 * ndc_scale;
            const long map_index_in = (axis == 0) ? img + (long)d1_in * is + d0 : img + (long)d0 * is + d1_in;
            const bool visible = __ldg(face_index_map + map_index_in) == fn;
            float d0_cross2;
            if ((d0 - p00) * (d0 - p20) < 0.f) d0_cross2 = (p21 - p01) / (p20 - p00) * (d0 - p00) + p01;
            else d0_cross2 = (p11 - p21) / (p10 - p20) * (d0 - p20) + p21;
            const int in_limit = (0 < direction) ? (int)ceilf(d0_cross2) : (int)floorf(d0_cross2);
            const long fim_base = (axis == 0) ? img + d0 : img + (long)d0 * is;
            const long fim_stride = (axis == 0) ? is : 1;

#pragma unroll 1
            for (int scan = 0; scan < 2; scan++) {
                // scan 0: outwards from the edge (:461-521), only if the face owns the inside pixel;
                // scan 1: inwards up to the opposite edge (:523-602), pixels owned by the face
                int lim, start;
                float r0, r1, r2, ra;
                if (scan == 0) {
                    if (!visible) continue;
                    lim = (0 < direction) ? is - 1 : 0;
                    start = d1_out;
                    r0 = ia.x; r1 = ia.y; r2 = ia.z; ra = ic.z;
                } else {
                    lim = in_limit;
                    start = d1_in;
                    r0 = oa.x; r1 = oa.y; r2 = oa.z; ra = oc.z;
                }
                const int d1_from = max(min(start, lim), 0);
                const int d1_to = min(max(start, lim), is - 1);
                // U pixels per lane per trip, every load issued before the first use: the scan is
                // latency-bound (one dependent L1/L2 round trip per trip), not bandwidth-bound
                float d1f = (float)(d1_from + lane);  // exact below 2^24
                for (int d1b = d1_from + lane; d1b <= d1_to; d1b += 32 * U, d1f += 32.f * U) {
                    float4 qa[U], qc[U];
                    int owner[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int d1 = min(d1b + 32 * u, d1_to);  // clamped: always a valid address
                        owner[u] = (scan == 1) ? __ldg(face_index_map + fim_base + (long)d1 * fim_stride) : fn;
                        qa[u] = __ldg(prow + d1 * 2);
                        qc[u] = __ldg(prow + d1 * 2 + 1);
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int d1 = d1b + 32 * u;
                        if (d1 > d1_to || owner[u] != fn) continue;
                        float diff_grad = 0.f;
                        if (return_alpha) diff_grad += (qc[u].z - ra) * qc[u].w;
                        if (return_rgb) {
                            diff_grad += (qa[u].x - r0) * qa[u].w;
                            diff_grad += (qa[u].y - r1) * qc[u].x;
                            diff_grad += (qa[u].z - r2) * qc[u].y;
                        }
                        if (diff_grad <= 0.f) continue;
                        const float delta = (d1f + 32.f * u) - d1_cross;
                        if (has0) {
                            float dist = delta * k0;
                            dist = (0.f < dist) ? dist + eps : dist - eps;
                            acc0 = __fmaf_rn(-diff_grad, rcp_approx(dist), acc0);
                        }
                        if (has1) {
                            float dist = delta * k1;
                            dist = (0.f < dist) ? dist + eps : dist - eps;
                            acc1 = __fmaf_rn(-diff_grad, rcp_approx(dist), acc1);
                        }
                    }
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

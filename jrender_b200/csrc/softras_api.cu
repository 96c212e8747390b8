// softras_api.cu -- C ABI entry points for the SoftRas path (see include/b200raster.h).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/b200raster.h"
#include "api_util.cuh"
#include "softras_backward.cuh"
#include "softras_forward.cuh"
#include "softras_setup.cuh"

using namespace b200r;

namespace {

std::atomic<int> g_fwd_variant{1};     // 0: warp-uniform face loop, 1: per-lane face lists
std::atomic<int> g_fwd_persistent{1};  // 0: one CTA per tile, 1: persistent grid + atomic tile queue

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

int gcd_i(long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return (int)a; }

int validate(const char* fn, int B, int nf, int T, int is, int K, int dist, int rgb, int alpha, int tex,
             float sigma, float gamma) {
    if (B <= 0 || nf <= 0 || T <= 0 || is <= 0) return b200r_fail(B200R_EINVAL, "%s: non-positive size (B=%d nf=%d T=%d image_size=%d)", fn, B, nf, T, is);
    if (is > 4096) return b200r_fail(B200R_EUNSUPPORTED, "%s: image_size %d > 4096", fn, is);
    if (K <= 0 || K > B200R_MAX_FACES_PER_PIXEL)
        return b200r_fail(B200R_EINVAL, "%s: max_faces_per_pixel %d outside [1, %d] (reference kMaxPointsPerPixel)", fn, K, B200R_MAX_FACES_PER_PIXEL);
    if (dist < 0 || dist > 2 || rgb < 0 || rgb > 2 || alpha < 0 || alpha > 2 || tex < 0 || tex > 1)
        return b200r_fail(B200R_EINVAL, "%s: enum out of range (dist=%d rgb=%d alpha=%d tex=%d)", fn, dist, rgb, alpha, tex);
    if (tex == 1 && T != 3) return b200r_fail(B200R_EINVAL, "%s: vertex textures need texture_size 3, got %d", fn, T);
    if (tex == 0) {
        const int R = (int)std::sqrt((double)T);
        if (R * R != T) return b200r_fail(B200R_EINVAL, "%s: surface texture_size %d is not a square", fn, T);
    }
    if (!(sigma != 0.f) || !(gamma != 0.f)) return b200r_fail(B200R_EINVAL, "%s: sigma_val/gamma_val must be non-zero", fn);
    return 0;
}

SoftRasParams make_params(int B, int nf, int T, int is, int K, float near_, float far_, float eps, float sigma,
                          float gamma, float dist_eps, int dist, int rgb, int alpha, int tex, int double_side) {
    SoftRasParams P;
    P.B = B; P.nf = nf; P.T = T; P.R = (int)std::sqrt((double)T); P.is = is; P.K = K;
    P.near_ = near_; P.far_ = far_; P.eps = eps; P.sigma = sigma; P.gamma = gamma; P.dist_eps = dist_eps;
    P.dist_func = dist; P.rgb_func = rgb; P.alpha_func = alpha; P.tex_type = tex; P.double_side = double_side ? 1 : 0;
    b200r_geometry(is, &P.ntx, &P.coarse_px, &P.ncs);
    const long long total = (long long)P.ntx * P.ntx * B;
    long long stride = (long long)(total * 0.6180339887) | 1;
    while (stride > 1 && gcd_i(stride, total) != 1) stride += 2;
    P.tile_stride = (int)(stride % total == 0 ? 1 : stride);
    return P;
}

template <int DIST, int RGB, int VARIANT>
cudaError_t launch_forward_v(const SoftRasParams& P, const SoftRasWorkspace& W, const float* textures,
                             float* soft_colors, float* aggrs_info, int32_t* ids, cudaStream_t st) {
    size_t smem = sizeof(FwdSmem) + (size_t)P.K * B200R_TILE_THREADS * 8;
    if (VARIANT == 1) smem += (size_t)B200R_CHUNK * B200R_TILE_THREADS;
    cudaError_t e = cudaFuncSetAttribute(k_softras_forward<DIST, RGB, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int tiles = P.ntx * P.ntx;
    const bool persistent = g_fwd_persistent.load() != 0;
    int* counter = nullptr;
    dim3 grid(tiles, P.B);
    if (persistent) {
        counter = W.counters;
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_softras_forward<DIST, RGB, VARIANT>, B200R_TILE_THREADS, smem);
        if (occ < 1) occ = 1;
        const long long total = (long long)tiles * P.B;
        const long long slots = (long long)sm_count() * occ;
        grid = dim3((unsigned)(total < slots ? total : slots), 1);
    }
    {
        B200rProfScope prof(B200R_K_SOFTRAS_FWD, st);
        k_softras_forward<DIST, RGB, VARIANT><<<grid, B200R_TILE_THREADS, smem, st>>>(
            P, W.recs, W.rects, W.coarse_cnt, W.coarse_ids, textures, soft_colors, aggrs_info, ids, counter, W.tile_order);
    }
    return cudaGetLastError();
}

template <int DIST, int RGB>
cudaError_t launch_forward(const SoftRasParams& P, const SoftRasWorkspace& W, const float* textures,
                           float* soft_colors, float* aggrs_info, int32_t* ids, cudaStream_t st) {
    if (g_fwd_variant.load() == 0) return launch_forward_v<DIST, RGB, 0>(P, W, textures, soft_colors, aggrs_info, ids, st);
    return launch_forward_v<DIST, RGB, 1>(P, W, textures, soft_colors, aggrs_info, ids, st);
}

template <int DIST, int RGB>
cudaError_t launch_backward(const SoftRasParams& P, const SoftRasWorkspace& W, const float* textures,
                            const float* soft_colors, const float* aggrs_info, const int32_t* ids,
                            const float* grad_soft_colors, float* grad_faces, float* grad_textures, cudaStream_t st) {
    const size_t smem = (size_t)P.K * B200R_TILE_THREADS * 4 + 8 * sizeof(FaceRec);
    cudaError_t e = cudaFuncSetAttribute(k_softras_backward<DIST, RGB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(P.ntx * P.ntx, P.B);
    {
        B200rProfScope prof(B200R_K_SOFTRAS_BWD, st);
        k_softras_backward<DIST, RGB><<<grid, B200R_TILE_THREADS, smem, st>>>(P, W.recs, textures, soft_colors, aggrs_info, ids,
                                                                               grad_soft_colors, grad_faces, grad_textures);
    }
    return cudaGetLastError();
}

#define DISPATCH(FN, ...)                                                                   \
    do {                                                                                    \
        switch (P.dist_func * 3 + P.rgb_func) {                                             \
            case 0: e = FN<0, 0>(__VA_ARGS__); break;                                       \
            case 1: e = FN<0, 1>(__VA_ARGS__); break;                                       \
            case 2: e = FN<0, 2>(__VA_ARGS__); break;                                       \
            case 3: e = FN<1, 0>(__VA_ARGS__); break;                                       \
            case 4: e = FN<1, 1>(__VA_ARGS__); break;                                       \
            case 5: e = FN<1, 2>(__VA_ARGS__); break;                                       \
            case 6: e = FN<2, 0>(__VA_ARGS__); break;                                       \
            case 7: e = FN<2, 1>(__VA_ARGS__); break;                                       \
            default: e = FN<2, 2>(__VA_ARGS__); break;                                      \
        }                                                                                   \
    } while (0)

}  // namespace

extern "C" {

const char* b200r_version(void) { return "b200raster 0.1 (sm_100a)"; }

int b200r_set_option(const char* name, int value) {
    if (!name) return b200r_fail(B200R_EINVAL, "b200r_set_option: NULL name");
    if (!strcmp(name, "softras_fwd_variant")) { g_fwd_variant.store(value ? 1 : 0); return 0; }
    if (!strcmp(name, "softras_fwd_persistent")) { g_fwd_persistent.store(value ? 1 : 0); return 0; }
    return b200r_fail(B200R_EINVAL, "b200r_set_option: unknown option '%s'", name);
}

size_t b200r_softras_workspace_bytes(int batch_size, int num_faces, int image_size) {
    if (batch_size <= 0 || num_faces <= 0 || image_size <= 0) return 0;
    return b200r_carve(nullptr, batch_size, num_faces, image_size).bytes;
}

int b200r_softras_forward(const float* face_vertices, const float* textures, float* soft_colors,
                          float* aggrs_info, int32_t* faces_id_buffer, float* faces_info, void* workspace,
                          size_t workspace_bytes, int B, int nf, int T, int is, int K, float near_, float far_,
                          float eps, float sigma_val, float gamma_val, float dist_eps_logit, int dist_func,
                          int rgb_func, int alpha_func, int texture_type, int double_side, void* stream) {
    int rc = validate("b200r_softras_forward", B, nf, T, is, K, dist_func, rgb_func, alpha_func, texture_type, sigma_val, gamma_val);
    if (rc) return rc;
    if (!face_vertices || !textures || !soft_colors || !aggrs_info || !faces_id_buffer || !workspace)
        return b200r_fail(B200R_EINVAL, "b200r_softras_forward: NULL pointer argument");
    const SoftRasWorkspace W = b200r_carve(workspace, B, nf, is);
    if (workspace_bytes < W.bytes)
        return b200r_fail(B200R_EWORKSPACE, "b200r_softras_forward: workspace %zu < required %zu bytes", workspace_bytes, W.bytes);
    const SoftRasParams P = make_params(B, nf, T, is, K, near_, far_, eps, sigma_val, gamma_val, dist_eps_logit,
                                        dist_func, rgb_func, alpha_func, texture_type, double_side);
    cudaStream_t st = (cudaStream_t)stream;
    const float border = sqrtf(dist_eps_logit * sigma_val);  // sqrt(threshold), :289,:316 (IEEE sqrt on host == device)

    const int total = B * nf;
    {
        B200rProfScope prof(B200R_K_FACE_SETUP, st);
        k_face_setup<<<(total + 255) / 256, 256, 0, st>>>(face_vertices, textures, W.recs, W.rects, faces_info, total, nf, T, texture_type, is, border);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return b200r_cuda_fail(e, "k_face_setup");

    e = cudaMemsetAsync(W.counters, 0, 256 * sizeof(int), st);
    if (e != cudaSuccess) return b200r_cuda_fail(e, "memset counters");
    {
        B200rProfScope prof(B200R_K_COARSE_BIN, st);
        k_coarse_bin<<<dim3(P.ncs * P.ncs, B), 256, 0, st>>>(W.rects, W.coarse_cnt, W.coarse_ids, W.tile_cost, W.counters + 64,
                                                             nf, is, P.coarse_px, P.ncs, P.ntx);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return b200r_cuda_fail(e, "k_coarse_bin");
    {
        const int total_tiles = P.ntx * P.ntx * B;
        B200rProfScope prof(B200R_K_TILE_ORDER, st);
        k_tile_order<<<(total_tiles + 255) / 256, 256, 0, st>>>(W.tile_cost, W.counters + 64, W.counters + 128, W.tile_order, total_tiles);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return b200r_cuda_fail(e, "k_tile_order");

    DISPATCH(launch_forward, P, W, textures, soft_colors, aggrs_info, faces_id_buffer, st);
    if (e != cudaSuccess) return b200r_cuda_fail(e, "k_softras_forward");
    return 0;
}

int b200r_softras_backward(const float* face_vertices, const float* textures, const float* soft_colors,
                           const float* aggrs_info, const int32_t* faces_id_buffer, const void* workspace,
                           size_t workspace_bytes, const float* grad_soft_colors, float* grad_face_vertices,
                           float* grad_textures, int B, int nf, int T, int is, int K, float near_, float far_,
                           float eps, float sigma_val, float gamma_val, float dist_eps_logit, int dist_func,
                           int rgb_func, int alpha_func, int texture_type, int double_side, void* stream) {
    int rc = validate("b200r_softras_backward", B, nf, T, is, K, dist_func, rgb_func, alpha_func, texture_type, sigma_val, gamma_val);
    if (rc) return rc;
    if (!face_vertices || !textures || !soft_colors || !aggrs_info || !faces_id_buffer || !workspace ||
        !grad_soft_colors || !grad_face_vertices || !grad_textures)
        return b200r_fail(B200R_EINVAL, "b200r_softras_backward: NULL pointer argument");
    const SoftRasWorkspace W = b200r_carve(const_cast<void*>(workspace), B, nf, is);
    if (workspace_bytes < W.bytes)
        return b200r_fail(B200R_EWORKSPACE, "b200r_softras_backward: workspace %zu < required %zu bytes", workspace_bytes, W.bytes);
    const SoftRasParams P = make_params(B, nf, T, is, K, near_, far_, eps, sigma_val, gamma_val, dist_eps_logit,
                                        dist_func, rgb_func, alpha_func, texture_type, double_side);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad_face_vertices, 0, sizeof(float) * 9 * (size_t)B * nf, st);  // :1374
    if (e != cudaSuccess) return b200r_cuda_fail(e, "memset grad_faces");
    e = cudaMemsetAsync(grad_textures, 0, sizeof(float) * 3 * (size_t)T * B * nf, st);  // :1375
    if (e != cudaSuccess) return b200r_cuda_fail(e, "memset grad_textures");
    DISPATCH(launch_backward, P, W, textures, soft_colors, aggrs_info, faces_id_buffer, grad_soft_colors,
             grad_face_vertices, grad_textures, st);
    if (e != cudaSuccess) return b200r_cuda_fail(e, "k_softras_backward");
    return 0;
}

}  // extern "C"

// softras_bwd.cu -- backward kernel instantiations.
#include <atomic>

#include "api_util.cuh"
#include "softras_backward.cuh"
#include "softras_launch.cuh"

using namespace b200r;

namespace {
template <int DIST, int RGB>
cudaError_t launch_union(const SoftRasParams& P, const SoftRasWorkspace& W, const float* textures, const float* soft_colors,
                         const float* aggrs_info, const int32_t* ids, const float* grad_soft_colors, float* grad_faces,
                         float* grad_textures, cudaStream_t st) {
    const size_t smem = (size_t)P.K * B200R_TILE_THREADS * 4 + 8 * sizeof(FaceRec);
    static std::atomic<size_t> cfg_smem{0};
    if (cfg_smem.load() != smem) {
        cudaError_t e = cudaFuncSetAttribute(k_softras_backward<DIST, RGB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cfg_smem.store(smem);
    }
    dim3 grid(P.ntx * P.ntx, P.B);
    {
        B200rProfScope prof(B200R_K_SOFTRAS_BWD, st);
        k_softras_backward<DIST, RGB><<<grid, B200R_TILE_THREADS, smem, st>>>(P, W.recs, textures, soft_colors, aggrs_info, ids,
                                                                               grad_soft_colors, grad_faces, grad_textures);
    }
    return cudaGetLastError();
}

template <int DIST, int RGB, bool EXACT>
cudaError_t launch_lane(const SoftRasParams& P, const SoftRasWorkspace& W, const float* textures, const float* soft_colors,
                        const float* aggrs_info, const int32_t* ids, const float* grad_soft_colors, float* grad_textures,
                        cudaStream_t st) {
    dim3 grid(P.ntx * P.ntx, P.B);
    {
        B200rProfScope prof(B200R_K_SOFTRAS_BWD, st);
        k_softras_backward_lane<DIST, RGB, EXACT><<<grid, B200R_TILE_THREADS, 0, st>>>(P, W.recs, textures, soft_colors, aggrs_info, ids,
                                                                                 grad_soft_colors, W.gacc, grad_textures);
    }
    return cudaGetLastError();
}
}  // namespace

// variant 1: memset(gacc) -> per-lane kernel -> finalize (writes every element of grad_faces, and of
// grad_textures when T == 1 surface); variant 0: memset(outputs) -> union-walk kernel.
cudaError_t b200r_launch_backward(const SoftRasParams& P, const SoftRasWorkspace& W, const float* textures,
                                  const float* soft_colors, const float* aggrs_info, const int32_t* ids,
                                  const float* grad_soft_colors, float* grad_faces, float* grad_textures, int variant,
                                  int exact, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    const size_t nfaces = (size_t)P.B * P.nf;
    const bool tex_in_acc = (P.tex_type == 0 && P.T == 1);
    if (variant == 0) {
        e = cudaMemsetAsync(grad_faces, 0, sizeof(float) * 9 * nfaces, st);   // reference :1374
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(grad_textures, 0, sizeof(float) * 3 * (size_t)P.T * nfaces, st);  // :1375
        if (e != cudaSuccess) return e;
        B200R_DISPATCH_DIST_RGB((e = launch_union<D, R>(P, W, textures, soft_colors, aggrs_info, ids, grad_soft_colors, grad_faces, grad_textures, st)))
        return e;
    }
    e = cudaMemsetAsync(W.gacc, 0, sizeof(float) * 12 * nfaces, st);
    if (e != cudaSuccess) return e;
    if (!tex_in_acc) {
        e = cudaMemsetAsync(grad_textures, 0, sizeof(float) * 3 * (size_t)P.T * nfaces, st);
        if (e != cudaSuccess) return e;
    }
    B200R_DISPATCH_DIST_RGB((e = (D == 2 && !exact) ? launch_lane<D, R, (D != 2)>(P, W, textures, soft_colors, aggrs_info, ids, grad_soft_colors, grad_textures, st) : launch_lane<D, R, true>(P, W, textures, soft_colors, aggrs_info, ids, grad_soft_colors, grad_textures, st)))
    if (e != cudaSuccess) return e;
    {
        B200rProfScope prof(B200R_K_SOFTRAS_BWD_FINALIZE, st);
        k_softras_bwd_finalize<<<(unsigned)((nfaces + 255) / 256), 256, 0, st>>>(W.gacc, grad_faces, grad_textures, (int)nfaces, tex_in_acc ? 1 : 0);
    }
    return cudaGetLastError();
}

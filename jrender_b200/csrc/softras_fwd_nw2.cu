// softras_fwd_nw2.cu -- forward kernel instantiations for the 2x1 warp layout (16x4 pixel tiles).
#include <atomic>

#include "api_util.cuh"
#include "softras_forward.cuh"
#include "softras_launch.cuh"

using namespace b200r;

namespace {
template <int DIST, int RGB, int VARIANT, bool EXACT>
cudaError_t launch_v(const SoftRasParams& P, const SoftRasWorkspace& W, const float* textures, float* soft_colors,
                     float* aggrs_info, int32_t* ids, int persistent, cudaStream_t st) {
    constexpr int WX = 2, WY = 1, NW = WX * WY, NT = 32 * NW;
    const size_t smem = fwd_smem_bytes<NW>(P.K, VARIANT);
    // attribute + occupancy are host calls: queried once per (instantiation, smem size)
    static std::atomic<size_t> cfg_smem{0};
    static std::atomic<int> cfg_occ{0};
    if (cfg_smem.load() != smem) {
        cudaError_t e = cudaFuncSetAttribute(k_softras_forward<DIST, RGB, VARIANT, WX, WY, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // The persistent grid is sized for the occupancy the FULL shared-memory carve-out allows.  Left to the driver's
        // per-function heuristic the carve-out sometimes came out smaller: the same kernel then ran with fewer resident
        // CTAs than the grid assumed and was 25-35 % slower for the lifetime of the process (seen on C2 / C5).
        e = cudaFuncSetAttribute(k_softras_forward<DIST, RGB, VARIANT, WX, WY, EXACT>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        int occ = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_softras_forward<DIST, RGB, VARIANT, WX, WY, EXACT>, NT, smem);
        if (e != cudaSuccess) return e;
        cfg_occ.store(occ < 1 ? 1 : occ);
        cfg_smem.store(smem);
    }
    const int tiles = P.fntx * P.fnty;
    int* counter = nullptr;
    dim3 grid(tiles, P.B);
    if (persistent) {
        counter = W.counters;
        const long long total = (long long)tiles * P.B;
        const long long slots = (long long)b200r_sm_count() * cfg_occ.load();
        grid = dim3((unsigned)(total < slots ? total : slots), 1);
    }
    {
        B200rProfScope prof(B200R_K_SOFTRAS_FWD, st);
        k_softras_forward<DIST, RGB, VARIANT, WX, WY, EXACT><<<grid, NT, smem, st>>>(
            P, W.recs, W.rects, W.coarse_cnt, W.coarse_ids, textures, soft_colors, aggrs_info, ids, counter, W.tile_order);
    }
    return cudaGetLastError();
}
}  // namespace

cudaError_t b200r_launch_forward_nw2(const SoftRasParams& P, const SoftRasWorkspace& W, const float* textures, float* soft_colors,
                                     float* aggrs_info, int32_t* ids, int variant, int persistent, int exact, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    (void)variant;  // only the per-lane-list variant is built for this layout
    B200R_DISPATCH_DIST_RGB((e = (D == 2 && R == 1 && !exact) ? launch_v<D, R, 1, (D != 2 || R != 1)>(P, W, textures, soft_colors, aggrs_info, ids, persistent, st) : launch_v<D, R, 1, true>(P, W, textures, soft_colors, aggrs_info, ids, persistent, st)))
    return e;
}

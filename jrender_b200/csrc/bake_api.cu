// bake_api.cu -- texture baking for SoftRas surface textures.
//
//   b200r_bake_textures_softras  <-  _load_textures_for_softras / load_textures_cuda_kernel,
//                                    jrender/io/utils/load_textures.py:3-101
//
// One thread per texel of [nf, R*R, 3]: the texel's barycentric sample point (two triangles of R x R cells per face,
// :29-37), its position in the texture image through the face's UVs (:42-45) and a bilinear fetch (:46-58).  The
// reference kernel promotes through double literals (`1. / 3.`, `1. - w0 - w1`) and narrows to float on assignment;
// the same promotions are made here.  Deviation, documented: the reference reads the image without bounds checks (a UV
// of exactly 1 reads one texel past the row / the image); out-of-image taps are clamped to the last valid element here.
#include "../../include/b200raster.h"
#include "api_util.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_bake_textures_softras(const float* __restrict__ image, const float* __restrict__ faces_uv, const int32_t* __restrict__ is_update,
                        float* __restrict__ textures, long ntexels, int R, int H, int W) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntexels) return;
    const int fn = (int)(i / (R * R));
    if (__ldg(is_update + fn) == 0) return;                    // :41
    const int w_y = (int)(i % (R * R)) / R, w_x = (int)(i % R);
    float w0, w1, w2;
    if (w_x + w_y < R) {                                       // :29-37
        w0 = (float)((w_x + 1. / 3.) / R);
        w1 = (float)((w_y + 1. / 3.) / R);
    } else {
        w0 = (float)(((R - 1. - w_x) + 2. / 3.) / R);
        w1 = (float)(((R - 1. - w_y) + 2. / 3.) / R);
    }
    w2 = (float)(1. - (double)w0 - (double)w1);
    const float* face = faces_uv + (size_t)fn * 6;
    const float pos_x = (__ldg(face + 0) * w0 + __ldg(face + 2) * w1 + __ldg(face + 4) * w2) * (float)(W - 1);   // :42-43
    const float pos_y = (__ldg(face + 1) * w0 + __ldg(face + 3) * w1 + __ldg(face + 5) * w2) * (float)(H - 1);   // :44-45
    const float weight_x1 = pos_x - (int)pos_x, weight_x0 = 1 - weight_x1;
    const float weight_y1 = pos_y - (int)pos_y, weight_y0 = 1 - weight_y1;
    const long n = (long)H * W;
    const long i00 = (long)(int)pos_y * W + (int)pos_x, i10 = (long)(int)(pos_y + 1) * W + (int)pos_x;
    const long i01 = i00 + 1, i11 = i10 + 1;
    auto px = [&](long j, int k) { return __ldg(image + (j < 0 ? 0 : (j >= n ? n - 1 : j)) * 3 + k); };
    float* texture = textures + i * 3;
#pragma unroll
    for (int k = 0; k < 3; k++) {                              // :50-57, same accumulation order
        float c = 0;
        c += px(i00, k) * (weight_x0 * weight_y0);
        c += px(i10, k) * (weight_x0 * weight_y1);
        c += px(i01, k) * (weight_x1 * weight_y0);
        c += px(i11, k) * (weight_x1 * weight_y1);
        texture[k] = c;
    }
}

}  // namespace

extern "C" int b200r_bake_textures_softras(const float* image, const float* faces_uv, const int32_t* is_update, float* textures,
                                           int nf, int texture_res, int image_height, int image_width, void* stream) {
    if (!image || !faces_uv || !is_update || !textures) return b200r_fail(B200R_EINVAL, "b200r_bake_textures_softras: NULL pointer argument");
    if (nf <= 0 || texture_res <= 0 || image_height <= 0 || image_width <= 0)
        return b200r_fail(B200R_EINVAL, "b200r_bake_textures_softras: non-positive size");
    cudaStream_t st = (cudaStream_t)stream;
    const long ntexels = (long)nf * texture_res * texture_res;
    {
        B200rProfScope prof(B200R_K_BAKE, st);
        k_bake_textures_softras<<<(unsigned)((ntexels + 255) / 256), 256, 0, st>>>(image, faces_uv, is_update, textures, ntexels, texture_res,
                                                                                  image_height, image_width);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return b200r_cuda_fail(e, "k_bake_textures_softras");
    return 0;
}

// lighting_api.cu -- fused surface-mode lighting of the pre-raster stage.
//
//   b200r_surface_lighting_forward / _backward  <-  Lighting.execute, light_mode='surface', no normal map, no SSS
//        jrender/renderer/lighting/lighting.py:186-204         (zeros -> ambient -> directional -> clamp(tex * diffuse + specular))
//        jrender/renderer/lighting/ambient_lighting.py:4-9
//        jrender/renderer/lighting/directional_lighting.py:54-135 (diffuse branch and the Cook-Torrance branch the
//                                                                reference takes by default: Mesh.with_specular = True)
//        jrender/structures/mesh.py:213-229                    (surface normals: float64 cross product, normalised)
//
// One thread per (batch, face): gathers the face's three world-space vertices, forms the normal and the centroid,
// evaluates the two light terms (6 floats) and streams the face's T texels through  clamp(tex * diffuse + specular, 0, 1).
// The reference runs this as ~30 tensor ops over [B, nf, 3] intermediates (and as many again backward); in the
// launch-bound demo loops that was ~230 launches per iteration.  Backward: the same expression re-evaluated in
// forward-mode automatic differentiation over the 9 vertex coordinates (dual.cuh) -- the exact gradient of the
// expression the forward evaluates, including the path through the normal, with 9 atomics per face into grad_vertices.
#include "../../include/b200raster.h"
#include "api_util.cuh"
#include "dual.cuh"

using namespace b200r;

namespace {

struct LightParams {
    float amb_i, amb_c[3];
    float dir_i, dir_c[3], dir[3];   // dir: normalised on the host exactly like F.normalize(light_direction, eps=1e-12)
    int with_specular;               // Cook-Torrance branch (needs eye + metallic + roughness)
};

template <class S> __device__ __forceinline__ S lift(float x);
template <> __device__ __forceinline__ float lift<float>(float x) { return x; }
template <> __device__ __forceinline__ Dual<9> lift<Dual<9>>(float x) { return dconst<9>(x); }

template <class S>
__device__ __forceinline__ S dot3(const S a[3], const S b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }   // torch.sum(a * b, dim): x, y, z order
template <class S>
__device__ __forceinline__ S dot3f(const S a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// x / max(||x||, eps)   (F.normalize / jt.normalize)
template <class S>
__device__ __forceinline__ void normalize3(S x[3], float eps) {
    const S n = dmax(dsqrt(dot3(x, x)), eps);
#pragma unroll
    for (int k = 0; k < 3; k++) x[k] = x[k] / n;
}

// diffuse[3], specular[3] of one face from its normal N and centroid pos (directional_lighting.py:63-135, 3-D normals)
template <class S>
__device__ __forceinline__ void face_light(const S N[3], const S pos[3], const LightParams& L, const float* eye, float metallic,
                                           float roughness, S diffuse[3], S specular[3]) {
    const S cosine = dmax(dot3f(N, L.dir), 0.f);   // relu(sum(normals * light_direction)) :66
#pragma unroll
    for (int c = 0; c < 3; c++) diffuse[c] = lift<S>(0.f + L.amb_i * L.amb_c[c]);   // zeros + ambient (ambient_lighting.py:8)
    if (L.with_specular && eye != nullptr) {
        S V[3] = {eye[0] - pos[0], eye[1] - pos[1], eye[2] - pos[2]};
        normalize3(V, 1e-12f);
        S H[3] = {V[0] + L.dir[0], V[1] + L.dir[1], V[2] + L.dir[2]};
        normalize3(H, 1e-12f);
        const float a = roughness * roughness, a2 = a * a;
        const S NdotH = dmax(dot3(N, H), 0.f);
        S den = (NdotH * NdotH) * (a2 - 1.0f) + 1.0f;          // GGX :17
        den = (den * 3.1415f) * den;                            // 3.1415 * denom * denom :18
        const S NDF = lift<S>(a2) / den;
        const float r1 = roughness + 1.0f, k = (r1 * r1) / 8.0f;
        const S NdotV = dmax(dot3(N, V), 0.f), NdotL = dmax(dot3f(N, L.dir), 0.f);
        const S G = (NdotL / (NdotL * (1.0f - k) + k)) * (NdotV / (NdotV * (1.0f - k) + k));   // GeometrySmith :34-46
        const S hv = 1.0f - dmax(dot3(H, V), 0.f);
        const S hv2 = hv * hv;
        const S p5 = (hv2 * hv2) * hv;                          // pow(1 - cosTheta, 5)
        S den4 = dmax((NdotV * 4.0f) * NdotL, 0.01f);          // clamp(4 N.V N.L, 0.01) :125-128
        const S ndf_g = NDF * G;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float F0 = 0.4f * (1.f - metallic) + 1.0f * metallic;   // :103
            const S Fr = F0 + (1.0f - F0) * p5;                           // fresnelSchlick :48-53
            const S radiance = (cosine * L.dir_c[c]) * L.dir_i;           // :105
            const S KD = (1.0f - Fr) * (1.0f - metallic);                 // :117-119
            diffuse[c] = diffuse[c] + KD * radiance;
            specular[c] = ((ndf_g * Fr) / den4) * radiance;               // :122-129
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            diffuse[c] = diffuse[c] + (cosine * L.dir_c[c]) * L.dir_i;    // :131-135
            specular[c] = lift<S>(0.f);
        }
    }
}

__device__ __forceinline__ float mean_tm(const float* p, int Tm) {   // torch.sum(x, dim=2) / total
    float s = 0.f;
    for (int t = 0; t < Tm; t++) s += __ldg(p + t);
    return s / (float)Tm;
}

// surface normal the way the mirror forms it: float64 cross product of float32 edge vectors, normalised, back to float32
__device__ __forceinline__ void normal_f64(const float v[9], float N[3]) {
    const double ax = (double)(v[0] - v[3]), ay = (double)(v[1] - v[4]), az = (double)(v[2] - v[5]);   // v10 = v0 - v1
    const double bx = (double)(v[6] - v[3]), by = (double)(v[7] - v[4]), bz = (double)(v[8] - v[5]);   // v12 = v2 - v1
    const double cx = by * az - bz * ay, cy = bz * ax - bx * az, cz = bx * ay - by * ax;               // cross(v12, v10)
    const double n = fmax(sqrt(cx * cx + cy * cy + cz * cz), 1e-12);
    N[0] = (float)(cx / n); N[1] = (float)(cy / n); N[2] = (float)(cz / n);
}

__global__ void __launch_bounds__(128)
k_surface_lighting_fwd(const float* __restrict__ vertices, const int32_t* __restrict__ faces, const float* __restrict__ textures,
                       const float* __restrict__ metallic, const float* __restrict__ roughness, const float* __restrict__ eye,
                       float* __restrict__ out, LightParams L, int B, int Bv, int Bf, int Be, int nv, int nf, int T, int Tm) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * nf) return;
    const int b = (int)(i / nf), f = (int)(i % nf);
    const int32_t* fi = faces + ((size_t)(Bf == 1 ? 0 : b) * nf + f) * 3;
    const float* vb = vertices + (size_t)(Bv == 1 ? 0 : b) * nv * 3;
    float v[9];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float* q = vb + (size_t)__ldg(fi + k) * 3;
        v[3 * k] = __ldg(q); v[3 * k + 1] = __ldg(q + 1); v[3 * k + 2] = __ldg(q + 2);
    }
    float N[3], pos[3], d[3], s[3];
    normal_f64(v, N);
#pragma unroll
    for (int k = 0; k < 3; k++) pos[k] = ((v[k] + v[3 + k]) + v[6 + k]) / 3.0f;   // torch.sum(face_vertices, dim=2) / 3.0
    const bool spec = L.with_specular && metallic != nullptr && roughness != nullptr && eye != nullptr;
    LightParams Lf = L;
    Lf.with_specular = spec ? 1 : 0;
    const float m = spec ? mean_tm(metallic + (size_t)i * Tm, Tm) : 0.f;
    const float r = spec ? mean_tm(roughness + (size_t)i * Tm, Tm) : 1.f;
    face_light<float>(N, pos, Lf, spec ? eye + (size_t)(Be == 1 ? 0 : b) * 3 : nullptr, m, r, d, s);
    const float* tx = textures + (size_t)i * T * 3;
    float* o = out + (size_t)i * T * 3;
    for (int t = 0; t < T; t++)
#pragma unroll
        for (int c = 0; c < 3; c++) o[t * 3 + c] = fminf(fmaxf(__ldg(tx + t * 3 + c) * d[c] + 1.0f * s[c], 0.0f), 1.0f);   // lighting.py:199-204
}

__global__ void __launch_bounds__(128)
k_surface_lighting_bwd(const float* __restrict__ vertices, const int32_t* __restrict__ faces, const float* __restrict__ textures,
                       const float* __restrict__ metallic, const float* __restrict__ roughness, const float* __restrict__ eye,
                       const float* __restrict__ grad_out, float* __restrict__ grad_textures, float* __restrict__ grad_vertices,
                       LightParams L, int B, int Bv, int Bf, int Be, int nv, int nf, int T, int Tm) {
    typedef Dual<9> D;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * nf) return;
    const int b = (int)(i / nf), f = (int)(i % nf);
    const int32_t* fi = faces + ((size_t)(Bf == 1 ? 0 : b) * nf + f) * 3;
    const size_t vbase = (size_t)(Bv == 1 ? 0 : b) * nv * 3;
    int idx[3];
    float v[9];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        idx[k] = __ldg(fi + k);
        const float* q = vertices + vbase + (size_t)idx[k] * 3;
        v[3 * k] = __ldg(q); v[3 * k + 1] = __ldg(q + 1); v[3 * k + 2] = __ldg(q + 2);
    }
    const bool spec = L.with_specular && metallic != nullptr && roughness != nullptr && eye != nullptr;
    LightParams Lf = L;
    Lf.with_specular = spec ? 1 : 0;
    const float m = spec ? mean_tm(metallic + (size_t)i * Tm, Tm) : 0.f;
    const float r = spec ? mean_tm(roughness + (size_t)i * Tm, Tm) : 1.f;
    const float* ey = spec ? eye + (size_t)(Be == 1 ? 0 : b) * 3 : nullptr;
    // values exactly as the forward formed them (masks of the clamp, d out / d tex)
    float N[3], pos[3], d[3], s[3];
    normal_f64(v, N);
#pragma unroll
    for (int k = 0; k < 3; k++) pos[k] = ((v[k] + v[3 + k]) + v[6 + k]) / 3.0f;
    face_light<float>(N, pos, Lf, ey, m, r, d, s);
    // upstream gradient folded over the texels: Gd[c] = sum_t g * mask * tex, Gs[c] = sum_t g * mask
    float Gd[3] = {0.f, 0.f, 0.f}, Gs[3] = {0.f, 0.f, 0.f};
    const float* tx = textures + (size_t)i * T * 3;
    const float* go = grad_out + (size_t)i * T * 3;
    float* gt = grad_textures + (size_t)i * T * 3;
    for (int t = 0; t < T; t++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float x = __ldg(tx + t * 3 + c), val = x * d[c] + 1.0f * s[c];
            const float g = (val >= 0.0f && val <= 1.0f) ? __ldg(go + t * 3 + c) : 0.f;   // torch.clamp passes the gradient on [min, max]
            gt[t * 3 + c] = g * d[c];
            Gd[c] += g * x;
            Gs[c] += g;
        }
    if (grad_vertices == nullptr) return;
    // derivatives of (diffuse, specular) w.r.t. the 9 vertex coordinates
    D dv[9];
#pragma unroll
    for (int k = 0; k < 9; k++) dv[k] = dvar<9>(v[k], k);
    D a[3] = {dv[0] - dv[3], dv[1] - dv[4], dv[2] - dv[5]};    // v10
    D bb[3] = {dv[6] - dv[3], dv[7] - dv[4], dv[8] - dv[5]};   // v12
    D Nd[3] = {bb[1] * a[2] - bb[2] * a[1], bb[2] * a[0] - bb[0] * a[2], bb[0] * a[1] - bb[1] * a[0]};
    normalize3(Nd, 1e-12f);
    D pd[3];
#pragma unroll
    for (int k = 0; k < 3; k++) pd[k] = ((dv[k] + dv[3 + k]) + dv[6 + k]) / 3.0f;
    D dd[3], sd[3];
    face_light<D>(Nd, pd, Lf, ey, m, r, dd, sd);
    float* gv = grad_vertices + vbase;
#pragma unroll
    for (int k = 0; k < 9; k++) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) acc += Gd[c] * dd[c].d[k] + Gs[c] * sd[c].d[k];
        if (acc != 0.f) atomicAdd(gv + (size_t)idx[k / 3] * 3 + (k % 3), acc);
    }
}

int check_args(const char* fn, const void* vertices, const void* faces, const void* textures, int B, int Bv, int Bf, int Be, int nv,
               int nf, int T, int Tm) {
    if (!vertices || !faces || !textures) return b200r_fail(B200R_EINVAL, "%s: NULL pointer argument", fn);
    if (B <= 0 || nv <= 0 || nf <= 0 || T <= 0 || Tm < 0) return b200r_fail(B200R_EINVAL, "%s: non-positive size", fn);
    if ((Bv != 1 && Bv != B) || (Bf != 1 && Bf != B) || (Be != 0 && Be != 1 && Be != B))
        return b200r_fail(B200R_EINVAL, "%s: vertices / faces / eye batch must be 1 or B", fn);
    return 0;
}

LightParams make(float amb_i, const float* amb_c, float dir_i, const float* dir_c, const float* dir, int with_specular) {
    LightParams L;
    L.amb_i = amb_i; L.dir_i = dir_i; L.with_specular = with_specular;
    for (int k = 0; k < 3; k++) { L.amb_c[k] = amb_c[k]; L.dir_c[k] = dir_c[k]; L.dir[k] = dir[k]; }
    return L;
}

}  // namespace

extern "C" {

int b200r_surface_lighting_forward(const float* vertices, const int32_t* faces, const float* textures, const float* metallic,
                                   const float* roughness, const float* eye, float* out_textures, int B, int vertices_batch,
                                   int faces_batch, int eye_batch, int nv, int nf, int T, int Tm, float ambient_intensity,
                                   const float* ambient_color, float directional_intensity, const float* directional_color,
                                   const float* direction, int with_specular, void* stream) {
    int rc = check_args("b200r_surface_lighting_forward", vertices, faces, textures, B, vertices_batch, faces_batch, eye_batch, nv, nf, T, Tm);
    if (rc) return rc;
    if (!out_textures || !ambient_color || !directional_color || !direction)
        return b200r_fail(B200R_EINVAL, "b200r_surface_lighting_forward: NULL pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    const LightParams L = make(ambient_intensity, ambient_color, directional_intensity, directional_color, direction, with_specular);
    const long total = (long)B * nf;
    {
        B200rProfScope prof(B200R_K_LIGHTING, st);
        k_surface_lighting_fwd<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(vertices, faces, textures, metallic, roughness,
                                                                             eye_batch ? eye : nullptr, out_textures, L, B, vertices_batch,
                                                                             faces_batch, eye_batch ? eye_batch : 1, nv, nf, T, Tm);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return b200r_cuda_fail(e, "k_surface_lighting_fwd");
    return 0;
}

int b200r_surface_lighting_backward(const float* vertices, const int32_t* faces, const float* textures, const float* metallic,
                                    const float* roughness, const float* eye, const float* grad_out, float* grad_textures,
                                    float* grad_vertices, int B, int vertices_batch, int faces_batch, int eye_batch, int nv, int nf,
                                    int T, int Tm, float ambient_intensity, const float* ambient_color, float directional_intensity,
                                    const float* directional_color, const float* direction, int with_specular, void* stream) {
    int rc = check_args("b200r_surface_lighting_backward", vertices, faces, textures, B, vertices_batch, faces_batch, eye_batch, nv, nf, T, Tm);
    if (rc) return rc;
    if (!grad_out || !grad_textures || !ambient_color || !directional_color || !direction)
        return b200r_fail(B200R_EINVAL, "b200r_surface_lighting_backward: NULL pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (grad_vertices != nullptr) {
        cudaError_t e = cudaMemsetAsync(grad_vertices, 0, sizeof(float) * 3 * (size_t)vertices_batch * nv, st);
        if (e != cudaSuccess) return b200r_cuda_fail(e, "memset grad_vertices");
    }
    const LightParams L = make(ambient_intensity, ambient_color, directional_intensity, directional_color, direction, with_specular);
    const long total = (long)B * nf;
    {
        B200rProfScope prof(B200R_K_LIGHTING, st);
        k_surface_lighting_bwd<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(vertices, faces, textures, metallic, roughness,
                                                                             eye_batch ? eye : nullptr, grad_out, grad_textures,
                                                                             grad_vertices, L, B, vertices_batch, faces_batch,
                                                                             eye_batch ? eye_batch : 1, nv, nf, T, Tm);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return b200r_cuda_fail(e, "k_surface_lighting_bwd");
    return 0;
}

}  // extern "C"

// sh.cu -- the spherical-harmonics colour branch of the operator API (shs / sh_degree > 0;
// computeColorFromSH forward X/cuda_rasterizer/forward.cu:20-71 and backward X/cuda_rasterizer/backward.cu:20-139).
// SplaTAM never uses it (it always passes colors_precomp, R/utils/slam_helpers.py:131-138); it is here so that
// the other callers of the reference operator (R/scripts/gaussian_splatting.py-style SH models) drop in too.
// Two standalone kernels: SH -> RGB (+ clamp flags) before the render, and RGB-gradient -> SH-gradient (+ the
// view-direction term of dL/dmean) after the per-Gaussian backward.
#include "common.cuh"

namespace sb {

namespace {

__device__ const float kC0 = 0.28209479177387814f;
__device__ const float kC1 = 0.4886025119029199f;
__device__ const float kC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                 -1.0925484305920792f, 0.5462742152960396f};
__device__ const float kC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                 -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 operator*(float a, V3 v) { return {a * v.x, a * v.y, a * v.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

__global__ void __launch_bounds__(256)
sh_forward_kernel(int P, int deg, int M, const float* __restrict__ means, const float* __restrict__ campos,
                  const float* __restrict__ shs, float* __restrict__ rgb, uint8_t* __restrict__ clamped) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const V3 pos = {means[3 * idx], means[3 * idx + 1], means[3 * idx + 2]};
    V3 dir = pos - V3{campos[0], campos[1], campos[2]};
    const float len = sqrtf(dot(dir, dir));
    dir = {dir.x / len, dir.y / len, dir.z / len};
    const V3* sh = reinterpret_cast<const V3*>(shs) + (size_t)idx * M;
    V3 result = kC0 * sh[0];
    if (deg > 0) {
        const float x = dir.x, y = dir.y, z = dir.z;
        result = result - (kC1 * y) * sh[1] + (kC1 * z) * sh[2] - (kC1 * x) * sh[3];
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            result = result + (kC2[0] * xy) * sh[4] + (kC2[1] * yz) * sh[5] + (kC2[2] * (2.0f * zz - xx - yy)) * sh[6] +
                     (kC2[3] * xz) * sh[7] + (kC2[4] * (xx - yy)) * sh[8];
            if (deg > 2) {
                result = result + (kC3[0] * y * (3.0f * xx - yy)) * sh[9] + (kC3[1] * xy * z) * sh[10] +
                         (kC3[2] * y * (4.0f * zz - xx - yy)) * sh[11] +
                         (kC3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy)) * sh[12] +
                         (kC3[4] * x * (4.0f * zz - xx - yy)) * sh[13] + (kC3[5] * z * (xx - yy)) * sh[14] +
                         (kC3[6] * x * (xx - 3.0f * yy)) * sh[15];
            }
        }
    }
    result = result + V3{0.5f, 0.5f, 0.5f};
    // colours are clamped to >= 0; remember where, the gradient is zero there (forward.cu:64-70)
    clamped[idx] = (uint8_t)((result.x < 0 ? 1 : 0) | (result.y < 0 ? 2 : 0) | (result.z < 0 ? 4 : 0));
    rgb[3 * idx] = fmaxf(result.x, 0.f); rgb[3 * idx + 1] = fmaxf(result.y, 0.f); rgb[3 * idx + 2] = fmaxf(result.z, 0.f);
}

__global__ void __launch_bounds__(256)
sh_backward_kernel(int P, int deg, int M, const float* __restrict__ means, const float* __restrict__ campos,
                   const float* __restrict__ shs, const uint8_t* __restrict__ clamped,
                   const float* __restrict__ dL_drgb, float* __restrict__ dL_dshs, float* __restrict__ dL_dmeans) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const V3 pos = {means[3 * idx], means[3 * idx + 1], means[3 * idx + 2]};
    const V3 dir_orig = pos - V3{campos[0], campos[1], campos[2]};
    const float len = sqrtf(dot(dir_orig, dir_orig));
    const V3 dir = {dir_orig.x / len, dir_orig.y / len, dir_orig.z / len};
    const V3* sh = reinterpret_cast<const V3*>(shs) + (size_t)idx * M;
    const uint8_t cl = clamped[idx];
    V3 g = {dL_drgb[3 * idx], dL_drgb[3 * idx + 1], dL_drgb[3 * idx + 2]};
    g = {(cl & 1) ? 0.f : g.x, (cl & 2) ? 0.f : g.y, (cl & 4) ? 0.f : g.z};
    V3* out = reinterpret_cast<V3*>(dL_dshs) + (size_t)idx * M;
    V3 dx = {0, 0, 0}, dy = {0, 0, 0}, dz = {0, 0, 0};
    const float x = dir.x, y = dir.y, z = dir.z;
    out[0] = kC0 * g;
    if (deg > 0) {
        out[1] = (-kC1 * y) * g; out[2] = (kC1 * z) * g; out[3] = (-kC1 * x) * g;
        dx = (-kC1) * sh[3]; dy = (-kC1) * sh[1]; dz = kC1 * sh[2];
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            out[4] = (kC2[0] * xy) * g; out[5] = (kC2[1] * yz) * g; out[6] = (kC2[2] * (2.f * zz - xx - yy)) * g;
            out[7] = (kC2[3] * xz) * g; out[8] = (kC2[4] * (xx - yy)) * g;
            dx = dx + (kC2[0] * y) * sh[4] + (kC2[2] * 2.f * -x) * sh[6] + (kC2[3] * z) * sh[7] + (kC2[4] * 2.f * x) * sh[8];
            dy = dy + (kC2[0] * x) * sh[4] + (kC2[1] * z) * sh[5] + (kC2[2] * 2.f * -y) * sh[6] + (kC2[4] * 2.f * -y) * sh[8];
            dz = dz + (kC2[1] * y) * sh[5] + (kC2[2] * 2.f * 2.f * z) * sh[6] + (kC2[3] * x) * sh[7];
            if (deg > 2) {
                out[9] = (kC3[0] * y * (3.f * xx - yy)) * g; out[10] = (kC3[1] * xy * z) * g;
                out[11] = (kC3[2] * y * (4.f * zz - xx - yy)) * g;
                out[12] = (kC3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * g;
                out[13] = (kC3[4] * x * (4.f * zz - xx - yy)) * g; out[14] = (kC3[5] * z * (xx - yy)) * g;
                out[15] = (kC3[6] * x * (xx - 3.f * yy)) * g;
                dx = dx + (kC3[0] * 3.f * 2.f * xy) * sh[9] + (kC3[1] * yz) * sh[10] + (kC3[2] * -2.f * xy) * sh[11] +
                     (kC3[3] * -3.f * 2.f * xz) * sh[12] + (kC3[4] * (-3.f * xx + 4.f * zz - yy)) * sh[13] +
                     (kC3[5] * 2.f * xz) * sh[14] + (kC3[6] * 3.f * (xx - yy)) * sh[15];
                dy = dy + (kC3[0] * 3.f * (xx - yy)) * sh[9] + (kC3[1] * xz) * sh[10] +
                     (kC3[2] * (-3.f * yy + 4.f * zz - xx)) * sh[11] + (kC3[3] * -3.f * 2.f * yz) * sh[12] +
                     (kC3[4] * -2.f * xy) * sh[13] + (kC3[5] * -2.f * yz) * sh[14] + (kC3[6] * -3.f * 2.f * xy) * sh[15];
                dz = dz + (kC3[1] * xy) * sh[10] + (kC3[2] * 4.f * 2.f * yz) * sh[11] +
                     (kC3[3] * 3.f * (2.f * zz - xx - yy)) * sh[12] + (kC3[4] * 4.f * 2.f * xz) * sh[13] +
                     (kC3[5] * (xx - yy)) * sh[14];
            }
        }
    }
    for (int k = (deg + 1) * (deg + 1); k < M; ++k) out[k] = {0.f, 0.f, 0.f};
    // view direction depends on the mean: dnormvdv (auxiliary.h:107-117)
    const V3 dd = {dot(dx, g), dot(dy, g), dot(dz, g)};
    const V3 v = dir_orig;
    const float sum2 = dot(v, v), inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dL_dmeans[3 * idx] += ((sum2 - v.x * v.x) * dd.x - v.y * v.x * dd.y - v.z * v.x * dd.z) * inv32;
    dL_dmeans[3 * idx + 1] += (-v.x * v.y * dd.x + (sum2 - v.y * v.y) * dd.y - v.z * v.y * dd.z) * inv32;
    dL_dmeans[3 * idx + 2] += (-v.x * v.z * dd.x - v.y * v.z * dd.y + (sum2 - v.z * v.z) * dd.z) * inv32;
}

}  // namespace

}  // namespace sb

using namespace sb;

extern "C" {

SB_API int sb_sh_forward(int P, int sh_degree, int max_coeffs, const float* means3D, const float* campos,
                         const float* shs, float* rgb, uint8_t* clamped, void* stream) {
    if (P < 0 || sh_degree < 0 || sh_degree > 3 || max_coeffs < (sh_degree + 1) * (sh_degree + 1)) return SB_ERR_BAD_ARG;
    if (P == 0) return SB_OK;
    if (!means3D || !campos || !shs || !rgb || !clamped) return SB_ERR_BAD_ARG;
    sh_forward_kernel<<<(P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(P, sh_degree, max_coeffs, means3D,
                                                                                     campos, shs, rgb, clamped);
    SB_LAUNCH_CHECK("sh_forward_kernel");
    return SB_OK;
}

SB_API int sb_sh_backward(int P, int sh_degree, int max_coeffs, const float* means3D, const float* campos,
                          const float* shs, const uint8_t* clamped, const float* dL_drgb, float* dL_dshs,
                          float* dL_dmeans3D_accumulate, void* stream) {
    if (P < 0 || sh_degree < 0 || sh_degree > 3 || max_coeffs < (sh_degree + 1) * (sh_degree + 1)) return SB_ERR_BAD_ARG;
    if (P == 0) return SB_OK;
    if (!means3D || !campos || !shs || !clamped || !dL_drgb || !dL_dshs || !dL_dmeans3D_accumulate) return SB_ERR_BAD_ARG;
    sh_backward_kernel<<<(P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        P, sh_degree, max_coeffs, means3D, campos, shs, clamped, dL_drgb, dL_dshs, dL_dmeans3D_accumulate);
    SB_LAUNCH_CHECK("sh_backward_kernel");
    return SB_OK;
}

}  // extern "C"

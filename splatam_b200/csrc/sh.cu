// sh.cu -- the spherical-harmonics colour branch of the operator API (shs / sh_degree > 0;
// computeColorFromSH forward X/cuda_rasterizer/forward.cu:20-71 and backward X/cuda_rasterizer/backward.cu:20-139).
// SplaTAM never uses it (it always passes colors_precomp, R/utils/slam_helpers.py:131-138); it is here so that
// the other callers of the reference operator (R/scripts/gaussian_splatting.py-style SH models) drop in too.
// Two standalone kernels: SH -> RGB (+ clamp flags) before the render, and RGB-gradient -> SH-gradient (+ the
// view-direction term of dL/dmean) after the per-Gaussian backward.  The basis is a monomial TABLE walked by one
// generic evaluator (values and mechanically differentiated gradients), not a hand-expanded polynomial.
#include "common.cuh"

namespace sb {

namespace {

// Real spherical-harmonics basis up to degree 3 as DATA: basis function k is  K[k] * sum_t m_t x^a_t y^b_t z^c_t
// over its monomials (3DGS sign convention).  One generic evaluator walks the table for the values and -- by
// differentiating each monomial mechanically -- for the gradient with respect to the direction, so forward and
// backward share one description of the basis and no hand-expanded derivative exists anywhere.
struct Mono { int k; float m; int a, b, c; };
constexpr int kNumMono = 28;
__host__ __device__ constexpr int mono_end(int deg) { return deg == 0 ? 1 : deg == 1 ? 4 : deg == 2 ? 12 : 28; }   // monomials up to degree
__host__ __device__ constexpr int num_basis(int deg) { return (deg + 1) * (deg + 1); }
__host__ __device__ constexpr Mono mono(int t) {
    constexpr Mono tab[kNumMono] = {
        {0, 1.f, 0, 0, 0},
        {1, 1.f, 0, 1, 0}, {2, 1.f, 0, 0, 1}, {3, 1.f, 1, 0, 0},
        {4, 1.f, 1, 1, 0}, {5, 1.f, 0, 1, 1}, {6, 2.f, 0, 0, 2}, {6, -1.f, 2, 0, 0}, {6, -1.f, 0, 2, 0}, {7, 1.f, 1, 0, 1},
        {8, 1.f, 2, 0, 0}, {8, -1.f, 0, 2, 0},
        {9, 3.f, 2, 1, 0}, {9, -1.f, 0, 3, 0}, {10, 1.f, 1, 1, 1}, {11, 4.f, 0, 1, 2}, {11, -1.f, 2, 1, 0}, {11, -1.f, 0, 3, 0},
        {12, 2.f, 0, 0, 3}, {12, -3.f, 2, 0, 1}, {12, -3.f, 0, 2, 1}, {13, 4.f, 1, 0, 2}, {13, -1.f, 3, 0, 0}, {13, -1.f, 1, 2, 0},
        {14, 1.f, 2, 0, 1}, {14, -1.f, 0, 2, 1}, {15, 1.f, 3, 0, 0}, {15, -3.f, 1, 2, 0}};
    return tab[t];
}
__host__ __device__ constexpr float norm_of(int k) {
    constexpr float tab[16] = {
        0.28209479177387814f,
        -0.4886025119029199f, 0.4886025119029199f, -0.4886025119029199f,
        1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f, 0.5462742152960396f,
        -0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f, -0.4570457994644658f,
        1.445305721320277f, -0.5900435899266435f};
    return tab[k];
}

// One monomial of the table (compile-time index T): adds its value to B[k] and, if GRAD, its partial derivatives
// (a x^(a-1) y^b z^c, ...) to dB[k]/d(x,y,z); then recurses to T + 1.  Every index is a constant expression, so the
// 4 x 16 accumulators live in registers.
template <bool GRAD, int T>
struct MonoStep {
    static __device__ __forceinline__ void run(int n, const float (&px)[4], const float (&py)[4], const float (&pz)[4],
                                               float (&B)[16], float (&Bx)[16], float (&By)[16], float (&Bz)[16]) {
        if (T < n) {
            constexpr Mono mo = mono(T);
            constexpr float w = norm_of(mo.k) * mo.m;
            B[mo.k] = fmaf(w, px[mo.a] * py[mo.b] * pz[mo.c], B[mo.k]);
            if (GRAD) {
                if (mo.a > 0) Bx[mo.k] = fmaf(w * (float)mo.a, px[mo.a > 0 ? mo.a - 1 : 0] * py[mo.b] * pz[mo.c], Bx[mo.k]);
                if (mo.b > 0) By[mo.k] = fmaf(w * (float)mo.b, px[mo.a] * py[mo.b > 0 ? mo.b - 1 : 0] * pz[mo.c], By[mo.k]);
                if (mo.c > 0) Bz[mo.k] = fmaf(w * (float)mo.c, px[mo.a] * py[mo.b] * pz[mo.c > 0 ? mo.c - 1 : 0], Bz[mo.k]);
            }
        }
        MonoStep<GRAD, T + 1>::run(n, px, py, pz, B, Bx, By, Bz);
    }
};
template <bool GRAD>
struct MonoStep<GRAD, kNumMono> {
    static __device__ __forceinline__ void run(int, const float (&)[4], const float (&)[4], const float (&)[4], float (&)[16],
                                               float (&)[16], float (&)[16], float (&)[16]) {}
};

// Basis values B[k] and, if GRAD, dB[k]/d(x,y,z) at the unit direction (x, y, z).
template <bool GRAD>
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float (&B)[16], float (&Bx)[16], float (&By)[16],
                                         float (&Bz)[16]) {
    const float px[4] = {1.f, x, x * x, x * x * x}, py[4] = {1.f, y, y * y, y * y * y}, pz[4] = {1.f, z, z * z, z * z * z};
#pragma unroll
    for (int k = 0; k < 16; ++k) { B[k] = 0.f; if (GRAD) { Bx[k] = 0.f; By[k] = 0.f; Bz[k] = 0.f; } }
    MonoStep<GRAD, 0>::run(mono_end(deg), px, py, pz, B, Bx, By, Bz);
}

// Unit view direction mean - campos; a Gaussian sitting exactly on the camera centre has no direction: use zero.
__device__ __forceinline__ void view_dir(const float* __restrict__ means, const float* __restrict__ campos, int idx,
                                         float& vx, float& vy, float& vz, float& x, float& y, float& z) {
    vx = means[3 * idx] - campos[0]; vy = means[3 * idx + 1] - campos[1]; vz = means[3 * idx + 2] - campos[2];
    const float len2 = vx * vx + vy * vy + vz * vz;
    const float inv = len2 > 0.f ? 1.0f / sqrtf(len2) : 0.f;
    x = vx * inv; y = vy * inv; z = vz * inv;
}

// colour = max(0, 0.5 + sum_k B_k(dir) sh_k)   (computeColorFromSH, X/cuda_rasterizer/forward.cu:20-71)
__global__ void __launch_bounds__(256)
sh_forward_kernel(int P, int deg, int M, const float* __restrict__ means, const float* __restrict__ campos,
                  const float* __restrict__ shs, float* __restrict__ rgb, uint8_t* __restrict__ clamped) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float vx, vy, vz, x, y, z;
    view_dir(means, campos, idx, vx, vy, vz, x, y, z);
    float B[16], u0[16], u1[16], u2[16];
    sh_basis<false>(deg, x, y, z, B, u0, u1, u2);
    const float* sh = shs + (size_t)idx * M * 3;
    float c[3] = {0.5f, 0.5f, 0.5f};
    const int nb = num_basis(deg);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k < nb) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) c[ch] = fmaf(B[k], sh[3 * k + ch], c[ch]);
        }
    // colours are clamped to >= 0; remember where: the gradient is zero there (forward.cu:64-70)
    clamped[idx] = (uint8_t)((c[0] < 0 ? 1 : 0) | (c[1] < 0 ? 2 : 0) | (c[2] < 0 ? 4 : 0));
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) rgb[3 * idx + ch] = fmaxf(c[ch], 0.f);
}

// dL/dsh_k = B_k dL/drgb;  dL/ddir = sum_k dB_k/ddir (sh_k . dL/drgb), then through the normalisation
// dir = v/|v|:  dL/dv = (dL/ddir - dir (dir . dL/ddir)) / |v|   (X/cuda_rasterizer/backward.cu:20-139).
// Gaussians that were culled (radii <= 0, when `radii` is given) get zeros, as the reference leaves them.
__global__ void __launch_bounds__(256)
sh_backward_kernel(int P, int deg, int M, const float* __restrict__ means, const float* __restrict__ campos,
                   const float* __restrict__ shs, const uint8_t* __restrict__ clamped,
                   const float* __restrict__ dL_drgb, const int32_t* __restrict__ radii, float* __restrict__ dL_dshs,
                   float* __restrict__ dL_dmeans) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float* out = dL_dshs + (size_t)idx * M * 3;
    if (radii != nullptr && radii[idx] <= 0) {
        for (int k = 0; k < 3 * M; ++k) out[k] = 0.f;
        return;
    }
    float vx, vy, vz, x, y, z;
    view_dir(means, campos, idx, vx, vy, vz, x, y, z);
    float B[16], Bx[16], By[16], Bz[16];
    sh_basis<true>(deg, x, y, z, B, Bx, By, Bz);
    const uint8_t cl = clamped[idx];
    float g[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) g[ch] = ((cl >> ch) & 1) ? 0.f : dL_drgb[3 * idx + ch];
    const float* sh = shs + (size_t)idx * M * 3;
    const int nb = num_basis(deg);
    float dx = 0.f, dy = 0.f, dz = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k < nb) {
            const float s = sh[3 * k] * g[0] + sh[3 * k + 1] * g[1] + sh[3 * k + 2] * g[2];
            dx = fmaf(Bx[k], s, dx); dy = fmaf(By[k], s, dy); dz = fmaf(Bz[k], s, dz);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) out[3 * k + ch] = B[k] * g[ch];
        }
    for (int k = 3 * nb; k < 3 * M; ++k) out[k] = 0.f;
    const float len2 = vx * vx + vy * vy + vz * vz;
    if (len2 > 0.f) {
        const float inv = 1.0f / sqrtf(len2), along = x * dx + y * dy + z * dz;
        dL_dmeans[3 * idx] += (dx - x * along) * inv;
        dL_dmeans[3 * idx + 1] += (dy - y * along) * inv;
        dL_dmeans[3 * idx + 2] += (dz - z * along) * inv;
    }
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
                          const float* shs, const uint8_t* clamped, const float* dL_drgb, const int32_t* radii,
                          float* dL_dshs, float* dL_dmeans3D_accumulate, void* stream) {
    if (P < 0 || sh_degree < 0 || sh_degree > 3 || max_coeffs < (sh_degree + 1) * (sh_degree + 1)) return SB_ERR_BAD_ARG;
    if (P == 0) return SB_OK;
    if (!means3D || !campos || !shs || !clamped || !dL_drgb || !dL_dshs || !dL_dmeans3D_accumulate) return SB_ERR_BAD_ARG;
    sh_backward_kernel<<<(P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        P, sh_degree, max_coeffs, means3D, campos, shs, clamped, dL_drgb, radii, dL_dshs, dL_dmeans3D_accumulate);
    SB_LAUNCH_CHECK("sh_backward_kernel");
    return SB_OK;
}

}  // extern "C"

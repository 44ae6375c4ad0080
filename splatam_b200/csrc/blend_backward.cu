// blend_backward.cu -- backward of the alpha-composite (BACKWARD::renderCUDA,
// X/cuda_rasterizer/backward.cu:399-557): walks every tile's sorted list back-to-front from the last
// contributor, rebuilds T by division exactly as the reference does, and produces per-Gaussian
// dL/d{mean2D.xy, conic.xx/xy/yy, opacity, colour rgb}.
//
// Structure mirrors blend_forward.cu (TMA producer warp + 8 autonomous consumer warps, one 8x4
// pixel rectangle each, warp-ballot culling with the same conservative boxes).  The reference issues
// 9 global float atomics per contributing pixel-pair (backward.cu:523-554); here the 9 partial sums of
// a (warp, Gaussian) pair are reduced across the 32 lanes with a value-splitting butterfly (12 shuffles
// for 9 values: each xor step halves the number of values a lane still carries) and then written with
// ONE reduction instruction whose 9 active lanes hit 9 consecutive floats of the Gaussian's 48-B
// accumulator row -- 32x fewer L2 atomic operations.
#include "common.cuh"
#include "pipeline.cuh"

namespace sb {

namespace {

constexpr int kBatch = 128;
constexpr int kStages = 4;
constexpr int kConsumerWarps = 8;
constexpr int kBlendThreads = (kConsumerWarps + 1) * 32;

struct __align__(128) BwdSmem {
    float4 A[kStages][kBatch];
    float4 B[kStages][kBatch];
    float4 C[kStages][kBatch];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint32_t nmax;
};

// Reduces v[0..8] over the warp; on return lane `slot_lane(q)` holds the total of quantity q in the
// returned value: lanes {0,2,4,8,10,16,18,20,24} <-> q {0..8}; every other lane returns garbage/zero.
__device__ __forceinline__ float butterfly9(float v0, float v1, float v2, float v3, float v4, float v5,
                                            float v6, float v7, float v8, int lane) {
    constexpr uint32_t full = 0xffffffffu;
    const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2;
    // xor 16: 9 -> 5
    float k0 = u16 ? v5 : v0, k1 = u16 ? v6 : v1, k2 = u16 ? v7 : v2, k3 = u16 ? v8 : v3, k4 = u16 ? 0.f : v4;
    k0 += __shfl_xor_sync(full, u16 ? v0 : v5, 16);
    k1 += __shfl_xor_sync(full, u16 ? v1 : v6, 16);
    k2 += __shfl_xor_sync(full, u16 ? v2 : v7, 16);
    k3 += __shfl_xor_sync(full, u16 ? v3 : v8, 16);
    k4 += __shfl_xor_sync(full, u16 ? v4 : 0.f, 16);
    // xor 8: 5 -> 3
    float m0 = u8 ? k3 : k0, m1 = u8 ? k4 : k1, m2 = u8 ? 0.f : k2;
    m0 += __shfl_xor_sync(full, u8 ? k0 : k3, 8);
    m1 += __shfl_xor_sync(full, u8 ? k1 : k4, 8);
    m2 += __shfl_xor_sync(full, u8 ? k2 : 0.f, 8);
    // xor 4: 3 -> 2
    float n0 = u4 ? m2 : m0, n1 = u4 ? 0.f : m1;
    n0 += __shfl_xor_sync(full, u4 ? m0 : m2, 4);
    n1 += __shfl_xor_sync(full, u4 ? m1 : 0.f, 4);
    // xor 2: 2 -> 1
    float o = u2 ? n1 : n0;
    o += __shfl_xor_sync(full, u2 ? n0 : n1, 2);
    // xor 1
    o += __shfl_xor_sync(full, o, 1);
    return o;
}

__global__ void __launch_bounds__(kBlendThreads)
blend_backward_kernel(const uint2* __restrict__ ranges, const float4* __restrict__ recA,
                      const float4* __restrict__ recB, const float4* __restrict__ recC,
                      int W, int H, uint32_t grid_x, const float* __restrict__ bg,
                      const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                      const float* __restrict__ dL_dpix, float* __restrict__ accum) {
    __shared__ BwdSmem sm;
    const uint32_t tile = blockIdx.x;
    const uint2 range = ranges[tile];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kConsumerWarps); }
        sm.nmax = 0u;
        mbar_fence_init();
    }
    __syncthreads();

    // pixel state (consumer warps)
    const uint32_t tx = tile % grid_x, ty = tile / grid_x;
    const int x0 = (int)tx * kTile + (warp & 1) * 8, y0 = (int)ty * kTile + (warp >> 1) * 4;
    const int px = x0 + (lane & 7), py = y0 + (lane >> 3);
    const bool inside = (warp < kConsumerWarps) && px < W && py < H;
    const size_t pix = (size_t)py * W + px, hw = (size_t)H * W;
    const float T_final = inside ? final_T[pix] : 0.f;
    const uint32_t nc = inside ? n_contrib[pix] : 0u;
    const uint32_t warp_nc = __reduce_max_sync(0xffffffffu, nc);
    if (lane == 0 && warp_nc > 0u) atomicMax(&sm.nmax, warp_nc);
    __syncthreads();
    const int m = (int)sm.nmax;                 // entries [0, m) of the tile list can matter
    const int nb = (m + kBatch - 1) / kBatch;

    if (warp == kConsumerWarps) {
        if (lane == 0) {
            for (int k = 0; k < nb; ++k) {
                const int s = k % kStages;
                if (k >= kStages) mbar_wait(&sm.empty[s], ((k / kStages) - 1) & 1);
                const int hi = m - k * kBatch, cnt = min(kBatch, hi), lo = hi - cnt;
                const uint32_t bytes = (uint32_t)cnt * 16u;
                const size_t src = (size_t)range.x + (size_t)lo;
                mbar_arrive_expect_tx(&sm.full[s], 3u * bytes);
                tma_load_1d(sm.A[s], recA + src, bytes, &sm.full[s]);
                tma_load_1d(sm.B[s], recB + src, bytes, &sm.full[s]);
                tma_load_1d(sm.C[s], recC + src, bytes, &sm.full[s]);
            }
        }
        return;
    }

    const float pxf = (float)px, pyf = (float)py;
    const float fx0 = (float)x0, fx1 = (float)(x0 + 7), fy0 = (float)y0, fy1 = (float)(y0 + 3);
    float dL0 = 0.f, dL1 = 0.f, dL2 = 0.f;
    if (inside) { dL0 = dL_dpix[pix]; dL1 = dL_dpix[hw + pix]; dL2 = dL_dpix[2 * hw + pix]; }
    const float bg_dot = __ldg(bg) * dL0 + __ldg(bg + 1) * dL1 + __ldg(bg + 2) * dL2;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;  // pixel -> NDC (backward.cu:452-453)
    float T = T_final;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;          // accum_rec
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;
    // which accumulator slot this lane flushes after the butterfly
    int slot = -1;
    switch (lane) { case 0: slot = 0; break; case 2: slot = 1; break; case 4: slot = 2; break;
                    case 8: slot = 3; break; case 10: slot = 4; break; case 16: slot = 5; break;
                    case 18: slot = 6; break; case 20: slot = 7; break; case 24: slot = 8; break; default: break; }

    for (int k = 0; k < nb; ++k) {
        const int s = k % kStages;
        mbar_wait(&sm.full[s], (k / kStages) & 1);
        const int hi = m - k * kBatch, cnt = min(kBatch, hi), lo = hi - cnt;
        if (lo < (int)warp_nc) {
            for (int c = 0; c < cnt; c += 32) {
                const int jl = cnt - 1 - (c + lane);
                bool hit = false;
                if (jl >= 0 && lo + jl < (int)warp_nc) {
                    const float4 a = sm.A[s][jl];
                    hit = (a.x + a.z >= fx0) && (a.x - a.z <= fx1) && (a.y + a.w >= fy0) && (a.y - a.w <= fy1);
                }
                uint32_t mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int j = cnt - 1 - (c + (__ffs(mask) - 1));
                    mask &= mask - 1;
                    const float4 a = sm.A[s][j];
                    const float4 q = sm.B[s][j];
                    // same float sequence as the forward so the skip decisions agree (backward.cu:491-500)
                    const float dx = __fsub_rn(a.x, pxf), dy = __fsub_rn(a.y, pyf);
                    const float sxy = __fmaf_rn(dx, __fmul_rn(dx, q.x), __fmul_rn(dy, __fmul_rn(dy, q.z)));
                    const float power = __fmaf_rn(sxy, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, q.y)));
                    const float G = expf(power);
                    const float alpha = fminf(__fmul_rn(q.w, G), 0.99f);
                    const bool active = ((uint32_t)(lo + j) < nc) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                    if (!__any_sync(0xffffffffu, active)) continue;
                    const float4 col = sm.C[s][j];
                    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f, v8 = 0.f;
                    if (active) {
                        T = T / (1.f - alpha);
                        const float dchannel_dcolor = alpha * T;
                        acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0;
                        acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1;
                        acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2;
                        lc0 = col.x; lc1 = col.y; lc2 = col.z;
                        float dL_dalpha = (col.x - acc0) * dL0 + (col.y - acc1) * dL1 + (col.z - acc2) * dL2;
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                        const float dL_dG = q.w * dL_dalpha;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * q.x - gdy * q.y;
                        const float dG_ddely = -gdy * q.z - gdx * q.y;
                        v0 = dL_dG * dG_ddelx * ddelx_dx;
                        v1 = dL_dG * dG_ddely * ddely_dy;
                        v2 = -0.5f * gdx * dx * dL_dG;
                        v3 = -0.5f * gdx * dy * dL_dG;
                        v4 = -0.5f * gdy * dy * dL_dG;
                        v5 = G * dL_dalpha;
                        v6 = dchannel_dcolor * dL0;
                        v7 = dchannel_dcolor * dL1;
                        v8 = dchannel_dcolor * dL2;
                    }
                    const float total = butterfly9(v0, v1, v2, v3, v4, v5, v6, v7, v8, lane);
                    if (slot >= 0)
                        atomicAdd(accum + (size_t)__float_as_uint(col.w) * kAccumStride + slot, total);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
    }
}

}  // namespace

int launch_blend_backward(const sb_settings& s, int R, const BinningWs& b, const ImageWs& img,
                          const float* dL_dout_color, float* accum, cudaStream_t st) {
    if (R <= 0) return SB_OK;
    const int W = s.image_width, H = s.image_height;
    const uint32_t gx = (W + kTile - 1) / kTile, gy = (H + kTile - 1) / kTile;
    ScopedStage _p(kStBlendBwd, st);
    blend_backward_kernel<<<gx * gy, kBlendThreads, 0, st>>>(img.ranges, b.recA, b.recB, b.recC, W, H, gx,
                                                             s.bg, img.final_T, img.n_contrib,
                                                             dL_dout_color, accum);
    SB_LAUNCH_CHECK("blend_backward_kernel");
    return SB_OK;
}

}  // namespace sb

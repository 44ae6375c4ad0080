// blend_backward.cu -- backward of the alpha-composite (BACKWARD::renderCUDA,
// X/cuda_rasterizer/backward.cu:399-557): walks every tile's sorted list back-to-front from the last
// contributor, rebuilds T as the reference does (T /= 1-alpha), and produces per-Gaussian
// dL/d{mean2D.xy, conic.xx/xy/yy, opacity, colour rgb}.
//
// Structure mirrors blend_forward.cu (TMA producer warp + 8 autonomous consumer warps, one 8x4 pixel
// rectangle each, warp-ballot culling with the same conservative boxes).  The reference issues 9 global
// float atomics per contributing pixel pair (backward.cu:523-554).  Here the reduction over pixels is
// done in two phases per warp:
//   phase 1 (lane = pixel, sequential along the list): per surviving Gaussian each lane computes only
//     the two scalars that depend on the running per-pixel state, w = G*dL/dG and ca = alpha*T, and
//     parks them in a 16-slot shared-memory queue [slot][pixel];
//   phase 2 (lane = Gaussian, when 16 survivors are queued): two lanes per queued Gaussian sweep 16
//     pixels each and accumulate the nine sums  S{w, w dx, w dy, w dx^2, w dx dy, w dy^2}, S ca*dL/dpix[rgb]
//     in registers at full lane utilisation, combine with one xor-16 shuffle per sum, turn them into
//     the reference's nine gradients and flush each with ONE reduction per quantity per (warp, Gaussian).
// Phase 1 runs at the (low) lane utilisation the pixel footprint dictates but is short; the wide part of
// the arithmetic runs in phase 2 with all 32 lanes busy.
#include "common.cuh"
#include "pipeline.cuh"

namespace sb {

namespace {

constexpr int kBatch = 128;
constexpr int kStages = 4;
constexpr int kConsumerWarps = 8;
constexpr int kBlendThreads = (kConsumerWarps + 1) * 32;
constexpr int kQStride = 33;      // row stride of the queue in floats: conflict-free for both phases

// kQueue = queued survivors per warp before phase 2 runs (16: two lanes per Gaussian, 8: four lanes)
// NCH = 3 (reference operator) or 6 (fused two-colour-set render; the first set's share of dL/dmean2D is
// tracked separately because SplaTAM reads the means2D gradient of the RGB render only, splatam.py:250)
template <int NCH, int kQueue>
struct __align__(128) BwdSmem {
    float4 A[kStages][kBatch];
    float4 B[kStages][kBatch];
    float4 C[kStages][kBatch];
    float4 D[NCH == 6 ? kStages : 1][NCH == 6 ? kBatch : 1];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint32_t nmax;
    uint32_t pad[15];
    float4 meta[kConsumerWarps][kQueue][2];       // {gx-x0, gy-y0, conic.x, conic.y}, {conic.z, opacity, bits(id), -}
    float qw[kConsumerWarps][kQueue][kQStride];   // w  = G * dL/dG      per (slot, pixel)
    float qc[kConsumerWarps][kQueue][kQStride];   // ca = alpha * T      per (slot, pixel)
    float qr[NCH == 6 ? kConsumerWarps : 1][NCH == 6 ? kQueue : 1][kQStride];   // w of the first colour set only
    float dL[kConsumerWarps][NCH][32];            // dL/dpixel of the warp's 32 pixels
};

// 1/x for x in [0.01, 1]: MUFU.RCP + one Newton step (error < 1 ulp; the reference divides, IEEE).
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

// Phase 2: lanes (s, part) = (lane % kQueue, lane / kQueue) sweep pixels [kQueue*part, kQueue*(part+1))
// of queue slot s (kQueue pixels per lane, 32/kQueue lanes per queued Gaussian).
template <int NCH, int kQueue>
__device__ __forceinline__ void flush_queue(BwdSmem<NCH, kQueue>& sm, int warp, int lane, int count, float ddelx_dx,
                                            float ddely_dy, float* __restrict__ accum) {
    constexpr int kStride = NCH == 6 ? kAccumStride2 : kAccumStride;
    __syncwarp();
    const int s = lane % kQueue, part = lane / kQueue;
    const float4 m0 = sm.meta[warp][s][0], m1 = sm.meta[warp][s][1];
    const float gxr = m0.x, gyr = m0.y - (float)((kQueue / 8) * part);
    const float* qw = &sm.qw[warp][s][kQueue * part];
    const float* qc = &sm.qc[warp][s][kQueue * part];
    const float* d0 = &sm.dL[warp][0][kQueue * part];
    const float* d1 = &sm.dL[warp][1][kQueue * part];
    const float* d2 = &sm.dL[warp][2][kQueue * part];
    float Sw = 0.f, Swx = 0.f, Swy = 0.f, Sxx = 0.f, Sxy = 0.f, Syy = 0.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    float E0 = 0.f, E1 = 0.f, E2 = 0.f, Rx = 0.f, Ry = 0.f;   // NCH == 6 only
#pragma unroll
    for (int it = 0; it < kQueue; ++it) {
        const float w = qw[it], ca = qc[it];
        const float dx = gxr - (float)(it & 7), dy = gyr - (float)(it >> 3);
        const float wdx = w * dx, wdy = w * dy;
        Sw += w; Swx += wdx; Swy += wdy;
        Sxx = fmaf(wdx, dx, Sxx); Sxy = fmaf(wdx, dy, Sxy); Syy = fmaf(wdy, dy, Syy);
        C0 = fmaf(ca, d0[it], C0); C1 = fmaf(ca, d1[it], C1); C2 = fmaf(ca, d2[it], C2);
        if (NCH == 6) {
            const float wr = sm.qr[warp][s][kQueue * part + it];
            Rx = fmaf(wr, dx, Rx); Ry = fmaf(wr, dy, Ry);
            E0 = fmaf(ca, sm.dL[warp][NCH - 3][kQueue * part + it], E0);
            E1 = fmaf(ca, sm.dL[warp][NCH - 2][kQueue * part + it], E1);
            E2 = fmaf(ca, sm.dL[warp][NCH - 1][kQueue * part + it], E2);
        }
    }
    constexpr uint32_t full = 0xffffffffu;
#pragma unroll
    for (int off = kQueue; off < 32; off <<= 1) {
        Sw += __shfl_xor_sync(full, Sw, off);   Swx += __shfl_xor_sync(full, Swx, off);
        Swy += __shfl_xor_sync(full, Swy, off); Sxx += __shfl_xor_sync(full, Sxx, off);
        Sxy += __shfl_xor_sync(full, Sxy, off); Syy += __shfl_xor_sync(full, Syy, off);
        C0 += __shfl_xor_sync(full, C0, off);   C1 += __shfl_xor_sync(full, C1, off);
        C2 += __shfl_xor_sync(full, C2, off);
        if (NCH == 6) {
            E0 += __shfl_xor_sync(full, E0, off); E1 += __shfl_xor_sync(full, E1, off);
            E2 += __shfl_xor_sync(full, E2, off); Rx += __shfl_xor_sync(full, Rx, off);
            Ry += __shfl_xor_sync(full, Ry, off);
        }
    }
    if (lane < count) {
        // dG/ddelx = -G (dx a + dy b), dG/ddely = -G (dy c + dx b)   (backward.cu:539-546)
        const float a = m0.z, b = m0.w, c = m1.x, op = m1.y;
        float* row = accum + (size_t)__float_as_uint(m1.z) * kStride;
        atomicAdd(row + 0, -(a * Swx + b * Swy) * ddelx_dx);
        atomicAdd(row + 1, -(c * Swy + b * Swx) * ddely_dy);
        atomicAdd(row + 2, -0.5f * Sxx);
        atomicAdd(row + 3, -0.5f * Sxy);
        atomicAdd(row + 4, -0.5f * Syy);
        atomicAdd(row + 5, Sw / op);           // G dL/dalpha = w / opacity  (dL/dG = opacity dL/dalpha)
        atomicAdd(row + 6, C0);
        atomicAdd(row + 7, C1);
        atomicAdd(row + 8, C2);
        if (NCH == 6) {
            atomicAdd(row + 9, E0); atomicAdd(row + 10, E1); atomicAdd(row + 11, E2);
            atomicAdd(row + 12, -(a * Rx + b * Ry) * ddelx_dx);
            atomicAdd(row + 13, -(c * Ry + b * Rx) * ddely_dy);
        }
    }
    __syncwarp();
}

template <int NCH, int kQueue, int kMinBlocks>
__global__ void __launch_bounds__(kBlendThreads, kMinBlocks)
blend_backward_kernel(const uint2* __restrict__ ranges, const float4* __restrict__ recA,
                      const float4* __restrict__ recB, const float4* __restrict__ recC,
                      const float4* __restrict__ recD,
                      int W, int H, uint32_t grid_x, const float* __restrict__ bg,
                      const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                      const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix2,
                      float* __restrict__ accum) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BwdSmem<NCH, kQueue>& sm = *reinterpret_cast<BwdSmem<NCH, kQueue>*>(smem_raw);
    const uint32_t tile = blockIdx.x;
    const uint2 range = ranges[tile];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kConsumerWarps); }
        sm.nmax = 0u;
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t tx = tile % grid_x, ty = tile / grid_x;
    const int x0 = (int)tx * kTile + (warp & 1) * 8, y0 = (int)ty * kTile + (warp >> 1) * 4;
    const int px = x0 + (lane & 7), py = y0 + (lane >> 3);
    const bool inside = (warp < kConsumerWarps) && px < W && py < H;
    const size_t pix = (size_t)py * W + px, hw = (size_t)H * W;
    const float T_final = inside ? final_T[pix] : 0.f;
    const uint32_t nc = inside ? n_contrib[pix] : 0u;
    const uint32_t warp_nc = __reduce_max_sync(0xffffffffu, nc);
    if (lane == 0 && warp_nc > 0u) atomicMax(&sm.nmax, warp_nc);
    float dL0 = 0.f, dL1 = 0.f, dL2 = 0.f;
    float dL3 = 0.f, dL4 = 0.f, dL5 = 0.f;
    if (inside) { dL0 = dL_dpix[pix]; dL1 = dL_dpix[hw + pix]; dL2 = dL_dpix[2 * hw + pix]; }
    if (NCH == 6 && inside) { dL3 = dL_dpix2[pix]; dL4 = dL_dpix2[hw + pix]; dL5 = dL_dpix2[2 * hw + pix]; }
    if (warp < kConsumerWarps) {
        sm.dL[warp][0][lane] = dL0; sm.dL[warp][1][lane] = dL1; sm.dL[warp][2][lane] = dL2;
        if (NCH == 6) { sm.dL[warp][NCH - 3][lane] = dL3; sm.dL[warp][NCH - 2][lane] = dL4; sm.dL[warp][NCH - 1][lane] = dL5; }
    }
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
                mbar_arrive_expect_tx(&sm.full[s], (NCH == 6 ? 4u : 3u) * bytes);
                tma_load_1d(sm.A[s], recA + src, bytes, &sm.full[s]);
                tma_load_1d(sm.B[s], recB + src, bytes, &sm.full[s]);
                tma_load_1d(sm.C[s], recC + src, bytes, &sm.full[s]);
                if (NCH == 6) tma_load_1d(sm.D[s], recD + src, bytes, &sm.full[s]);
            }
        }
        return;
    }

    const float pxf = (float)px, pyf = (float)py;
    const float fx0 = (float)x0, fx1 = (float)(x0 + 7), fy0 = (float)y0, fy1 = (float)(y0 + 3);
    const float bg_dot = __ldg(bg) * dL0 + __ldg(bg + 1) * dL1 + __ldg(bg + 2) * dL2;
    const float bg_dot2 = NCH == 6 ? __ldg(bg) * dL3 + __ldg(bg + 1) * dL4 + __ldg(bg + 2) * dL5 : 0.f;
    float acc3 = 0.f, acc4 = 0.f, acc5 = 0.f, lc3 = 0.f, lc4 = 0.f, lc5 = 0.f;   // second colour set
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;  // pixel -> NDC (backward.cu:452-453)
    float T = T_final;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;          // accum_rec
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;
    int qn = 0;                                        // queued survivors (warp-uniform)
    // this lane's queue column and the warp's meta rows, hoisted so the inner loop only adds an offset
    float* const qw_lane = &sm.qw[warp][0][lane];
    float* const qc_lane = &sm.qc[warp][0][lane];
    float* const qr_lane = &sm.qr[NCH == 6 ? warp : 0][0][lane];
    float4* const meta_row = &sm.meta[warp][0][0];
    uint32_t qoff = 0;                                 // qn * kQStride

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
                    float w = 0.f, ca = 0.f, wr = 0.f;
                    if (active) {
                        const float ir = fast_rcp(1.f - alpha);
                        T = T * ir;
                        ca = alpha * T;
                        const float om = 1.f - last_alpha;
                        acc0 = fmaf(last_alpha, lc0, om * acc0);
                        acc1 = fmaf(last_alpha, lc1, om * acc1);
                        acc2 = fmaf(last_alpha, lc2, om * acc2);
                        lc0 = col.x; lc1 = col.y; lc2 = col.z;
                        float dL_dalpha = (col.x - acc0) * dL0 + (col.y - acc1) * dL1 + (col.z - acc2) * dL2;
                        dL_dalpha = fmaf(dL_dalpha, T, -(T_final * ir) * bg_dot);
                        w = G * (q.w * dL_dalpha);
                        if (NCH == 6) {
                            const float4 ex = sm.D[s][j];
                            acc3 = fmaf(last_alpha, lc3, om * acc3);
                            acc4 = fmaf(last_alpha, lc4, om * acc4);
                            acc5 = fmaf(last_alpha, lc5, om * acc5);
                            lc3 = ex.x; lc4 = ex.y; lc5 = ex.z;
                            float d2 = (ex.x - acc3) * dL3 + (ex.y - acc4) * dL4 + (ex.z - acc5) * dL5;
                            d2 = fmaf(d2, T, -(T_final * ir) * bg_dot2);
                            wr = w;
                            w = fmaf(G, q.w * d2, w);
                        }
                        last_alpha = alpha;
                    }
                    qw_lane[qoff] = w;
                    qc_lane[qoff] = ca;
                    if (NCH == 6) qr_lane[qoff] = wr;
                    if (lane == 0) {
                        meta_row[2 * qn] = make_float4(a.x - fx0, a.y - fy0, q.x, q.y);
                        meta_row[2 * qn + 1] = make_float4(q.z, q.w, col.w, 0.f);
                    }
                    qoff += kQStride;
                    if (++qn == kQueue) {
                        flush_queue<NCH, kQueue>(sm, warp, lane, kQueue, ddelx_dx, ddely_dy, accum);
                        qn = 0; qoff = 0;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
    }
    // rows >= qn hold stale data; flush_queue only writes rows < qn
    if (qn > 0) flush_queue<NCH, kQueue>(sm, warp, lane, qn, ddelx_dx, ddely_dy, accum);
}

}  // namespace

int launch_blend_backward(const sb_settings& s, int R, const BinningWs& b, const ImageWs& img,
                          const float* dL_dout_color, const float* dL_dout_color2, float* accum,
                          cudaStream_t st) {
    if (R <= 0) return SB_OK;
    const int W = s.image_width, H = s.image_height;
    const uint32_t gx = (W + kTile - 1) / kTile, gy = (H + kTile - 1) / kTile;
#define SB_LAUNCH_BWD(NCH, Q, MB)                                                                               \
    do {                                                                                                        \
        SB_CUDA_CHECK(cudaFuncSetAttribute(blend_backward_kernel<NCH, Q, MB>,                                   \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,                         \
                                           (int)sizeof(BwdSmem<NCH, Q>)));                                      \
        ScopedStage _p(kStBlendBwd, st);                                                                        \
        blend_backward_kernel<NCH, Q, MB><<<gx * gy, kBlendThreads, sizeof(BwdSmem<NCH, Q>), st>>>(             \
            img.ranges, b.recA, b.recB, b.recC, b.recD, W, H, gx, s.bg, img.final_T, img.n_contrib, dL_dout_color, \
            dL_dout_color2, accum);                                                                             \
    } while (0)
    if (dL_dout_color2 != nullptr) SB_LAUNCH_BWD(6, 8, 3);
    else SB_LAUNCH_BWD(3, 16, 3);
#undef SB_LAUNCH_BWD
    SB_LAUNCH_CHECK("blend_backward_kernel");
    return SB_OK;
}

}  // namespace sb

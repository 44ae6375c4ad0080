// binning.cu -- stage 2a of the forward: per-tile duplication, tile sort, tile ranges and the
// sorted SoA splat records (duplicateWithKeys + SortPairs + identifyTileRanges,
// X/cuda_rasterizer/rasterizer_impl.cu:70-138,284-319).
//
// Two-level ordering instead of the reference's single 64-bit sort: the Gaussians were already
// ordered by (depth bits, index) in launch_depth_order; instances are emitted in that order (tiles
// row-major inside one Gaussian, as the reference emits them) and then STABLY sorted by tile id
// alone.  The result -- instance list ordered by (tile, depth bits, index) -- is bit-identical to the
// reference's sorted point_list / point_list_keys, at 2 radix passes over 8-B pairs instead of 6
// passes over 12-B pairs.  The sort itself is this repo's own (radix_sort.cu).
#include "common.cuh"
#include "radix_sort.cuh"

namespace sb {

namespace {

__global__ void __launch_bounds__(256)
emit_instances_kernel(int P, const uint32_t* __restrict__ sorted_idx, const uint32_t* __restrict__ offsets,
                      const uint32_t* __restrict__ tiles_touched, const uint2* __restrict__ rect,
                      uint32_t grid_x, uint32_t cap, uint32_t* __restrict__ tile_out, uint32_t* __restrict__ val_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t g = __ldg(sorted_idx + i);
    const uint32_t n = __ldg(tiles_touched + g);
    if (n == 0u) return;
    uint32_t off = __ldg(offsets + i) - n;  // inclusive scan -> start of this Gaussian's run
    const uint2 r = __ldg(rect + g);
    const uint32_t x0 = r.x & 0xFFFFu, y0 = r.x >> 16, x1 = r.y & 0xFFFFu, y1 = r.y >> 16;
    for (uint32_t y = y0; y < y1; ++y)
        for (uint32_t x = x0; x < x1; ++x) {
            if (off < cap) {              // cap = capacity of the instance arrays (sync-free mode), else 2^32-1
                tile_out[off] = y * grid_x + x;
                val_out[off] = g;
            }
            ++off;
        }
}

// One thread per sorted instance: tile range boundaries (identifyTileRanges) fused with the gather
// of the instance's splat record into the three sorted SoA arrays the blend kernels stage by TMA.
__global__ void __launch_bounds__(256)
ranges_and_records_kernel(int R_host, const int32_t* __restrict__ count_dev, const uint32_t* __restrict__ tile_sorted,
                          const uint32_t* __restrict__ point_list,
                          const float4* __restrict__ geomA, const float4* __restrict__ geomB,
                          const float* __restrict__ colors, const float* __restrict__ colors2,
                          uint2* __restrict__ ranges, float4* __restrict__ recA, float4* __restrict__ recB,
                          float4* __restrict__ recC, float4* __restrict__ recD) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int R = count_dev ? min(*count_dev, R_host) : R_host;   // sync-free mode: R_host is the capacity
    if (i >= R) return;
    const uint32_t tile = tile_sorted[i];
    if (i == 0) ranges[tile].x = 0u;
    else {
        const uint32_t prev = tile_sorted[i - 1];
        if (prev != tile) { ranges[prev].y = (uint32_t)i; ranges[tile].x = (uint32_t)i; }
    }
    if (i == R - 1) ranges[tile].y = (uint32_t)R;
    const uint32_t g = __ldg(point_list + i);
    recA[i] = __ldg(geomA + g);
    recB[i] = __ldg(geomB + g);
    const float* c = colors + 3 * (size_t)g;
    recC[i] = make_float4(__ldg(c), __ldg(c + 1), __ldg(c + 2), __uint_as_float(g));
    if (colors2 != nullptr) {
        const float* e = colors2 + 3 * (size_t)g;
        recD[i] = make_float4(__ldg(e), __ldg(e + 1), __ldg(e + 2), 0.f);
    }
}

}  // namespace

int launch_binning(const sb_settings& s, int P, int R, const float* colors, const float* colors2,
                   const GeometryWs& g, const BinningWs& b, const ImageWs& img, const int32_t* count_dev,
                   cudaStream_t st) {
    const int gx = (s.image_width + kTile - 1) / kTile, gy = (s.image_height + kTile - 1) / kTile;
    const int tiles = gx * gy;
    SB_CUDA_CHECK(cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)tiles, st));  // rasterizer_impl.cu:311
    if (R <= 0) return SB_OK;
    const int bits = (int)higher_msb((uint32_t)tiles);  // same bit count as the reference sort uses
    { ScopedStage _p(kStEmit, st);
      emit_instances_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, g.sorted_idx, g.offsets, g.tiles_touched, g.rect,
                                                             (uint32_t)gx, count_dev ? (uint32_t)R : 0xFFFFFFFFu,
                                                             b.tile_unsorted, b.val_unsorted); }
    SB_LAUNCH_CHECK("emit_instances_kernel");
    { ScopedStage _p(kStTileSort, st);
      // sync-free mode: the pair count is read on the device (count_dev), so only the live instances are sorted --
      // no sentinel padding of the unused capacity, no extra key bits
      const int rc = radix_sort_pairs(b.tile_unsorted, b.val_unsorted, b.tile_sorted, b.point_list, R, count_dev, 0, bits,
                                      b.sort_temp, b.sort_temp_bytes, st);
      if (rc != SB_OK) return rc; }
    ScopedStage _p(kStRecords, st);
    ranges_and_records_kernel<<<(R + 255) / 256, 256, 0, st>>>(R, count_dev, b.tile_sorted, b.point_list, g.geomA,
                                                               g.geomB, colors, colors2, img.ranges,
                                                               b.recA, b.recB, b.recC, b.recD);
    SB_LAUNCH_CHECK("ranges_and_records_kernel");
    return SB_OK;
}

}  // namespace sb

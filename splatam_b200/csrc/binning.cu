// binning.cu -- stage 2a of the forward: per-tile duplication, tile sort, tile ranges and the
// sorted SoA splat records (duplicateWithKeys + SortPairs + identifyTileRanges,
// X/cuda_rasterizer/rasterizer_impl.cu:70-138,284-319).
//
// Two-level ordering instead of the reference's single 64-bit sort: the Gaussians were already
// ordered by (depth bits, index) in launch_depth_order; instances are emitted in that order (tiles
// row-major inside one Gaussian, as the reference emits them) and then STABLY sorted by tile id
// alone.  The result -- instance list ordered by (tile, depth bits, index) -- is bit-identical to the
// reference's sorted point_list / point_list_keys, at 2 radix passes over 6-B pairs instead of 6
// passes over 12-B pairs.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace sb {

namespace {

template <typename KeyT>
__global__ void __launch_bounds__(256)
emit_instances_kernel(int P, const uint32_t* __restrict__ sorted_idx, const uint32_t* __restrict__ offsets,
                      const uint32_t* __restrict__ tiles_touched, const uint2* __restrict__ rect,
                      uint32_t grid_x, uint32_t cap, KeyT* __restrict__ tile_out, uint32_t* __restrict__ val_out) {
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
                tile_out[off] = (KeyT)(y * grid_x + x);
                val_out[off] = g;
            }
            ++off;
        }
}

// Sync-free mode: slots [num_rendered, capacity) get the all-ones sentinel tile id so they sort to the end.
template <typename KeyT>
__global__ void __launch_bounds__(256)
pad_tiles_kernel(int cap, const int32_t* __restrict__ count_dev, KeyT* __restrict__ tile_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap && i >= *count_dev) tile_out[i] = (KeyT)~(KeyT)0;
}

// One thread per sorted instance: tile range boundaries (identifyTileRanges) fused with the gather
// of the instance's splat record into the three sorted SoA arrays the blend kernels stage by TMA.
template <typename KeyT>
__global__ void __launch_bounds__(256)
ranges_and_records_kernel(int R_host, const int32_t* __restrict__ count_dev, const KeyT* __restrict__ tile_sorted,
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

template <typename KeyT>
int run_binning(const sb_settings& s, int P, int R, int bits, const float* colors, const float* colors2,
                const GeometryWs& g, const BinningWs& b, const ImageWs& img, const int32_t* count_dev,
                cudaStream_t st) {
    const uint32_t gx = (s.image_width + kTile - 1) / kTile;
    KeyT* tile_unsorted = reinterpret_cast<KeyT*>(b.tile_unsorted);
    KeyT* tile_sorted = reinterpret_cast<KeyT*>(b.tile_sorted);
    { ScopedStage _p(kStEmit, st);
    emit_instances_kernel<KeyT><<<(P + 255) / 256, 256, 0, st>>>(P, g.sorted_idx, g.offsets, g.tiles_touched, g.rect,
                                                                 gx, count_dev ? (uint32_t)R : 0xFFFFFFFFu,
                                                                 tile_unsorted, b.val_unsorted);
    if (count_dev) {
        pad_tiles_kernel<KeyT><<<(R + 255) / 256, 256, 0, st>>>(R, count_dev, tile_unsorted);
        bits = (int)sizeof(KeyT) * 8;      // the sentinel must take part in the sort
    }
    }
    SB_LAUNCH_CHECK("emit_instances_kernel");
    size_t tb = b.cub_temp_bytes;
    { ScopedStage _p(kStTileSort, st);
      SB_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(b.cub_temp, tb, tile_unsorted, tile_sorted, b.val_unsorted,
                                                    b.point_list, R, 0, bits, st)); }
    ScopedStage _p(kStRecords, st);
    ranges_and_records_kernel<KeyT><<<(R + 255) / 256, 256, 0, st>>>(R, count_dev, tile_sorted, b.point_list, g.geomA,
                                                                     g.geomB, colors, colors2, img.ranges,
                                                                     b.recA, b.recB, b.recC, b.recD);
    SB_LAUNCH_CHECK("ranges_and_records_kernel");
    return SB_OK;
}

}  // namespace

size_t binning_cub_temp_bytes(int R, bool keys16) {
    size_t bytes = 0;
    if (keys16)
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint16_t*)nullptr, (uint16_t*)nullptr,
                                        (const uint32_t*)nullptr, (uint32_t*)nullptr, R);
    else
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                        (const uint32_t*)nullptr, (uint32_t*)nullptr, R);
    return bytes;
}

int launch_binning(const sb_settings& s, int P, int R, const float* colors, const float* colors2,
                   const GeometryWs& g, const BinningWs& b, const ImageWs& img, const int32_t* count_dev,
                   cudaStream_t st) {
    const int gx = (s.image_width + kTile - 1) / kTile, gy = (s.image_height + kTile - 1) / kTile;
    const int tiles = gx * gy;
    SB_CUDA_CHECK(cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)tiles, st));  // rasterizer_impl.cu:311
    if (R <= 0) return SB_OK;
    const int bits = (int)higher_msb((uint32_t)tiles);  // same bit count as the reference sort uses
    if (bits <= 16 && (count_dev == nullptr || tiles < 65535))
        return run_binning<uint16_t>(s, P, R, bits, colors, colors2, g, b, img, count_dev, st);
    return run_binning<uint32_t>(s, P, R, bits, colors, colors2, g, b, img, count_dev, st);
}

}  // namespace sb

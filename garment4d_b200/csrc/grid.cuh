// Layout of the per-cloud uniform grid record produced by g4d_grid_build (spatial_grid.cu) and consumed by the
// grid ball query / three_nn (spatial_grid.cu) and the pruned FPS (fps_pruned.cu).
#pragma once
#include "common.cuh"

namespace g4d {

constexpr int GRID_MAX_CELLS = 4096;
constexpr int GRID_HDR = 16;                 // 4-byte words
// per-cloud record: [hdr 16 words][cell_start GRID_MAX_CELLS+1 ints][pad to 16 B][sorted float4 (x,y,z,bits(k)) x n]
struct GridHdr {
    float ox, oy, oz, inv_h;
    int dx, dy, dz, ncells;
    float h;
    float eps;                               // absolute slack >= any rounding error of a cell-face coordinate or of the binning (8 ulp of max |coord|)
    int pad[6];
};
static_assert(sizeof(GridHdr) == GRID_HDR * 4, "GridHdr layout");

__host__ __device__ inline size_t grid_cloud_words(int n) {
    size_t w = GRID_HDR + (GRID_MAX_CELLS + 1);
    w = (w + 3) / 4 * 4;
    return w + (size_t)n * 4;
}
__device__ __forceinline__ const GridHdr* grid_hdr(const float* g) { return reinterpret_cast<const GridHdr*>(g); }
__device__ __forceinline__ const int* grid_cell_start(const float* g) { return reinterpret_cast<const int*>(g) + GRID_HDR; }
__host__ __device__ __forceinline__ const float4* grid_sorted(const float* g) {
    return reinterpret_cast<const float4*>(g + (GRID_HDR + GRID_MAX_CELLS + 1 + 3) / 4 * 4);
}

}  // namespace g4d

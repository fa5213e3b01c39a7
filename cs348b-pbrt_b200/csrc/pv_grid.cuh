// pv_grid.cuh -- cell addressing shared by the map build (pv_build.cu) and the gather (pv_gather.cu).
#pragma once
#include "pv_ctx.h"

__host__ __device__ __forceinline__ uint32_t pv_part1by1(uint32_t x) {
    x &= 0x0000ffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
// Z-order (Morton) index of a row (cy, cz)
__host__ __device__ __forceinline__ uint32_t pv_morton2(uint32_t y, uint32_t z) { return pv_part1by1(y) | (pv_part1by1(z) << 1); }
__host__ __device__ __forceinline__ uint32_t pv_cell_key(int xbits, int cx, int cy, int cz) {
    return (pv_morton2((uint32_t)cy, (uint32_t)cz) << xbits) | (uint32_t)cx;
}
// Cell coordinate of a coordinate value.  floor((p - o) * inv_h) is monotone in p, which is all the
// search-radius guarantee needs (DESIGN.md "Exactness of the grid search").
__device__ __forceinline__ int pv_cell_coord(float p, float o, float inv_h, int n) {
    int c = (int)floorf((p - o) * inv_h);
    return min(max(c, 0), n - 1);
}

// sa_tile_host.h -- host-side entry points of the register-tiled long-pair kernels (sa_tile.cuh), compiled in their own
// translation unit (sa_tile.cu) so that the kernels build in parallel with the rest of the library.
#pragma once
#include "sa_long.cuh"

namespace sa {

constexpr int TILE_WARPS = 4;        // warps (strips) per block: one per SM sub-partition
__host__ __device__ constexpr int tile_nwt(int R, int C) { return R * C / 16; }      // direction words per lane and macro-step

// score scale of a tiled kernel: 4*H (sa_cell.cuh), except local kernels whose lanes own more than four rows -- their arg-max
// key needs log2(R) free bits below the score (sa_tile.cuh).  The one-byte profile then holds scale*S: |scale*S| <= 127.
__host__ __device__ constexpr int tile_scale(int R, bool local) { return !local || R <= 4 ? 4 : R <= 8 ? 8 : 16; }

bool tile_cfg_exists(int R, int C);
size_t tile_smem_bytes(int R, int C, int alpha);
// resident blocks per SM (0 when the configuration does not exist or does not fit)
int tile_occupancy(int R, int C, bool local, bool linked, size_t smem);
// cooperative launch (every strip's warp must be resident: the strip chain spins on its producer)
cudaError_t tile_launch(int R, int C, bool local, bool linked, const LongArgs &A, int grid, size_t smem, cudaStream_t st);

} // namespace sa

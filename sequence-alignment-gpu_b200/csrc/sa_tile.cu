// sa_tile.cu -- instantiations and launch wrappers of tile_fill_kernel (sa_tile.cuh).
// Compiled twice (SA_TILE_PART = 1 / 2, sa_tile.o and sa_tile2.o) so that the 21 kernel instantiations build in parallel:
// part 1 holds the shapes the library picks by itself and the host-side wrappers, part 2 the shapes reached with SA_TILE.
#include "sa_tile.cuh"
#include "sa_tile_host.h"

namespace sa {

#ifndef SA_TILE_PART
#define SA_TILE_PART 1
#endif

// (R rows, C columns) per lane and macro-step
#define SA_TILE_CFG_LIST_1(X) X(4, 4) X(8, 2) X(8, 4)
#ifdef SA_TILE_FEW
#define SA_TILE_CFG_LIST_2(X)
#else
#define SA_TILE_CFG_LIST_2(X) X(4, 8) X(2, 8) X(16, 4) X(8, 8)
#endif

template <int R, int C>
static const void *tile_fn(bool local, bool linked)
{
    return local ? (const void *)tile_fill_kernel<R, C, true, TILE_WARPS, false>
                 : linked ? (const void *)tile_fill_kernel<R, C, false, TILE_WARPS, true> : (const void *)tile_fill_kernel<R, C, false, TILE_WARPS, false>;
}

#if SA_TILE_PART == 2
const void *tile_fn_part2(int R, int C, bool local, bool linked)
{
#define X(r, c) if (R == r && C == c) return tile_fn<r, c>(local, linked);
    SA_TILE_CFG_LIST_2(X)
#undef X
    (void)R; (void)C; (void)local; (void)linked;
    return nullptr;
}
#else
const void *tile_fn_part2(int R, int C, bool local, bool linked);

static const void *tile_fn_of(int R, int C, bool local, bool linked)
{
#define X(r, c) if (R == r && C == c) return tile_fn<r, c>(local, linked);
    SA_TILE_CFG_LIST_1(X)
#undef X
    return tile_fn_part2(R, C, local, linked);
}

bool tile_cfg_exists(int R, int C) { return tile_fn_of(R, C, false, false) != nullptr; }

size_t tile_smem_bytes(int R, int C, int alpha) { return 32 * MAX_ALPHA + (size_t)TILE_WARPS * tile_warp_smem(R, C, alpha); }

int tile_occupancy(int R, int C, bool local, bool linked, size_t smem)
{
    const void *fn = tile_fn_of(R, C, local, linked);
    if (!fn) return 0;
    int nb = 0;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, TILE_WARPS * 32, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return nb;
}

cudaError_t tile_launch(int R, int C, bool local, bool linked, const LongArgs &A, int grid, size_t smem, cudaStream_t st)
{
    const void *fn = tile_fn_of(R, C, local, linked);
    if (!fn) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    void *args[] = {(void *)&A};
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(TILE_WARPS * 32), args, smem, st);
}
#endif

} // namespace sa

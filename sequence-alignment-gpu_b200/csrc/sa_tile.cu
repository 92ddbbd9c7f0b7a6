// sa_tile.cu -- instantiations and launch wrappers of tile_fill_kernel (sa_tile.cuh).
#include "sa_tile.cuh"
#include "sa_tile_host.h"

namespace sa {

// (R rows, C columns) per lane and macro-step
#ifdef SA_TILE_FEW
#define SA_TILE_CFG_LIST(X) X(4, 4) X(8, 2)
#else
#define SA_TILE_CFG_LIST(X) X(4, 4) X(8, 4) X(8, 2) X(4, 8) X(2, 8) X(16, 4) X(8, 8)
#endif

bool tile_cfg_exists(int R, int C)
{
#define X(r, c) if (R == r && C == c) return true;
    SA_TILE_CFG_LIST(X)
#undef X
    return false;
}

size_t tile_smem_bytes(int R, int C, int alpha) { return 32 * MAX_ALPHA + (size_t)TILE_WARPS * tile_warp_smem(R, C, alpha); }

template <int R, int C>
static const void *tile_fn(bool local, bool linked)
{
    return local ? (const void *)tile_fill_kernel<R, C, true, TILE_WARPS, false>
                 : linked ? (const void *)tile_fill_kernel<R, C, false, TILE_WARPS, true> : (const void *)tile_fill_kernel<R, C, false, TILE_WARPS, false>;
}

static const void *tile_fn_of(int R, int C, bool local, bool linked)
{
#define X(r, c) if (R == r && C == c) return tile_fn<r, c>(local, linked);
    SA_TILE_CFG_LIST(X)
#undef X
    return nullptr;
}

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

} // namespace sa

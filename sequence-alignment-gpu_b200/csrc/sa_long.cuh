// sa_long.cuh -- one long pair (BASELINE configs 1-3, and each GPU's column strip of config 5).
//
// The (m+1) x (n+1) matrix is cut into horizontal STRIPS of 32*R rows.  A strip is swept left to
// right by one warp exactly like a batch group with L = 32 (lane l owns R rows, one __shfl_up per
// step).  All strips form ONE systolic chain: strip s consumes the bottom row of strip s-1 a few
// columns behind its producer.  The kernel is persistent and cooperative (every warp resident):
// warp w sweeps strips w, w+W, w+2W, ... so the chain never deadlocks.
//
// Tile-boundary H rows pass through HBM/L2 as 64-bit {4*H, tag} words (tag = strip id + 1): the
// 8-byte store is atomic, so the consumer needs no fence and no separate flag -- it polls the data
// word itself (the cross-strip analogue of the reference's columnState hand-off,
// alignSequenceGPU.cu:14-40, without the per-column atomics).  Rows live in a ring of W+1 buffers.
//
// Directions: same warp-step-major packed 2-bit layout as the batch kernel, one region per strip:
//     word(s, kb, w, lane) at s*stripStride + (kb*NW + w)*32 + lane,   kb = step / CB.
#pragma once
#include "sa_cell.cuh"
#include <type_traits>

namespace sa {

struct LongArgs {
    const uint8_t *text;     uint32_t n;        // this launch's text slice (columns)
    const uint8_t *pattern;  uint32_t m;
    uint32_t *dirs;          uint64_t strip_stride;     // words per strip
    unsigned long long *rowbuf;  uint32_t ring;  uint64_t row_stride;   // ring x row_stride entries
    const int8_t *S4;        // 32x32 bytes, 4*S[p][t]
    int alpha, gap;
    uint32_t n_strips;
    // Left border of this column slice (multi-GPU strips): nullptr => the matrix border.
    // left_col[i] = 4*H(i, col0) for DP rows i = 0..m  (col0 = first DP column of the slice minus 1)
    const int *left_col;
    int *right_col;          // optional output: 4*H(i, last column) for i = 0..m
    uint32_t col0;           // global DP column index of the slice's column 0 (for NW borders / arg-max)
    // Row chunks of a slice (the multi-GPU pipeline hands the border column over chunk by chunk): this launch
    // covers DP rows row_base+1 .. row_base+m; top_row[j] = 4*H(row_base, col0+1+j) replaces the matrix border
    // when row_base > 0, bottom_row[j] receives 4*H(row_base+m, col0+1+j) (m must then be a multiple of 32*R).
    uint32_t row_base;
    const int *top_row;
    int *bottom_row;
    // Slices LINKED inside the launch (the multi-GPU wavefront): the border column is a buffer of 64-bit
    // {4*H(i, col0), tag} words in THIS GPU's memory that the left neighbour's kernel fills over NVLink with plain
    // 8-byte stores (right_col64 is the peer's buffer, mapped through CUDA IPC).  A strip polls its own rows before
    // it starts and publishes its right-most column when it ends, so GPU k works on strip s while GPU k+1 works on
    // strip s-1 -- the cross-GPU analogue of the strip ring.  xtag identifies the call; abort_flag (device int) is
    // raised by a strip that waited longer than the limit (dead neighbour) and makes every other wait give up.
    const unsigned long long *left_col64;
    unsigned long long *right_col64;
    uint32_t xtag;
    int *abort_flag;
    unsigned long long *dbg;       // optional: per strip {start ns, border-ready ns, end ns} (globaltimer)
    // results
    int32_t *score;          // NW: H(m, n) of this slice
    int *cand_v; uint32_t *cand_i; uint32_t *cand_j;   // SW: per-strip arg-max candidates
    uint32_t tag_base;       // per-call epoch << 21: ring entries of earlier calls never match
    int *gmax;               // SW: running alignment-wide maximum (4*H), zeroed per call
    // Optional (tiled kernel): only the first *n_dev columns are filled (a device value, <= n).  The second pass of the
    // checkpointed traceback needs a chunk only up to the column where the path entered it from below.
    const int *n_dev;
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// the same store under a predicate instead of a branch: a divergent `if (lane == 31)` in every step splits the
// step into basic blocks and keeps the scheduler from interleaving the bookkeeping with the cell chain
__device__ __forceinline__ void st_volatile_u64_if(const bool pred, unsigned long long *p, unsigned long long v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.volatile.global.u64 [%0], %1;\n\t}" ::"l"(p), "l"(v), "r"((unsigned)pred) : "memory");
}

constexpr int PB = 8;   // boundary-row prefetch block (columns)

// LINKED: the variant with the cross-GPU border hand-off and the per-strip timestamps compiled in (kept out of the
// plain kernel: even outside the step loop the extra code costs it ~5 % through register allocation).
// WIDE: two profile planes for score matrices beyond +-31 (sweep_column_wide); A.S4 then holds the low plane followed
// by the high plane (32*MAX_ALPHA bytes each).
template <int R, bool LOCAL, int WARPS, bool LINKED = false, bool WIDE = false>
__global__ void __launch_bounds__(WARPS * 32) long_fill_kernel(const LongArgs A)
{
    static_assert(R % 2 == 0, "R must be even");
    constexpr int CB = cb_for(R);
    constexpr int NW = R * CB / 16;
    constexpr int RPAD = rpad_for(R);
    constexpr int PS = 32 * RPAD;
    constexpr int NPW = (R + 3) / 4;
    constexpr int ROWS = 32 * R;

    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int alpha = A.alpha;
    int8_t *S4s = reinterpret_cast<int8_t *>(smem);
    constexpr uint32_t snapBytes = LOCAL ? ((R + 3) / 4) * 32 * 16 : 0;
    constexpr uint32_t winBytes = 64 + 2 * PB * 4;                     // text window + top-row window
    constexpr int PLANES = WIDE ? 2 : 1;
    unsigned char *profS = smem + PLANES * 32 * MAX_ALPHA + (size_t)warp * (PLANES * alpha * PS + snapBytes + winBytes);
    unsigned char *profHS = profS + alpha * PS;                      // WIDE: the high plane of the profile
    uint4 *snap = reinterpret_cast<uint4 *>(profS + PLANES * alpha * PS);
    unsigned char *textWin = profS + PLANES * alpha * PS + snapBytes;
    int *topWin = reinterpret_cast<int *>(textWin + 64);
    for (int i = threadIdx.x; i < PLANES * 32 * MAX_ALPHA; i += blockDim.x) S4s[i] = A.S4[i];
    __syncthreads();

    const int KL = 2 - SCALE * A.gap, KT = 1 - SCALE * A.gap;
    const uint32_t W = gridDim.x * WARPS;
    const int n = (int)A.n, m = (int)A.m;
    const int nSteps = n + 31;

    for (uint32_t s = blockIdx.x * WARPS + warp; s < A.n_strips; s += W) {
        const int row0 = (int)s * ROWS;                 // pattern index of the strip's first row
        // ---- query profile of this strip ----
        __syncwarp();
        for (int i = lane; i < ROWS; i += 32) {
            const int off = (i / R) * RPAD + (i % R);
            const int gi = row0 + i;
            if (gi < m) {
                const int8_t *srow = S4s + 32 * min((int)A.pattern[gi], alpha - 1);
                for (int a = 0; a < alpha; ++a) profS[a * PS + off] = (unsigned char)srow[a];
                if (WIDE) for (int a = 0; a < alpha; ++a) profHS[a * PS + off] = (unsigned char)srow[32 * MAX_ALPHA + a];
            } else {
                for (int a = 0; a < alpha; ++a) profS[a * PS + off] = WIDE ? (unsigned char)0 : (unsigned char)0x80;
                if (WIDE) for (int a = 0; a < alpha; ++a) profHS[a * PS + off] = (unsigned char)0x80;
            }
        }
        __syncwarp();

        // ---- boundary state (left border of the slice) ----
        // linked slices: wait until the left neighbour's kernel has delivered the word (tag == xtag)
        // The wait is WARP-UNIFORM (all lanes leave the loop together on a vote).  A per-lane loop -- every lane
        // spinning on its own word -- leaves the warp split into lane groups that the later __syncwarp / shuffles
        // synchronise but do not merge again: such a strip then issues every instruction once per group and sweeps
        // 3.4x slower for its whole life (seen as activemask 0xff00ffff after the steady loop), and the chain queues
        // up behind it.
        auto gtime = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
        if (LINKED && A.dbg && lane == 0) A.dbg[3 * s] = gtime();
        auto linked_border = [&](const int gi) -> int {
            const bool need = gi > 0 && gi <= m;
            unsigned long long v = need ? ld_volatile_u64(A.left_col64 + gi) : 0ull;
            const long long t0 = clock64();
            for (unsigned it = 1;; ++it) {
                const bool ok = !need || (uint32_t)(v >> 32) == A.xtag;
                if (__all_sync(0xffffffffu, ok)) break;
                if ((it & 1023u) == 0) {      // rarely: thousands of lanes reading ONE word every microsecond would saturate its L2 slice
                    bool dead = *reinterpret_cast<volatile int *>(A.abort_flag) != 0;
                    if (clock64() - t0 > 20000000000ll) { atomicExch(A.abort_flag, 1); dead = true; }     // ~10 s: the neighbour is gone
                    if (__any_sync(0xffffffffu, dead)) break;
                }
                __nanosleep(1000);
                if (!ok) v = ld_volatile_u64(A.left_col64 + gi);
            }
            return gi == 0 ? (LOCAL ? 0 : -SCALE * A.gap * (int)A.col0) : (int)(uint32_t)v;
        };
        int c[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int gi = row0 + lane * R + r + 1;     // DP row
            if (LINKED && A.left_col64) c[r] = linked_border(gi);
            else if (A.left_col) c[r] = gi <= m ? A.left_col[gi] : 0;
            else c[r] = LOCAL ? 0 : -SCALE * A.gap * (gi + (int)A.row_base);
        }
        int prevTop;                                    // 4*H(i0-1, col0)
        {
            const int gi = row0 + lane * R;
            if (LINKED && A.left_col64) prevTop = linked_border(gi);
            else if (A.left_col) prevTop = gi <= m ? A.left_col[gi] : 0;
            else prevTop = LOCAL ? 0 : -SCALE * A.gap * (gi + (int)A.row_base);
        }
        if (LINKED && A.dbg) { __syncwarp(); if (lane == 0) A.dbg[3 * s + 1] = gtime(); }
        int bottom = 0;
        int bestv = 0, besti = 0, bestj = 0;
        int gmCached = 0;                               // lane-local copy of *A.gmax (a lower bound)
        const bool hasUp = s > 0, hasDown = s + 1 < A.n_strips;
        // lane 31 publishes the strip's bottom row for the strip below -- or, for the last strip of a row chunk, for
        // long_bottom_row_kernel, which copies it out of the ring
        const bool writesRow = lane == 31 && (hasDown || A.bottom_row != nullptr);
        const unsigned long long *rowIn = A.rowbuf + (size_t)((s + A.ring - 1) % A.ring) * A.row_stride;
        unsigned long long *rowOut = A.rowbuf + (size_t)(s % A.ring) * A.row_stride;
        const unsigned long long wantTag = (unsigned long long)(A.tag_base | s);          // producer s-1 writes (s-1)+1
        const unsigned long long myTag = (unsigned long long)(A.tag_base | (s + 1));
        uint32_t *dbase = A.dirs + (size_t)s * A.strip_stride + lane;

        // Everything a step needs besides `up` is fetched ONE STEP AHEAD (text letter -> profile words,
        // lane 0's top value), so the per-step critical path is just SHFL.UP -> R cells.
        //   textWin : 64-byte window of the text (two 32-letter blocks), refilled every 32 steps
        //   topWin  : 2*PB top-row values of the strip above (validated {value, tag} words)
        __syncwarp();
        if (lane < n) textWin[lane] = (unsigned char)min((int)A.text[lane], alpha - 1);
        int tnext = (32 + lane < n) ? min((int)A.text[32 + lane], alpha - 1) : 0;    // block 1, stored at step 31
        unsigned long long nextEnt = 0;
        // Top-row blocks are requested FOUR blocks (32 columns) ahead so that the L2 round trip of the
        // tagged words never sits on the critical path: lanes 8g..8g+7 own the blocks b with b%4 == g.
        auto take_top_block = [&](const int blk) {
            const bool mine = (lane >> 3) == (blk & 3);
            const int col = blk * PB + (lane & 7);
            const bool need = mine && col < n;
            if (hasUp) {
                unsigned long long cur = nextEnt;
                while (true) {
                    const bool ok = !need || (cur >> 32) == wantTag;
                    if (__all_sync(0xffffffffu, ok)) break;
                    if (need && (cur >> 32) != wantTag) { __nanosleep(20); cur = ld_volatile_u64(rowIn + col); }
                }
                if (need) topWin[(blk & 1) * PB + (lane & 7)] = (int)(uint32_t)cur;
                const int ncol = col + 4 * PB;
                if (mine && ncol < n) nextEnt = ld_volatile_u64(rowIn + ncol);
            } else if (need) {
                topWin[(blk & 1) * PB + (lane & 7)] = A.top_row ? A.top_row[col] : LOCAL ? 0 : -SCALE * A.gap * (col + 1 + (int)A.col0);
            }
        };
        // window upkeep that must precede the prefetch for step k1 (= k+1)
        auto upkeep = [&](const int k1) {
            if ((k1 & 31) == 0) {
                textWin[((k1 >> 5) & 1) * 32 + lane] = (unsigned char)tnext;
                const int tcol = k1 + 32 + lane;
                tnext = tcol < n ? min((int)A.text[tcol], alpha - 1) : 0;
                if (LOCAL) gmCached = max(gmCached, *reinterpret_cast<volatile int *>(A.gmax));
            }
            if ((k1 % PB) == 0) {
                if (k1 < n) take_top_block(k1 / PB);
                __syncwarp();
            }
        };
        if (hasUp && lane < n) nextEnt = ld_volatile_u64(rowIn + lane);      // blocks 0..3
        take_top_block(0);
        __syncwarp();
        uint32_t profN[NPW], profHN[NPW];
#pragma unroll
        for (int q = 0; q < NPW; ++q) { profN[q] = 0; profHN[q] = 0; }
        if (lane == 0 && n > 0) {
            load_profile_words<R>(profS + (int)textWin[0] * PS, profN);
            if (WIDE) load_profile_words<R>(profHS + (int)textWin[0] * PS, profHN);
        }
        int topN = topWin[0];
        uint32_t acc[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) acc[w] = 0;

        // One wavefront step.  MODE 1 = steady state (every lane has a column, upkeep done by the
        // caller), MODE 2 = ramp-up (same, but lanes whose first column has not arrived yet idle),
        // MODE 0 = generic (all bounds checked; short texts and the drain).
        auto step = [&](const int k, const int kk, auto modeTag) {
            constexpr int MODE = decltype(modeTag)::value;
            constexpr bool FAST = MODE != 0;
            const int jt = k - lane, k1 = k + 1, jn = k1 - lane;
            uint32_t prof[NPW], profH[NPW];
#pragma unroll
            for (int q = 0; q < NPW; ++q) { prof[q] = profN[q]; profH[q] = profHN[q]; }
            const int topv = topN;
            if (!FAST) upkeep(k1);
            // next step's letter first: its shared-memory latency hides behind the sweep below
            const bool nextActive = MODE == 1 || (MODE == 2 ? jn >= 0 : (jn >= 0 && jn < n));
            const int letterN = nextActive ? (int)textWin[jn & 63] : 0;
            topN = topWin[k1 & (2 * PB - 1)];
            const int up = __shfl_up_sync(0xffffffffu, bottom, 1);
            const bool active = MODE == 1 || (MODE == 2 ? jt >= 0 : (jt >= 0 && jt < n));
            int bmax[nblk_for(R)];
            if (active) {
                const int top = (lane == 0) ? topv : up;
                if (WIDE) sweep_column_wide<R, LOCAL, NW>(c, top, prevTop, prof, profH, KL, KT, acc, 2 * R * kk, bmax);
                else sweep_column<R, LOCAL, NW>(c, top, prevTop, prof, KL, KT, acc, 2 * R * kk, bmax);
                prevTop = top;
                bottom = c[R - 1];
            }
            // (issuing the next step's shuffle here, right after the sweep, makes a lone strip 25 % faster but a long
            // chain of strips slower -- 15.1 vs 14.2 ms at 100 k x 95 k -- so the exchange stays at the top of the step)
            if (MODE == 1) st_volatile_u64_if(writesRow, rowOut + jt, (myTag << 32) | (unsigned long long)(uint32_t)bottom);
            else if (active) st_volatile_u64_if(writesRow, rowOut + jt, (myTag << 32) | (unsigned long long)(uint32_t)bottom);
            if (active) {
                if (LOCAL) {
                    const int colmax = max_of_blocks(bmax);
                    if (row0 + lane * R < m &&
                        track_argmax<R>(c, colmax, jt + 1 + (int)A.col0, snap, lane, gmCached, bestv, bestj) &&
                        colmax > gmCached) {
                        atomicMax(A.gmax, colmax);
                        gmCached = colmax;
                    }
                }
            }
            if (nextActive) {
                load_profile_words<R>(profS + letterN * PS + lane * RPAD, profN);
                if (WIDE) load_profile_words<R>(profHS + letterN * PS + lane * RPAD, profHN);
            }
            if (kk == CB - 1) {
#pragma unroll
                for (int w = 0; w < NW; ++w) { dbase[(size_t)((k / CB) * NW + w) * 32] = acc[w]; acc[w] = 0; }
            }
        };

        // ramp-up (first 32 steps) and steady state in sub-blocks of 8 steps, then the generic drain.
        // The ramp is on the critical path of the whole strip chain (strip s+1 starts when strip s has
        // produced its first columns), so it runs the lean path too.
        const int kFast1 = (n >= 64) ? (n / 32) * 32 : 0;
        const int nStepsPad = (nSteps + CB - 1) / CB * CB;
        int k = 0;
        for (; k < min(32, kFast1); k += 8) {
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
                if (k8 == 7) upkeep(k + 8);
                step(k + k8, k8 % CB, std::integral_constant<int, 2>{});
            }
        }
        for (; k < kFast1; k += 8) {
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
                if (k8 == 7) upkeep(k + 8);
                step(k + k8, k8 % CB, std::integral_constant<int, 1>{});
            }
        }
        if (LINKED && A.dbg) { const unsigned am = __activemask(); if (lane == 0) A.dbg[3 * A.n_strips + s] = am; }
        for (; k < nStepsPad; k += CB) {
#pragma unroll
            for (int kk = 0; kk < CB; ++kk) step(k + kk, kk, std::integral_constant<int, 0>{});
        }

        // ---- strip results ----
        if (LINKED && A.dbg && lane == 0) A.dbg[3 * s + 2] = gtime();
        if (LINKED && A.right_col64) {          // linked slices: straight into the right neighbour's memory
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int gi = row0 + lane * R + r + 1;
                if (gi <= m) st_volatile_u64(A.right_col64 + gi, ((unsigned long long)A.xtag << 32) | (unsigned long long)(uint32_t)c[r]);
            }
        }
        if (A.right_col) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int gi = row0 + lane * R + r + 1;
                if (gi <= m) A.right_col[gi] = c[r];
            }
            if (s == 0 && lane == 0 && A.row_base == 0) A.right_col[0] = LOCAL ? 0 : -SCALE * A.gap * (int)(A.col0 + A.n);
        }
        if (LOCAL) {
            besti = bestv > 0 ? row0 + lane * R + snapshot_first_row<R>(snap, lane, bestv) + 1 : 0;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const int ov = __shfl_xor_sync(0xffffffffu, bestv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
                const bool take = ov > bestv || (ov == bestv && (oi < besti || (oi == besti && oj < bestj)));
                if (take) { bestv = ov; besti = oi; bestj = oj; }
            }
            if (lane == 0) { A.cand_v[s] = bestv; A.cand_i[s] = bestv > 0 ? besti : 0; A.cand_j[s] = bestv > 0 ? bestj : 0; }
        } else {
            const int lm = (m - 1 - row0) / R;
            if (m - 1 >= row0 && m - 1 < row0 + ROWS && lane == lm) {
                const int rm = (m - 1 - row0) % R;
                int v = c[0];
#pragma unroll
                for (int r = 1; r < R; ++r) v = (r == rm) ? c[r] : v;
                *A.score = v / SCALE;
            }
        }
    }
}

// bottom row of a row chunk: the last strip wrote it into its ring row as {4H, tag} words
static __global__ void long_bottom_row_kernel(const unsigned long long *row, int *bottom_row, const uint32_t n)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) bottom_row[j] = (int)(uint32_t)row[j];
}

// ---------------------------------------------------------------------------------------------
// Serial device traceback over the strip layout (one thread).  Same semantics as
// batch_traceback_kernel; used for single long pairs.  Walks backwards writing into the END of
// the output buffers (capacity cap), result offset = cap - len.
struct LongTraceArgs {
    const uint8_t *text;     uint32_t n;
    const uint8_t *pattern;  uint32_t m;
    const uint32_t *dirs;    uint64_t strip_stride;
    const int32_t *S;  int alpha, gap, local;
    int R, CB;
    int C;                   // > 0: tile layout of sa_tile.cuh with C columns per tile (then CB is unused)
    uint32_t n_strips;
    const int *cand_v; const uint32_t *cand_i; const uint32_t *cand_j;
    int32_t *score;          // in (NW) / out (SW)
    char alphabet[MAX_ALPHA + 1];
    uint64_t cap;
    char *out_text; char *out_pattern;
    uint64_t *res;           // [0]=len [1]=start_text [2]=start_pattern [3]=argmax linear index
    int emit;                // 0: score / arg-max only (the reference's BENCHMARK mode)
    unsigned long long *stats;   // optional: {identity, gaps} of the emitted alignment
};

static __global__ void long_traceback_kernel(const LongTraceArgs A)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int NW = A.R * A.CB / 16;
    const int ROWS = 32 * A.R;
    const int n = (int)A.n, m = (int)A.m;
    int i, j, H;
    if (A.local) {
        int bv = 0; uint32_t bi = 0, bj = 0;
        for (uint32_t s = 0; s < A.n_strips; ++s) {
            const int v = A.cand_v[s];
            const uint32_t ci = A.cand_i[s], cj = A.cand_j[s];
            if (v > bv || (v == bv && v > 0 && (ci < bi || (ci == bi && cj < bj)))) { bv = v; bi = ci; bj = cj; }
        }
        H = bv / SCALE; i = (int)bi; j = (int)bj;
        *A.score = H;
        A.res[3] = (uint64_t)i * (uint64_t)(n + 1) + (uint64_t)j;
    } else {
        H = *A.score; i = m; j = n;
        A.res[3] = 0;
    }
    if (!A.emit) { A.res[0] = 0; A.res[1] = 0; A.res[2] = 0; return; }
    const char GAPC = A.alphabet[A.alpha];
    char *oT = A.out_text + A.cap, *oP = A.out_pattern + A.cap;
    uint64_t len = 0, nIdent = 0, nGap = 0;
    size_t cachedAddr = ~(size_t)0; uint32_t cachedWord = 0;
    auto fetch = [&](int ii, int jj) -> int {
        const int s = (ii - 1) / ROWS, rr = (ii - 1) % ROWS;
        const int ll = rr / A.R, r = rr % A.R;
        int bit; size_t addr;
        if (A.C) {
            const int k = (jj - 1) / A.C + ll, cc = (jj - 1) % A.C;
            bit = (cc * A.R + r) * 2;
            addr = (size_t)s * A.strip_stride + ((size_t)k * 32 + ll) * (A.R * A.C / 16) + (bit >> 5);
        } else {
            const int k = (jj - 1) + ll;
            const int kb = k / A.CB, kk = k % A.CB;
            bit = (kk * A.R + r) * 2;
            addr = (size_t)s * A.strip_stride + (size_t)(kb * NW + (bit >> 5)) * 32 + ll;
        }
        if (addr != cachedAddr) { cachedAddr = addr; cachedWord = A.dirs[addr]; }
        return (cachedWord >> (bit & 31)) & 3;
    };
    auto count = [&](const bool takeT, const bool takeP, const int tIdx, const int pIdx) {
        if (takeT && takeP) nIdent += A.text[tIdx] == A.pattern[pIdx]; else ++nGap;
    };
    int ti, pi;
    if (!A.local) {
        ti = n - 1; pi = m - 1;
        while (i > 0 || j > 0) {
            int tag;
            if (j == 0) tag = TAG_TOP;
            else if (i == 0) tag = TAG_LEFT;
            else tag = fetch(i, j);
            const bool takeT = tag != TAG_TOP, takeP = tag != TAG_LEFT;
            ++len;
            oT[-(int64_t)len] = takeT ? A.alphabet[A.text[ti]] : GAPC;
            oP[-(int64_t)len] = takeP ? A.alphabet[A.pattern[pi]] : GAPC;
            count(takeT, takeP, ti, pi);
            ti = max(0, ti - (int)takeT);
            pi = max(0, pi - (int)takeP);
            i -= takeP; j -= takeT;
        }
    } else {
        ti = j - 1; pi = i - 1;
        while (H > 0) {
            const int tag = fetch(i, j);
            const bool takeT = tag != TAG_TOP, takeP = tag != TAG_LEFT;
            ++len;
            oT[-(int64_t)len] = takeT ? A.alphabet[A.text[ti]] : GAPC;
            oP[-(int64_t)len] = takeP ? A.alphabet[A.pattern[pi]] : GAPC;
            count(takeT, takeP, ti, pi);
            if (tag == TAG_DIAG) H -= A.S[A.pattern[i - 1] * A.alpha + A.text[j - 1]]; else H += A.gap;
            i -= takeP; j -= takeT;
            if (i == 0 || j == 0) break;
            ti = max(0, ti - (int)takeT);
            pi = max(0, pi - (int)takeP);
        }
    }
    A.res[0] = len;
    A.res[1] = (uint64_t)(int64_t)ti;
    A.res[2] = (uint64_t)(int64_t)pi;
    if (A.stats) { A.stats[0] = nIdent; A.stats[1] = nGap; }
}

// ---------------------------------------------------------------------------------------------
// Traceback through ONE column slice of a pair that is split over several GPUs (config 5,
// global alignment).  The path enters the slice on its right edge at DP row start_row and is
// followed (reference rules, alignSequenceCPU.cpp:64-114) until it reaches the slice's left edge --
// or, in the first slice (col0 == 0), the matrix origin.  The piece is written backwards into the
// end of the output buffers; the neighbour on the left continues from res[1].
struct StripTraceArgs {
    const uint8_t *text;     uint32_t n;        // the slice's letters
    const uint8_t *pattern;  uint32_t m;
    const uint32_t *dirs;    uint64_t strip_stride;
    int alpha;
    int R, CB;
    int C;                   // > 0: tile layout (see LongTraceArgs)
    uint32_t col0;
    uint64_t start_row;
    char alphabet[MAX_ALPHA + 1];
    uint64_t cap;
    char *out_text; char *out_pattern;
    uint64_t *res;           // [0]=len of the piece  [1]=row where the path leaves the slice  [2],[3]= text/pattern index state
};

static __global__ void strip_traceback_kernel(const StripTraceArgs A)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int NW = A.R * A.CB / 16;
    const int ROWS = 32 * A.R;
    int i = (int)A.start_row, j = (int)A.n;
    const bool first = A.col0 == 0;
    const char GAPC = A.alphabet[A.alpha];
    char *oT = A.out_text + A.cap, *oP = A.out_pattern + A.cap;
    uint64_t len = 0;
    size_t cachedAddr = ~(size_t)0; uint32_t cachedWord = 0;
    auto fetch = [&](int ii, int jj) -> int {
        const int s = (ii - 1) / ROWS, rr = (ii - 1) % ROWS;
        const int ll = rr / A.R, r = rr % A.R;
        int bit; size_t addr;
        if (A.C) {
            const int k = (jj - 1) / A.C + ll, cc = (jj - 1) % A.C;
            bit = (cc * A.R + r) * 2;
            addr = (size_t)s * A.strip_stride + ((size_t)k * 32 + ll) * (A.R * A.C / 16) + (bit >> 5);
        } else {
            const int k = (jj - 1) + ll;
            const int kb = k / A.CB, kk = k % A.CB;
            bit = (kk * A.R + r) * 2;
            addr = (size_t)s * A.strip_stride + (size_t)(kb * NW + (bit >> 5)) * 32 + ll;
        }
        if (addr != cachedAddr) { cachedAddr = addr; cachedWord = A.dirs[addr]; }
        return (cachedWord >> (bit & 31)) & 3;
    };
    int ti = j - 1, pi = i - 1;
    while (j > 0 || (first && i > 0)) {
        int tag;
        if (j == 0) tag = TAG_TOP;
        else if (i == 0) tag = TAG_LEFT;
        else tag = fetch(i, j);
        const bool takeT = tag != TAG_TOP, takeP = tag != TAG_LEFT;
        ++len;
        oT[-(int64_t)len] = takeT ? A.alphabet[A.text[max(ti, 0)]] : GAPC;
        oP[-(int64_t)len] = takeP ? A.alphabet[A.pattern[max(pi, 0)]] : GAPC;
        ti -= (int)takeT;
        pi -= (int)takeP;
        i -= takeP; j -= takeT;
    }
    A.res[0] = len;
    A.res[1] = (uint64_t)i;
    A.res[2] = (uint64_t)(int64_t)max(ti, 0);
    A.res[3] = (uint64_t)(int64_t)max(pi, 0);
}

} // namespace sa

// sa_tile.cuh -- one long pair, register-tiled (BASELINE configs 1-3 and each GPU's column slice of config 5).
//
// Replaces long_fill_kernel (sa_long.cuh) for every score matrix that fits the one-byte profile.  Same strip chain --
// the (m+1) x (n+1) matrix is cut into horizontal strips of 32*R rows, one warp per strip, lane l owns R rows, all
// strips form one systolic chain through {4H, tag} words in L2 -- but a lane now advances by a TILE of R rows x C
// columns per "macro-step" instead of one column:
//
//   * the R*C cells of a tile are independent along its anti-diagonals, so ONE warp has enough instruction-level
//     parallelism to keep its scheduler issuing (the one-column step was a serial chain of R cells: 3 cycles per
//     instruction, 6 % of the warp slots active -- profiles/r01_long_c3_ncu.txt);
//   * the per-step bookkeeping (neighbour exchange, text / profile / top-row fetch, boundary store, direction store,
//     loop control) is paid once per R*C cells: C shuffles, C profile loads, one direction store of R*C/16 words;
//   * the bottom value of column cc is handed to the lane below as soon as that column is done (the shuffle of
//     column 0 overlaps the sweep of columns 1..C-1), so the exchange latency is off the critical path;
//   * lane 31 publishes the strip's bottom row with 128-bit stores (two {4H, tag} words each), the strip below fetches
//     it with 128-bit loads, TG macro-steps (TG*C columns) ahead of use.
//
// Cell arithmetic, tie-breaking and the 2-bit tags are those of sa_cell.cuh ("tagged max"); results are bit-identical
// to long_fill_kernel and to the reference's alignSequenceCPU.cpp (tests/test_gpu_parity.py).
//
// Direction words ("tile layout"): word(s, k, lane, w) at  s*strip_stride + (k*32 + lane)*NWT + w,  k = macro-step =
// tile index + lane, NWT = R*C/16; cell (r, cc) of the tile sits at bit (cc*R + r)*2 of the lane's NWT words.  One
// 4/8/16-byte store per lane and macro-step, 128..512 contiguous bytes per warp; 0.25 B/cell.
//
// SW arg-max (first maximum in row-major order, alignSequenceCPU.cpp:191-192), without a branch in the macro-step.  The
// local kernel of 8-row lanes carries 8*H instead of 4*H (tags in bits 0..1, bit 2 zero; tile_scale), so a masked cell value
// has three free low bits: every cell enters the tile's maximum as KEY = 8*H + (R-1-r) -- one VIADDMNMX per cell, the add is free -- and a
// larger key means a larger score OR the same score in a smaller row of the lane, which is exactly when a later tile
// replaces an earlier one.  A lane whose tile key beats its best only keeps the tile's inputs (R left values, C top values,
// the corner, the macro-step: register selects); the cell is located once, when the strip is over, by recomputing that one
// tile.  (The first version located on the spot behind a warp vote: the head of a growing local alignment sets a record in
// almost every tile it crosses, and any branch splits the straight-line macro-step -- the SW sweep cost 2.8x the NW sweep.)
#pragma once
#include "sa_tile_host.h"

namespace sa {

#ifndef SA_TILE_TG
#define SA_TILE_TG 8
#endif
#ifndef SA_TILE_SLEEP
#define SA_TILE_SLEEP 20
#endif
constexpr int TG = SA_TILE_TG;     // macro-steps per top-row group (prefetch distance of the strip hand-off)
#ifndef SA_TILE_TG_REQ
#define SA_TILE_TG_REQ 3
#endif
constexpr int TG_REQ = SA_TILE_TG_REQ;   // macro-step of a group after which the next group is requested
constexpr int TEXT_RING = 128;     // tiles of text kept in shared memory per warp
#ifndef SA_TILE_WATCHDOG
#define SA_TILE_WATCHDOG 0
#endif
#ifndef SA_TILE_DBG_PLAIN
#define SA_TILE_DBG_PLAIN 0        // 1 (dev builds): the per-strip timestamps of SA_LONG_DBG also in the plain kernels
#endif
// In-block hand-off (compile-time option, OFF): the strips of one block are neighbours in the chain, so lane 31 can store
// the {4H, tag} words into a ring of HRING tiles in the NEXT warp's shared memory and that warp's top-row upkeep can poll
// the ring instead of L2.  Built, bit-exact (the GPU parity tests pass with it) and measured slower on B200: config 3 fills
// in 11.6 ms against 10.2 ms, a lone strip sweeps 9 % slower (the extra predicated store and the two-source upkeep change
// the schedule of the straight-line macro-step) and the lag per strip does not shrink (11.8 us against 10.5 us in the
// instrumented build): the lag is not the L2 round trip -- see DESIGN.md 4.2.  -DSA_TILE_HANDOFF=1 turns it on.
#ifndef SA_TILE_HANDOFF
#define SA_TILE_HANDOFF 0
#endif
#ifndef SA_TILE_RING_SLEEP
#define SA_TILE_RING_SLEEP 40
#endif
constexpr int HRING_SHIFT = 8;
constexpr int HRING = 1 << HRING_SHIFT;   // tiles of the in-block hand-off ring (the producer runs 9..60 tiles ahead)

__host__ __device__ constexpr size_t tile_warp_smem(int R, int C, int alpha)
{
    // profile, text ring, top-row window (2*TG tiles), exchange buffer (C/2 parts of 32 x 2 ints), hand-off ring of
    // HRING tiles x C {4H, tag} words + the consumer's progress word
    return ((size_t)alpha * 32 * rpad_for(R) + (size_t)TEXT_RING * C + (size_t)2 * TG * C * 4 + (size_t)(C / 2) * 256 +
            (SA_TILE_HANDOFF ? (size_t)HRING * C * 8 + 16 : 0) + 15) & ~(size_t)15;
}

__device__ __forceinline__ void ld_volatile_v2u64(const unsigned long long *p, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// C bottom values of lane 31 as {4H, tag} words, two per 128-bit store, under ONE predicate (flag != 0) instead of a branch
template <int C>
__device__ __forceinline__ void st_row_words_if(const uint32_t flag, unsigned long long *p, const int (&v)[C], const uint32_t tagHi)
{
    if (C == 2)
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.volatile.global.v4.u32 [%1], {%2, %4, %3, %4};\n\t}"
                     ::"r"(flag), "l"(p), "r"(v[0]), "r"(v[1 % C]), "r"(tagHi) : "memory");
    else {
#pragma unroll
        for (int cc = 0; cc < C; cc += 4)
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.volatile.global.v4.u32 [%1], {%2, %6, %3, %6};\n\t"
                         "@q st.volatile.global.v4.u32 [%1+16], {%4, %6, %5, %6};\n\t}"
                         ::"r"(flag), "l"(p + cc), "r"(v[cc]), "r"(v[(cc + 1) % C]), "r"(v[(cc + 2) % C]), "r"(v[(cc + 3) % C]), "r"(tagHi) : "memory");
    }
}

// the same words into the hand-off ring of the next warp (shared memory), under one predicate
template <int C>
__device__ __forceinline__ void st_ring_words_if(const uint32_t flag, const uint32_t saddr, const int (&v)[C], const uint32_t tag)
{
#pragma unroll
    for (int cc = 0; cc < C; cc += 2)
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.volatile.shared.v4.u32 [%1], {%2, %4, %3, %4};\n\t}"
                     ::"r"(flag), "r"(saddr + cc * 8), "r"(v[cc]), "r"(v[(cc + 1) % C]), "r"(tag) : "memory");
}
__device__ __forceinline__ void lds_volatile_v2u64(const uint32_t saddr, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(saddr) : "memory");
}
__device__ __forceinline__ unsigned long long lds_volatile_u64(const uint32_t saddr)
{
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile_u64(const uint32_t saddr, const unsigned long long v)
{
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(saddr), "l"(v) : "memory");
}

// SW arg-max, once per strip and lane: recompute one tile (plain cells, no tags) and return r*C + cc of the row-major first
// cell (within the tile's first ncols columns) whose value equals v (R*C when there is none).
template <int R, int C, int MASK>
__device__ __noinline__ int tile_locate(const int v, const int KL, const int KT, const int corner,
                                        const int (&cin)[R], const int (&top)[C], const uint32_t (&pw)[C][(R + 3) / 4], const int ncols)
{
    int c[R];
#pragma unroll
    for (int r = 0; r < R; ++r) c[r] = cin[r];
    int key = R * C;
    int d0 = corner;
#pragma unroll
    for (int cc = 0; cc < C; ++cc) {
        int t = top[cc], d = d0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int x = __dp4a((int)pw[cc][r >> 2], onehot(r), d);
            const int cn = viaddmax_relu(t, KT, viaddmax(c[r], KL, x)) & MASK;
            d = c[r]; t = cn; c[r] = cn;
            if (cn == v && cc < ncols) key = min(key, r * C + cc);
        }
        d0 = top[cc];
    }
    return key;
}

template <int R, int C, bool LOCAL, int WARPS, bool LINKED = false>
__global__ void __launch_bounds__(WARPS * 32) tile_fill_kernel(const LongArgs A)
{
    static_assert((R * C) % 16 == 0, "a tile must fill whole direction words");
    static_assert(C == 2 || C == 4 || C == 8, "tile width");
    constexpr bool DBG = LINKED || SA_TILE_DBG_PLAIN;
    constexpr int NWT = tile_nwt(R, C);
    static_assert(NWT == 1 || NWT == 2 || NWT == 4, "direction words per lane and macro-step: one vector store");
    constexpr int RPAD = rpad_for(R);
    constexpr int PS = 32 * RPAD;
    constexpr int NPW = (R + 3) / 4;
    constexpr int ROWS = 32 * R;

    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int alpha = A.alpha;
    int8_t *S4s = reinterpret_cast<int8_t *>(smem);
    unsigned char *wbase = smem + 32 * MAX_ALPHA + (size_t)warp * tile_warp_smem(R, C, alpha);
    unsigned char *profS = wbase;
    unsigned char *textRing = profS + alpha * PS;
    int *topWin = reinterpret_cast<int *>(textRing + TEXT_RING * C);
    int *xbuf = topWin + 2 * TG * C;
    // in-block hand-off: this warp's ring (filled by the warp above) and progress word (read by the warp above)
    const uint32_t sRing = (uint32_t)__cvta_generic_to_shared(xbuf + (C / 2) * 64);
    const uint32_t sProg = sRing + HRING * C * 8;
    const uint32_t sRingDown = sRing + (uint32_t)tile_warp_smem(R, C, alpha), sProgDown = sProg + (uint32_t)tile_warp_smem(R, C, alpha);
    if (SA_TILE_HANDOFF) {
        // ring words of an earlier launch must never match: the ring starts out zeroed and every tag has bit 31 set
        for (int i = lane; i < HRING * C / 2; i += 32)
            asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(sRing + i * 16), "r"(0u) : "memory");
        if (lane == 0) sts_volatile_u64(sProg, ~0ull);
    }
    for (int i = threadIdx.x; i < 32 * MAX_ALPHA; i += blockDim.x) S4s[i] = A.S4[i];
    __syncthreads();
    const unsigned char *profL = profS + lane * RPAD;       // this lane's R profile bytes of letter 0

    constexpr int SC = tile_scale(R, LOCAL);     // the local kernels carry SC*H: log2(SC) free low bits for the arg-max key
    constexpr int VMASK = ~(SC - 1);
    static_assert(R <= SC || !LOCAL, "the arg-max key holds the lane's row in the free low bits");
    const int KL = 2 - SC * A.gap, KT = 1 - SC * A.gap;
    const uint32_t W = gridDim.x * WARPS;
    const int n = A.n_dev ? max(1, min((int)A.n, *A.n_dev)) : (int)A.n, m = (int)A.m;
    const int nTiles = (n + C - 1) / C;          // tiles per lane
    const int nFull = n / C;                     // tiles that lie completely inside the text
    const int kEnd = nTiles + 31;                // macro-steps of a strip

    uint32_t wave = 0;
    for (uint32_t s = blockIdx.x * WARPS + warp; s < A.n_strips; s += W, ++wave) {
        const int row0 = (int)s * ROWS;                 // pattern index of the strip's first row
        // the strip above / below belongs to the neighbouring warp of this block (same wave): hand-off in shared memory
        const bool upInBlock = SA_TILE_HANDOFF == 1 && warp > 0 && s > 0;          // (2: dev aid, the ring is written but not read)
        const bool downInBlock = SA_TILE_HANDOFF && warp + 1 < WARPS && s + 1 < A.n_strips;
        // this warp is done with its previous strip: the warp above may overwrite the ring (progress = {wave, tiles copied})
        if (SA_TILE_HANDOFF && lane == 0) sts_volatile_u64(sProg, (unsigned long long)wave << 32);
        // ---- query profile of this strip: prof[a][lane*RPAD + r] = 4*S[p_row][a], padding rows -128 ----
        __syncwarp();
        for (int i = lane; i < ROWS; i += 32) {
            const int off = (i / R) * RPAD + (i % R);
            const int gi = row0 + i;
            if (gi < m) {
                const int8_t *srow = S4s + 32 * min((int)A.pattern[gi], alpha - 1);
                for (int a = 0; a < alpha; ++a) profS[a * PS + off] = (unsigned char)((SC / SCALE) * srow[a]);      // (host: |SC*S| <= 127)
            } else {
                for (int a = 0; a < alpha; ++a) profS[a * PS + off] = (unsigned char)0x80;
            }
        }
        // ---- text ring: tiles 0..63 now, 64..95 in flight ----
        auto load_tile_letters = [&](const int tile) -> unsigned long long {
            unsigned long long w = 0;
#pragma unroll
            for (int cc = 0; cc < C; ++cc) {
                const int j = tile * C + cc;
                const unsigned letter = j < n ? (unsigned)min((int)A.text[j], alpha - 1) : 0u;
                w |= (unsigned long long)letter << (8 * cc);
            }
            return w;
        };
        auto store_tile_letters = [&](const int tile, const unsigned long long w) {
            unsigned char *p = textRing + (tile & (TEXT_RING - 1)) * C;
            if (C == 2) *reinterpret_cast<unsigned short *>(p) = (unsigned short)w;
            else if (C == 4) *reinterpret_cast<uint32_t *>(p) = (uint32_t)w;
            else *reinterpret_cast<unsigned long long *>(p) = w;
        };
        store_tile_letters(lane, load_tile_letters(lane));
        store_tile_letters(32 + lane, load_tile_letters(32 + lane));
        store_tile_letters(64 + lane, 0ull);
        store_tile_letters(96 + lane, 0ull);            // tiles -32..-1 of the ramp read these slots
        unsigned long long tnext = load_tile_letters(64 + lane);
        __syncwarp();

        // ---- boundary state (left border of the slice) ----
        auto gtime = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
        if (DBG && A.dbg && lane == 0) A.dbg[3 * s] = gtime();
        // linked slices: wait (warp-uniformly, see sa_long.cuh) until the left neighbour's kernel has delivered the word
        auto linked_border = [&](const int gi) -> int {
            const bool need = gi > 0 && gi <= m;
            unsigned long long v = need ? ld_volatile_u64(A.left_col64 + gi) : 0ull;
            const long long t0 = clock64();
            for (unsigned it = 1;; ++it) {
                const bool ok = !need || (uint32_t)(v >> 32) == A.xtag;
                if (__all_sync(0xffffffffu, ok)) break;
                if ((it & 1023u) == 0) {
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
        int corner;                                     // 4*H(i0-1, column before the tile)
        {
            const int gi = row0 + lane * R;
            if (LINKED && A.left_col64) corner = linked_border(gi);
            else if (A.left_col) corner = gi <= m ? A.left_col[gi] : 0;
            else corner = LOCAL ? 0 : -SCALE * A.gap * (gi + (int)A.row_base);
        }
        if (DBG && A.dbg) { __syncwarp(); if (lane == 0) A.dbg[3 * s + 1] = gtime(); }

        int bestv = 0, besti = 0, bestj = 0;
        int bestKey = 0;                                // SW: 8*H + (R-1-row in lane) of the lane's best cell so far
        // inputs of the lane's best tile so far (straight-line modes): see "SW arg-max" above
        int snapC[R], snapTop[C], snapCorner = 0, snapK = -1;
#pragma unroll
        for (int r = 0; r < R; ++r) snapC[r] = 0;
#pragma unroll
        for (int cc = 0; cc < C; ++cc) snapTop[cc] = 0;
        const bool rowsValid = row0 + lane * R < m;
        const bool hasUp = s > 0, hasDown = s + 1 < A.n_strips;
        // lane 31 publishes the strip's bottom row for the strip below -- or, for the last strip of a row chunk, for
        // long_bottom_row_kernel, which copies it out of the ring
        const bool writesRow = lane == 31 && ((hasDown && !(downInBlock && SA_TILE_HANDOFF == 1)) || (!hasDown && A.bottom_row != nullptr));
        const uint32_t writesFlag = writesRow ? 1u : 0u;
        const uint32_t ringFlag = (lane == 31 && downInBlock) ? 1u : 0u;
        // ring words carry {wave, tile / HRING}: a slot read before its tile has arrived never matches
        const uint32_t ringTagBase = 0x80000000u | ((wave & 0x7fu) << 24);
        const unsigned long long *rowIn = A.rowbuf + (size_t)((s + A.ring - 1) % A.ring) * A.row_stride;
        unsigned long long *rowOut = A.rowbuf + (size_t)(s % A.ring) * A.row_stride;
        const uint32_t wantTag = A.tag_base | s;          // producer s-1 writes (s-1)+1
        const unsigned long long myTagHi = (unsigned long long)(A.tag_base | (s + 1)) << 32;
        uint32_t *dbase = A.dirs + (size_t)s * A.strip_stride + lane * NWT;

        // ---- top row of the strip: groups of TG tiles, fetched one group ahead by lanes 0..TG-1 into a window of 2*TG
        // tiles (lane 0 reads tile k+1 while macro-step k runs, so a group must survive the arrival of the next one) ----
        unsigned long long pend[C];
#pragma unroll
        for (int cc = 0; cc < C; ++cc) pend[cc] = 0;
        unsigned long long dbgSpins = 0, dbgSpinNs = 0, dbgStalls = 0, dbgRampStalls = 0, dbgEnter = 0, dbgExit = 0, dbgWrite = 0;      // (LINKED + dbg only)
        uint32_t dbgEvents = 0;
#if SA_TILE_WATCHDOG
        bool gaveUp = false;                            // the watchdog of the top-row wait fired (this launch's results are void)
#endif
        // group with first tile `first`: lanes q < TG own tile first + q (window slot (first + q) % (2*TG))
        auto request_top = [&](const int first) {
            if (upInBlock) return;                       // shared memory: read when needed
            const int col = (first + lane) * C;
            if (lane < TG && col >= 0 && col < n) {
#pragma unroll
                for (int cc = 0; cc < C; cc += 2) ld_volatile_v2u64(rowIn + col + cc, pend[cc], pend[cc + 1]);
            }
        };
        auto ring_load = [&](const int tile) {
            const uint32_t a = sRing + (uint32_t)(tile & (HRING - 1)) * (C * 8);
#pragma unroll
            for (int cc = 0; cc < C; cc += 2) lds_volatile_v2u64(a + cc * 8, pend[cc], pend[cc + 1]);
        };
        auto top_upkeep = [&](const int first) {
            const int col = (first + lane) * C;
            int tv[C];
            if (hasUp) {
                const bool mine = lane < TG && col >= 0 && col < n;
                const uint32_t want = upInBlock ? (ringTagBase | (((uint32_t)(first + lane) >> HRING_SHIFT) & 0xffffffu)) : wantTag;
                if (upInBlock && mine) ring_load(first + lane);
                const unsigned long long tSpin = (DBG && A.dbg) ? gtime() : 0ull;
                bool spun = false;
                if (DBG && A.dbg && first == 8001) dbgEnter = tSpin;
#if SA_TILE_WATCHDOG
                unsigned polls = 0;
                long long tWait = 0;
#endif
                while (true) {
                    bool ok = true;
#pragma unroll
                    for (int cc = 0; cc < C; ++cc) ok = ok && (!(mine && col + cc < n) || (uint32_t)(pend[cc] >> 32) == want);
#if SA_TILE_WATCHDOG
                    if (__all_sync(0xffffffffu, ok) || gaveUp) break;
                    // Watchdog (compile-time option, off): the cooperative launch keeps every producer resident and a
                    // strip that gave up on its left neighbour GPU (linked_border) still publishes its row, so this wait
                    // cannot deadlock by construction.  With -DSA_TILE_WATCHDOG=1 a wait that outlives ~20 s raises the
                    // call's abort flag (the host then reports SA_ERR_LAUNCH) and every other wait of the launch runs
                    // out.  Off because the extra state costs the straight-line macro-steps 2.5-4 % (config 3: 10.48
                    // against 10.23 ms) -- this kernel's schedule is that sensitive.
                    if (polls == 0) tWait = clock64();
                    if ((++polls & 0x3fffu) == 0 && A.abort_flag) {
                        bool dead = *reinterpret_cast<volatile int *>(A.abort_flag) != 0;
                        if (clock64() - tWait > 40000000000ll) { atomicExch(A.abort_flag, 1); dead = true; }
                        if (__any_sync(0xffffffffu, dead)) { gaveUp = true; break; }
                    }
#else
                    if (__all_sync(0xffffffffu, ok)) break;
#endif
                    if (DBG && A.dbg) { spun = true; ++dbgSpins; }
                    if (!ok) {
                        if (upInBlock) { if (SA_TILE_RING_SLEEP > 0) __nanosleep(SA_TILE_RING_SLEEP); ring_load(first + lane); }
                        else {
                            if (SA_TILE_SLEEP > 0) __nanosleep(SA_TILE_SLEEP);
#pragma unroll
                            for (int cc = 0; cc < C; cc += 2) ld_volatile_v2u64(rowIn + col + cc, pend[cc], pend[cc + 1]);
                        }
                    }
                }
                if (DBG && A.dbg && spun) {
                    const unsigned long long dt = gtime() - tSpin;
                    dbgSpinNs += dt; ++dbgStalls; if (first <= 33) ++dbgRampStalls;
                    // the first 64 stalls after the ramp: {first tile of the group, ns}
                    if (first > 33 && dbgEvents < 64 && lane == 0) {
                        unsigned long long *ev = A.dbg + 16 * (size_t)A.n_strips + 128 * (size_t)s + 2 * dbgEvents;
                        ev[0] = (unsigned long long)first; ev[1] = dt;
                    }
                    if (first > 33) ++dbgEvents;
                }
                if (DBG && A.dbg && first == 8001) dbgExit = gtime();
#pragma unroll
                for (int cc = 0; cc < C; ++cc) tv[cc] = (int)(uint32_t)pend[cc];
            } else {
#pragma unroll
                for (int cc = 0; cc < C; ++cc)
                    tv[cc] = A.top_row ? (col >= 0 && col + cc < n ? A.top_row[col + cc] : 0) : LOCAL ? 0 : -SCALE * A.gap * (col + cc + 1 + (int)A.col0);
            }
            if (lane < TG) {
#pragma unroll
                for (int cc = 0; cc < C; ++cc) topWin[((first + lane) & (2 * TG - 1)) * C + cc] = tv[cc];
            }
            __syncwarp();
            // the tiles below first + TG are out of the ring: tell the warp above
            if (upInBlock && lane == 0) sts_volatile_u64(sProg, ((unsigned long long)wave << 32) | (uint32_t)max(0, first + TG));
            // The next group is NOT requested here but in the middle of this one (TG_REQ macro-steps later): requested
            // right away, the words of a producer that is only ~40 tiles ahead are not written yet, the reload at the
            // next group boundary costs an L2 round trip, and the lag a strip picks up that way while it starts stays
            // with it for the whole sweep (both run at the same speed) -- times the number of strips in the chain.
        };
        // Producer side of the in-block hand-off, once per group of macro-steps k .. k+TG-1: lane 31 is about to store the
        // tiles up to k+TG-32 into the ring of the warp below.  That warp must have started this wave's strip (it is done
        // with the previous one) and have copied the tiles that still occupy those slots.  (It runs ~41 tiles behind the
        // stores, the ring holds 256: the wait only bites when the warp below is held up, e.g. by its left neighbour GPU.)
        auto ring_space = [&](const int k) {
            const int last = k + TG - 32;                // last tile stored during this group
            if (!downInBlock || last < 0 || SA_TILE_HANDOFF != 1) return;
            while (true) {               // (warp-uniform exit: lanes that leave a spin loop at different times stay split)
                const unsigned long long pr = lds_volatile_u64(sProgDown);
                if (__all_sync(0xffffffffu, (uint32_t)(pr >> 32) == wave && (int)(uint32_t)pr + HRING > last)) break;
                __nanosleep(100);
            }
        };
        // text ring upkeep, every 32 macro-steps: tiles k+32..k+63 become readable, k+64..k+95 are requested
        auto text_upkeep = [&](const int k) {
            store_tile_letters(k + 32 + lane, tnext);
            tnext = load_tile_letters(k + 64 + lane);
            __syncwarp();
        };

        // ---- neighbour exchange through shared memory ----
        // Lane l leaves the bottom values of its tile in xbuf, two columns ("part") per 64-bit store, as soon as the
        // second column of the part is done; lane l+1 picks them up for its next macro-step.  Lane 0 reads the top-row
        // window instead -- the same code with another address -- so there is no shuffle, no lane-0 select, and the
        // exchange latency hides behind the rest of the tile: the parts 0..NP-2 of the next macro-step are loaded just
        // before the last cell of this one, the last part right at its start (it is needed C-2 diagonals later).
        // (Shuffles could not be placed: ptxas gathers them at the end of the macro-step, where each one stalls the
        // first cells of the next tile -- profiles/r02_tile44_ncu.txt.)  All accesses are volatile: they stay in
        // program order, and the lanes of the warp run in lock-step through the straight-line modes.
        constexpr int NP = C / 2;
        const uint32_t sX = (uint32_t)__cvta_generic_to_shared(xbuf);
        const uint32_t wrBase = sX + lane * 8;                                          // part p at + p*256
        const uint32_t rdBase = lane ? sX + (lane - 1) * 8 : (uint32_t)__cvta_generic_to_shared(topWin);
        const uint32_t rdPart = lane ? 256u : 8u;                                       // bytes between parts
        const uint32_t rdTile = lane ? 0u : (uint32_t)(C * 4);                          // bytes between tiles (lane 0 only)
        auto read_part = [&](const int tile, const int p, int &a, int &b) {
            const uint32_t addr = rdBase + p * rdPart + (uint32_t)(tile & (2 * TG - 1)) * rdTile;
            asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory");
        };
        auto write_part = [&](const int p, const int a, const int b) {
            asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(wrBase + p * 256), "r"(a), "r"(b) : "memory");
        };

        // Lane 0 reads the top values of tile k+1 during macro-step k, so the groups are tiles 8g+1 .. 8g+8 and tile 0
        // comes first
        if (hasUp) request_top(1 - TG);
        top_upkeep(1 - TG);
        if (hasUp) request_top(1);
        // profile words of the macro-step about to run (prefetched one macro-step ahead), text word of the one after
        uint32_t pw[C][NPW];
        auto text_word = [&](const int kb) -> unsigned long long {
            const unsigned char *p = textRing + ((kb & (TEXT_RING - 1)) * C);
            if (C == 2) return *reinterpret_cast<const unsigned short *>(p);
            if (C == 4) return *reinterpret_cast<const uint32_t *>(p);
            return *reinterpret_cast<const unsigned long long *>(p);
        };
        auto load_profile = [&](const unsigned long long tw, uint32_t (&dst)[C][NPW]) {
#pragma unroll
            for (int cc = 0; cc < C; ++cc) {
                const uint32_t letter = (uint32_t)(tw >> (8 * cc)) & 0xffu;
                const uint32_t *p = reinterpret_cast<const uint32_t *>(profL + letter * PS);
#pragma unroll
                for (int q = 0; q < NPW; ++q) dst[cc][q] = p[q];
            }
        };
        // SW: the row-major first cell of the kept tile that holds the lane's best value, as r*C + cc
        auto locate_snapshot = [&]() -> int {
            const int kbS = snapK - lane;
            uint32_t pwS[C][NPW];
            load_profile(load_tile_letters(kbS), pwS);
            // (copies: tile_locate is not inlined and takes its arrays by reference -- the kept state itself must stay in registers)
            int cS[R], tS[C];
#pragma unroll
            for (int r = 0; r < R; ++r) cS[r] = snapC[r];
#pragma unroll
            for (int cc = 0; cc < C; ++cc) tS[cc] = snapTop[cc];
            return tile_locate<R, C, VMASK>(bestKey & VMASK, KL, KT, snapCorner, cS, tS, pwS, min(C, n - kbS * C));
        };
        load_profile(text_word(0 - lane), pw);
        unsigned long long twN = text_word(1 - lane);
        int top[C];                                     // values above the lane's first row for the coming macro-step
#pragma unroll
        for (int p = 0; p < NP; ++p) read_part(0, p, top[2 * p], top[2 * p + 1]);      // (lanes > 0: not used before their ramp ends)

        // Ramp-up: a lane whose first tile has not arrived yet computes on (its tile index is negative, nothing it writes is
        // ever read) and takes its real border state when its tile 0 starts -- so the ramp runs the same straight-line
        // code as the steady state.  The ramp is on the critical path of the whole strip chain: strip s+1 starts once
        // strip s has produced its first columns.
        int cb[R];
#pragma unroll
        for (int r = 0; r < R; ++r) cb[r] = c[r];
        const int cornerB = corner;
        unsigned long long *rowW = rowOut - (long long)lane * C;          // where lane 31 stores tile k - lane of macro-step k
        uint32_t *dirW = dbase;

        // Drain: the mirror image.  A lane that has finished its last tile computes on (tiles past the text, never read)
        // after parking its final column in cfin; the columns of the last, partial tile that lie past the text pass their
        // left value through.  The drain is on the critical path as well: every strip ends 32 macro-steps after the one
        // above it, so a slow drain (the generic mode cost 2.5x a steady macro-step) delays the end of the whole chain by
        // that much per strip (profiles/README.md, round 2).
        int cfin[R];
#pragma unroll
        for (int r = 0; r < R; ++r) cfin[r] = c[r];

        // One macro-step.  MODE 1 = steady state (every lane has a full tile), MODE 2 = ramp-up, MODE 3 = drain (see above),
        // MODE 0 = generic (all bounds checked, divergent: texts too short for the straight-line modes).
        auto macro_step = [&](const int k, auto modeTag) {
            constexpr int MODE = decltype(modeTag)::value;
            const int kb = k - lane;
            if (MODE == 2 && kb == 0) {
#pragma unroll
                for (int r = 0; r < R; ++r) c[r] = cb[r];
                corner = cornerB;
            }
            // prefetch for the next macro-step: its profile words, and the text word of the one after
            uint32_t pwN[C][NPW];
            load_profile(twN, pwN);
            twN = text_word(kb + 2);
            if (MODE == 0) {
#pragma unroll
                for (int p = 0; p < NP; ++p) read_part(k, p, top[2 * p], top[2 * p + 1]);
            } else read_part(k, NP - 1, top[C - 2], top[C - 1]);
            const bool active = MODE == 1 || (MODE == 2 ? kb >= 0 : MODE == 3 ? kb < nTiles : (kb >= 0 && kb < nTiles));
            // valid columns of the tile (drain: the tile after the last full one is partial, or empty when C divides n)
            const int ncols = (MODE == 1 || MODE == 2) ? C : MODE == 3 ? (kb == nFull ? n - nFull * C : C) : min(C, n - kb * C);
            int cin[R], topIn[C];
#pragma unroll
            for (int r = 0; r < R; ++r) cin[r] = c[r];
#pragma unroll
            for (int cc = 0; cc < C; ++cc) topIn[cc] = top[cc];
            const int cornerIn = corner;
            // tags of the tile: one partial word per column, so that the deposits do not form one long dependency chain
            uint32_t accp[C][NWT];
#pragma unroll
            for (int cc = 0; cc < C; ++cc)
#pragma unroll
                for (int w = 0; w < NWT; ++w) accp[cc][w] = 0;
            int bot[C];
#pragma unroll
            for (int cc = 0; cc < C; ++cc) bot[cc] = c[R - 1];
            int tmax = 0;
            if (MODE != 0) {
                // Straight-line modes: the cells in ANTI-DIAGONAL order (d = r + cc).  The cells of one diagonal are
                // independent, so consecutive instructions of the stream do not wait for each other: a lone warp per
                // scheduler is the normal case of this kernel (column order left ~1 stall cycle per instruction,
                // profiles/r02_tile44_ncu.txt).
                int v[R][C];
#pragma unroll
                for (int d = 0; d < R + C - 1; ++d) {
                    if (d == R + C - 2) {          // before the last cell: the early parts of the next macro-step's top values
#pragma unroll
                        for (int p = 0; p + 1 < NP; ++p) read_part(k + 1, p, top[2 * p], top[2 * p + 1]);
                    }
#pragma unroll
                    for (int cc = 0; cc < C; ++cc) {
                        const int r = d - cc;
                        if (r < 0 || r >= R) continue;
                        const int left = cc ? v[r][cc ? cc - 1 : 0] : c[r];
                        const int tp = r ? v[r ? r - 1 : 0][cc] : topIn[cc];
                        const int dg = (r && cc) ? v[r ? r - 1 : 0][cc ? cc - 1 : 0] : r ? c[r ? r - 1 : 0] : cc ? topIn[cc ? cc - 1 : 0] : corner;
                        const int x = __dp4a((int)pw[cc][r >> 2], onehot(r), dg);         // cD = 4*(D + s)
                        const int mx = viaddmax(left, KL, x);                             // max(cL, cD)
                        const int h = LOCAL ? viaddmax_relu(tp, KT, mx) : viaddmax(tp, KT, mx);
                        const int cn = h & VMASK;
                        const int bit = 2 * (cc * R + r);
                        deposit_tag(accp[cc][bit >> 5], h, cn, bit & 31);
                        const int cv = (MODE == 3 && cc >= ncols) ? left : cn;             // drain: past the text, keep the last column
                        v[r][cc] = cv;
                        if (r == R - 1) {
                            bot[cc] = cv;
                            if (cc & 1) write_part(cc >> 1, bot[cc - 1 >= 0 ? cc - 1 : 0], cv);      // hand-off as soon as the part exists
                        }
                    }
                }
                corner = topIn[C - 1];
#pragma unroll
                for (int r = 0; r < R; ++r) c[r] = v[r][C - 1];
                if (MODE == 3) {
#pragma unroll
                    for (int r = 0; r < R; ++r) cfin[r] = (kb == nTiles - 1) ? c[r] : cfin[r];
                }
                if (LOCAL) {
#pragma unroll
                    for (int cc = 0; cc < C; ++cc) {
                        int w = tmax;          // (two chains per column would be shorter; the tree is off the critical path)
#pragma unroll
                        for (int r = 0; r < R; ++r) w = viaddmax(v[r][cc], R - 1 - r, w);
                        tmax = w;
                    }
                }
            } else {
                if (active) {
                    int d0 = corner;
#pragma unroll
                    for (int cc = 0; cc < C; ++cc) {
                        if (cc < ncols) {
                            int t = topIn[cc], d = d0;
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const int x = __dp4a((int)pw[cc][r >> 2], onehot(r), d);
                                const int mx = viaddmax(c[r], KL, x);
                                const int h = LOCAL ? viaddmax_relu(t, KT, mx) : viaddmax(t, KT, mx);
                                const int cn = h & VMASK;
                                const int bit = 2 * (cc * R + r);
                                deposit_tag(accp[cc][bit >> 5], h, cn, bit & 31);
                                d = c[r]; t = cn; c[r] = cn;
                            }
                            d0 = topIn[cc];
                            corner = topIn[cc];
                            if (LOCAL) {
#pragma unroll
                                for (int r = 0; r < R; ++r) tmax = viaddmax(c[r], R - 1 - r, tmax);
                            }
                            bot[cc] = c[R - 1];
                        }
                    }
                }
                __syncwarp();          // (the region above is divergent) every lane has read its top values: publish the new ones
#pragma unroll
                for (int p = 0; p < NP; ++p) write_part(p, bot[2 * p], bot[2 * p + 1]);
                __syncwarp();
            }
            // direction words of the tile: one vector store per lane, 32*NWT contiguous words per warp
            {
                uint32_t acc[NWT];
#pragma unroll
                for (int w = 0; w < NWT; ++w) {
                    uint32_t a = accp[0][w];
#pragma unroll
                    for (int cc = 1; cc < C; ++cc) if ((2 * (cc * R) >> 5) <= w && (2 * (cc * R + R - 1) >> 5) >= w) a += accp[cc][w];
                    acc[w] = a;
                }
                if (NWT == 1) dirW[0] = acc[0];
                else if (NWT == 2) *reinterpret_cast<uint2 *>(dirW) = make_uint2(acc[0], acc[1 % NWT]);
                else *reinterpret_cast<uint4 *>(dirW) = make_uint4(acc[0], acc[1 % NWT], acc[2 % NWT], acc[3 % NWT]);
                dirW += 32 * NWT;
            }
            // bottom row of the strip -> ring (lane 31): two {4H, tag} words per 128-bit store
            if (MODE != 0) {
                const uint32_t wr = MODE == 1 ? writesFlag : MODE == 2 ? (kb >= 0 ? writesFlag : 0u) : (kb < nTiles ? writesFlag : 0u);
                st_row_words_if<C>(wr, rowW, bot, (uint32_t)(myTagHi >> 32));
                if (SA_TILE_HANDOFF) {
                    const uint32_t wq = MODE == 1 ? ringFlag : MODE == 2 ? (kb >= 0 ? ringFlag : 0u) : (kb < nTiles ? ringFlag : 0u);
                    st_ring_words_if<C>(wq, sRingDown + (uint32_t)(kb & (HRING - 1)) * (C * 8), bot, ringTagBase | (((uint32_t)kb >> HRING_SHIFT) & 0xffffffu));
                }
            } else if (active && writesRow) {
#pragma unroll
                for (int cc = 0; cc < C; ++cc)
                    if (cc < ncols) st_volatile_u64(rowW + cc, myTagHi | (uint32_t)bot[cc]);
            } else if (SA_TILE_HANDOFF && active && ringFlag) {
#pragma unroll
                for (int cc = 0; cc < C; ++cc)
                    if (cc < ncols) sts_volatile_u64(sRingDown + (uint32_t)(kb & (HRING - 1)) * (C * 8) + cc * 8,
                                                     ((unsigned long long)(ringTagBase | (((uint32_t)kb >> HRING_SHIFT) & 0xffffffu)) << 32) | (uint32_t)bot[cc]);
            }
            rowW += C;
            if (LOCAL) {
                // tmax is the tile's largest KEY (see the head of the file); scores of 0 never count
                const bool take = active && rowsValid && tmax > bestKey && tmax >= SC;
                bestKey = take ? tmax : bestKey;
                snapK = take ? k : snapK;
                snapCorner = take ? cornerIn : snapCorner;
#pragma unroll
                for (int r = 0; r < R; ++r) snapC[r] = take ? cin[r] : snapC[r];
#pragma unroll
                for (int cc = 0; cc < C; ++cc) snapTop[cc] = take ? topIn[cc] : snapTop[cc];
            }
#pragma unroll
            for (int cc = 0; cc < C; ++cc)
#pragma unroll
                for (int q = 0; q < NPW; ++q) pw[cc][q] = pwN[cc][q];
        };

        // groups of TG macro-steps: ramp-up (first 32), steady state, then the generic tail (last partial tile + drain)
        // (the last tile of a lane -- full or partial -- always belongs to the drain, which parks the final column)
        const int nStraight = nTiles - 1;
        const int kRamp = nStraight >= 32 ? 32 : 0;
        if (DBG && A.dbg && lane == 0) A.dbg[3 * A.n_strips + 4 * s] = gtime();
        for (int k = 0; k < kEnd; k += TG) {
            top_upkeep(k + 1);
            if (SA_TILE_HANDOFF) ring_space(k);
            if (k >= 32 && (k & 31) == 0) text_upkeep(k);
            if (DBG && A.dbg && lane == 0) {
                if (k == 32) A.dbg[3 * A.n_strips + 4 * s + 1] = gtime();
                if (k + TG > nStraight && k < nStraight + TG) A.dbg[3 * A.n_strips + 4 * s + 2] = gtime();
                if (k + TG >= kEnd) A.dbg[3 * A.n_strips + 4 * s + 3] = gtime();
            }
            if (k + TG <= kRamp) {
#pragma unroll
                for (int u = 0; u < TG; ++u) {
                    macro_step(k + u, std::integral_constant<int, 2>{});
                    if (u == TG_REQ && hasUp) request_top(k + 1 + TG);
                }
            } else if (k >= kRamp && kRamp > 0 && k + TG <= nStraight) {
#pragma unroll
                for (int u = 0; u < TG; ++u) {
                    macro_step(k + u, std::integral_constant<int, 1>{});
                    if (u == TG_REQ && hasUp) request_top(k + 1 + TG);
                }
            } else if (kRamp > 0) {
#pragma unroll
                for (int u = 0; u < TG; ++u) {
                    if (k + u < kEnd) macro_step(k + u, std::integral_constant<int, 3>{});
                    if (u == TG_REQ && hasUp) request_top(k + 1 + TG);
                }
            } else {
#pragma unroll 1
                for (int u = 0; u < TG; ++u) {
                    if (k + u < kEnd) macro_step(k + u, std::integral_constant<int, 0>{});
                    if (u == TG_REQ && hasUp) request_top(k + 1 + TG);
                }
            }
        }
        if (kRamp == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) cfin[r] = c[r];
        }

        // ---- strip results ----
        if (DBG && A.dbg && lane == 0) {
            A.dbg[3 * s + 2] = gtime();
            unsigned long long *x = A.dbg + 7 * (size_t)A.n_strips + 4 * (size_t)s;
            x[0] = dbgSpins; x[1] = dbgSpinNs; x[2] = dbgStalls; x[3] = dbgRampStalls;
            unsigned long long *y = A.dbg + 11 * (size_t)A.n_strips + 2 * (size_t)s;
            y[0] = dbgEnter; y[1] = dbgExit;
        }
        if (DBG && A.dbg && lane == 31) A.dbg[13 * (size_t)A.n_strips + s] = dbgWrite;
        if (LINKED && A.right_col64) {          // linked slices: straight into the right neighbour's memory
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int gi = row0 + lane * R + r + 1;
                if (gi <= m) st_volatile_u64(A.right_col64 + gi, ((unsigned long long)A.xtag << 32) | (unsigned long long)(uint32_t)cfin[r]);
            }
        }
        if (A.right_col) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int gi = row0 + lane * R + r + 1;
                if (gi <= m) A.right_col[gi] = cfin[r];
            }
            if (s == 0 && lane == 0 && A.row_base == 0) A.right_col[0] = LOCAL ? 0 : -SCALE * A.gap * (int)(A.col0 + A.n);
        }
        if (LOCAL) {
            bestv = (bestKey & VMASK) / (SC / SCALE);          // 4*H, the form the candidates of a strip are reduced in
            if (bestv > 0) {
                const int key = locate_snapshot();
                besti = row0 + lane * R + key / C + 1;
                bestj = (snapK - lane) * C + key % C + 1 + (int)A.col0;
            }
            __syncwarp();
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
                int v = cfin[0];
#pragma unroll
                for (int r = 1; r < R; ++r) v = (r == rm) ? cfin[r] : v;
                *A.score = v / SCALE;
            }
        }
    }
}

} // namespace sa

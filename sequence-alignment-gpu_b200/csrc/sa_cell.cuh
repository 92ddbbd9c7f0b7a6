// sa_cell.cuh -- the DP cell recurrence shared by the batch and the long-pair kernels.
//
// Reference semantics (alignSequenceCPU.cpp:170-192 SW, :254-273 NW):
//     L = H(i,j-1) - g ; T = H(i-1,j) - g ; D = H(i-1,j-1) + S[p][t]
//     dir = D > max(L,T) ? DIAG : (L >= T ? LEFT : TOP)     (diag only on a STRICT win, gap tie -> LEFT)
//     NW: H = max(D,L,T)          SW: H = max(0, D, L, T), dir = STOP when max <= 0
//
// "Tagged max" formulation (measured 1.7x faster on B200 than compare+select,
// profiles/r01_pipe_peaks.jsonl): every score is carried as c = 4*H and the three
// candidates get a 2-bit tie-break tag in the low bits
//     cL = 4*(L) + 2        cT = 4*(T) + 1        cD = 4*(D) + 0
// so one integer max picks the winner AND its direction with exactly the
// reference's priority (equal H: LEFT beats TOP beats DIAG):
//     h   = max(cL, cT, cD)  [SW: max(.., 0)]      -> 2 x VIADDMNMX(.RELU)
//     tag = h & 3   (2 = LEFT, 1 = TOP, 0 = DIAG)  -> the stored 2-bit direction code
//     c'  = h & ~3  (= 4*H(i,j))                   -> 1 x LOP3, carried to the 3 successors
// The adds of the constants KL = 2-4g, KT = 1-4g are absorbed by VIADDMNMX and the
// substitution score by IDP.4A on a shared-memory profile of bytes 4*S[p][t], so a
// cell costs 3 ALU-pipe + 3 FMA-pipe instructions (the two IMADs deposit the tag:
// acc += h<<pos; acc -= c'<<pos).  SW needs no STOP code: the traceback carries
// the running score and stops when it reaches 0 (H==0 <=> reference STOP).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sa {

// Stored 2-bit codes (tag values) and the reference enum they decode to
// (SequenceAlignment.hpp:122: LEFT=0, DIAG=1, TOP=2, STOP=3).
enum : int { TAG_DIAG = 0, TAG_TOP = 1, TAG_LEFT = 2 };

constexpr int SCALE = 4;          // scores are carried as SCALE*H
constexpr int MAX_ALPHA = 32;     // alphabet_size limit of the packed profile

__device__ __forceinline__ int viaddmax(int a, int b, int c) { return __viaddmax_s32(a, b, c); }
__device__ __forceinline__ int viaddmax_relu(int a, int b, int c) { return __viaddmax_s32_relu(a, b, c); }

// acc += h * 2^pos ; acc -= cn * 2^pos   (pos is a compile-time constant after unrolling)
__device__ __forceinline__ void deposit_tag(uint32_t &acc, int h, int cn, const int pos)
{
    const int mul = (int)(1u << pos);
    asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc) : "r"(h), "r"(mul));
    asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc) : "r"(cn), "r"(-mul));
}

// One-hot byte selector for IDP.4A: picks profile byte (r & 3) of a packed word.
__host__ __device__ constexpr int onehot(int r) { return 1 << (8 * (r & 3)); }

// Profile slice of one lane: R bytes (4*S[p_row][letter] for the lane's R rows), padded to
// RPAD bytes so that the widest aligned shared-memory load can be used.
__host__ __device__ constexpr int rpad_for(int R) { return (R + 3) / 4 * 4; }

template <int R>
__device__ __forceinline__ void load_profile_words(const unsigned char *p, uint32_t (&prof)[(R + 3) / 4])
{
    constexpr int RPAD = rpad_for(R);
    constexpr int NPW = (R + 3) / 4;
    if constexpr (RPAD % 8 != 0) {
#pragma unroll
        for (int q = 0; q < RPAD / 4; ++q)
            if (q < NPW) prof[q < NPW ? q : 0] = reinterpret_cast<const uint32_t *>(p)[q];
    } else if constexpr (RPAD % 16 == 0) {
#pragma unroll
        for (int q = 0; q < RPAD / 16; ++q) {
            const uint4 v = reinterpret_cast<const uint4 *>(p)[q];
            if (4 * q + 0 < NPW) prof[4 * q + 0 < NPW ? 4 * q + 0 : 0] = v.x;
            if (4 * q + 1 < NPW) prof[4 * q + 1 < NPW ? 4 * q + 1 : 0] = v.y;
            if (4 * q + 2 < NPW) prof[4 * q + 2 < NPW ? 4 * q + 2 : 0] = v.z;
            if (4 * q + 3 < NPW) prof[4 * q + 3 < NPW ? 4 * q + 3 : 0] = v.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < RPAD / 8; ++q) {
            const uint2 v = reinterpret_cast<const uint2 *>(p)[q];
            if (2 * q + 0 < NPW) prof[2 * q + 0 < NPW ? 2 * q + 0 : 0] = v.x;
            if (2 * q + 1 < NPW) prof[2 * q + 1 < NPW ? 2 * q + 1 : 0] = v.y;
        }
    }
}

// Steps accumulated per direction store block: R*CB*2 bits must be a whole number of words.
__host__ __device__ constexpr int cb_for(int R) { return (R % 16 == 0) ? 1 : (R % 8 == 0) ? 2 : (R % 4 == 0) ? 4 : 8; }

// Per-lane sweep of one DP column over R consecutive rows.
//   c[r]   in : 4*H(i_r, j-1)      out: 4*H(i_r, j)
//   top       : 4*H(i_0-1, j)      (bottom row of the lane above, or the matrix border)
//   diag      : 4*H(i_0-1, j-1)
//   prof[w]   : packed profile bytes 4*S[p_{i_r}][t_j], 4 rows per 32-bit word
//   acc[]     : direction bits; cell r deposits its tag at bit (BITBASE + 2r); BITBASE is a
//               compile-time constant after unrolling the caller's step loop
// Returns the lane's new bottom value c[R-1] (the `top` of the lane below).
// SW arg-max bookkeeping is hierarchical: the sweep keeps the maximum of every block of RB rows
// (free: it is the same max3 tree), and only the rare update path searches one block for the row.
constexpr int RB = 4;
__host__ __device__ constexpr int nblk_for(int R) { return (R + RB - 1) / RB; }

template <int R, bool LOCAL, int NACC>
__device__ __forceinline__ void sweep_column(int (&c)[R], int top, int diag, const uint32_t (&prof)[(R + 3) / 4],
                                             const int KL, const int KT, uint32_t (&acc)[NACC], const int BITBASE,
                                             int (&bmax)[nblk_for(R)])
{
    int t = top, d = diag;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int x = __dp4a((int)prof[r >> 2], onehot(r), d);           // cD = 4*(D + s)
        const int m = viaddmax(c[r], KL, x);                              // max(cL, cD)
        const int h = LOCAL ? viaddmax_relu(t, KT, m) : viaddmax(t, KT, m);
        const int cn = h & ~3;
        const int bit = BITBASE + 2 * r;
        // acc += (h - cn) << pos, written as two multiply-adds so that they issue on the FMA pipe
        // (IMAD) next to the ALU-pipe VIADDMNMX/LOP3: balanced 3 + 3 instead of 5 + 1.
        deposit_tag(acc[bit >> 5], h, cn, bit & 31);
        d = c[r];
        t = cn;
        c[r] = cn;
    }
    if (LOCAL) {
#pragma unroll
        for (int b = 0; b < nblk_for(R); ++b) {
            int v = c[b * RB];
#pragma unroll
            for (int q = 1; q < RB; q += 2) {
                const int r1 = b * RB + q, r2 = b * RB + q + 1;
                if (r2 < R && q + 1 < RB) v = __vimax3_s32(v, c[r1], c[r2]);
                else if (r1 < R) v = max(v, c[r1]);
            }
            bmax[b] = v;
        }
    }
}

// Wide scores (|4*S| > 127): the profile has a second byte plane, 4*S = 128*hi + lo with lo in 0..127 and hi a signed
// byte (|S| <= 4064); the cell adds them with two dot products -- the second one multiplies by an UNSIGNED 128, which
// only the mixed-sign PTX form of dp4a offers.  One more FMA-pipe instruction per cell, otherwise the same cell.
__device__ __forceinline__ int dp4a_s8_u8(const uint32_t a, const uint32_t b, const int c)
{
    int d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
template <int R, bool LOCAL, int NACC>
__device__ __forceinline__ void sweep_column_wide(int (&c)[R], int top, int diag, const uint32_t (&prof)[(R + 3) / 4],
                                                  const uint32_t (&profH)[(R + 3) / 4], const int KL, const int KT,
                                                  uint32_t (&acc)[NACC], const int BITBASE, int (&bmax)[nblk_for(R)])
{
    int t = top, d = diag;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int x = dp4a_s8_u8(profH[r >> 2], 128u << (8 * (r & 3)), __dp4a((int)prof[r >> 2], onehot(r), d));
        const int m = viaddmax(c[r], KL, x);
        const int h = LOCAL ? viaddmax_relu(t, KT, m) : viaddmax(t, KT, m);
        const int cn = h & ~3;
        const int bit = BITBASE + 2 * r;
        deposit_tag(acc[bit >> 5], h, cn, bit & 31);
        d = c[r];
        t = cn;
        c[r] = cn;
    }
    if (LOCAL) {
#pragma unroll
        for (int b = 0; b < nblk_for(R); ++b) {
            int v = c[b * RB];
#pragma unroll
            for (int q = 1; q < RB; q += 2) {
                const int r1 = b * RB + q, r2 = b * RB + q + 1;
                if (r2 < R && q + 1 < RB) v = __vimax3_s32(v, c[r1], c[r2]);
                else if (r1 < R) v = max(v, c[r1]);
            }
            bmax[b] = v;
        }
    }
}

// SW arg-max tracking (row-major-first maximum, alignSequenceCPU.cpp:191-192), per lane:
//   fast path  : colmax = max over the lane's rows (free max3 tree) and one compare per column;
//   new record : (colmax > bestv) the lane only SNAPSHOTS its R column values into shared memory
//                with R/4 STS.128 -- the row is resolved once, when the sweep is over;
//   tie        : (colmax == bestv) rare; resolved on the spot against the snapshot.
// Snapshot layout: uint4 snap[(R+3)/4][32], element [q][lane] holds rows 4q..4q+3 of that lane.
template <int R>
__device__ __forceinline__ void snapshot_store(uint4 *snap, const int lane, const int (&c)[R])
{
#pragma unroll
    for (int q = 0; q < (R + 3) / 4; ++q) {
        uint4 v;
        v.x = (uint32_t)c[4 * q];
        v.y = (uint32_t)(4 * q + 1 < R ? c[4 * q + 1 < R ? 4 * q + 1 : 0] : 0);
        v.z = (uint32_t)(4 * q + 2 < R ? c[4 * q + 2 < R ? 4 * q + 2 : 0] : 0);
        v.w = (uint32_t)(4 * q + 3 < R ? c[4 * q + 3 < R ? 4 * q + 3 : 0] : 0);
        snap[q * 32 + lane] = v;
    }
}

// smallest row r of the snapshot whose value equals v (R if none)
template <int R>
__device__ __forceinline__ int snapshot_first_row(const uint4 *snap, const int lane, const int v)
{
    int rfirst = R;
#pragma unroll
    for (int q = (R + 3) / 4 - 1; q >= 0; --q) {
        const uint4 x = snap[q * 32 + lane];
        if (4 * q + 3 < R && (int)x.w == v) rfirst = 4 * q + 3;
        if (4 * q + 2 < R && (int)x.z == v) rfirst = 4 * q + 2;
        if (4 * q + 1 < R && (int)x.y == v) rfirst = 4 * q + 1;
        if ((int)x.x == v) rfirst = 4 * q;
    }
    return rfirst;
}

template <int R>
__device__ __forceinline__ int first_row_with(const int (&c)[R], const int v)
{
    int rfirst = R;
#pragma unroll
    for (int r = R - 1; r >= 0; --r) rfirst = (c[r] == v) ? r : rfirst;
    return rfirst;
}

// One column of bookkeeping.  bestv/bestj: lane's record value (4*H) and its DP column; the row is
// in the snapshot.  Padding rows never exceed the valid rows above them (profile byte -128,
// gap >= 0), so the first row attaining the column maximum is always a valid one.
// `floorv` is a lower bound of the alignment-wide maximum known so far (shared by the lanes that
// work on the same pair): candidates below it can never be the arg-max, which removes the
// frequent low-score records/ties of lanes far away from the best local alignment.
// Returns true when the lane set a new record (the caller then raises the shared bound).
template <int R>
__device__ __forceinline__ bool track_argmax(const int (&c)[R], const int colmax, const int jcol, uint4 *snap,
                                             const int lane, const int floorv, int &bestv, int &bestj)
{
    if (colmax <= 0 || colmax < floorv || colmax < bestv) return false;
    if (colmax > bestv) {
        bestv = colmax;
        bestj = jcol;
        snapshot_store<R>(snap, lane, c);
        return true;
    }
    // equal value: it only wins if it sits in a smaller row than the recorded one
    if (first_row_with<R>(c, colmax) < snapshot_first_row<R>(snap, lane, bestv)) {
        bestj = jcol;
        snapshot_store<R>(snap, lane, c);
    }
    return false;
}

template <int NB>
__device__ __forceinline__ int max_of_blocks(const int (&bmax)[NB])
{
    int v = bmax[0];
#pragma unroll
    for (int b = 1; b < NB; b += 2) {
        if (b + 1 < NB) v = __vimax3_s32(v, bmax[b], bmax[b + 1]);
        else v = max(v, bmax[b]);
    }
    return v;
}

} // namespace sa

// sa_batch16.cuh -- s16x2 batch fill: TWO pairs per lane group, one in each 16-bit half.
//
// Same mapping as batch_fill_kernel (group of L lanes, lane l owns rows l*R+1..l*R+R, systolic column
// sweep, one __shfl_up per step) but every register carries the scores of two different pairs that
// sit next to each other in the class' sorted order: low half = position 2u, high half = 2u+1.
// The DPX s16x2 forms then process both cells in one instruction:
//     s2 = PRMT(profA, profB)                    sign-extended 4*S of both pairs
//     cL = VIADD.16x2(cLeft, KL)
//     m  = VIADDMNMX.S16x2(cDiag, s2, cL)
//     h  = VIADDMNMX.S16x2(.RELU)(cTop, KT, m)
//     c' = h & 0xFFFCFFFC ;  tags (h - c') deposited with two IMADs, 8 cells per half-word
// 7 instructions per TWO cells instead of 6 per cell.  Scores must fit 16 bits; the host checks
// 4*max|H| < 32000 per chunk and otherwise uses the s32 kernel (bit-identical results).
#pragma once
#include "sa_batch.cuh"

namespace sa {

__device__ __forceinline__ uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// selector: byte0 = a.byte[b], byte1 = sign(a.byte[b]), byte2 = b.byte[b], byte3 = sign(b.byte[b])
__host__ __device__ constexpr uint32_t sel16(int b) { return (uint32_t)(b | ((b | 8) << 4) | ((4 + b) << 8) | (((4 + b) | 8) << 12)); }

__host__ __device__ constexpr int cb16_for(int R) { return (R % 8 == 0) ? 1 : (R % 4 == 0) ? 2 : 4; }   // R*CB % 8 == 0

template <int R, bool LOCAL, int NACC>
__device__ __forceinline__ void sweep_column16(uint32_t (&c)[R], uint32_t top, uint32_t diag,
                                               const uint32_t (&pa)[(R + 3) / 4], const uint32_t (&pb)[(R + 3) / 4],
                                               const uint32_t KL2, const uint32_t KT2, uint32_t (&acc)[NACC],
                                               const int CELLBASE, uint32_t (&bmax)[nblk_for(R)])
{
    uint32_t t = top, d = diag;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t s2 = prmt_sx(pa[r >> 2], pb[r >> 2], sel16(r & 3));
        const uint32_t cl = __vadd2(c[r], KL2);
        const uint32_t m = __viaddmax_s16x2(d, s2, cl);
        const uint32_t h = LOCAL ? __viaddmax_s16x2_relu(t, KT2, m) : __viaddmax_s16x2(t, KT2, m);
        const uint32_t cn = h & 0xFFFCFFFCu;
        const int cell = CELLBASE + r;
        deposit_tag(acc[cell >> 3], (int)h, (int)cn, 2 * (cell & 7));
        d = c[r];
        t = cn;
        c[r] = cn;
    }
    if (LOCAL) {
#pragma unroll
        for (int b = 0; b < nblk_for(R); ++b) {
            uint32_t v = c[b * RB];
#pragma unroll
            for (int q = 1; q < RB; q += 2) {
                const int r1 = b * RB + q, r2 = b * RB + q + 1;
                if (r2 < R && q + 1 < RB) v = __vimax3_s16x2(v, c[r1 < R ? r1 : 0], c[r2 < R ? r2 : 0]);
                else if (r1 < R) v = __vmaxs2(v, c[r1 < R ? r1 : 0]);
            }
            bmax[b] = v;
        }
    }
}

template <int HALF>
__device__ __forceinline__ int half_of(uint32_t v) { return HALF ? (int)(short)(v >> 16) : (int)(short)(v & 0xffffu); }

// smallest row whose HALF equals v in the registers / in the snapshot
template <int R, int HALF>
__device__ __forceinline__ int first_row_with16(const uint32_t (&c)[R], const int v)
{
    int rfirst = R;
#pragma unroll
    for (int r = R - 1; r >= 0; --r) rfirst = (half_of<HALF>(c[r]) == v) ? r : rfirst;
    return rfirst;
}
template <int R, int HALF>
__device__ __forceinline__ int snapshot_first_row16(const uint4 *snap, const int lane, const int v)
{
    int rfirst = R;
#pragma unroll
    for (int q = (R + 3) / 4 - 1; q >= 0; --q) {
        const uint4 x = snap[q * 32 + lane];
        if (4 * q + 3 < R && half_of<HALF>(x.w) == v) rfirst = 4 * q + 3;
        if (4 * q + 2 < R && half_of<HALF>(x.z) == v) rfirst = 4 * q + 2;
        if (4 * q + 1 < R && half_of<HALF>(x.y) == v) rfirst = 4 * q + 1;
        if (half_of<HALF>(x.x) == v) rfirst = 4 * q;
    }
    return rfirst;
}
template <int R>
__device__ __forceinline__ void snapshot_store16(uint4 *snap, const int lane, const uint32_t (&c)[R])
{
#pragma unroll
    for (int q = 0; q < (R + 3) / 4; ++q) {
        uint4 v;
        v.x = c[4 * q];
        v.y = 4 * q + 1 < R ? c[4 * q + 1 < R ? 4 * q + 1 : 0] : 0;
        v.z = 4 * q + 2 < R ? c[4 * q + 2 < R ? 4 * q + 2 : 0] : 0;
        v.w = 4 * q + 3 < R ? c[4 * q + 3 < R ? 4 * q + 3 : 0] : 0;
        snap[q * 32 + lane] = v;
    }
}
// same policy as track_argmax (sa_cell.cuh), for one half
template <int R, int HALF>
__device__ __forceinline__ bool track_argmax16(const uint32_t (&c)[R], const int colmax, const int jcol, uint4 *snap,
                                               const int lane, const int floorv, int &bestv, int &bestj)
{
    if (colmax <= 0 || colmax < floorv || colmax < bestv) return false;
    if (colmax > bestv) {
        bestv = colmax;
        bestj = jcol;
        snapshot_store16<R>(snap, lane, c);
        return true;
    }
    if (first_row_with16<R, HALF>(c, colmax) < snapshot_first_row16<R, HALF>(snap, lane, bestv)) {
        bestj = jcol;
        snapshot_store16<R>(snap, lane, c);
    }
    return false;
}

template <int R, int L, bool LOCAL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) batch_fill16_kernel(const BatchArgs A)
{
    static_assert(R % 2 == 0, "R must be even");
    constexpr int G = 32 / L;
    constexpr int CB = cb16_for(R);
    constexpr int NW = R * CB / 8;          // words per store block (each word: 8 cells x 2 pairs)
    constexpr int RPAD = rpad_for(R);
    constexpr int PS = L * RPAD;
    constexpr int NPW = (R + 3) / 4;
    constexpr int NQ = (R + 3) / 4;

    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / L, l = lane % L;
    const int alpha = A.alpha;
    const uint32_t textPad = (A.max_n + 15u) & ~15u;
    const uint32_t pairBytes = (alpha * PS + textPad + 15u) & ~15u;      // profile + text of ONE pair
    constexpr uint32_t gmBytes = LOCAL ? 16 * G : 0;

    int8_t *S4s = reinterpret_cast<int8_t *>(smem);
    unsigned char *wbase = smem + 32 * MAX_ALPHA + (size_t)warp * (2 * G * pairBytes + gmBytes);
    unsigned char *profA = wbase + (size_t)(2 * g) * pairBytes, *profB = profA + pairBytes;
    unsigned char *textA = profA + alpha * PS, *textB = profB + alpha * PS;
    int *gmS = reinterpret_cast<int *>(wbase + 2 * G * pairBytes) + 4 * g;       // [0] pair A, [1] pair B
    uint4 *snapA = A.snap_ws + (size_t)(blockIdx.x * WARPS + warp) * (2 * NQ * 32), *snapB = snapA + NQ * 32;

    for (int i = threadIdx.x; i < 32 * MAX_ALPHA; i += blockDim.x) S4s[i] = A.S4[i];
    __syncthreads();

    const int KL = 2 - SCALE * A.gap, KT = 1 - SCALE * A.gap;
    const uint32_t KL2 = (uint32_t)(KL & 0xffff) * 0x10001u, KT2 = (uint32_t)(KT & 0xffff) * 0x10001u;
    const uint32_t first_pos = A.dyn->first, n_pos = A.dyn->count;
    uint32_t *const dirs = A.dirs + A.dyn->dir_base;
    const uint32_t nTasks = (n_pos + 2 * G - 1) / (2 * G);

    for (uint32_t task = blockIdx.x * WARPS + warp; task < nTasks; task += gridDim.x * WARPS) {
        const uint32_t posA = (task * G + g) * 2, posB = posA + 1;
        const bool validA = posA < n_pos, validB = posB < n_pos;
        uint32_t pairA = 0, pairB = 0; int nA = 0, mA = 0, nB = 0, mB = 0;
        const uint8_t *txA = nullptr, *ptA = nullptr, *txB = nullptr, *ptB = nullptr;
        if (validA) {
            pairA = A.order[first_pos + posA];
            const int64_t t0 = A.text_off[pairA], p0 = A.pattern_off[pairA];
            nA = (int)(A.text_off[pairA + 1] - t0); mA = (int)(A.pattern_off[pairA + 1] - p0);
            txA = A.text + t0; ptA = A.pattern + p0;
        }
        if (validB) {
            pairB = A.order[first_pos + posB];
            const int64_t t0 = A.text_off[pairB], p0 = A.pattern_off[pairB];
            nB = (int)(A.text_off[pairB + 1] - t0); mB = (int)(A.pattern_off[pairB + 1] - p0);
            txB = A.text + t0; ptB = A.pattern + p0;
        }
        __syncwarp();
        if (LOCAL && l == 0) { gmS[0] = 0; gmS[1] = 0; }
        const int nG = max(nA, nB);
        // stage both texts (padded with letter 0 up to the longer one) and build both profiles
        for (int j = l; j < nG; j += L) {
            textA[j] = j < nA ? (unsigned char)min((int)txA[j], alpha - 1) : 0;
            textB[j] = j < nB ? (unsigned char)min((int)txB[j], alpha - 1) : 0;
        }
        for (int i = l; i < L * R; i += L) {
            const int off = (i / R) * RPAD + (i % R);
            const int8_t *sa_ = i < mA ? S4s + 32 * min((int)ptA[i], alpha - 1) : nullptr;
            const int8_t *sb_ = i < mB ? S4s + 32 * min((int)ptB[i], alpha - 1) : nullptr;
            for (int a = 0; a < alpha; ++a) {
                profA[a * PS + off] = sa_ ? (unsigned char)sa_[a] : (unsigned char)0x80;
                profB[a * PS + off] = sb_ ? (unsigned char)sb_[a] : (unsigned char)0x80;
            }
        }
        __syncwarp();

        uint32_t c[R];
#pragma unroll
        for (int r = 0; r < R; ++r) c[r] = LOCAL ? 0u : (uint32_t)((-SCALE * A.gap * (l * R + r + 1)) & 0xffff) * 0x10001u;
        uint32_t prevTop = LOCAL ? 0u : (uint32_t)((-SCALE * A.gap * (l * R)) & 0xffff) * 0x10001u;
        uint32_t bottom = 0;
        int bestvA = 0, bestjA = 0, bestvB = 0, bestjB = 0;
        int finA = 0, finB = 0;                  // NW: 4*H(m, n) captured at each pair's last column

        int nmax = nG;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
        const int nSteps = nmax + L - 1;
        uint32_t *dbase = dirs + (size_t)task * A.task_stride + lane;
        const int rmA = mA > 0 ? (mA - 1) % R : 0, lmA = mA > 0 ? (mA - 1) / R : -1;
        const int rmB = mB > 0 ? (mB - 1) % R : 0, lmB = mB > 0 ? (mB - 1) / R : -1;

        // the profile words of a step are fetched one step ahead (text -> profile is two dependent
        // shared-memory loads that would otherwise sit in front of every column sweep)
        uint32_t paN[NPW], pbN[NPW];
#pragma unroll
        for (int q = 0; q < NPW; ++q) { paN[q] = 0; pbN[q] = 0; }
        if (validA && l == 0 && nG > 0) {
            load_profile_words<R>(profA + (int)textA[0] * PS, paN);
            load_profile_words<R>(profB + (int)textB[0] * PS, pbN);
        }

        uint32_t upN = 0;
        for (int kb = 0; kb * CB < nSteps; ++kb) {
            uint32_t acc[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) acc[w] = 0;
#pragma unroll
            for (int kk = 0; kk < CB; ++kk) {
                const int jt = kb * CB + kk - l;
                const uint32_t up = upN;                     // exchanged at the end of the previous step
                uint32_t pa[NPW], pb[NPW];
#pragma unroll
                for (int q = 0; q < NPW; ++q) { pa[q] = paN[q]; pb[q] = pbN[q]; }
                if (validA && jt + 1 >= 0 && jt + 1 < nG) {
                    const int la = textA[jt + 1], lb = textB[jt + 1];
                    load_profile_words<R>(profA + la * PS + l * RPAD, paN);
                    load_profile_words<R>(profB + lb * PS + l * RPAD, pbN);
                }
                int flA = 0, flB = 0;
                if (LOCAL) { flA = *reinterpret_cast<volatile int *>(gmS); flB = *reinterpret_cast<volatile int *>(gmS + 1); }
                const bool active = validA && jt >= 0 && jt < nG;
                uint32_t bmax[nblk_for(R)];
                if (active) {
                    const uint32_t top = (l == 0) ? (LOCAL ? 0u : (uint32_t)((-SCALE * A.gap * (jt + 1)) & 0xffff) * 0x10001u) : up;
                    sweep_column16<R, LOCAL, NW>(c, top, prevTop, pa, pb, KL2, KT2, acc, R * kk, bmax);
                    prevTop = top;
                    bottom = c[R - 1];
                }
                // the neighbour exchange for the NEXT step is issued before the bookkeeping below, so that
                // the shuffle latency overlaps it (every lane of the warp executes it)
                upN = __shfl_up_sync(0xffffffffu, bottom, 1);
                if (active) {
                    if (LOCAL) {
                        uint32_t cm = bmax[0];
#pragma unroll
                        for (int b = 1; b < nblk_for(R); ++b) cm = __vmaxs2(cm, bmax[b]);
                        const int cmA = half_of<0>(cm), cmB = half_of<1>(cm);
                        if (jt < nA && l * R < mA && track_argmax16<R, 0>(c, cmA, jt + 1, snapA, lane, flA, bestvA, bestjA))
                            atomicMax(gmS, cmA);
                        if (jt < nB && l * R < mB && track_argmax16<R, 1>(c, cmB, jt + 1, snapB, lane, flB, bestvB, bestjB))
                            atomicMax(gmS + 1, cmB);
                    } else {
                        if (jt == nA - 1 && l == lmA) {
                            uint32_t v = c[0];
#pragma unroll
                            for (int r = 1; r < R; ++r) v = (r == rmA) ? c[r] : v;
                            finA = half_of<0>(v);
                        }
                        if (jt == nB - 1 && l == lmB) {
                            uint32_t v = c[0];
#pragma unroll
                            for (int r = 1; r < R; ++r) v = (r == rmB) ? c[r] : v;
                            finB = half_of<1>(v);
                        }
                    }
                }
            }
#pragma unroll
            for (int w = 0; w < NW; ++w) dbase[(size_t)(kb * NW + w) * 32] = acc[w];
        }

        if (LOCAL) {
            int bestiA = bestvA > 0 ? l * R + snapshot_first_row16<R, 0>(snapA, lane, bestvA) + 1 : 0;
            int bestiB = bestvB > 0 ? l * R + snapshot_first_row16<R, 1>(snapB, lane, bestvB) + 1 : 0;
#pragma unroll
            for (int o = L / 2; o >= 1; o >>= 1) {
                int ov = __shfl_xor_sync(0xffffffffu, bestvA, o), oi = __shfl_xor_sync(0xffffffffu, bestiA, o),
                    oj = __shfl_xor_sync(0xffffffffu, bestjA, o);
                if (ov > bestvA || (ov == bestvA && (oi < bestiA || (oi == bestiA && oj < bestjA)))) { bestvA = ov; bestiA = oi; bestjA = oj; }
                ov = __shfl_xor_sync(0xffffffffu, bestvB, o); oi = __shfl_xor_sync(0xffffffffu, bestiB, o);
                oj = __shfl_xor_sync(0xffffffffu, bestjB, o);
                if (ov > bestvB || (ov == bestvB && (oi < bestiB || (oi == bestiB && oj < bestjB)))) { bestvB = ov; bestiB = oi; bestjB = oj; }
            }
            if (l == 0) {
                if (validA) { A.score[pairA] = bestvA / SCALE; A.end_i[pairA] = bestvA > 0 ? bestiA : 0; A.end_j[pairA] = bestvA > 0 ? bestjA : 0; }
                if (validB) { A.score[pairB] = bestvB / SCALE; A.end_i[pairB] = bestvB > 0 ? bestiB : 0; A.end_j[pairB] = bestvB > 0 ? bestjB : 0; }
            }
        } else {
            if (validA && l == lmA) { A.score[pairA] = finA / SCALE; A.end_i[pairA] = mA; A.end_j[pairA] = nA; }
            if (validB && l == lmB) { A.score[pairB] = finB / SCALE; A.end_i[pairB] = mB; A.end_j[pairB] = nB; }
        }
        __syncwarp();
    }
}

} // namespace sa

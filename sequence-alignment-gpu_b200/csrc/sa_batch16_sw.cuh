// sa_batch16_sw.cuh -- packed (s16x2) Smith-Waterman batch fill, straight-line version.
//
// Same cell arithmetic, lane mapping and direction layout as batch_fill16_kernel<R, 32, true>
// (sa_batch16.cuh), but the column loop has no per-step control flow:
//
//   * columns outside a pair's text (before a lane's first column, after the last one) read a
//     SENTINEL letter whose profile row is all 0x80 (= -128).  For SW that keeps not-yet-started
//     lanes at H = 0 and makes the columns past the end decay (every value there is < the value
//     it was derived from), so neither needs a predicate, and nothing they write is ever read by
//     the traceback.
//   * the text is fetched as one aligned word per FOUR steps (funnel-shifted by the lane's byte
//     phase) instead of a byte per step.
//   * the arg-max (reference: first maximum in row-major order, alignSequenceCPU.cpp:150-168) is
//     tracked per OCTET of eight columns (two quads of the direction layout): each step only reduces
//     the column to its maximum (VIMNMX3 tree); once per octet the lane compares the octet maximum
//     with its best and, on a strict improvement, keeps the octet's START state (R cells + the 8 tops
//     + the diagonal) with one PRMT per register -- branch-free, the two pairs of the word select
//     independently.  After the sweep the winning octet is replayed (8 columns, once per task) to
//     locate the exact first cell.  Equal maxima in different octets take a slow path that replays
//     both; it is only entered while the value is the pair-wide maximum so far (one REDUX.MAX per half
//     and octet).  (Per quad the copies and selects of this bookkeeping were 8 % of the loop.)
//
// Results are bit-identical to the s32 kernel; tests/test_gpu_parity.py compares all three.
#pragma once
#include "sa_batch16.cuh"

namespace sa {

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t lds_u32_off(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(saddr), "n"(OFF));
    return v;
}

// Direction words of the straight-line kernels ("octet layout", BatchClassTable.packed == 2): the tags of one lane
// for EIGHT columns (8*R cells x 2 pairs = R words) are contiguous, padded to PO = R rounded up to an even count, and an
// octet of a warp is 32*PO words:   word(task, o, lane, w) at task*stride + (o*32 + lane)*PO + w,  cell = (k&7)*R + r.
// A path that crosses a lane's rows then reads one or two 32-byte sectors per octet instead of a new 128-byte
// line on almost every step (the warp-step-major layout), which is what bounds the batch traceback.  (Until round 2
// the unit was a quad of four columns padded to 2/4/8 words: R = 9..12 stored 5-6 words as 8, 1.33-1.6 x the bytes.)
__host__ __device__ constexpr int po_for(int R) { return (R + 1) & ~1; }

constexpr int NKEEP = 8;         // columns per arg-max bookkeeping interval (two quads)
template <int R>
struct Quad16 {
    uint32_t c[R];       // the lane's R cells in the column before the octet (packed A|B)
    uint32_t top[NKEEP]; // the value above the lane's first row in the octet's eight columns
    uint32_t diag;       // ... and in the column before
};

// plain packed SW column (no direction deposit); used by the replay only
template <int R>
__device__ __forceinline__ void sw16_column_plain(uint32_t (&c)[R], uint32_t top, uint32_t diag, const uint32_t (&pa)[(R + 3) / 4],
                                                  const uint32_t (&pb)[(R + 3) / 4], const uint32_t KL2, const uint32_t KT2)
{
    uint32_t t = top, d = diag;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t s2 = prmt_sx(pa[r >> 2], pb[r >> 2], sel16(r & 3));
        const uint32_t m = __viaddmax_s16x2(d, s2, __vadd2(c[r], KL2));
        const uint32_t cn = __viaddmax_s16x2_relu(t, KT2, m) & 0xFFFCFFFCu;
        d = c[r];
        t = cn;
        c[r] = cn;
    }
}

// Re-run the eight columns of an octet and return, per half, the key 8*row + column of the row-major
// first cell equal to that half of v2 (0xFFFF when there is none).
//   lettersA/B: the eight letters of the octet (byte k = column k), sprofA/B: the lane's profile base.
template <int R>
__device__ __noinline__ uint32_t replay_quad16(const Quad16<R> s, const uint32_t v2, const unsigned long long lettersA, const unsigned long long lettersB,
                                               const uint32_t sprofA, const uint32_t sprofB, const uint32_t KL2, const uint32_t KT2)
{
    constexpr int NPW = (R + 3) / 4;
    constexpr int PS = 32 * rpad_for(R);
    uint32_t c[R];
#pragma unroll
    for (int r = 0; r < R; ++r) c[r] = s.c[r];
    uint32_t diag = s.diag;
    const int vA = half_of<0>(v2), vB = half_of<1>(v2);
    int keyA = 0xFFFF, keyB = 0xFFFF;
#pragma unroll 1
    for (int k = 0; k < NKEEP; ++k) {
        uint32_t top = s.top[0];
#pragma unroll
        for (int kk = 1; kk < NKEEP; ++kk) top = (k == kk) ? s.top[kk] : top;
        const uint32_t la = (uint32_t)(lettersA >> (8 * k)) & 0xffu, lb = (uint32_t)(lettersB >> (8 * k)) & 0xffu;
        uint32_t pa[NPW], pb[NPW];
#pragma unroll
        for (int q = 0; q < NPW; ++q) { pa[q] = lds_u32(sprofA + la * PS + 4 * q); pb[q] = lds_u32(sprofB + lb * PS + 4 * q); }
        sw16_column_plain<R>(c, top, diag, pa, pb, KL2, KT2);
        diag = top;
#pragma unroll
        for (int r = R - 1; r >= 0; --r) {
            if (half_of<0>(c[r]) == vA) keyA = min(keyA, NKEEP * r + k);
            if (half_of<1>(c[r]) == vB) keyB = min(keyB, NKEEP * r + k);
        }
    }
    return (uint32_t)keyA | ((uint32_t)keyB << 16);
}

// Shared-memory layout of a block: per warp two profiles (alpha rows of PS bytes each) and two padded texts; ONE
// sentinel row (all 0x80) behind the last warp serves every warp -- its "letter" is the distance to it in rows, so
// the warp areas are padded to a multiple of PS.  (Falls back to a sentinel row per warp when that distance does
// not fit a byte.)  Every KB counts here: the block must leave room for the traceback blocks of the previous chunk.
__host__ __device__ constexpr uint32_t sw16_text_bytes(uint32_t max_n) { return (max_n + 80u + 15u) & ~15u; }
struct Sw16Layout { uint32_t warpBytes, blockBytes, sharedSentinel; };
__host__ __device__ constexpr Sw16Layout sw16_layout(int R, int alpha, uint32_t max_n, int warps)
{
    const uint32_t PS = 32u * (uint32_t)rpad_for(R);
    const uint32_t text2 = 2u * sw16_text_bytes(max_n);
    const uint32_t wbShared = 2u * alpha * PS + (text2 + PS - 1) / PS * PS;
    if ((uint32_t)warps * (wbShared / PS) <= 255u) return Sw16Layout{wbShared, (uint32_t)warps * wbShared + PS, 1u};
    const uint32_t wb = (2u * alpha + 1u) * PS + text2;
    return Sw16Layout{wb, (uint32_t)warps * wb, 0u};
}


// Query profile of one pair, built by the lane that owns the rows: byte [a*PS + lane*RPAD + r] = 4*S[p_(lane*R+r)][a].
// The lane reads the 4*S row of each of its R pattern letters four text letters at a time (one aligned word of the
// 32-byte table row, L1-resident), transposes 4 rows x 4 letters with eight PRMTs and stores one word per letter:
// R loads + 8*NPW PRMTs + 4*NPW stores per four letters instead of a byte load and a byte store per (row, letter).
// Rows past the end of the pattern read as 0x80 (-128, the sentinel).  `prof` points at the lane's own RPAD bytes.
template <int R>
__device__ __forceinline__ void build_profile_rows(unsigned char *prof, const uint8_t *pt, const int m, const int l, const int alpha,
                                                   const int8_t *S4)
{
    constexpr int RPAD = rpad_for(R);
    constexpr int PS = 32 * RPAD;
    constexpr int NPW = (R + 3) / 4;
    const uint32_t *S4w = reinterpret_cast<const uint32_t *>(S4);
    uint32_t rowoff[R];                 // word offset of the letter's table row, ~0u past the end of the pattern
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = l * R + r;
        rowoff[r] = i < m ? 8u * (uint32_t)min((int)pt[i], alpha - 1) : ~0u;
    }
    const int na4 = (alpha + 3) >> 2;
#pragma unroll 1
    for (int a4 = 0; a4 < na4; ++a4) {
        uint32_t W[4 * NPW];
#pragma unroll
        for (int r = 0; r < 4 * NPW; ++r) W[r] = 0x80808080u;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (rowoff[r] != ~0u) W[r] = __ldg(S4w + rowoff[r] + a4);
        uint32_t *dst = reinterpret_cast<uint32_t *>(prof + (size_t)(4 * a4) * PS);
        const int left = alpha - 4 * a4;            // letters of this group that exist (1..4 in the last group)
#pragma unroll
        for (int q = 0; q < NPW; ++q) {
            const uint32_t t0 = __byte_perm(W[4 * q], W[4 * q + 1], 0x5140), t1 = __byte_perm(W[4 * q + 2], W[4 * q + 3], 0x5140);
            const uint32_t t2 = __byte_perm(W[4 * q], W[4 * q + 1], 0x7362), t3 = __byte_perm(W[4 * q + 2], W[4 * q + 3], 0x7362);
            dst[q] = __byte_perm(t0, t1, 0x5410);
            if (left > 1) dst[PS / 4 + q] = __byte_perm(t0, t1, 0x7632);
            if (left > 2) dst[2 * (PS / 4) + q] = __byte_perm(t2, t3, 0x5410);
            if (left > 3) dst[3 * (PS / 4) + q] = __byte_perm(t2, t3, 0x7632);
        }
    }
}

// What a task needs to know about one of its two pairs; loaded one task ahead of its use.
struct PairMeta { uint32_t pair; int n, m; int64_t t0, p0; };
__device__ __forceinline__ PairMeta load_pair_meta(const BatchArgs &A, const uint32_t pair, const bool valid)
{
    PairMeta M; M.pair = pair; M.n = 0; M.m = 0; M.t0 = 0; M.p0 = 0;
    if (valid) {
        M.t0 = A.text_off[pair]; M.p0 = A.pattern_off[pair];
        M.n = (int)(A.text_off[pair + 1] - M.t0); M.m = (int)(A.pattern_off[pair + 1] - M.p0);
    }
    return M;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int R, bool LOCAL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) batch_line16_kernel(const BatchArgs A)
{
    static_assert(R >= 2 && R <= 16, "strip heights of the packed classes");
    constexpr int PO = po_for(R);            // direction words per lane and octet
    constexpr int SW = (PO % 4 == 0) ? 4 : 2; // words per store: 128-bit when the lane stride keeps them aligned
    constexpr int RPAD = rpad_for(R);
    constexpr int PS = 32 * RPAD;
    constexpr int NPW = (R + 3) / 4;
    constexpr int TPAD = 32;                // sentinel letters in front of the text

    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = lane;
    const int alpha = A.alpha;
    const uint32_t textBytes = sw16_text_bytes(A.max_n);
    const Sw16Layout lay = sw16_layout(R, alpha, A.max_n, WARPS);

    // (the 4*S table is read from global memory / L1 while staging: shared memory is kept for profiles)
    const int8_t *S4s = A.S4;
    unsigned char *wbase = smem + (size_t)warp * lay.warpBytes;
    unsigned char *profA = wbase, *profB = profA + alpha * PS;
    unsigned char *textA = profB + alpha * PS + (lay.sharedSentinel ? 0 : PS), *textB = textA + textBytes;
    unsigned char *sent = lay.sharedSentinel ? smem + (size_t)WARPS * lay.warpBytes : profB + alpha * PS;
    // letter whose row is the sentinel, relative to each profile
    const uint32_t sentA = lay.sharedSentinel ? (uint32_t)(WARPS - warp) * (lay.warpBytes / PS) : 2u * alpha, sentB = sentA - alpha;
    const uint32_t sprofA = (uint32_t)__cvta_generic_to_shared(profA) + l * RPAD, sprofB = sprofA + alpha * PS;
    const uint32_t stextA = (uint32_t)__cvta_generic_to_shared(textA), stextB = stextA + textBytes;

    if (!lay.sharedSentinel || warp == 0)
        for (int i = lane; i < PS / 4; i += 32) reinterpret_cast<uint32_t *>(sent)[i] = 0x80808080u;
    __syncthreads();

    const int KL = 2 - SCALE * A.gap, KT = 1 - SCALE * A.gap;
    const uint32_t KL2 = (uint32_t)(KL & 0xffff) * 0x10001u, KT2 = (uint32_t)(KT & 0xffff) * 0x10001u;
    const uint32_t first_pos = A.dyn->first, n_pos = A.dyn->count;
    uint32_t *const dirs = A.dirs + A.dyn->dir_base;
    const uint32_t nTasks = (n_pos + 1) / 2;
    const int phase = (TPAD - l) & 3;                 // byte phase of the lane's column inside a text word
    const int word0 = (TPAD - l) >> 2;                // word that holds the lane's column at step 0

    // Task metadata runs ahead of its use so that the dependent loads (order -> offsets -> residues) of a task are
    // not paid when it starts: the pair indices two tasks ahead, the offsets one task ahead, and the next task's
    // residues are pulled into L2 while this task's matrix is filled.
    const uint32_t taskStride = gridDim.x * WARPS;
    const uint32_t task0 = blockIdx.x * WARPS + warp;
    auto order_at = [&](const uint32_t pos) -> uint32_t { return pos < n_pos ? A.order[first_pos + pos] : 0u; };
    uint32_t ordA = order_at(task0 * 2), ordB = order_at(task0 * 2 + 1);
    PairMeta nextA = load_pair_meta(A, ordA, task0 * 2 < n_pos), nextB = load_pair_meta(A, ordB, task0 * 2 + 1 < n_pos);
    ordA = order_at((task0 + taskStride) * 2); ordB = order_at((task0 + taskStride) * 2 + 1);

    for (uint32_t task = task0; task < nTasks; task += taskStride) {
        const PairMeta MA = nextA, MB = nextB;
        {
            const uint32_t t1 = task + taskStride, t2 = t1 + taskStride;
            nextA = load_pair_meta(A, ordA, t1 * 2 < n_pos); nextB = load_pair_meta(A, ordB, t1 * 2 + 1 < n_pos);
            ordA = order_at(t2 * 2); ordB = order_at(t2 * 2 + 1);
        }
        const bool validB = task * 2 + 1 < n_pos;
        const uint32_t pairA = MA.pair, pairB = MB.pair;
        const int nA = MA.n, mA = MA.m, nB = MB.n, mB = MB.m;
        const uint8_t *txA = A.text + MA.t0, *ptA = A.pattern + MA.p0, *txB = A.text + MB.t0, *ptB = A.pattern + MB.p0;
        __syncwarp();
        const int nG = max(nA, nB);
        const int nSteps = nG + 31;
        const int nOct = (nSteps + NKEEP - 1) / NKEEP;          // whole octets (two quads): the bookkeeping interval
        const int nQuads = 2 * nOct;
        // texts: TPAD sentinels, the letters, sentinels up to the last byte any quad can touch
        for (int j = l; j < 4 * nQuads + TPAD + 8; j += 32) {
            const int t = j - TPAD;
            textA[j] = (t >= 0 && t < nA) ? (unsigned char)min((int)txA[t], alpha - 1) : (unsigned char)sentA;
            textB[j] = (t >= 0 && t < nB) ? (unsigned char)min((int)txB[t], alpha - 1) : (unsigned char)sentB;
        }
        build_profile_rows<R>(profA + l * RPAD, ptA, mA, l, alpha, S4s);
        build_profile_rows<R>(profB + l * RPAD, ptB, mB, l, alpha, S4s);
        {
            // the next task's residues: text and pattern of both pairs, eight 128-byte lines each
            const PairMeta &M = (l & 16) ? nextB : nextA;
            const uint8_t *base = (l & 8) ? A.pattern + M.p0 : A.text + M.t0;
            const int len = (l & 8) ? M.m : M.n;
            if ((l & 7) * 128 < len + 127) prefetch_l2(base + (l & 7) * 128);
        }
        __syncwarp();

        // NW: a lane that has not reached its first column yet must keep its border column H(i,0) = -g*i.  Its
        // sentinel letters take DIAG out (-128) and, while k < l, the LEFT constant is +2 instead of 2-4g: then
        // cL = c+2 beats cT (= 4H of the row above - 4g + 1 = c+1) and the cell reproduces itself.  (g <= 31.)
        uint32_t c[R];
#pragma unroll
        for (int r = 0; r < R; ++r) c[r] = LOCAL ? 0u : (uint32_t)((-SCALE * A.gap * (l * R + r + 1)) & 0xffff) * 0x10001u;
        uint32_t prevTop = LOCAL ? 0u : (uint32_t)((-SCALE * A.gap * (l * R)) & 0xffff) * 0x10001u;
        uint32_t bottom = c[R - 1];
        const uint32_t G2 = (uint32_t)((SCALE * A.gap) & 0xffff) * 0x10001u;
        uint32_t border = 0u;                         // NW: 4*H(0, column) of the step, both halves
        uint32_t best2 = 0u;
        int bestqA = 0, bestqB = 0;                   // octet of the kept state, per half
        Quad16<R> snap;
#pragma unroll
        for (int r = 0; r < R; ++r) snap.c[r] = 0u;
#pragma unroll
        for (int k = 0; k < NKEEP; ++k) snap.top[k] = 0u;
        snap.diag = 0u;

        uint32_t *dptr = dirs + (size_t)task * A.task_stride + lane * PO;
        uint32_t pTextA = stextA + 4 * word0, pTextB = stextB + 4 * word0;
        uint32_t wA0 = lds_u32(pTextA), wB0 = lds_u32(pTextB);
        // the eight letters of octet o for this lane (byte k = column 8*o + k)
        auto octet_letters = [&](const uint32_t stext, const int o) -> unsigned long long {
            const int a = NKEEP * o - l + TPAD;
            const uint32_t w0 = lds_u32(stext + (a & ~3)), w1 = lds_u32(stext + (a & ~3) + 4), w2 = lds_u32(stext + (a & ~3) + 8);
            return (unsigned long long)__funnelshift_r(w0, w1, 8 * (a & 3)) | ((unsigned long long)__funnelshift_r(w1, w2, 8 * (a & 3)) << 32);
        };

        for (int o = 0; o < nOct; ++o) {
            Quad16<R> cur;
#pragma unroll
            for (int r = 0; r < R; ++r) cur.c[r] = c[r];
            cur.diag = prevTop;
            uint32_t qm = 0u;                // maximum of the octet's cells: one VIMNMX3 chain through all eight columns
            uint32_t acc[PO];
#pragma unroll
            for (int w = 0; w < PO; ++w) acc[w] = 0;
#pragma unroll
            for (int h = 0; h < NKEEP / 4; ++h) {
                pTextA += 4; pTextB += 4;
                const uint32_t wA1 = lds_u32(pTextA), wB1 = lds_u32(pTextB);
                const uint32_t la4 = __funnelshift_r(wA0, wA1, 8 * phase), lb4 = __funnelshift_r(wB0, wB1, 8 * phase);
                wA0 = wA1; wB0 = wB1;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int kc = 4 * h + k;                  // column inside the octet
                    const uint32_t up = __shfl_up_sync(0xffffffffu, bottom, 1);
                    if (!LOCAL) border = __vsub2(border, G2);
                    const uint32_t top = (l == 0) ? (LOCAL ? 0u : border) : up;
                    const uint32_t KLk = (LOCAL || NKEEP * o + kc >= l) ? KL2 : 0x00020002u;
                    const uint32_t la = (la4 >> (8 * k)) & 0xffu, lb = (lb4 >> (8 * k)) & 0xffu;
                    const uint32_t aA = sprofA + la * PS, aB = sprofB + lb * PS;
                    uint32_t pa[NPW], pb[NPW];
#pragma unroll
                    for (int w = 0; w < NPW; ++w) { pa[w] = lds_u32(aA + 4 * w); pb[w] = lds_u32(aB + 4 * w); }
                    uint32_t bmax[nblk_for(R)];
                    sweep_column16<R, LOCAL, PO>(c, top, prevTop, pa, pb, KLk, KT2, acc, R * kc, bmax);
                    cur.top[kc] = top;
                    prevTop = top;
                    bottom = c[R - 1];
                    if (LOCAL) {
#pragma unroll
                        for (int r = 0; r + 1 < R; r += 2) qm = __vimax3_s16x2(qm, c[r], c[r + 1 < R ? r + 1 : r]);
                        if (R & 1) qm = __vmaxs2(qm, c[R - 1]);
                    }
                    // the octet's tags go out as soon as a store's worth of words is complete (word w holds cells
                    // 8w .. 8w+7): few accumulators are alive at a time
#pragma unroll
                    for (int w = 0; w < PO; w += SW) {
                        const int lastCell = (8 * (w + SW) - 1 < 8 * R - 1) ? 8 * (w + SW) - 1 : 8 * R - 1;
                        if (lastCell / R == kc) {
                            if (SW == 4) *reinterpret_cast<uint4 *>(dptr + w) = make_uint4(acc[w], acc[w + 1], acc[w + 2], acc[w + 3]);
                            else *reinterpret_cast<uint2 *>(dptr + w) = make_uint2(acc[w], acc[w + 1]);
                        }
                    }
                }
            }
            dptr += 32 * PO;

            if (!LOCAL) continue;        // global: the end cell is (m, n) and the traceback re-derives the score
            // ---- arg-max bookkeeping, once per octet
            const uint32_t nb = __vmaxs2(best2, qm);
            const uint32_t chg = nb ^ best2;
            const uint32_t eq = qm ^ best2;
            const bool impA = (chg & 0xffffu) != 0u, impB = (chg >> 16) != 0u;
            // an equal maximum only matters while it is the pair-wide maximum (REDUX over the warp): the small
            // scores in front of the alignment repeat all the time
            const uint32_t fl2 = (uint32_t)__reduce_max_sync(0xffffffffu, (int)(nb & 0xffffu)) |
                                 ((uint32_t)__reduce_max_sync(0xffffffffu, (int)(nb >> 16)) << 16);
            const uint32_t te = eq | (nb ^ fl2);
            const bool tieA = (te & 0xffffu) == 0u && (qm & 0xffffu) != 0u;
            const bool tieB = (te >> 16) == 0u && (qm >> 16) != 0u;
            const uint32_t sel = (impA ? 0x54u : 0x10u) | (impB ? 0x7600u : 0x3200u);
#pragma unroll
            for (int r = 0; r < R; ++r) snap.c[r] = __byte_perm(snap.c[r], cur.c[r], sel);
#pragma unroll
            for (int k = 0; k < NKEEP; ++k) snap.top[k] = __byte_perm(snap.top[k], cur.top[k], sel);
            snap.diag = __byte_perm(snap.diag, cur.diag, sel);
            bestqA = impA ? o : bestqA;
            bestqB = impB ? o : bestqB;
            best2 = nb;
            if (tieA || tieB) {
                // the same maximum again in a later octet: it only replaces the kept one if it sits in a smaller row
                const uint32_t kOld = replay_quad16<R>(snap, best2, octet_letters(stextA, bestqA), octet_letters(stextB, bestqB), sprofA, sprofB, KL2, KT2);
                const uint32_t kNew = replay_quad16<R>(cur, best2, octet_letters(stextA, o), octet_letters(stextB, o), sprofA, sprofB, KL2, KT2);
                const bool repA = tieA && ((kNew & 0xffffu) / NKEEP) < ((kOld & 0xffffu) / NKEEP);
                const bool repB = tieB && ((kNew >> 16) / NKEEP) < ((kOld >> 16) / NKEEP);
                const uint32_t sel2 = (repA ? 0x54u : 0x10u) | (repB ? 0x7600u : 0x3200u);
#pragma unroll
                for (int r = 0; r < R; ++r) snap.c[r] = __byte_perm(snap.c[r], cur.c[r], sel2);
#pragma unroll
                for (int k = 0; k < NKEEP; ++k) snap.top[k] = __byte_perm(snap.top[k], cur.top[k], sel2);
                snap.diag = __byte_perm(snap.diag, cur.diag, sel2);
                bestqA = repA ? o : bestqA;
                bestqB = repB ? o : bestqB;
            }
        }

        if (!LOCAL) {
            if (l == 0) {
                A.end_i[pairA] = mA; A.end_j[pairA] = nA;
                if (validB) { A.end_i[pairB] = mB; A.end_j[pairB] = nB; }
            }
        } else {
            // ---- locate the first maximum inside each lane's kept octet, then reduce over the lanes
            const uint32_t key = replay_quad16<R>(snap, best2, octet_letters(stextA, bestqA), octet_letters(stextB, bestqB), sprofA, sprofB, KL2, KT2);
            const int keyA = (int)(key & 0xffffu), keyB = (int)(key >> 16);
            int bestvA = half_of<0>(best2), bestvB = half_of<1>(best2);
            int bestiA = l * R + keyA / NKEEP + 1, bestjA = NKEEP * bestqA + (keyA % NKEEP) - l + 1;
            int bestiB = l * R + keyB / NKEEP + 1, bestjB = NKEEP * bestqB + (keyB % NKEEP) - l + 1;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                int ov = __shfl_xor_sync(0xffffffffu, bestvA, o), oi = __shfl_xor_sync(0xffffffffu, bestiA, o),
                    oj = __shfl_xor_sync(0xffffffffu, bestjA, o);
                if (ov > bestvA || (ov == bestvA && (oi < bestiA || (oi == bestiA && oj < bestjA)))) { bestvA = ov; bestiA = oi; bestjA = oj; }
                ov = __shfl_xor_sync(0xffffffffu, bestvB, o); oi = __shfl_xor_sync(0xffffffffu, bestiB, o);
                oj = __shfl_xor_sync(0xffffffffu, bestjB, o);
                if (ov > bestvB || (ov == bestvB && (oi < bestiB || (oi == bestiB && oj < bestjB)))) { bestvB = ov; bestiB = oi; bestjB = oj; }
            }
            if (l == 0) {
                A.score[pairA] = bestvA / SCALE; A.end_i[pairA] = bestvA > 0 ? bestiA : 0; A.end_j[pairA] = bestvA > 0 ? bestjA : 0;
                if (validB) { A.score[pairB] = bestvB / SCALE; A.end_i[pairB] = bestvB > 0 ? bestiB : 0; A.end_j[pairB] = bestvB > 0 ? bestjB : 0; }
            }
        }
        __syncwarp();
    }
}

} // namespace sa

// sa_traceback.cuh -- PARALLEL device traceback for one long pair (strip layout of sa_long.cuh).
//
// A serial pointer chase over the packed direction matrix costs one dependent L2/HBM access every
// few steps (~0.1-0.3 us per step): 200 k steps at 100 k x 100 k would take longer than the fill.
// The traceback is therefore split at the strip boundaries ("lines", DP rows b_s = s*ROWS):
//
//   A  walkers   : for every line s and every candidate column c_q = min(q*Wd, n) a thread follows
//                  the direction tags from (b_s, c_q) up to line s-1 and records the arrival column
//                  fa[s][q].  All (s, q) are independent -> hundreds of thousands of threads.
//   B  resolve   : one thread chains the lines from the start cell: the true crossing X[s] lies
//                  between two candidates; optimal-path trees never cross, so if both candidates
//                  arrive at the same column the true path arrives there too (paths coalesce within
//                  a few dozen rows on real data).  If they differ, that one segment is walked
//                  serially (exact fallback, never wrong, only slower).
//   C  segments  : one thread per strip walks its segment (entry cell known from B), counting
//                  steps and, for local alignments, the score change (total and running minimum).
//   D  offsets   : prefix sums -> position of every segment in the output; local alignments are
//                  CUT where the running score reaches 0 (H == 0 <=> the reference's STOP).
//   E  emit      : every segment is walked once more and writes its characters straight into the
//                  final (backwards-filled) output.
//
// Semantics are those of traceBackNW / traceBackSW (alignSequenceCPU.cpp:64-114 / :10-62).
#pragma once
#include "sa_cell.cuh"

namespace sa {

struct TbLayout {
    const uint32_t *dirs;  uint64_t strip_stride;
    int R, CB, NW, ROWS, cbShift;
    int n, m;
    // tile layout of tile_fill_kernel (sa_tile.cuh): C > 0 columns per tile, NW = R*C/16 words per lane and macro-step;
    // C == 0 selects the warp-step-major layout of long_fill_kernel
    int C, cShift;
};

struct TbState {            // device-resident scalars of one traceback
    int i0, j0;             // start cell (DP coordinates)
    int H0;                 // score at the start cell (true score, not scaled)
    int s0;                 // strip containing row i0
    int cut_seg;            // local: segment in which the running score reaches 0 (-1: none, runs to the border)
    int exit_row;           // column slices: DP row at which the path leaves through the slice's left edge
    unsigned long long total_len;
    unsigned long long fallbacks;   // segments resolved by the serial fallback (diagnostic)
    // the two counts of prettyAlignmentPrint (utilities.cpp:262-283), accumulated by the emitting segments
    unsigned long long identity, gaps;
};

struct TbArgs {
    TbLayout Lay;
    const uint8_t *text;  const uint8_t *pattern;
    const int32_t *S;  int alpha, gap, local;
    uint32_t n_strips;
    int Wd, Q;                       // candidate spacing and count-1 (candidates q = 0..Q)
    int BQ;                          // walkers only cover candidates within BQ of the predicted crossing
    double slope;                    // predicted columns per row along the path (n/m global, 1 local)
    // column slice of a global alignment (sa_strip_traceback): the path starts at (start_row, n) on the right
    // edge; with `slice` set (not the first slice) it ends on the left edge instead of running up column 0
    int start_given, start_row, slice;
    // row chunk of a checkpointed (linear-space) global traceback: the path enters on the chunk's bottom row at the
    // column *start_col_dev (device value: it is where the chunk below was left) and, with chunk_top set, ends on the
    // chunk's top row instead of running along the matrix border; the column where it leaves goes to *exit_col_dev and
    // the piece is appended in front of what the chunks below have emitted (*global_off steps so far)
    const int *start_col_dev;  int chunk_top;  unsigned long long *global_off;  int *exit_col_dev;
    // fill results
    const int *cand_v; const uint32_t *cand_i; const uint32_t *cand_j;
    int32_t *score;
    // workspaces
    TbState *st;
    uint32_t *fa;                    // n_strips x (2*BQ+1) arrival columns of the banded candidates
    int *X;                          // n_strips+1 crossing columns, X[s] on row b_s
    unsigned long long *seg_len;     // per segment
    long long *seg_delta;            // local: total score change of the segment
    long long *seg_min;              // local: minimum running change inside the segment
    unsigned long long *seg_off;     // per segment: steps emitted before it
    // output
    char alphabet[MAX_ALPHA + 1];
    uint64_t cap;
    char *out_text;  char *out_pattern;
    uint64_t *res;                   // [0]=len [1]=start_text [2]=start_pattern [3]=argmax linear index
};

struct TbCursor {
    size_t cachedAddr;  uint32_t cachedWord;
    __device__ TbCursor() : cachedAddr(~(size_t)0), cachedWord(0) {}
};

// direction tag of cell (i, j), 1 <= i <= m, 1 <= j <= n
__device__ __forceinline__ int tb_fetch(const TbLayout &L, TbCursor &cur, const int i, const int j)
{
    // one division by R; ROWS = 32*R and CB is a power of two
    const int gl = (i - 1) / L.R, r = (i - 1) - gl * L.R;
    const int s = gl >> 5, ll = gl & 31;
    int bit;
    size_t addr;
    if (L.C) {
        const int k = ((j - 1) >> L.cShift) + ll, cc = (j - 1) & (L.C - 1);
        bit = (cc * L.R + r) * 2;
        addr = (size_t)s * L.strip_stride + ((size_t)k * 32 + ll) * L.NW + (bit >> 5);
    } else {
        const int k = (j - 1) + ll;
        const int kb = k >> L.cbShift, kk = k & (L.CB - 1);
        bit = (kk * L.R + r) * 2;
        addr = (size_t)s * L.strip_stride + (size_t)(kb * L.NW + (bit >> 5)) * 32 + ll;
    }
    if (addr != cur.cachedAddr) { cur.cachedAddr = addr; cur.cachedWord = __ldg(L.dirs + addr); }
    return (cur.cachedWord >> (bit & 31)) & 3;
}

// tag with the border override of traceBackNW (alignSequenceCPU.cpp:78-81)
__device__ __forceinline__ int tb_tag(const TbLayout &L, TbCursor &cur, const int i, const int j)
{
    if (j == 0) return TAG_TOP;
    if (i == 0) return TAG_LEFT;
    return tb_fetch(L, cur, i, j);
}

// Follow the tags from (i, j) until row `stop_row` is reached (column 0 then runs straight up).
// max_steps > 0 bounds the walk (speculative walkers only): a candidate far from the optimal path follows long gap
// runs before it rejoins it; such a walker gives up and reports TB_UNKNOWN, which the resolver never matches.
constexpr int TB_UNKNOWN = -1;
__device__ __forceinline__ int tb_walk_to_row(const TbLayout &L, int i, int j, const int stop_row, int max_steps = 0)
{
    TbCursor cur;
    while (i > stop_row) {
        if (j == 0) break;                       // forced TOP all the way: arrival column 0
        const int tag = tb_fetch(L, cur, i, j);
        i -= (tag != TAG_LEFT);
        j -= (tag != TAG_TOP);
        if (max_steps > 0 && --max_steps == 0 && i > stop_row) return TB_UNKNOWN;
    }
    return j;
}

// ---- start cell ------------------------------------------------------------------------------
__global__ void tb_prepare_kernel(const TbArgs A)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n = A.Lay.n, m = A.Lay.m;
    int i0, j0, H0;
    if (A.local) {
        int bv = 0; uint32_t bi = 0, bj = 0;
        for (uint32_t s = 0; s < A.n_strips; ++s) {
            const int v = A.cand_v[s];
            const uint32_t ci = A.cand_i[s], cj = A.cand_j[s];
            if (v > bv || (v == bv && v > 0 && (ci < bi || (ci == bi && cj < bj)))) { bv = v; bi = ci; bj = cj; }
        }
        H0 = bv / SCALE; i0 = (int)bi; j0 = (int)bj;
        *A.score = H0;
        A.res[3] = (uint64_t)i0 * (uint64_t)(n + 1) + (uint64_t)j0;
    } else if (A.start_given) {
        H0 = 0; i0 = A.start_row; j0 = A.start_col_dev ? *A.start_col_dev : n;
        A.res[3] = 0;
    } else {
        H0 = *A.score; i0 = m; j0 = n;
        A.res[3] = 0;
    }
    A.st->exit_row = 0;
    A.st->i0 = i0; A.st->j0 = j0; A.st->H0 = H0;
    A.st->s0 = i0 > 0 ? (i0 - 1) / A.Lay.ROWS : 0;
    A.st->cut_seg = -1;
    A.st->total_len = 0;
    A.st->fallbacks = 0;
    A.st->identity = 0;
    A.st->gaps = 0;
}

// First candidate index of line s: the band is centred on the straight-line prediction of the
// crossing (start cell, slope); a crossing outside the band only costs a serial fallback segment.
__device__ __forceinline__ int tb_band_q0(const TbArgs &A, const int s)
{
    // local alignments: the slope of the line from the end cell to the origin (clamped) -- exact for an alignment that
    // spans the matrix, harmless for a short one; with slope 1 a 100 k local alignment drifted 5 k columns out of
    // the band and was walked serially (19.5 ms instead of 1.2)
    const double slope = A.local ? fmin(1.25, fmax(0.8, (double)A.st->j0 / fmax(1.0, (double)A.st->i0))) : A.slope;
    const double pred = (double)A.st->j0 - (double)(A.st->i0 - s * A.Lay.ROWS) * slope;
    const int qc = (int)(fmax(0.0, fmin(pred, (double)A.Lay.n)) / (double)A.Wd);
    return max(0, min(qc - A.BQ, A.Q - 2 * A.BQ));
}

// ---- A: speculative walkers ----------------------------------------------------------------------
__global__ void __launch_bounds__(128) tb_walkers_kernel(const TbArgs A)
{
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = 2 * A.BQ + 1;
    const long long total = (long long)A.n_strips * nb;
    if (id >= total) return;
    const int s = (int)(id / nb), qi = (int)(id % nb);
    if (s == 0 || s > A.st->s0) return;          // lines above the start cell only
    const int q = tb_band_q0(A, s) + qi;
    if (q > A.Q) return;
    // (row chunk of a checkpointed traceback: the chunk is only filled up to the column where the path enters it)
    const int c = min(q * A.Wd, A.start_col_dev ? A.st->j0 : A.Lay.n);
    A.fa[(size_t)s * nb + qi] = (uint32_t)tb_walk_to_row(A.Lay, s * A.Lay.ROWS, c, (s - 1) * A.Lay.ROWS, 3 * A.Lay.ROWS + 64);
}

// ---- B: chain the lines ---------------------------------------------------------------------------
__global__ void tb_resolve_kernel(const TbArgs A)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int ROWS = A.Lay.ROWS, nb = 2 * A.BQ + 1;
    const int s0 = A.st->s0;
    int x = tb_walk_to_row(A.Lay, A.st->i0, A.st->j0, s0 * ROWS);
    A.X[s0] = x;
    unsigned long long fb = 0;
    for (int s = s0; s >= 1; --s) {
        int nx;
        if (x == 0) nx = 0;
        else {
            const int q = x / A.Wd, q0 = tb_band_q0(A, s);
            const int clo = min(q * A.Wd, A.start_col_dev ? A.st->j0 : A.Lay.n);
            const bool exact = clo == x;
            if (q < q0 || q + (exact ? 0 : 1) > min(q0 + 2 * A.BQ, A.Q)) {
                nx = tb_walk_to_row(A.Lay, s * ROWS, x, (s - 1) * ROWS); ++fb;      // outside the band
            } else {
                const uint32_t lo = A.fa[(size_t)s * nb + (q - q0)];
                const uint32_t hi = exact ? lo : A.fa[(size_t)s * nb + (q + 1 - q0)];
                if (lo == hi && lo != (uint32_t)TB_UNKNOWN) nx = (int)lo;     // exact hit, or sandwiched between two merged paths
                else { nx = tb_walk_to_row(A.Lay, s * ROWS, x, (s - 1) * ROWS); ++fb; }
            }
        }
        x = nx;
        A.X[s - 1] = x;
    }
    A.st->fallbacks = fb;
}

// One segment: from its entry cell until row b_t (segment 0: until (0,0) for global alignments).
// MODE 0: count steps / score changes.  MODE 1: emit characters at out[cap - 1 - (off + k)].
template <int MODE>
__device__ __forceinline__ void tb_segment(const TbArgs &A, const int t)
{
    const TbLayout &L = A.Lay;
    const int s0 = A.st->s0;
    if (t > s0) return;
    int i = (t == s0) ? A.st->i0 : (t + 1) * L.ROWS;
    int j = (t == s0) ? A.st->j0 : A.X[t + 1];
    const int stop_row = t * L.ROWS;
    TbCursor cur;
    unsigned long long len = 0, nIdent = 0, nGap = 0;
    long long delta = 0, dmin = 0;
    long long budget = 0;                       // MODE 1, local: steps this segment may still emit
    unsigned long long off = 0;
    if (MODE == 1) {
        off = A.seg_off[t];
        if (A.local) {
            const int cut = A.st->cut_seg;
            if (cut >= 0 && t < cut) return;                     // beyond the end of the local alignment
            budget = (cut >= 0 && t == cut) ? (long long)A.seg_len[t] : -1;   // seg_len[cut] was trimmed in D
        } else budget = -1;
    }
    const char GAPC = A.alphabet[A.alpha];
    char *oT = A.out_text + A.cap, *oP = A.out_pattern + A.cap;
    while (true) {
        if (A.local) { if (i <= stop_row || i == 0 || j == 0) break; }
        else if (A.slice) { if (j == 0 || (t > 0 && i <= stop_row)) break; }      // row 0 runs LEFT to the slice edge
        else if (A.chunk_top) { if (i <= stop_row) break; }                        // row chunk: ends on its top row
        else if (t == 0) { if (i == 0 && j == 0) break; }
        else if (i <= stop_row) break;
        if (MODE == 1 && budget == 0) break;
        const int tag = tb_tag(L, cur, i, j);
        const bool takeT = tag != TAG_TOP, takeP = tag != TAG_LEFT;
        if (MODE == 0) {
            if (A.local) {
                delta += (tag == TAG_DIAG) ? -(long long)A.S[A.pattern[i - 1] * A.alpha + A.text[j - 1]] : (long long)A.gap;
                dmin = min(dmin, delta);
            }
        } else {
            const unsigned long long pos = off + len + 1;
            oT[-(long long)pos] = takeT ? A.alphabet[A.text[j - 1]] : GAPC;
            oP[-(long long)pos] = takeP ? A.alphabet[A.pattern[i - 1]] : GAPC;
            if (takeT && takeP) nIdent += A.text[j - 1] == A.pattern[i - 1]; else ++nGap;
            if (budget > 0) --budget;
        }
        ++len;
        i -= takeP; j -= takeT;
    }
    if (MODE == 1 && len) {
        atomicAdd(&A.st->identity, nIdent);
        atomicAdd(&A.st->gaps, nGap);
    }
    if (MODE == 0) {
        A.seg_len[t] = len;
        A.seg_delta[t] = delta;
        A.seg_min[t] = dmin;
        if (A.slice && j == 0 && len > 0) A.st->exit_row = i;     // exactly one segment steps onto the left edge
    }
}

__global__ void __launch_bounds__(64) tb_count_kernel(const TbArgs A)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (int)A.n_strips) tb_segment<0>(A, t);
}

__global__ void __launch_bounds__(64) tb_emit_kernel(const TbArgs A)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (int)A.n_strips) tb_segment<1>(A, t);
}

// ---- D: offsets, local cut, result fields -----------------------------------------------------
__global__ void tb_offsets_kernel(const TbArgs A)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const TbLayout &L = A.Lay;
    const int s0 = A.st->s0;
    if (A.local && A.st->H0 <= 0) {
        // best score 0: nothing is emitted and the int(-1) indices wrap (alignSequenceCPU.cpp:13-14,56-57)
        for (int t = s0; t >= 0; --t) { A.seg_off[t] = 0; A.seg_len[t] = 0; }
        A.st->cut_seg = s0;
        A.st->total_len = 0;
        A.res[0] = 0; A.res[1] = ~0ull; A.res[2] = ~0ull;
        return;
    }
    unsigned long long off = A.global_off ? *A.global_off : 0ull;
    long long H = A.st->H0;
    int cut = -1;
    int ei = 0, ej = 0;                 // cell reached after the last emitted step
    int li = 0, lj = 0;                 // last emitted cell
    bool haveExit = false;
    for (int t = s0; t >= 0; --t) {
        A.seg_off[t] = off;
        if (!A.local) { off += A.seg_len[t]; continue; }
        if (H + A.seg_min[t] > 0) {     // every cell of the segment has H > 0, and so has the next entry cell
            off += A.seg_len[t];
            H += A.seg_delta[t];
            continue;
        }
        // the running score reaches 0 inside (or right at the end of) this segment: replay it
        int i = (t == s0) ? A.st->i0 : (t + 1) * L.ROWS;
        int j = (t == s0) ? A.st->j0 : A.X[t + 1];
        TbCursor cur;
        unsigned long long k = 0;
        while (H > 0 && i > t * L.ROWS && j > 0) {
            const int tag = tb_tag(L, cur, i, j);
            H += (tag == TAG_DIAG) ? -(long long)A.S[A.pattern[i - 1] * A.alpha + A.text[j - 1]] : (long long)A.gap;
            li = i; lj = j;
            i -= (tag != TAG_LEFT); j -= (tag != TAG_TOP);
            ++k;
        }
        A.seg_len[t] = k;
        off += k;
        cut = t;
        ei = i; ej = j;
        haveExit = k > 0;
        break;
    }
    A.st->cut_seg = cut;
    A.st->total_len = off;
    A.res[0] = off;
    if (A.global_off) *A.global_off = off;
    if (A.exit_col_dev) *A.exit_col_dev = A.X[0];
    if (!A.local && A.start_given) { A.res[1] = (uint64_t)A.st->exit_row; A.res[2] = 0; A.res[3] = 0; return; }
    if (!A.local) { A.res[1] = 0; A.res[2] = 0; return; }      // clamped indices end at 0 (alignSequenceCPU.cpp:100-101)
    if (!haveExit) {
        // the alignment ran into the matrix border, or ended exactly on a strip line: the last
        // emitted cell is the last step of the last non-empty segment at or above the cut
        int t = cut >= 0 ? cut : 0;
        while (t < s0 && A.seg_len[t] == 0) ++t;
        int i = (t == s0) ? A.st->i0 : (t + 1) * L.ROWS;
        int j = (t == s0) ? A.st->j0 : A.X[t + 1];
        TbCursor cur;
        for (unsigned long long k = 0; k < A.seg_len[t]; ++k) {
            const int tag = tb_tag(L, cur, i, j);
            li = i; lj = j;
            i -= (tag != TAG_LEFT); j -= (tag != TAG_TOP);
        }
        ei = i; ej = j;
    }
    // start indices of traceBackSW (:45-57): leaving through the border keeps the indices of the last
    // emitted cell (break before the update), an interior STOP cell contributes its own indices
    if (ei == 0 || ej == 0) { A.res[1] = (uint64_t)(int64_t)(lj - 1); A.res[2] = (uint64_t)(int64_t)(li - 1); }
    else { A.res[1] = (uint64_t)(int64_t)(ej - 1); A.res[2] = (uint64_t)(int64_t)(ei - 1); }
}

} // namespace sa

// sa_batch.cuh -- batch of independent short pairs (BASELINE config 4): fill + traceback kernels.
//
// Mapping.  A sub-warp GROUP of L lanes owns one pair; lane l of the group owns
// the R consecutive DP rows  l*R+1 .. l*R+R  (pattern residues l*R .. l*R+R-1) and
// sweeps the columns as a systolic wavefront: at step k lane l processes text
// index k-l.  The bottom value of lane l-1 reaches lane l with one __shfl_up per
// step.  A warp holds G = 32/L groups that run in lock-step.
//
// Shared memory per group: the pair's query profile  prof[a][row] = 4*S[p_row][a]
// as bytes (alpha x L*Rpad), and the text residues.  A lane fetches the 16-row
// slices of its profile row with LDS.128 and selects bytes with IDP.4A.
//
// Direction matrix: packed 2-bit tags, written in WARP-STEP-MAJOR order so that
// every store instruction is a fully coalesced 128-byte line:
//     word(task, kb, w, lane)  at  task*stride + (kb*NW + w)*32 + lane
// where kb = step / CB (CB steps are accumulated per store block), NW = R*CB/16
// and bit (kk*R + r)*2 of the block belongs to step kb*CB+kk, row r of the lane.
// 0.25 B per cell, no write amplification (see DESIGN.md "direction layout").
#pragma once
#include "sa_cell.cuh"
#include "../../include/sa_b200.h"

namespace sa {

// Pairs are binned on the device by pattern length into CLASSES (one (R, L) instantiation each,
// rows covered = R*L) and ordered by descending text length inside a class, so that the lanes of a
// group are filled and the groups of a warp finish together.  The host never sees the bins: every
// class kernel reads its own range from BatchClassDyn.
constexpr int MAX_CLASSES = 16;
constexpr int SORT_BUCKETS = 512;

struct BatchClassTable {            // static description, passed by value
    int n_classes;
    int R[MAX_CLASSES], L[MAX_CLASSES];
    int packed[MAX_CLASSES];        // 1: s16x2 kernel, two pairs per lane group (sa_batch16.cuh); 2: same, straight-line
                                    //    kernel with the octet direction layout (sa_batch16_sw.cuh)
    uint32_t max_rows[MAX_CLASSES];
    unsigned long long stride[MAX_CLASSES];     // direction words per task
    uint32_t max_text;              // pairs with a longer text are skipped (host aligns them one by one)
    int bucket_shift;               // text length -> bucket
};

struct BatchClassDyn {              // device-computed per chunk
    uint32_t first;                 // first position of the class in `order`
    uint32_t count;                 // pairs in the class
    unsigned long long dir_base;    // word offset of the class' direction region
};

struct BatchArgs {
    const uint8_t *text;      const int64_t *text_off;
    const uint8_t *pattern;   const int64_t *pattern_off;
    const uint32_t *order;    // position -> pair index
    const BatchClassDyn *dyn; // this launch's class: positions [first, first+count)
    uint32_t *dirs;           // direction workspace of the chunk
    uint64_t task_stride;     // words per task
    int32_t *score;           // per pair
    uint32_t *end_i;          // per pair: SW arg-max row (DP coordinates), NW: m
    uint32_t *end_j;          // per pair: SW arg-max column,               NW: n
    const int8_t *S4;         // 32x32 bytes: 4*S[p][t] at [p*32+t]
    int alpha;
    int gap;
    uint32_t max_n;           // max text length of this launch (sizes the text stage)
    uint4 *snap_ws;           // SW: per-warp arg-max snapshot area in global memory (L2-resident)
};


template <int R, int L, bool LOCAL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) batch_fill_kernel(const BatchArgs A)
{
    static_assert(R % 2 == 0, "R must be even");
    static_assert(L == 4 || L == 8 || L == 16 || L == 32, "L must divide 32");
    constexpr int G = 32 / L;
    constexpr int CB = cb_for(R);
    constexpr int NW = R * CB / 16;
    constexpr int RPAD = rpad_for(R);
    constexpr int PS = L * RPAD;            // profile stride per letter (bytes)
    constexpr int NPW = (R + 3) / 4;        // profile words per lane per column

    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / L, l = lane % L;
    const int alpha = A.alpha;
    const uint32_t textPad = (A.max_n + 15u) & ~15u;
    const uint32_t groupBytes = (alpha * PS + textPad + 15u) & ~15u;
    constexpr uint32_t snapBytes = LOCAL ? 16 * G : 0;                   // per-group arg-max bound

    int8_t *S4s = reinterpret_cast<int8_t *>(smem);                          // alpha x 32
    unsigned char *wbase = smem + 32 * MAX_ALPHA + (size_t)warp * (G * groupBytes + snapBytes);
    unsigned char *profS = wbase + (size_t)g * groupBytes;
    unsigned char *textS = profS + alpha * PS;
    // the snapshots live in global memory: written on records only, read back by the same lane
    uint4 *snap = A.snap_ws + (size_t)(blockIdx.x * WARPS + warp) * (((R + 3) / 4) * 32);
    int *gmS = reinterpret_cast<int *>(wbase + G * groupBytes) + 4 * g;

    for (int i = threadIdx.x; i < 32 * MAX_ALPHA; i += blockDim.x) S4s[i] = A.S4[i];
    __syncthreads();

    const int KL = 2 - SCALE * A.gap, KT = 1 - SCALE * A.gap;
    const uint32_t first_pos = A.dyn->first, n_pos = A.dyn->count;
    uint32_t *const dirs = A.dirs + A.dyn->dir_base;
    const uint32_t nTasks = (n_pos + G - 1) / G;

    for (uint32_t task = blockIdx.x * WARPS + warp; task < nTasks; task += gridDim.x * WARPS) {
        const uint32_t pos = task * G + g;
        const bool valid = pos < n_pos;
        uint32_t pair = 0; int n = 0, m = 0;
        const uint8_t *tx = nullptr, *pt = nullptr;
        if (valid) {
            pair = A.order[first_pos + pos];
            const int64_t t0 = A.text_off[pair], p0 = A.pattern_off[pair];
            n = (int)(A.text_off[pair + 1] - t0);
            m = (int)(A.pattern_off[pair + 1] - p0);
            tx = A.text + t0; pt = A.pattern + p0;
        }
        __syncwarp();
        // ---- stage text and build the query profile ----
        if (LOCAL && l == 0) *gmS = 0;
        for (int j = l; j < n; j += L) textS[j] = tx[j];
        for (int i = l; i < L * R; i += L) {
            // row i of the pair lives at byte (i/R)*RPAD + i%R of every letter's profile row
            const int off = (i / R) * RPAD + (i % R);
            if (i < m) {
                const int8_t *srow = S4s + 32 * min((int)pt[i], alpha - 1);
                for (int a = 0; a < alpha; ++a) profS[a * PS + off] = (unsigned char)srow[a];
            } else {
                // padding rows never beat the row above them (keeps the SW arg-max filter quiet)
                for (int a = 0; a < alpha; ++a) profS[a * PS + off] = (unsigned char)0x80;
            }
        }
        __syncwarp();

        // ---- boundary state ----
        int c[R];
#pragma unroll
        for (int r = 0; r < R; ++r) c[r] = LOCAL ? 0 : -SCALE * A.gap * (l * R + r + 1);   // H(i,0)
        int prevTop = LOCAL ? 0 : -SCALE * A.gap * (l * R);                                 // H(i0-1,0)
        int bottom = 0;
        int bestv = 0, besti = 0, bestj = 0;   // SW arg-max (4*H, DP row, DP column); row resolved at the end

        int nmax = n;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
        const int nSteps = nmax + L - 1;
        uint32_t *dbase = dirs + (size_t)task * A.task_stride + lane;

        // the profile words of a step are fetched one step ahead (text -> profile is two dependent
        // shared-memory loads that would otherwise sit in front of every column sweep)
        uint32_t profN[NPW];
#pragma unroll
        for (int q = 0; q < NPW; ++q) profN[q] = 0;
        if (valid && l == 0 && n > 0) load_profile_words<R>(profS + min((int)textS[0], alpha - 1) * PS, profN);

        for (int kb = 0; kb * CB < nSteps; ++kb) {
            uint32_t acc[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) acc[w] = 0;
#pragma unroll
            for (int kk = 0; kk < CB; ++kk) {
                const int jt = kb * CB + kk - l;                 // text index of this lane at this step
                const int up = __shfl_up_sync(0xffffffffu, bottom, 1);
                uint32_t prof[NPW];
#pragma unroll
                for (int q = 0; q < NPW; ++q) prof[q] = profN[q];
                if (valid && jt + 1 >= 0 && jt + 1 < n)
                    load_profile_words<R>(profS + min((int)textS[jt + 1], alpha - 1) * PS + l * RPAD, profN);
                int floorv = 0;
                if (LOCAL) floorv = *reinterpret_cast<volatile int *>(gmS);
                if (valid && jt >= 0 && jt < n) {
                    const int top = (l == 0) ? (LOCAL ? 0 : -SCALE * A.gap * (jt + 1)) : up;
                    int bmax[nblk_for(R)];
                    sweep_column<R, LOCAL, NW>(c, top, prevTop, prof, KL, KT, acc, 2 * R * kk, bmax);
                    prevTop = top;
                    bottom = c[R - 1];
                    if (LOCAL) {
                        const int colmax = max_of_blocks(bmax);
                        if (l * R < m && track_argmax<R>(c, colmax, jt + 1, snap, lane, floorv, bestv, bestj))
                            atomicMax(gmS, colmax);
                    }
                }
            }
#pragma unroll
            for (int w = 0; w < NW; ++w) dbase[(size_t)(kb * NW + w) * 32] = acc[w];
        }

        // ---- per-pair result of the fill ----
        if (LOCAL) {
            besti = bestv > 0 ? l * R + snapshot_first_row<R>(snap, lane, bestv) + 1 : 0;
#pragma unroll
            for (int o = L / 2; o >= 1; o >>= 1) {
                const int ov = __shfl_xor_sync(0xffffffffu, bestv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
                const bool take = ov > bestv || (ov == bestv && (oi < besti || (oi == besti && oj < bestj)));
                if (take) { bestv = ov; besti = oi; bestj = oj; }
            }
            if (valid && l == 0) {
                A.score[pair] = bestv / SCALE;
                A.end_i[pair] = bestv > 0 ? besti : 0;
                A.end_j[pair] = bestv > 0 ? bestj : 0;
            }
        } else {
            if (valid && m >= 1 && l == (m - 1) / R) {
                const int rm = (m - 1) % R;
                int v = c[0];
#pragma unroll
                for (int r = 1; r < R; ++r) v = (r == rm) ? c[r] : v;
                A.score[pair] = v / SCALE;      // exact: c is a multiple of 4
                A.end_i[pair] = m;
                A.end_j[pair] = n;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Device-side binning: counting sort on key = class * SORT_BUCKETS + bucket(descending text length).
struct BatchSortArgs {
    const int64_t *text_off;  const int64_t *pattern_off;
    uint32_t base, count;             // pairs [base, base+count) of the batch
    BatchClassTable table;
    uint32_t *hist;                   // MAX_CLASSES * SORT_BUCKETS counters (zeroed by the host)
    uint32_t *key;                    // count entries
    uint32_t *order;                  // count entries (output)
    BatchClassDyn *dyn;               // n_classes + 1 entries (output)
    // pairs no class takes (empty side, text above table.max_text, pattern above the top class) get a sentinel
    // result {score INT32_MIN, aln_len 0} instead of stale memory; the host path aligns such members afterwards
    sa_result *skipped_results;  unsigned long long *skipped_alnoff;
};

__global__ void __launch_bounds__(256) batch_classify_kernel(const BatchSortArgs A)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.count) return;
    const uint32_t p = A.base + i;
    const uint32_t n = (uint32_t)(A.text_off[p + 1] - A.text_off[p]);
    const uint32_t m = (uint32_t)(A.pattern_off[p + 1] - A.pattern_off[p]);
    uint32_t key = 0xffffffffu;
    if (n >= 1 && m >= 1 && n <= A.table.max_text) {
        for (int c = 0; c < A.table.n_classes; ++c)
            if (m <= A.table.max_rows[c]) {
                const uint32_t b = min(n >> A.table.bucket_shift, (uint32_t)SORT_BUCKETS - 1);
                key = c * SORT_BUCKETS + (SORT_BUCKETS - 1 - b);
                break;
            }
    }
    A.key[i] = key;
    if (key != 0xffffffffu) atomicAdd(&A.hist[key], 1u);
    else if (A.skipped_results) {
        sa_result r; r.score = (int32_t)0x80000000; r.aln_len = 0; r.start_text = 0; r.start_pattern = 0;
        A.skipped_results[p] = r;
        A.skipped_alnoff[p] = 0;
    }
}

__global__ void __launch_bounds__(1024) batch_scan_kernel(const BatchSortArgs A)
{
    // one block: exclusive scan of the (at most 8192) counters, then the per-class ranges
    __shared__ uint32_t part[1024];
    const int nKeys = A.table.n_classes * SORT_BUCKETS;
    const int per = (nKeys + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(lo + per, nKeys);
    uint32_t sum = 0;
    for (int k = lo; k < hi; ++k) sum += A.hist[k];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const uint32_t v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = threadIdx.x ? part[threadIdx.x - 1] : 0;
    for (int k = lo; k < hi; ++k) { const uint32_t c = A.hist[k]; A.hist[k] = run; run += c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long dirBase = 0;
        for (int c = 0; c < A.table.n_classes; ++c) {
            const uint32_t first = A.hist[c * SORT_BUCKETS];
            const uint32_t end = (c + 1 < A.table.n_classes) ? A.hist[(c + 1) * SORT_BUCKETS] : part[1023];
            A.dyn[c].first = first;
            A.dyn[c].count = end - first;
            A.dyn[c].dir_base = dirBase;
            const uint32_t G = (32 / A.table.L[c]) * (A.table.packed[c] ? 2 : 1);      // pairs per task
            dirBase += (unsigned long long)((end - first + G - 1) / G) * A.table.stride[c];
        }
        A.dyn[A.table.n_classes].first = part[1023];
        A.dyn[A.table.n_classes].count = 0;
        A.dyn[A.table.n_classes].dir_base = dirBase;
    }
}

__global__ void __launch_bounds__(256) batch_scatter_kernel(const BatchSortArgs A)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.count) return;
    const uint32_t key = A.key[i];
    if (key == 0xffffffffu) return;
    const uint32_t pos = atomicAdd(&A.hist[key], 1u);
    A.order[pos] = A.base + i;
}

// ---------------------------------------------------------------------------------------------
// Traceback + alignment-string emission for the batch: one thread per pair, pointer-chasing the
// packed tags (L2-resident right after the fill).  Reproduces traceBackNW / traceBackSW
// (alignSequenceCPU.cpp:64-114 / :10-62) including the border override, the clamped indices,
// the break-before-update quirk and the uint64 wrap of a zero-score local alignment.
// Strings are written BACKWARDS into the pair's slot so that no reverse pass is needed: pair p
// owns bytes [slot, slot+cap) with slot = text_off[p]+pattern_off[p], cap = n+m, and its
// alignment ends at slot+cap; aln_off[p] = slot + cap - len.
struct BatchTraceArgs {
    const uint8_t *text;      const int64_t *text_off;
    const uint8_t *pattern;   const int64_t *pattern_off;
    const uint32_t *order;
    const BatchClassDyn *dyn;      // n_classes entries + one trailing entry whose .first is the total
    BatchClassTable table;
    const uint32_t *dirs;
    const int32_t *score;     const uint32_t *end_i;  const uint32_t *end_j;
    const int32_t *S;         // alpha x alpha, [pattern][text]
    int alpha;  int gap;  int local;
    char alphabet[MAX_ALPHA + 1];
    // outputs (per pair)
    sa_result *results;  uint64_t *aln_off;
    char *out_text;  char *out_pattern;
    // optional (may be nullptr): per pair {identity, gaps} -- the two counts of prettyAlignmentPrint
    // (utilities.cpp:262-283: columns with equal letters / columns with a gap), taken while the strings are emitted
    uint32_t *stats;
};

// shared-memory bytes by 32-bit shared address (the generic form re-derives the shared window in every iteration)
__device__ __forceinline__ uint32_t tb_lds_u8(uint32_t a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int tb_lds_s8(uint32_t a) { int v; asm("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

// Residues read at a moving index, an aligned word at a time: the byte comes out with one PRMT.
struct TbWordReader {
    const uint32_t *base; int off; int curw; uint32_t w;
    __device__ TbWordReader(const uint8_t *p) : base(reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3)),
                                                 off((int)(reinterpret_cast<uintptr_t>(p) & 3)), curw(-1), w(0) {}
    __device__ __forceinline__ int get(const int q) {
        const int a = q + off, wi = a >> 2;
        if (wi != curw) { curw = wi; w = __ldg(base + wi); }
        return (int)__byte_perm(w, 0u, 0x4440u | (uint32_t)(a & 3));
    }
};

// Characters written BACKWARDS from `end`, four at a time and at the same steps in every lane of the warp (the lanes
// walk in lock-step, so `if ((len & 3) == 0)` is a uniform branch and the store code is issued on every fourth step
// only).  `end` has any alignment a = end & 3: the group of four newest characters lies at end - len (address & 3 == a),
// the aligned word stored at a flush is made of the older a characters of this group and the newest 4 - a of the one
// before (a funnel shift); the a bytes right below `end` go out as bytes at the first flush, and finish() writes what
// is still pending (the low 4 - a bytes of the last group -- all four when a = 0 -- and the len & 3 characters after the
// last flush).
struct TbBackWriter {
    char *end; uint32_t a, cur, prev;
    __device__ TbBackWriter(char *e) : end(e), a((uint32_t)(reinterpret_cast<uintptr_t>(e) & 3)), cur(0u), prev(0u) {}
    __device__ __forceinline__ void push(const uint32_t c) { cur = (cur << 8) | c; }       // byte 0 = newest = lowest address
    // len = 4, 8, ...: the same path in every lane (an aligned `end` is handled as a = 0 with the word one group late:
    // the clamped funnel shift by 32 bits returns the group before, so no lane takes a branch of its own)
    __device__ __forceinline__ void flush(const uint32_t len) {
        char *q = end - len;
        if (len == 4u) {
#pragma unroll
            for (uint32_t b = 1; b < 4u; ++b) if (b >= 4u - a) q[b] = (char)(cur >> (8u * b));
        } else *reinterpret_cast<uint32_t *>(q + (4u - a)) = __funnelshift_rc(cur, prev, 8u * (4u - a));
        prev = cur;
    }
    __device__ __forceinline__ void finish(const uint32_t len) {
        const uint32_t rem = len & 3u, full = len - rem;
        if (full >= 4u) {                  // the low 4 - a bytes of the last group
            char *q = end - full;
            if (a == 0u) *reinterpret_cast<uint32_t *>(q) = prev;
            else for (uint32_t b = 0; b < 4u - a; ++b) q[b] = (char)(prev >> (8u * b));
        }
        char *q = end - len;
        for (uint32_t b = 0; b < rem; ++b) q[b] = (char)(cur >> (8u * b));
    }
};

__global__ void __launch_bounds__(256) batch_traceback_kernel(const BatchTraceArgs A)
{
    // Alphabet (+ gap character at index alpha) and the score matrix as bytes in shared memory: per-lane indices
    // differ, and a divergent index into the constant bank (kernel parameters) would serialise.  ~1.1 KB per block,
    // so that the block still fits next to the fill blocks of the next chunk.
    __shared__ unsigned char alphS[MAX_ALPHA + 1];
    __shared__ signed char S8[MAX_ALPHA * MAX_ALPHA];
    for (int i = threadIdx.x; i <= A.alpha; i += blockDim.x) alphS[i] = (unsigned char)A.alphabet[i];
    for (int i = threadIdx.x; i < A.alpha * A.alpha; i += blockDim.x) S8[i] = (signed char)A.S[i];      // |S| <= 31 (upload_scoring)
    __syncthreads();
    const int alpha = A.alpha, gap = A.gap;
    uint32_t sAlph = (uint32_t)__cvta_generic_to_shared(alphS), sS8 = (uint32_t)__cvta_generic_to_shared(S8);
    asm volatile("" : "+r"(sAlph), "+r"(sS8));        // kept in registers: the compiler would re-derive the shared window per step
    // grid-stride over the pairs: a pipelined chunk launches only a few blocks per SM so that the fill blocks of the
    // next chunk fit next to them (a full grid would hold every SM until the traceback drains)
    for (uint32_t gpos = blockIdx.x * blockDim.x + threadIdx.x; gpos < A.dyn[A.table.n_classes].first; gpos += gridDim.x * blockDim.x) {
        int cls = 0;
        while (gpos >= A.dyn[cls].first + A.dyn[cls].count) ++cls;
        const int cR = A.table.R[cls], cL = A.table.L[cls];
        const int layout = A.table.packed[cls];                  // 0: s32, 1: s16x2 warp-step-major, 2: s16x2 octet layout
        const bool packed = layout != 0;
        const int cCB = layout == 1 ? ((cR % 8 == 0) ? 1 : (cR % 4 == 0) ? 2 : 4) : cb_for(cR);
        const int G = 32 / cL;
        const uint32_t pos = gpos - A.dyn[cls].first;
        const uint32_t pair = A.order[gpos];
        const uint32_t unit = packed ? pos >> 1 : pos;
        const int halfBit = packed ? (int)(pos & 1) * 16 : 0;
        const uint32_t task = unit / G, g = unit % G;
        const int64_t t0 = A.text_off[pair], p0 = A.pattern_off[pair];
        const int n = (int)(A.text_off[pair + 1] - t0), m = (int)(A.pattern_off[pair + 1] - p0);
        const uint8_t *tx = A.text + t0, *pt = A.pattern + p0;
        const uint32_t *dbase = A.dirs + A.dyn[cls].dir_base + (size_t)task * A.table.stride[cls] + g * cL;
        const uint64_t slotEnd = (uint64_t)(t0 + p0) + (uint64_t)(n + m);

        // One formula for the three direction layouts: step k = column-1 + lane, block kb = k >> kbShift, step in
        // block kk, cell = kk*R + r; the word of a cell is kb*KBS + lane*LS + (cell >> cs) << csh (32-bit: a task has
        // fewer than 2^32 words).
        const int cs = packed ? 3 : 4;
        const int kbShift = layout == 2 ? 3 : (cCB == 1 ? 0 : cCB == 2 ? 1 : cCB == 4 ? 2 : 3);
        const int kkMask = (1 << kbShift) - 1;
        const int PO = (cR + 1) & ~1;                                                     // po_for(R)
        const uint32_t KBS = layout == 2 ? 32u * PO : (uint32_t)((cR * cCB) >> cs) * 32u;
        const uint32_t LS = layout == 2 ? (uint32_t)PO : 1u;
        const int csh = layout == 2 ? 0 : 5;
        const int cmask = (1 << cs) - 1;

        int i = (int)A.end_i[pair], j = (int)A.end_j[pair];
        int H = A.local ? A.score[pair] : 0;       // global: accumulated along the path (the fill does not report it)
        uint32_t len = 0, nDiag = 0, nIdent = 0;
        uint32_t cachedIdx = ~0u, cachedWord = 0;
        // (lane, row-in-lane) of DP row i, kept incrementally: no integer division in the walk
        int ll = i > 0 ? (i - 1) / cR : 0, r = i > 0 ? (i - 1) % cR : 0;
        uint32_t llLS = (uint32_t)ll * LS;
        auto fetch = [&](const int jj) -> int {
            const int k = (jj - 1) + ll;
            const int cell = (k & kkMask) * cR + r;
            const uint32_t idx = (uint32_t)(k >> kbShift) * KBS + llLS + ((uint32_t)(cell >> cs) << csh);
            if (idx != cachedIdx) { cachedIdx = idx; cachedWord = __ldg(dbase + idx); }
            return (cachedWord >> (2 * (cell & cmask) + halfBit)) & 3;
        };
        auto row_up = [&]() { --i; if (r == 0) { r = cR - 1; --ll; llLS -= LS; } else --r; };

        TbWordReader rdT(tx), rdP(pt);
        TbBackWriter wT(A.out_text + slotEnd), wP(A.out_pattern + slotEnd);
        auto emit = [&](const uint32_t cT, const uint32_t cP) {
            wT.push(cT); wP.push(cP);
            ++len;
            if ((len & 3u) == 0u) { wT.flush(len); wP.flush(len); }
        };

        int ti, pi;
        if (!A.local) {
            ti = n - 1; pi = m - 1;
            while (i > 0 || j > 0) {
                int tag;
                if (j == 0) tag = TAG_TOP;             // alignSequenceCPU.cpp:78-79
                else if (i == 0) tag = TAG_LEFT;       // :80-81
                else tag = fetch(j);
                const bool takeT = tag != TAG_TOP, takeP = tag != TAG_LEFT;
                const int ct = rdT.get(ti), cp = rdP.get(pi);
                emit(tb_lds_u8(sAlph + (takeT ? ct : alpha)), tb_lds_u8(sAlph + (takeP ? cp : alpha)));
                H += (tag == TAG_DIAG) ? tb_lds_s8(sS8 + cp * alpha + ct) : -gap;      // the path's score is H(m, n)
                nDiag += (tag == TAG_DIAG); nIdent += (tag == TAG_DIAG && ct == cp);
                ti = max(0, ti - (int)takeT);
                pi = max(0, pi - (int)takeP);
                if (takeP) row_up();
                j -= takeT;
            }
        } else {
            // :13-14: ti = j-1, pi = i-1 (-1/-1 when the best score is 0).  Inside the matrix the clamped indices of the
            // reference stay at j-1 / i-1, so they are not carried: the residues are read at j-1 / i-1, and the break before
            // the index update (:45-46) leaves them one step behind in the direction(s) taken last (adjT / adjP).
            int adjT = 0, adjP = 0;
            while (H > 0) {                            // H(i,j) == 0  <=>  reference STOP
                const int tag = fetch(j);
                const bool takeT = tag != TAG_TOP, takeP = tag != TAG_LEFT;
                const int ct = rdT.get(j - 1), cp = rdP.get(i - 1);
                emit(tb_lds_u8(sAlph + (takeT ? ct : alpha)), tb_lds_u8(sAlph + (takeP ? cp : alpha)));
                H += (tag == TAG_DIAG) ? -tb_lds_s8(sS8 + cp * alpha + ct) : gap;
                nDiag += (tag == TAG_DIAG); nIdent += (tag == TAG_DIAG && ct == cp);
                if (takeP) row_up();
                j -= takeT;
                if (i == 0 || j == 0) { adjT = (int)takeT; adjP = (int)takeP; break; }
            }
            ti = j - 1 + adjT; pi = i - 1 + adjP;
        }
        wT.finish(len); wP.finish(len);
        sa_result res;
        res.score = A.local ? A.score[pair] : H;
        res.aln_len = len;
        res.start_text = (uint64_t)(int64_t)ti;
        res.start_pattern = (uint64_t)(int64_t)pi;
        A.results[pair] = res;
        A.aln_off[pair] = slotEnd - len;
        if (A.stats) { A.stats[2 * (size_t)pair] = nIdent; A.stats[2 * (size_t)pair + 1] = len - nDiag; }
    }
}

} // namespace sa

// sa_api.cu -- extern "C" layer (include/sa_b200.h) over the CUDA kernels.
// Host orchestration that replaces alignSequenceGPU.cu:356-653 (initMemory, the
// per-band launch loop, the D2H drain and the CPU traceback).  No CPU fallback:
// every compute entry fails with SA_ERR_NO_DEVICE when there is no GPU.
#include "../../include/sa_b200.h"
#include "sa_batch.cuh"
#include "sa_batch16.cuh"
#include "sa_batch16_sw.cuh"
#include "sa_long.cuh"
#include "sa_tile_host.h"
#include "sa_traceback.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

using namespace sa;

namespace {

#define SA_TRY(expr, code)                                        \
    do {                                                          \
        cudaError_t e__ = (expr);                                 \
        if (e__ != cudaSuccess) { ctx->last_cuda = (int)e__; return (code); } \
    } while (0)

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

constexpr int NSLOT = 3;   // pipeline depth of the host batch path (slot pipeline)
constexpr int NIO_MAX = 5; // staged pipeline: I/O buffer sets

struct Slot {              // per-chunk device buffers of the host batch path
    cudaStream_t stream = nullptr;
    DevBuf text, pattern, toff, poff, results, alnoff, outT, outP, dirs, fill, order, snap;
    DevBuf packT, packP, noff;             // staged host path: packed strings of the chunk
    DevBuf stats;                          // optional per-pair {identity, gaps}
    cudaEvent_t packed = nullptr;
    cudaEvent_t done = nullptr;
    cudaEvent_t in = nullptr, out = nullptr;      // staged host path: inputs landed / outputs drained
};

} // namespace

struct sa_context {
    int device = 0;
    int sms = 0;
    int smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {};
    // scoring tables on the device
    DevBuf dS4, dS;
    // single-pair / device-batch workspaces
    DevBuf dirs, rowbuf, fill, misc, dtext, dpat, doutT, doutP, sortbuf, tbbuf, snapbuf;
    DevBuf ckpt;                          // checkpoint rows of the linear-space traceback
    // device-batch pipeline: fills on the caller's stream, tracebacks on `stream`, two buffer sets
    DevBuf pdirs[2], psort[2];
    DevBuf packstate;                     // staged host path: running total of the packed strings
    DevBuf errflag;                       // host batch: set by validate_residues_kernel
    cudaEvent_t evFill[2] = {}, evTrace[2] = {};
    // the class kernels of one chunk are independent: they run on side streams so that the tail of
    // one class overlaps the body of the next
    static constexpr int NCLS_STREAMS = 4;
    cudaStream_t clsStream[NCLS_STREAMS] = {};
    cudaEvent_t evFork = nullptr, evSorted = nullptr, evJoin[NCLS_STREAMS] = {};
    DevBuf clsSnap[NCLS_STREAMS];
    PinBuf pin;
    Slot slot[NIO_MAX];
    uint32_t epoch = 0;
    bool wide = false;                      // the scoring uploaded last needs the two-plane profile
    int max_abs_score = 0;                  // ... and its largest |S|
    // sub-contexts for the long members of a host batch: several medium-size pairs run side by side (each is a
    // small cooperative launch), one host thread per sub-context
    std::vector<sa_context *> workers;
    size_t rowbuf_entries_valid = 0;
    // column-slice state of sa_strip_fill, consumed by sa_strip_traceback
    struct StripState {
        bool valid = false;
        int R = 0, CB = 0, C = 0, alpha = 0;
        uint32_t n_strips = 0;
        size_t strip_stride = 0;
        const uint8_t *d_text = nullptr, *d_pat = nullptr;
        uint64_t n = 0, m = 0, col0 = 0, total = 0, chunk = 0;
        size_t row_stride = 0, smem = 0;
        uint32_t dbg_strips = 0;
        int gap = 0;
        char alphabet[40] = {};
    } strip;
    sa_timing timing = {};
    const void *stats_src = nullptr;        // device: {identity, gaps} u64 of the last long-pair traceback
    sa_stats last_stats = {};               // identity / gaps of the last single-pair alignment
    bool have_stats = false;
    // per-kernel timing: (before fill, after fill, after traceback) event triples of the last call
    std::vector<cudaEvent_t> evpool;
    size_t evused = 0;
    bool timing_dirty = false;
    int last_cuda = 0;
    // bytes of direction workspace per chunk.  About 120 k pairs of 300 aa per chunk is the sweet spot on B200
    // (measured per 1 M pairs: 32.2 ms at 2 GB, 30.0 at 3 GB, 29.0 at 6 GB, 31.6 at 9 GB): the scattered tag reads of
    // the traceback start missing the TLB / L2 with larger chunks, smaller chunks pay more launches and tails.
    size_t dirs_budget = (size_t)6 << 30;         // host-buffer batches: the eighths of the graded schedule must fit one set
    // device-resident batches: chunks of ~80 000 pairs of config 4.  The traceback's scattered reads want the chunk's
    // directions small (1 M pairs on B200: 3383 / 3435 / 3471 / 3489 GCUPS with 6 / 8 / 10 / 12 chunks)
    size_t dev_dirs_budget = (size_t)3500 << 20;
    size_t host_dirs_budget = (size_t)8 << 30;    // host path: split over its NSLOT slots
    int tb_blocks_per_sm = 1;               // traceback blocks per SM while the next chunk's fill shares the GPU
    // sa_set_option (0 = automatic): chunk count floor of the device batch, checkpointed traceback of global alignments
    long long batch_min_chunks = 0, ckpt_rows = 0, ckpt_limit_mb = 0, ckpt_chunk_mb = 0;
};

namespace {

double ev_us(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0.0; }
    return (double)ms * 1000.0;
}

cudaEvent_t next_event(sa_context *ctx)
{
    if (ctx->evused == ctx->evpool.size()) {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        ctx->evpool.push_back(e);
    }
    return ctx->evpool[ctx->evused++];
}

void reset_timing(sa_context *ctx)
{
    ctx->timing = sa_timing{};
    ctx->evused = 0;
    ctx->timing_dirty = false;
}

// ------------------------------------------------------------------ scoring tables
// Scores beyond +-31 do not fit the one-byte profile of the packed kernels: such matrices run through the long-pair
// kernel with a two-plane profile (sweep_column_wide), whatever the pair size.
bool scoring_is_wide(const sa_scoring *sc)
{
    if (!sc || !sc->score_matrix) return false;
    const int a = sc->alphabet_size;
    if (a < 1 || a > MAX_ALPHA) return false;
    for (int i = 0; i < a * a; ++i)
        if (sc->score_matrix[i] * SCALE < -127 || sc->score_matrix[i] * SCALE > 127) return true;
    return false;
}

int upload_scoring(sa_context *ctx, const sa_scoring *sc, cudaStream_t st)
{
    if (!sc || !sc->score_matrix || !sc->alphabet) return SA_ERR_ARGUMENT;
    if (sc->alphabet_size < 2 || sc->alphabet_size > MAX_ALPHA) return SA_ERR_ARGUMENT;
    if (sc->mode != SA_GLOBAL && sc->mode != SA_LOCAL) return SA_ERR_ARGUMENT;
    if (sc->gap < 0 || sc->gap > (1 << 24)) return SA_ERR_SCORE_RANGE;
    const int a = sc->alphabet_size;
    // 4*S as bytes for IDP.4A: one plane when every |4*S| <= 127; otherwise ("wide" scoring, long-pair kernel only)
    // 4*S = 128*hi + lo with lo in 0..127 in the first plane and the signed hi in the second
    int8_t h4[2 * 32 * MAX_ALPHA];
    int32_t hs[MAX_ALPHA * MAX_ALPHA];
    std::memset(h4, 0, sizeof h4);
    const bool wide = scoring_is_wide(sc);
    for (int p = 0; p < a; ++p)
        for (int t = 0; t < a; ++t) {
            const int v = sc->score_matrix[p * a + t];
            if (v < -4064 || v > 4064) return SA_ERR_SCORE_RANGE;
            const int v4 = v * SCALE;
            if (!wide) h4[p * 32 + t] = (int8_t)v4;
            else {
                const int lo = v4 & 127;
                h4[p * 32 + t] = (int8_t)lo;
                h4[32 * MAX_ALPHA + p * 32 + t] = (int8_t)((v4 - lo) / 128);
            }
            hs[p * a + t] = v;
        }
    ctx->wide = wide;
    ctx->max_abs_score = 0;
    for (int i = 0; i < a * a; ++i) ctx->max_abs_score = std::max(ctx->max_abs_score, std::abs((int)sc->score_matrix[i]));
    SA_TRY(ctx->dS4.reserve(sizeof h4), SA_ERR_MEMORY);
    SA_TRY(ctx->dS.reserve(sizeof hs), SA_ERR_MEMORY);
    SA_TRY(cudaMemcpyAsync(ctx->dS4.p, h4, sizeof h4, cudaMemcpyHostToDevice, st), SA_ERR_COPY);
    SA_TRY(cudaMemcpyAsync(ctx->dS.p, hs, sizeof(int32_t) * a * a, cudaMemcpyHostToDevice, st), SA_ERR_COPY);
    // the two small host arrays live on this stack frame: make the copies complete before returning
    SA_TRY(cudaStreamSynchronize(st), SA_ERR_COPY);
    return SA_OK;
}

// ------------------------------------------------------------------ batch kernel dispatch
struct BatchCfg { int R, L; };

// Default classes: rows covered = R*L.  Short patterns use 16-lane groups (2 pairs per warp),
// longer ones a whole warp per pair.  SA_BATCH_CLASSES="R,L;R,L;..." overrides (dev tool).
const BatchCfg kBatchCfgs[] = {{8, 16}, {12, 16}, {16, 16}, {20, 16}, {24, 16}, {16, 32}, {24, 32}, {32, 32}, {48, 32}};
// every instantiation of batch_fill_kernel
#define SA_BATCH_CFG_LIST(X) X(8, 8) X(16, 8) X(32, 8) X(40, 8) X(48, 8) X(8, 16) X(12, 16) X(16, 16) X(20, 16) X(24, 16) \
    X(4, 32) X(6, 32) X(8, 32) X(10, 32) X(12, 32) X(16, 32) X(24, 32) X(32, 32) X(48, 32)
#ifndef SA_BATCH_WARPS
#define SA_BATCH_WARPS 4
#endif
constexpr int BATCH_WARPS = SA_BATCH_WARPS;   // warps per block
constexpr uint32_t BATCH_MAX_TEXT = 16384;
constexpr uint32_t BATCH_MAX_ROWS = 1536;

bool cfg_exists(int R, int L)
{
#define X(r, l) if (R == r && L == l) return true;
    SA_BATCH_CFG_LIST(X)
#undef X
    return false;
}

size_t batch_task_stride(const BatchCfg &c, uint32_t max_n, int packed)
{
    if (packed == 2) return (((size_t)max_n + 31 + 7) / 8) * 32 * po_for(c.R);       // octet layout (sa_batch16_sw.cuh), whole octets of columns
    const int CB = packed ? cb16_for(c.R) : cb_for(c.R), NW = c.R * CB / (packed ? 8 : 16);
    const size_t nblocks = ((size_t)max_n + c.L - 1 + CB - 1) / CB;
    return nblocks * NW * 32;
}

size_t batch_smem_bytes(const BatchCfg &c, int alpha, uint32_t max_n, bool local, bool packed)
{
    const int G = 32 / c.L;
    const size_t group = ((size_t)alpha * c.L * rpad_for(c.R) + ((max_n + 15u) & ~15u) + 15u) & ~(size_t)15;
    const size_t snap = local ? (size_t)16 * G : 0;
    return 32 * MAX_ALPHA + (size_t)BATCH_WARPS * ((packed ? 2 : 1) * G * group + snap);
}

// s16x2 kernels exist for the 16-lane classes
#define SA_BATCH16_CFG_LIST(X) X(8, 16) X(12, 16) X(16, 16) X(20, 16) X(24, 16) X(4, 32) X(6, 32) X(8, 32) X(10, 32) X(12, 32)
// default classes when the scores fit 16 bits: a whole warp per PAIR OF PAIRS (keeps 12 warps per SM)
const BatchCfg kBatchCfgs16[] = {{4, 32}, {6, 32}, {8, 32}, {10, 32}, {12, 32}, {16, 32}, {24, 32}, {32, 32}, {48, 32}};
bool cfg16_exists(int R, int L)
{
#define X(r, l) if (R == r && L == l) return true;
    SA_BATCH16_CFG_LIST(X)
#undef X
    return false;
}

// Classes needed for patterns up to max_m and texts up to max_n.  Returns false if max_m is not covered.
bool sw16_exists(const BatchCfg &cfg);
bool build_class_table(uint32_t max_n, uint32_t max_m, bool allow16, bool line, BatchClassTable *T)
{
    std::vector<BatchCfg> cfgs;
    if (const char *e = std::getenv("SA_BATCH_CLASSES")) {
        int R = 0, L = 0, used = 0;
        const char *q = e;
        while (std::sscanf(q, "%d,%d%n", &R, &L, &used) == 2) {
            if (cfg_exists(R, L)) cfgs.push_back(BatchCfg{R, L});
            q += used;
            if (*q == ';') ++q; else break;
        }
        std::sort(cfgs.begin(), cfgs.end(), [](const BatchCfg &x, const BatchCfg &y) { return x.R * x.L < y.R * y.L; });
    }
    if (cfgs.empty() || (uint32_t)(cfgs.back().R * cfgs.back().L) < max_m) {
        if (allow16) {
            cfgs.assign(std::begin(kBatchCfgs16), std::end(kBatchCfgs16));
            if (line && sw16_exists(BatchCfg{9, 32})) {       // the straight-line kernels also come in odd strip heights: finer classes, fuller lanes
                cfgs.push_back(BatchCfg{9, 32});
                cfgs.push_back(BatchCfg{11, 32});
                std::sort(cfgs.begin(), cfgs.end(), [](const BatchCfg &x, const BatchCfg &y) { return x.R * x.L < y.R * y.L; });
            }
        }
        else cfgs.assign(std::begin(kBatchCfgs), std::end(kBatchCfgs));
    }
    std::memset(T, 0, sizeof *T);
    for (const BatchCfg &c : cfgs) {
        if (T->n_classes == MAX_CLASSES) break;
        const int k = T->n_classes++;
        T->R[k] = c.R; T->L[k] = c.L; T->max_rows[k] = (uint32_t)(c.R * c.L);
        // 2: straight-line kernel with the octet direction layout, 1: batch_fill16_kernel, 0: s32 kernel
        T->packed[k] = !allow16 ? 0 : (line && sw16_exists(c)) ? 2 : cfg16_exists(c.R, c.L) ? 1 : 0;
        T->stride[k] = batch_task_stride(c, max_n, T->packed[k]);
        if ((uint32_t)(c.R * c.L) >= max_m) break;      // larger classes cannot occur
    }
    T->max_text = std::min<uint32_t>(BATCH_MAX_TEXT, max_n);      // windows, strides and the dirs bound are sized from max_n
    int shift = 0;
    while ((max_n >> shift) >= (uint32_t)SORT_BUCKETS) ++shift;
    T->bucket_shift = shift;
    return T->n_classes > 0 && T->max_rows[T->n_classes - 1] >= max_m;
}

// upper bound of the direction words a chunk of `count` pairs can need
size_t batch_dirs_bound(const BatchClassTable &T, uint64_t count)
{
    double perPair = 0;
    size_t round = 0;
    for (int c = 0; c < T.n_classes; ++c) {
        perPair = std::max(perPair, (double)T.stride[c] / ((32 / T.L[c]) * (T.packed[c] ? 2 : 1)));
        round += T.stride[c];
    }
    return (size_t)(perPair * (double)count) + round + 64;
}

size_t batch_sort_bytes(uint64_t count) { return (size_t)MAX_CLASSES * SORT_BUCKETS * 4 + count * 8 + (MAX_CLASSES + 1) * sizeof(BatchClassDyn) + 256; }

template <int R, int L>
cudaError_t launch_batch_fill_t(const BatchArgs &A, bool local, int grid, size_t smem, cudaStream_t st)
{
    if (local) batch_fill_kernel<R, L, true, BATCH_WARPS><<<grid, BATCH_WARPS * 32, smem, st>>>(A);
    else batch_fill_kernel<R, L, false, BATCH_WARPS><<<grid, BATCH_WARPS * 32, smem, st>>>(A);
    return cudaGetLastError();
}

template <int R, int L>
int occupancy_batch_t(bool local, size_t smem)
{
    int nb = 0;
    // the opt-in shared-memory limit must be raised BEFORE the query, otherwise it reports 0 blocks
    if (local) {
        cudaFuncSetAttribute(batch_fill_kernel<R, L, true, BATCH_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, batch_fill_kernel<R, L, true, BATCH_WARPS>, BATCH_WARPS * 32, smem);
    } else {
        cudaFuncSetAttribute(batch_fill_kernel<R, L, false, BATCH_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, batch_fill_kernel<R, L, false, BATCH_WARPS>, BATCH_WARPS * 32, smem);
    }
    return nb;
}

template <int R, int L>
cudaError_t launch_batch_fill16_t(const BatchArgs &A, bool local, int grid, size_t smem, cudaStream_t st)
{
    if (local) batch_fill16_kernel<R, L, true, BATCH_WARPS><<<grid, BATCH_WARPS * 32, smem, st>>>(A);
    else batch_fill16_kernel<R, L, false, BATCH_WARPS><<<grid, BATCH_WARPS * 32, smem, st>>>(A);
    return cudaGetLastError();
}
template <int R, int L>
int occupancy_batch16_t(bool local, size_t smem)
{
    int nb = 0;
    if (local) {
        cudaFuncSetAttribute(batch_fill16_kernel<R, L, true, BATCH_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, batch_fill16_kernel<R, L, true, BATCH_WARPS>, BATCH_WARPS * 32, smem);
    } else {
        cudaFuncSetAttribute(batch_fill16_kernel<R, L, false, BATCH_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, batch_fill16_kernel<R, L, false, BATCH_WARPS>, BATCH_WARPS * 32, smem);
    }
    return nb;
}
cudaError_t launch_batch_fill16(const BatchCfg &cfg, const BatchArgs &A, bool local, int grid, size_t smem, cudaStream_t st)
{
#define X(r, l) if (cfg.R == r && cfg.L == l) return launch_batch_fill16_t<r, l>(A, local, grid, smem, st);
    SA_BATCH16_CFG_LIST(X)
#undef X
    return cudaErrorInvalidValue;
}
int occupancy_batch16(const BatchCfg &cfg, bool local, size_t smem)
{
#define X(r, l) if (cfg.R == r && cfg.L == l) return occupancy_batch16_t<r, l>(local, smem);
    SA_BATCH16_CFG_LIST(X)
#undef X
    return 0;
}

// straight-line packed SW kernel (sa_batch16_sw.cuh): L = 32 classes only
#define SA_SW16_R_LIST(X) X(4) X(6) X(8) X(9) X(10) X(11) X(12)
bool sw16_exists(const BatchCfg &cfg)
{
    static const bool off = [] { const char *e = std::getenv("SA_BATCH_SW16"); return e && e[0] == '0'; }();
    if (off || cfg.L != 32) return false;
#define X(r) if (cfg.R == r) return true;
    SA_SW16_R_LIST(X)
#undef X
    return false;
}
// the NW form of the straight-line kernels keeps not-yet-started lanes on their border with a -128 sentinel score,
// which needs 4*gap - 128 < 2
bool line16_ok(const sa_scoring *sc) { return sc->mode == SA_LOCAL || sc->gap <= 31; }
size_t sw16_smem_bytes(const BatchCfg &c, int alpha, uint32_t max_n) { return sw16_layout(c.R, alpha, max_n, BATCH_WARPS).blockBytes; }
cudaError_t launch_sw16(const BatchCfg &cfg, const BatchArgs &A, bool local, int grid, size_t smem, cudaStream_t st)
{
#define X(r) if (cfg.R == r) { \
        if (local) batch_line16_kernel<r, true, BATCH_WARPS><<<grid, BATCH_WARPS * 32, smem, st>>>(A); \
        else batch_line16_kernel<r, false, BATCH_WARPS><<<grid, BATCH_WARPS * 32, smem, st>>>(A); \
        return cudaGetLastError(); }
    SA_SW16_R_LIST(X)
#undef X
    return cudaErrorInvalidValue;
}
int occupancy_sw16(const BatchCfg &cfg, bool local, size_t smem)
{
    int nb = 0;
#define X(r) if (cfg.R == r) { \
        const void *fn = local ? (const void *)batch_line16_kernel<r, true, BATCH_WARPS> : (const void *)batch_line16_kernel<r, false, BATCH_WARPS>; \
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, BATCH_WARPS * 32, smem); }
    SA_SW16_R_LIST(X)
#undef X
    return nb;
}

// 16-bit range guard of the s16x2 kernels: every value the packed arithmetic can produce must fit.
bool fits_s16(const sa_scoring *sc, uint32_t max_n, uint32_t max_m)
{
    if (const char *e = std::getenv("SA_BATCH_S16")) if (e[0] == '0') return false;
    long long smax = 0, smin = 0;
    const int a = sc->alphabet_size;
    for (int i = 0; i < a * a; ++i) { smax = std::max<long long>(smax, sc->score_matrix[i]); smin = std::min<long long>(smin, sc->score_matrix[i]); }
    const long long g = sc->gap, lo = std::min<long long>(max_n, max_m), sum = (long long)max_n + max_m;
    const long long up = smax * lo;                                     // best possible score
    const long long down = sc->mode == SA_LOCAL ? 0 : std::max(g, -smin) * sum;   // worst global score
    const long long slack = 4 * (std::max(g, std::max(smax, -smin)) + 40);
    return 4 * std::max(up, down) + slack < 32000;
}

cudaError_t launch_batch_fill(const BatchCfg &cfg, const BatchArgs &A, bool local, int grid, size_t smem, cudaStream_t st)
{
#define X(r, l) if (cfg.R == r && cfg.L == l) return launch_batch_fill_t<r, l>(A, local, grid, smem, st);
    SA_BATCH_CFG_LIST(X)
#undef X
    return cudaErrorInvalidValue;
}
int occupancy_batch(const BatchCfg &cfg, bool local, size_t smem)
{
#define X(r, l) if (cfg.R == r && cfg.L == l) return occupancy_batch_t<r, l>(local, smem);
    SA_BATCH_CFG_LIST(X)
#undef X
    return 0;
}

// ---- residue validation (sa_b200.h: a residue >= alphabet_size is SA_ERR_ARGUMENT) ----
// The kernels clamp letters to alphabet_size-1 while staging, so a bad byte never indexes out of bounds; the host
// entry points still refuse such input.  Single pairs are scanned on the host (a few hundred KB); the chunks of a
// host batch are scanned on the device right behind their copy (600 MB per 1 M pairs would cost the host tens of ms).
__global__ void __launch_bounds__(256) validate_residues_kernel(const uint8_t *p, const size_t n, const unsigned alpha, int *flag)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
    unsigned worst = 0;
    const size_t head = min(n, (size_t)((16 - ((uintptr_t)p & 15)) & 15));
    if (tid < head) worst = p[tid];
    const uint4 *q = reinterpret_cast<const uint4 *>(p + head);
    const size_t nv = (n - head) / 16;
    for (size_t i = tid; i < nv; i += nthreads) {
        const uint4 v = q[i];
        const unsigned m = __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w));
        worst = max(worst, max(max(m & 0xffu, (m >> 8) & 0xffu), max((m >> 16) & 0xffu, m >> 24)));
    }
    const size_t tail = head + nv * 16;
    if (tail + tid < n) worst = max(worst, (unsigned)p[tail + tid]);
    if (worst >= alpha) *flag = 1;
}

bool residues_ok(const uint8_t *p, uint64_t n, int alpha)
{
    unsigned worst = 0;
    for (uint64_t i = 0; i < n; ++i) worst = std::max<unsigned>(worst, p[i]);      // vectorised by the compiler
    return worst < (unsigned)alpha;
}

// ---- staged host path: strings of a chunk packed back to back before they cross PCIe ----
// The traceback writes pair p's strings at the END of its slot of text_len + pattern_len bytes (it walks backwards and
// does not know the length in advance), so a chunk's arenas are about half slack.  These kernels pack the used
// parts; aln_off is rewritten to the packed position (absolute over the whole call: `running` carries the total of the
// earlier chunks), and the chunk's packed size is stored straight into pinned host memory for the host's copy call.
struct CompactArgs {
    const sa_result *results; uint64_t *aln_off; uint32_t count;
    const char *srcT, *srcP;              // arenas as the traceback addressed them (aln_off-based)
    char *dstT, *dstP;                    // packed chunk buffers
    unsigned long long *noff;             // count entries: packed offsets inside the pair's block of 1024 pairs
    unsigned long long *block_base;       // per block of 1024 pairs: absolute packed offset; one more entry: the chunk's
    unsigned long long *running;          // device: total of the chunks before this one
    volatile unsigned long long *host_total;     // pinned host: this chunk's packed bytes
};

// All three kernels use 128-thread blocks without shared-memory demands worth mentioning, so that -- like the traceback
// blocks -- they find room on SMs that the fill blocks of the following chunk already occupy (a 1024-thread scan block
// had to wait for an SM to drain, which held up the traceback stream: 33.8 -> 36.5 ms per 1 M pairs).
constexpr int PACK_ITEMS = 8, PACK_BLOCK = 128 * PACK_ITEMS;      // pairs per block of the local scan

__device__ __forceinline__ unsigned long long pack_block_exscan(unsigned long long v, unsigned long long *total)
{
    __shared__ unsigned long long wsum[4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    unsigned long long before = 0, all = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (k < w) before += wsum[k]; all += wsum[k]; }
    *total = all;
    __syncthreads();
    return before + inc - v;
}

// (1) per block of 1024 pairs: offsets inside the block, and the block's total
__global__ void __launch_bounds__(128) pack_local_kernel(const CompactArgs A)
{
    const uint32_t b0 = blockIdx.x * PACK_BLOCK + threadIdx.x * PACK_ITEMS;
    unsigned long long len[PACK_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < PACK_ITEMS; ++k) { len[k] = b0 + k < A.count ? A.results[b0 + k].aln_len : 0ull; s += len[k]; }
    unsigned long long total;
    unsigned long long off = pack_block_exscan(s, &total);
#pragma unroll
    for (int k = 0; k < PACK_ITEMS; ++k) { if (b0 + k < A.count) A.noff[b0 + k] = off; off += len[k]; }
    if (threadIdx.x == 0) A.block_base[blockIdx.x] = total;
}

// (2) one block: block totals -> absolute bases (after the chunks before this one); chunk total to the host
__global__ void __launch_bounds__(128) pack_bases_kernel(const CompactArgs A)
{
    const uint32_t nb = (A.count + PACK_BLOCK - 1) / PACK_BLOCK;
    const unsigned long long start = *A.running;
    unsigned long long carry = start;
    for (uint32_t b0 = 0; b0 < nb; b0 += 128) {
        const uint32_t i = b0 + threadIdx.x;
        const unsigned long long v = i < nb ? A.block_base[i] : 0ull;
        unsigned long long total;
        const unsigned long long ex = pack_block_exscan(v, &total);
        if (i < nb) A.block_base[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) {
        A.block_base[nb] = start;             // the chunk's own base, for the copy kernel
        *A.running = carry;
        *A.host_total = carry - start;
        __threadfence_system();
    }
}

// (3) one warp per pair: used part of the slot -> packed position; aln_off follows
__global__ void __launch_bounds__(128) compact_copy_kernel(const CompactArgs A)
{
    const uint32_t pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (pair >= A.count) return;
    const uint32_t len = (uint32_t)A.results[pair].aln_len;
    const unsigned long long dstAbs = A.block_base[pair / PACK_BLOCK] + A.noff[pair];
    const unsigned long long base = A.block_base[(A.count + PACK_BLOCK - 1) / PACK_BLOCK];
    const char *sT = A.srcT + A.aln_off[pair], *sP = A.srcP + A.aln_off[pair];
    char *dT = A.dstT + (dstAbs - base), *dP = A.dstP + (dstAbs - base);
    for (uint32_t k = lane; k < len; k += 32) { dT[k] = sT[k]; dP[k] = sP[k]; }
    __syncwarp();
    if (lane == 0) A.aln_off[pair] = dstAbs;
}

// Enqueue binning + fill (one launch per class) + traceback for pairs [first, first+count) of a
// device-resident batch.  d_dirs: direction workspace (dirs_words words); d_fill: score/end_i/end_j
// (3 x n_pairs x 4 B); d_sort: batch_sort_bytes(count) bytes.  Pairs whose pattern exceeds
// BATCH_MAX_ROWS or whose text exceeds BATCH_MAX_TEXT are left untouched (the host path aligns
// them one by one through the long-pair kernels).
int enqueue_batch(sa_context *ctx, const sa_scoring *sc, const sa_batch *b, sa_result *d_results,
                  uint64_t *d_alnoff, uint32_t *d_stats, char *d_outT, char *d_outP, uint32_t max_n, uint32_t max_m,
                  uint32_t *d_dirs, size_t dirs_words, void *d_fill, void *d_sort, DevBuf *snapbuf, cudaStream_t st,
                  uint32_t first, uint32_t count, cudaStream_t stTrace = nullptr, cudaEvent_t evFillDone = nullptr,
                  bool tbShare = false)
{
    const bool split = evFillDone != nullptr;      // traceback on its own stream, overlapping the next fill
    if (!split) stTrace = st;
    BatchClassTable T;
    if (max_n > BATCH_MAX_TEXT || !build_class_table(max_n, std::min(max_m, BATCH_MAX_ROWS), fits_s16(sc, max_n, max_m), line16_ok(sc), &T))
        return SA_ERR_ARGUMENT;
    if (batch_dirs_bound(T, count) > dirs_words) return SA_ERR_MEMORY;
    const bool local = sc->mode == SA_LOCAL;

    int32_t *d_score = reinterpret_cast<int32_t *>(d_fill);
    uint32_t *d_ei = reinterpret_cast<uint32_t *>(d_fill) + b->n_pairs;
    uint32_t *d_ej = reinterpret_cast<uint32_t *>(d_fill) + 2 * b->n_pairs;

    // ---- binning (counting sort by class and descending text length) ----
    BatchSortArgs S{};
    S.text_off = b->text_off; S.pattern_off = b->pattern_off; S.base = first; S.count = count; S.table = T;
    S.hist = reinterpret_cast<uint32_t *>(d_sort);
    S.key = S.hist + MAX_CLASSES * SORT_BUCKETS;
    S.order = S.key + count;
    S.dyn = reinterpret_cast<BatchClassDyn *>(reinterpret_cast<char *>(S.order + count) + ((8 - ((uintptr_t)(S.order + count) & 7)) & 7));
    S.skipped_results = d_results; S.skipped_alnoff = reinterpret_cast<unsigned long long *>(d_alnoff);
    // With SA_BATCH_CLASS_STREAMS (default on for big chunks) the binning and the class kernels run on the context's
    // HIGH-PRIORITY side streams (fork/join by events): the classes run concurrently, and none of it queues behind the
    // traceback blocks of the previous chunk, which would otherwise hold every SM until they drain.
    const char *cse = std::getenv("SA_BATCH_CLASS_STREAMS");
    const bool forkClasses = snapbuf == &ctx->snapbuf && count >= 4096 && T.n_classes > 1 && !(cse && cse[0] == '0');
    cudaEvent_t e0 = next_event(ctx), e1 = next_event(ctx), e2 = next_event(ctx), e3 = next_event(ctx);
    cudaEventRecord(e0, st);
    cudaStream_t stSort = st;
    if (forkClasses) {
        cudaEventRecord(ctx->evFork, st);
        stSort = ctx->clsStream[0];
        cudaStreamWaitEvent(stSort, ctx->evFork, 0);
    }
    SA_TRY(cudaMemsetAsync(S.hist, 0, (size_t)MAX_CLASSES * SORT_BUCKETS * 4, stSort), SA_ERR_LAUNCH);
    batch_classify_kernel<<<(count + 255) / 256, 256, 0, stSort>>>(S);
    batch_scan_kernel<<<1, 1024, 0, stSort>>>(S);
    batch_scatter_kernel<<<(count + 255) / 256, 256, 0, stSort>>>(S);
    SA_TRY(cudaGetLastError(), SA_ERR_LAUNCH);
    ctx->timing.kernel_launches += 3;
    if (forkClasses) cudaEventRecord(ctx->evSorted, stSort);

    // ---- fill: one launch per class; empty classes exit at once ----
    const cudaStream_t stMain = st;
    for (int c = 0; c < T.n_classes; ++c) {
        const int lane_ = c % sa_context::NCLS_STREAMS;
        if (forkClasses) {
            st = ctx->clsStream[lane_];
            if (c > 0 && c < sa_context::NCLS_STREAMS) cudaStreamWaitEvent(st, ctx->evSorted, 0);
            snapbuf = &ctx->clsSnap[lane_];
        }
        const BatchCfg cfg{T.R[c], T.L[c]};
        const bool packed = T.packed[c] != 0;
        const int G = (32 / cfg.L) * (packed ? 2 : 1);                  // pairs per warp task
        const bool sw16 = T.packed[c] == 2;                                            // straight-line kernels (sa_batch16_sw.cuh)
        const size_t smem = sw16 ? sw16_smem_bytes(cfg, sc->alphabet_size, max_n) : batch_smem_bytes(cfg, sc->alphabet_size, max_n, local, packed);
        if (smem > (size_t)ctx->smem_optin) return SA_ERR_ARGUMENT;
        BatchArgs A{};
        A.text = b->text; A.text_off = b->text_off; A.pattern = b->pattern; A.pattern_off = b->pattern_off;
        A.order = S.order; A.dyn = S.dyn + c; A.dirs = d_dirs; A.task_stride = T.stride[c];
        A.score = d_score; A.end_i = d_ei; A.end_j = d_ej;
        A.S4 = ctx->dS4.as<int8_t>(); A.alpha = sc->alphabet_size; A.gap = sc->gap; A.max_n = max_n;
        int occ = sw16 ? occupancy_sw16(cfg, local, smem) : packed ? occupancy_batch16(cfg, local, smem) : occupancy_batch(cfg, local, smem);
        if (occ < 1) return SA_ERR_LAUNCH;
        const uint64_t nTasksMax = ((uint64_t)count + G - 1) / G;
        int grid = (int)std::min<uint64_t>((uint64_t)ctx->sms * occ, (nTasksMax + BATCH_WARPS - 1) / BATCH_WARPS);
        if (grid < 1) grid = 1;
        if (local && !sw16) {
            // arg-max snapshots: one area per resident warp (stays in L2)
            const size_t need = (size_t)grid * BATCH_WARPS * ((cfg.R + 3) / 4) * 32 * 16 * (packed ? 2 : 1);
            if (need > snapbuf->cap) {
                SA_TRY(cudaStreamSynchronize(st), SA_ERR_LAUNCH);          // earlier launches may still use the old buffer
                SA_TRY(snapbuf->reserve(need), SA_ERR_MEMORY);
            }
            A.snap_ws = snapbuf->as<uint4>();
        }
        SA_TRY(sw16 ? launch_sw16(cfg, A, local, grid, smem, st)
                    : packed ? launch_batch_fill16(cfg, A, local, grid, smem, st) : launch_batch_fill(cfg, A, local, grid, smem, st),
               SA_ERR_LAUNCH);
        ctx->timing.kernel_launches++;
    }
    st = stMain;
    if (forkClasses) {
        for (int k = 0; k < std::min(T.n_classes, (int)sa_context::NCLS_STREAMS); ++k) {
            cudaEventRecord(ctx->evJoin[k], ctx->clsStream[k]);
            cudaStreamWaitEvent(st, ctx->evJoin[k], 0);
        }
    }
    cudaEventRecord(e1, st);
    if (split) {
        cudaEventRecord(evFillDone, st);
        cudaStreamWaitEvent(stTrace, evFillDone, 0);
    }
    cudaEventRecord(e2, stTrace);

    BatchTraceArgs R{};
    R.text = b->text; R.text_off = b->text_off; R.pattern = b->pattern; R.pattern_off = b->pattern_off;
    R.order = S.order; R.dyn = S.dyn; R.table = T; R.dirs = d_dirs;
    R.score = d_score; R.end_i = d_ei; R.end_j = d_ej;
    R.S = ctx->dS.as<int32_t>(); R.alpha = sc->alphabet_size; R.gap = sc->gap; R.local = local;
    std::memcpy(R.alphabet, sc->alphabet, sc->alphabet_size + 1);
    R.results = d_results; R.aln_off = d_alnoff; R.out_text = d_outT; R.out_pattern = d_outP; R.stats = d_stats;
    // tbShare: another chunk's fill follows on the other stream -- one block per SM leaves it its three blocks per SM
    // (the carve-out preference follows: maximum shared memory while sharing SMs with fill blocks, the default --
    // more L1 for the scattered tag reads -- when the traceback has the GPU to itself)
    cudaFuncSetAttribute(batch_traceback_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         tbShare ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault);
    const unsigned tbFull = (count + 127) / 128;
    const unsigned tbGrid = tbShare ? std::min<unsigned>(tbFull, (unsigned)ctx->sms * ctx->tb_blocks_per_sm) : tbFull;
    static const int tbThreads = [] { const char *e = std::getenv("SA_TB_THREADS"); return e ? std::atoi(e) : 128; }();
    if (tbShare && tbThreads != 128)
        batch_traceback_kernel<<<std::min<unsigned>((count + tbThreads - 1) / tbThreads, (unsigned)ctx->sms * ctx->tb_blocks_per_sm), tbThreads, 0, stTrace>>>(R);
    else
    batch_traceback_kernel<<<tbGrid, 128, 0, stTrace>>>(R);
    SA_TRY(cudaGetLastError(), SA_ERR_LAUNCH);
    cudaEventRecord(e3, stTrace);
    ctx->timing.kernel_launches++;
    ctx->timing_dirty = true;
    return SA_OK;
}

// ------------------------------------------------------------------ long-pair dispatch
constexpr int LONG_WARPS = 4;
static_assert(LONG_WARPS == TILE_WARPS, "both long-pair kernels use the same block shape");
const int kLongR[] = {2, 4, 6, 8, 12, 16};

template <int R>
const void *long_kernel_fn(bool local, bool linked)
{
    return local ? (const void *)long_fill_kernel<R, true, LONG_WARPS, false>
                 : linked ? (const void *)long_fill_kernel<R, false, LONG_WARPS, true> : (const void *)long_fill_kernel<R, false, LONG_WARPS, false>;
}
constexpr int WIDE_R = 8;         // the wide-score variant exists for one strip height
const void *long_kernel_wide_fn(bool local)
{
    return local ? (const void *)long_fill_kernel<WIDE_R, true, LONG_WARPS, false, true> : (const void *)long_fill_kernel<WIDE_R, false, LONG_WARPS, false, true>;
}
template <int R>
cudaError_t launch_long_t(const LongArgs &A, bool local, int grid, size_t smem, cudaStream_t st)
{
    void *args[] = {(void *)&A};
    const void *fn = long_kernel_fn<R>(local, A.left_col64 || A.right_col64 || A.dbg);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(LONG_WARPS * 32), args, smem, st);
}
template <int R>
int occupancy_long_t(bool local, size_t smem, bool linked)
{
    int nb = 0;
    const void *fn = long_kernel_fn<R>(local, linked);
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, LONG_WARPS * 32, smem);
    return nb;
}
#define LONG_DISPATCH(FN, ...)                       \
    switch (R) {                                     \
    case 2: return FN<2>(__VA_ARGS__);               \
    case 4: return FN<4>(__VA_ARGS__);               \
    case 6: return FN<6>(__VA_ARGS__);               \
    case 8: return FN<8>(__VA_ARGS__);               \
    case 12: return FN<12>(__VA_ARGS__);             \
    case 16: return FN<16>(__VA_ARGS__);             \
    }
cudaError_t launch_long(int R, const LongArgs &A, bool local, int grid, size_t smem, cudaStream_t st, bool wide = false)
{
    if (wide) {
        void *args[] = {(void *)&A};
        const void *fn = long_kernel_wide_fn(local);
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(LONG_WARPS * 32), args, smem, st);
    }
    LONG_DISPATCH(launch_long_t, A, local, grid, smem, st);
    return cudaErrorInvalidValue;
}
int occupancy_long(int R, bool local, size_t smem, bool linked = false, bool wide = false)
{
    if (wide) {
        int nb = 0;
        const void *fn = long_kernel_wide_fn(local);
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, LONG_WARPS * 32, smem);
        return nb;
    }
    LONG_DISPATCH(occupancy_long_t, local, smem, linked);
    return 0;
}

struct LongPlan {
    int R, CB, NW, grid;
    int C;                   // > 0: register-tiled kernel (sa_tile.cuh) with C columns per tile; 0: long_fill_kernel
    uint32_t n_strips, ring;
    size_t strip_stride, row_stride, smem;
};

// Resident blocks per SM of a tiled fill: two (two warps per SM scheduler).  One tiled warp keeps its scheduler about
// half busy, so two saturate it; warps beyond that only stretch every macro-step -- and with it the hand-off lag that each
// of the n_strips strips pays (~55 macro-steps) and the time strip 0 needs to cross a column slice, which is
// the delay per GPU of the linked slices of config 5.  Measured: a 125 000-column slice (3 716 strips) fills in 76.8 / 57.0 /
// 63.5 / 72.6 ms with 1 / 2 / 3 / 4 blocks per SM and strip 0 crosses it in 9.5 / 11.8 / 13.7 / 14.8 ms; a 250 000-column
// slice in 105 / 102 / 105 ms with 2 / 3 / 4, and config 5 on 4 GPUs in 0.199 / 0.218 / 0.236 s.  SA_LONG_BLOCKS_PER_SM overrides.
// A fill whose strips all fit with three or four blocks per SM but not with two takes those (a second wave of a few
// strips costs a whole sweep: 1 000 000 x 317 000, a chunk of the checkpointed traceback, 192 ms with two and 137 ms
// with three).
int tile_blocks_per_sm(const sa_context *ctx, uint64_t n_strips, int slices, int occ)
{
    const uint64_t perBlockSm = (uint64_t)ctx->sms * TILE_WARPS;
    int perSm = 2;
    if (slices <= 1 && n_strips > 2 * perBlockSm && n_strips <= 4 * perBlockSm) perSm = n_strips <= 3 * perBlockSm ? 3 : 4;
    perSm = std::min(occ, perSm);
    if (const char *e = std::getenv("SA_LONG_BLOCKS_PER_SM")) { const int b = std::atoi(e); if (b >= 1) perSm = std::min(occ, b); }
    return perSm;
}

// Which long-pair kernel, measured on B200 (tools/probe_tile.py, profiles/README.md round 2).  The register-tiled kernel
// (sa_tile.cuh) is built for the latency-bound regime -- at most a few strips per SM scheduler:
// 100 000 x 95 217 fills in 10.3 ms with 8 x 2 tiles against 13.7 ms (one-column kernel), 4 000 x 3 800 in 0.41 against
// 0.61 ms.  With its residency capped by tile_blocks_per_sm it also wins on the tall shapes: a 125 000-column slice of
// config 5 (3 716 strips) in 51 ms against 59 ms (one-column kernel, strips of 512 rows), 1 000 000 x 390 000 (a chunk of
// the checkpointed traceback) in 147 against 181 ms; 500 000 x 475 000 is a tie (112 / 114 ms).
// SA_TILE="R,C" forces a tile shape; SA_LONG_R or SA_LONG_KERNEL=strip force the one-column kernel (always the path
// of wide score matrices).
bool pick_tile(const sa_context *ctx, const sa_scoring *sc, uint64_t n, uint64_t m, bool traceback, bool local, int slices, int *R, int *C)
{
    if (std::getenv("SA_LONG_R")) return false;
    if (const char *e = std::getenv("SA_LONG_KERNEL")) if (!std::strcmp(e, "strip")) return false;
    // 8 x 2 tiles fill fastest; when a traceback follows a small matrix, 4 x 4 (strips of 128 rows) keeps the per-strip
    // passes of the parallel traceback short (3 903 x 3 698: fill + traceback 0.82 ms against 0.99).  Local alignments
    // take the same kernels with the branch-free arg-max key of sa_tile.cuh, as long as tile_scale(R) * S fits the
    // one-byte profile and tile_scale(R) * H fits 32 bits; otherwise the one-column kernel.
    // Column slices (config 5): GPU k starts when strip 0 has crossed k slices, and a tiled strip crosses a 125 000-column
    // slice in 12 ms against 22 ms; 1 000 000 x 950 793 on 2 GPUs takes 0.274 s tiled, 0.328 s with the one-column kernel.
    (void)slices;
    int r = 8, c = 2;
    if (traceback && m <= 16000) { r = 4; c = 4; }
    bool use = true;
    if (const char *e = std::getenv("SA_TILE")) {
        int er = 0, ec = 0;
        if (std::sscanf(e, "%d,%d", &er, &ec) == 2 && tile_cfg_exists(er, ec)) { r = er; c = ec; }
    }
    if (local) {
        const long double scale = tile_scale(r, true), big = std::max<long double>(sc->gap, ctx->max_abs_score);
        if (ctx->max_abs_score * (int)scale > 127 || big * (long double)(n + m + 2) * scale >= 2147483000.0L) use = false;
    }
    *R = r; *C = c;
    return use;
}

cudaError_t launch_plan(const LongPlan &P, const LongArgs &A, bool local, int grid, cudaStream_t st, bool wide)
{
    if (P.C) return tile_launch(P.R, P.C, local, A.left_col64 || A.right_col64 || (A.dbg && !std::getenv("SA_LONG_DBG_PLAIN")), A, grid, P.smem, st);
    return launch_long(P.R, A, local, grid, P.smem, st, wide);
}

// The kernels carry 4*H in 32 bits: every value the recurrence can produce must fit.  |H| <= max(|S|, gap) * (n + m).
bool fits_s32(const sa_scoring *sc, uint64_t n, uint64_t m)
{
    long long big = sc->gap;
    for (int i = 0; i < sc->alphabet_size * sc->alphabet_size; ++i) big = std::max<long long>(big, std::llabs((long long)sc->score_matrix[i]));
    return (long double)big * (long double)(n + m + 2) * SCALE < 2147483000.0L;
}

int plan_long(sa_context *ctx, const sa_scoring *sc, uint64_t n, uint64_t m, LongPlan *P, bool traceback = true, bool linked = false, int slices = 1)
{
    const bool local = sc->mode == SA_LOCAL;
    if (!fits_s32(sc, n, m)) return SA_ERR_SCORE_RANGE;
    // Strip height, measured on B200 (tools/probe_r.py, bench.py --workload c1..c3, bench_c5.py).  Fill alone:
    // R = 8 is fastest from 4 k to 300 k rows (a step costs ~55 ns + 5 ns per row of the lane, the chain lag per
    // strip is ~80 steps); beyond that the tallest strips win because every strip stays resident in one or two
    // waves.  With the traceback the per-strip segment walks (32 R rows each, three passes) count as well, which
    // favours lower strips for small matrices: 3.9 k x 3.7 k takes 1.30 / 1.31 / 1.64 ms at R = 4 / 6 / 8.
    (void)ctx;
    int R = m <= 300000 ? 8 : m <= 450000 ? 12 : 16;
    if (traceback && m <= 16000) R = 4;
    else if (traceback && m <= 60000) R = 6;
    if (const char *e = std::getenv("SA_LONG_R")) {
        const int r = std::atoi(e);
        for (int k : kLongR) if (k == r) R = r;
    }
    const bool wide = ctx->wide;
    if (wide) R = WIDE_R;
    int tR = 0, tC = 0;
    P->C = 0;
    if (!wide && pick_tile(ctx, sc, n, m, traceback, local, slices, &tR, &tC)) {
        P->R = tR; P->C = tC; P->CB = 1; P->NW = tile_nwt(tR, tC);
        P->n_strips = (uint32_t)((m + 32ull * tR - 1) / (32ull * tR));
        P->smem = tile_smem_bytes(tR, tC, sc->alphabet_size);
        const int occ = tile_occupancy(tR, tC, local, linked, P->smem);
        if (occ < 1) return SA_ERR_LAUNCH;
        // every strip of a launch should be resident at once when it can be: up to 8 blocks (32 strips) per SM
        const int perSm = tile_blocks_per_sm(ctx, P->n_strips, slices, occ);
        const uint64_t maxBlocks = (uint64_t)ctx->sms * perSm;
        const uint64_t needBlocks = (P->n_strips + TILE_WARPS - 1) / TILE_WARPS;
        P->grid = (int)std::min(maxBlocks, needBlocks);
        const uint64_t W = (uint64_t)P->grid * TILE_WARPS;
        P->ring = (uint32_t)std::min<uint64_t>(P->n_strips, W + 1);
        P->row_stride = (n + 8 + 63) & ~(size_t)63;
        const size_t nTiles = (n + tC - 1) / tC;
        P->strip_stride = (nTiles + 31) * 32 * P->NW;
        return SA_OK;
    }
    P->R = R; P->CB = cb_for(R); P->NW = R * P->CB / 16;
    P->n_strips = (uint32_t)((m + 32ull * R - 1) / (32ull * R));
    const size_t planes = wide ? 2 : 1;
    P->smem = planes * 32 * MAX_ALPHA + (size_t)LONG_WARPS * (planes * (size_t)sc->alphabet_size * 32 * rpad_for(R) + (local ? ((R + 3) / 4) * 32 * 16 : 0) + 64 + 2 * PB * 4);
    int occ = occupancy_long(R, local, P->smem, false, wide);
    if (occ < 1) return SA_ERR_LAUNCH;
    int perSm = std::min(occ, 2);
    if (const char *e = std::getenv("SA_LONG_BLOCKS_PER_SM")) { const int b = std::atoi(e); if (b >= 1) perSm = std::min(occ, b); }
    const uint64_t maxBlocks = (uint64_t)ctx->sms * perSm;
    const uint64_t needBlocks = (P->n_strips + LONG_WARPS - 1) / LONG_WARPS;
    P->grid = (int)std::min(maxBlocks, needBlocks);
    const uint64_t W = (uint64_t)P->grid * LONG_WARPS;
    P->ring = (uint32_t)std::min<uint64_t>(P->n_strips, W + 1);
    P->row_stride = (n + 63) & ~(size_t)63;
    const size_t nblocks = (n + 31 + P->CB - 1) / P->CB;
    P->strip_stride = nblocks * P->NW * 32;
    return SA_OK;
}

// Parallel traceback over the strip layout of the last fill (sa_traceback.cuh).  start_row >= 0 selects the
// column-slice form: the path starts at (start_row, n) and, with `slice`, ends on the slice's left edge.
__global__ void tb_stats_accumulate_kernel(const unsigned long long *src, unsigned long long *dst) { dst[0] += src[0]; dst[1] += src[1]; }
__global__ void tb_finish_checkpointed_kernel(uint64_t *res, const unsigned long long *emitted) { res[0] = *emitted; res[1] = 0; res[2] = 0; res[3] = 0; }

struct TbChunk {          // row chunk of a checkpointed traceback (see TbArgs)
    const int *start_col_dev = nullptr; bool chunk_top = false; unsigned long long *global_off = nullptr; int *exit_col_dev = nullptr;
};
int enqueue_parallel_traceback(sa_context *ctx, const LongPlan &P, uint64_t n, uint64_t m, const uint8_t *d_text,
                               const uint8_t *d_pat, int alpha, int gap, const char *alphabet, bool local, int *d_cv,
                               uint32_t *d_ci, uint32_t *d_cj, int32_t *d_score, char *d_outT, char *d_outP, uint64_t cap,
                               uint64_t *d_res, bool traceback, long long start_row, bool slice, double slope, cudaStream_t st,
                               const TbChunk *chunk = nullptr)
{
    {
        TbArgs T{};
        T.Lay.dirs = ctx->dirs.as<uint32_t>(); T.Lay.strip_stride = P.strip_stride;
        T.Lay.R = P.R; T.Lay.CB = P.CB; T.Lay.NW = P.NW; T.Lay.ROWS = 32 * P.R;
        T.Lay.cbShift = P.CB == 1 ? 0 : P.CB == 2 ? 1 : P.CB == 4 ? 2 : 3; T.Lay.n = (int)n; T.Lay.m = (int)m;
        T.Lay.C = P.C; T.Lay.cShift = P.C == 2 ? 1 : P.C == 4 ? 2 : P.C == 8 ? 3 : 0;
        T.text = d_text; T.pattern = d_pat;
        T.S = ctx->dS.as<int32_t>(); T.alpha = alpha; T.gap = gap; T.local = local;
        T.n_strips = P.n_strips;
        T.start_given = start_row >= 0 ? 1 : 0; T.start_row = (int)std::max<long long>(0, start_row); T.slice = slice ? 1 : 0;
        // band of candidates around the predicted crossing: +-max(512 columns, n/50).  (A crossing outside the band
        // only costs a serial segment; a wide band costs walkers far from the path, whose own paths run long gap
        // stretches: +-2048 made the walkers 0.47 ms of a 1.3 ms call at 3.9 k x 3.7 k.)
        const uint64_t half = std::max<uint64_t>(512, n / 50);
        // candidate spacing: as fine as 8 columns -- two candidates 8 apart merge within a strip far more often than
        // two that are 64 apart (unrelated 32 k x 32 k sequences: traceback 6.5 -> 0.8 ms) -- but at most ~1000
        // candidates per line, and never more than half a strip height
        int Wd = 8;
        while ((uint64_t)Wd * 500 < half && Wd * 2 <= 16 * P.R && Wd < 256) Wd *= 2;
        // ... and at most ~200 k walkers per call: a 500 k-column slice of config 5 has 1857 strips, and 625 candidates
        // on each (spacing 32) cost 17 ms per slice against 78 candidates (spacing 256), 0.325 -> 0.360 s at N = 2
        while ((uint64_t)P.n_strips * (2 * half / Wd + 1) > 200000 && Wd * 2 <= 16 * P.R && Wd < 256) Wd *= 2;
        if (const char *e = std::getenv("SA_TB_WD")) { const int w = std::atoi(e); if (w >= 1) Wd = w; }
        T.Wd = Wd; T.Q = (int)((n + Wd - 1) / Wd);
        {
            int bq = (int)((half + Wd - 1) / Wd);
            if (const char *e = std::getenv("SA_TB_BAND")) { const int b = std::atoi(e); if (b >= 1) bq = b; }
            T.BQ = std::max(1, std::min(bq, std::max(1, T.Q / 2)));
            T.slope = slope > 0 ? slope : local ? 1.0 : (double)n / (double)m;
        }
        T.cand_v = d_cv; T.cand_i = d_ci; T.cand_j = d_cj; T.score = d_score;
        if (chunk) { T.start_col_dev = chunk->start_col_dev; T.chunk_top = chunk->chunk_top ? 1 : 0; T.global_off = chunk->global_off; T.exit_col_dev = chunk->exit_col_dev; }
        const size_t nq = (size_t)2 * T.BQ + 1, S = P.n_strips;
        size_t off = 0;
        auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
        const size_t oSt = carve(sizeof(TbState)), oX = carve((S + 1) * 4), oLen = carve(S * 8), oDel = carve(S * 8),
                     oMin = carve(S * 8), oOff = carve(S * 8), oFa = carve(S * nq * 4);
        SA_TRY(ctx->tbbuf.reserve(off), SA_ERR_MEMORY);
        char *base = ctx->tbbuf.as<char>();
        T.st = reinterpret_cast<TbState *>(base + oSt); T.X = reinterpret_cast<int *>(base + oX);
        ctx->stats_src = &T.st->identity;
        T.seg_len = reinterpret_cast<unsigned long long *>(base + oLen);
        T.seg_delta = reinterpret_cast<long long *>(base + oDel); T.seg_min = reinterpret_cast<long long *>(base + oMin);
        T.seg_off = reinterpret_cast<unsigned long long *>(base + oOff); T.fa = reinterpret_cast<uint32_t *>(base + oFa);
        std::memcpy(T.alphabet, alphabet, alpha + 1);
        T.cap = cap; T.out_text = d_outT; T.out_pattern = d_outP; T.res = d_res;
        tb_prepare_kernel<<<1, 32, 0, st>>>(T);
        ctx->timing.kernel_launches++;
        if (traceback) {
            const long long walkers = (long long)S * (long long)nq;
            tb_walkers_kernel<<<(unsigned)((walkers + 127) / 128), 128, 0, st>>>(T);
            tb_resolve_kernel<<<1, 32, 0, st>>>(T);
            tb_count_kernel<<<(unsigned)((S + 63) / 64), 64, 0, st>>>(T);
            tb_offsets_kernel<<<1, 32, 0, st>>>(T);
            tb_emit_kernel<<<(unsigned)((S + 63) / 64), 64, 0, st>>>(T);
            ctx->timing.kernel_launches += 5;
        } else {
            SA_TRY(cudaMemsetAsync(d_res, 0, 24, st), SA_ERR_LAUNCH);
        }
        SA_TRY(cudaGetLastError(), SA_ERR_LAUNCH);
    }
    return SA_OK;
}

// Align one long pair whose sequences are already on the device.  Results (len, starts) land
// in ctx->misc (device) and are read back by the caller.
int enqueue_long(sa_context *ctx, const sa_scoring *sc, const uint8_t *d_text, uint64_t n,
                 const uint8_t *d_pat, uint64_t m, char *d_outT, char *d_outP, uint64_t cap,
                 bool traceback, cudaStream_t st)
{
    LongPlan P;
    int rc = plan_long(ctx, sc, n, m, &P, traceback);
    if (rc) return rc;
    const bool local = sc->mode == SA_LOCAL;
    SA_TRY(ctx->dirs.reserve((size_t)P.n_strips * P.strip_stride * 4), SA_ERR_MEMORY);
    const size_t rowEntries = (size_t)P.ring * P.row_stride;
    const bool fresh = rowEntries * 8 > ctx->rowbuf.cap;
    SA_TRY(ctx->rowbuf.reserve(rowEntries * 8), SA_ERR_MEMORY);
    // tags carry a per-call epoch so the ring never needs clearing between calls
    ctx->epoch = (ctx->epoch + 1) & 0x7ff;
    if (fresh || ctx->epoch == 0) {
        SA_TRY(cudaMemsetAsync(ctx->rowbuf.p, 0, ctx->rowbuf.cap, st), SA_ERR_LAUNCH);
        if (ctx->epoch == 0) ctx->epoch = 1;
    }
    // misc layout: [0..3] u64 res, then score(int), then cand arrays
    const size_t miscBytes = 64 + (size_t)P.n_strips * 12 + 64;   // [48..51] = gmax
    SA_TRY(ctx->misc.reserve(miscBytes), SA_ERR_MEMORY);
    uint64_t *d_res = ctx->misc.as<uint64_t>();
    int32_t *d_score = reinterpret_cast<int32_t *>(ctx->misc.as<char>() + 32);
    int *d_cv = reinterpret_cast<int *>(ctx->misc.as<char>() + 64);
    uint32_t *d_ci = reinterpret_cast<uint32_t *>(d_cv + P.n_strips);
    uint32_t *d_cj = d_ci + P.n_strips;

    LongArgs A{};
    A.text = d_text; A.n = (uint32_t)n; A.pattern = d_pat; A.m = (uint32_t)m;
    A.dirs = ctx->dirs.as<uint32_t>(); A.strip_stride = P.strip_stride;
    A.rowbuf = ctx->rowbuf.as<unsigned long long>(); A.ring = P.ring; A.row_stride = P.row_stride;
    A.S4 = ctx->dS4.as<int8_t>(); A.alpha = sc->alphabet_size; A.gap = sc->gap;
    A.n_strips = P.n_strips; A.left_col = nullptr; A.right_col = nullptr; A.col0 = 0;
    A.score = d_score; A.cand_v = d_cv; A.cand_i = d_ci; A.cand_j = d_cj;
    A.tag_base = (uint32_t)ctx->epoch << 21;
    A.gmax = reinterpret_cast<int *>(ctx->misc.as<char>() + 48);
    A.abort_flag = reinterpret_cast<int *>(ctx->misc.as<char>() + 56);
    SA_TRY(cudaMemsetAsync(A.gmax, 0, 12, st), SA_ERR_LAUNCH);          // gmax, (unused), abort flag
    cudaEvent_t e0 = next_event(ctx), e1 = next_event(ctx), e2 = next_event(ctx), e3 = next_event(ctx);
    cudaEventRecord(e0, st);
    SA_TRY(launch_plan(P, A, local, P.grid, st, ctx->wide), SA_ERR_LAUNCH);
    cudaEventRecord(e1, st);
    cudaEventRecord(e2, st);
    ctx->timing.kernel_launches++;

    const char *tbmode = std::getenv("SA_TB");
    if (tbmode && !std::strcmp(tbmode, "serial")) {
        // reference implementation of the device traceback: one thread chasing pointers
        LongTraceArgs T{};
        T.text = d_text; T.n = (uint32_t)n; T.pattern = d_pat; T.m = (uint32_t)m;
        T.dirs = ctx->dirs.as<uint32_t>(); T.strip_stride = P.strip_stride;
        T.S = ctx->dS.as<int32_t>(); T.alpha = sc->alphabet_size; T.gap = sc->gap; T.local = local;
        T.R = P.R; T.CB = P.CB; T.C = P.C; T.n_strips = P.n_strips;
        T.cand_v = d_cv; T.cand_i = d_ci; T.cand_j = d_cj; T.score = d_score;
        std::memcpy(T.alphabet, sc->alphabet, sc->alphabet_size + 1);
        T.cap = cap; T.out_text = d_outT; T.out_pattern = d_outP; T.res = d_res;
        T.emit = traceback ? 1 : 0;
        T.stats = reinterpret_cast<unsigned long long *>(ctx->misc.as<char>() + 40);      // (bytes 40..55 are free: score at 32, gmax at 48 is zeroed before the fill and not used afterwards)
        ctx->stats_src = T.stats;
        long_traceback_kernel<<<1, 32, 0, st>>>(T);
        SA_TRY(cudaGetLastError(), SA_ERR_LAUNCH);
        ctx->timing.kernel_launches++;
    } else {
        // parallel traceback (sa_traceback.cuh): walkers -> resolve -> count -> offsets -> emit
        rc = enqueue_parallel_traceback(ctx, P, n, m, d_text, d_pat, sc->alphabet_size, sc->gap, sc->alphabet, local, d_cv, d_ci,
                                        d_cj, d_score, d_outT, d_outP, cap, d_res, traceback, -1, false, 0.0, st);
        if (rc) return rc;
    }
    cudaEventRecord(e3, st);
    ctx->timing_dirty = true;
    return SA_OK;
}

// One GLOBAL alignment whose packed direction matrix does not fit the device (or SA_CKPT_ROWS asks for it): linear-space
// traceback by checkpoints (the reference caps the matrix at host RAM, alignSequenceGPU.cu:410-416).
//   pass 1  the matrix is filled in row chunks of `chunk` rows; only the int32 H row between two chunks is kept
//           (n x 4 B per checkpoint), the direction words of a chunk are overwritten by the next one;
//   pass 2  from the last chunk upwards: the chunk is filled again from its checkpoint (the last one still holds its
//           directions), the path is followed from where the chunk below was left to the chunk's top row, and the
//           piece is emitted in front of what has been emitted so far.
// Twice the fill work, directions for ONE chunk in memory: 1 M x 0.95 M needs 16 GB at 64 k rows instead of 250 GB.
// Everything is enqueued on `st` without a host synchronisation.  Results land where enqueue_long puts them.
int enqueue_long_checkpointed(sa_context *ctx, const sa_scoring *sc, const uint8_t *d_text, uint64_t n, const uint8_t *d_pat,
                              uint64_t m, char *d_outT, char *d_outP, uint64_t cap, uint64_t chunk_hint, bool traceback, cudaStream_t st)
{
    if (sc->mode != SA_GLOBAL) return SA_ERR_MEMORY;
    LongPlan P;
    int rc = plan_long(ctx, sc, n, std::min<uint64_t>(m, chunk_hint), &P, true);
    if (rc) return rc;
    const uint64_t ROWS = 32ull * P.R;
    const uint64_t chunk = std::min<uint64_t>((chunk_hint + ROWS - 1) / ROWS * ROWS, (m + ROWS - 1) / ROWS * ROWS);
    const uint64_t K = (m + chunk - 1) / chunk;
    const uint32_t stripsPerChunk = (uint32_t)(chunk / ROWS);
    SA_TRY(ctx->dirs.reserve((size_t)stripsPerChunk * P.strip_stride * 4), SA_ERR_MEMORY);
    SA_TRY(ctx->ckpt.reserve((size_t)K * n * 4 + 64), SA_ERR_MEMORY);
    SA_TRY(ctx->misc.reserve(256), SA_ERR_MEMORY);
    uint64_t *d_res = ctx->misc.as<uint64_t>();
    int32_t *d_score = reinterpret_cast<int32_t *>(ctx->misc.as<char>() + 32);
    int *d_col = reinterpret_cast<int *>(ctx->misc.as<char>() + 64);
    unsigned long long *d_emitted = reinterpret_cast<unsigned long long *>(ctx->misc.as<char>() + 72);
    unsigned long long *d_stats = reinterpret_cast<unsigned long long *>(ctx->misc.as<char>() + 80);      // identity, gaps over all chunks
    const int ncol = (int)n;
    SA_TRY(cudaMemcpyAsync(d_col, &ncol, 4, cudaMemcpyHostToDevice, st), SA_ERR_COPY);
    SA_TRY(cudaMemsetAsync(d_emitted, 0, 24, st), SA_ERR_LAUNCH);
    SA_TRY(cudaMemsetAsync(ctx->misc.as<char>() + 56, 0, 4, st), SA_ERR_LAUNCH);          // abort flag of the fills
    auto ckrow = [&](uint64_t c) { return ctx->ckpt.as<int>() + (size_t)c * n; };          // top row of chunk c (c >= 1)
    int32_t *d_scratch_score = reinterpret_cast<int32_t *>(ctx->misc.as<char>() + 100);   // pass 2: a chunk's own corner value is not the score
    auto fill_chunk = [&](uint64_t c, bool keepBottom) -> int {
        const uint64_t row0 = c * chunk, rows = std::min<uint64_t>(chunk, m - row0);
        const uint32_t nStrips = (uint32_t)((rows + ROWS - 1) / ROWS);
        const int grid = (int)std::min<uint64_t>((uint64_t)P.grid, (nStrips + LONG_WARPS - 1) / LONG_WARPS);
        const uint32_t ring = (uint32_t)std::min<uint64_t>(nStrips, (uint64_t)grid * LONG_WARPS + 1);
        const size_t rowEntries = (size_t)ring * P.row_stride;
        const bool fresh = rowEntries * 8 > ctx->rowbuf.cap;
        if (fresh) SA_TRY(cudaStreamSynchronize(st), SA_ERR_LAUNCH);
        SA_TRY(ctx->rowbuf.reserve(rowEntries * 8), SA_ERR_MEMORY);
        ctx->epoch = (ctx->epoch + 1) & 0x7ff;
        if (fresh || ctx->epoch == 0) {
            SA_TRY(cudaMemsetAsync(ctx->rowbuf.p, 0, ctx->rowbuf.cap, st), SA_ERR_LAUNCH);
            if (ctx->epoch == 0) ctx->epoch = 1;
        }
        LongArgs A{};
        A.text = d_text; A.n = (uint32_t)n; A.pattern = d_pat + row0; A.m = (uint32_t)rows;
        A.dirs = ctx->dirs.as<uint32_t>(); A.strip_stride = P.strip_stride;
        A.rowbuf = ctx->rowbuf.as<unsigned long long>(); A.ring = ring; A.row_stride = P.row_stride;
        A.S4 = ctx->dS4.as<int8_t>(); A.alpha = sc->alphabet_size; A.gap = sc->gap;
        A.n_strips = nStrips; A.col0 = 0; A.row_base = (uint32_t)row0;
        A.top_row = c ? ckrow(c) : nullptr;
        A.bottom_row = (keepBottom && c + 1 < K) ? ckrow(c + 1) : nullptr;
        A.score = keepBottom ? d_score : d_scratch_score;
        // pass 2: the path enters this chunk at column *d_col (left there by the traceback of the chunk below) and only
        // moves left from there: the columns beyond are not filled again (1 000 000 x 950 793 in 3 chunks: 0.70 -> 0.57 s)
        A.n_dev = (!keepBottom && P.C) ? d_col : nullptr;
        A.tag_base = (uint32_t)ctx->epoch << 21;
        A.gmax = reinterpret_cast<int *>(ctx->misc.as<char>() + 48);
        A.abort_flag = reinterpret_cast<int *>(ctx->misc.as<char>() + 56);
        SA_TRY(launch_plan(P, A, false, grid, st, ctx->wide), SA_ERR_LAUNCH);
        ctx->timing.kernel_launches++;
        if (A.bottom_row) {
            const unsigned long long *row = A.rowbuf + (size_t)((nStrips - 1) % ring) * P.row_stride;
            long_bottom_row_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(row, A.bottom_row, (uint32_t)n);
            SA_TRY(cudaGetLastError(), SA_ERR_LAUNCH);
            ctx->timing.kernel_launches++;
        }
        return SA_OK;
    };
    cudaEvent_t e0 = next_event(ctx), e1 = next_event(ctx), e2 = next_event(ctx), e3 = next_event(ctx);
    cudaEventRecord(e0, st);
    for (uint64_t c = 0; c < K; ++c) { rc = fill_chunk(c, true); if (rc) return rc; }
    cudaEventRecord(e1, st);
    cudaEventRecord(e2, st);          // (the second pass interleaves fills and tracebacks: it is accounted as traceback time)
    // score only (sa_fill_only): the first pass is the whole job
    for (uint64_t c = traceback ? K : 0; c-- > 0;) {
        if (c + 1 < K) { rc = fill_chunk(c, false); if (rc) return rc; }
        const uint64_t row0 = c * chunk, rows = std::min<uint64_t>(chunk, m - row0);
        LongPlan Pc = P;
        Pc.n_strips = (uint32_t)((rows + ROWS - 1) / ROWS);
        TbChunk tc;
        tc.start_col_dev = d_col; tc.chunk_top = c > 0; tc.global_off = d_emitted; tc.exit_col_dev = d_col;
        rc = enqueue_parallel_traceback(ctx, Pc, n, rows, d_text, d_pat + row0, sc->alphabet_size, sc->gap, sc->alphabet, false, nullptr,
                                        nullptr, nullptr, nullptr, d_outT, d_outP, cap, d_res, true, (long long)rows, false,
                                        (double)n / (double)m, st, &tc);
        if (rc) return rc;
        tb_stats_accumulate_kernel<<<1, 1, 0, st>>>(reinterpret_cast<const unsigned long long *>(ctx->stats_src), d_stats);
        ctx->timing.kernel_launches++;
    }
    // results in the layout of enqueue_long: {len, start_text, start_pattern, argmax} -- a global alignment starts at 0 / 0
    tb_finish_checkpointed_kernel<<<1, 1, 0, st>>>(d_res, d_emitted);
    SA_TRY(cudaGetLastError(), SA_ERR_LAUNCH);
    ctx->stats_src = d_stats;
    cudaEventRecord(e3, st);
    ctx->timing_dirty = true;
    return SA_OK;
}

// When the packed directions of a GLOBAL alignment would not fit the device (or SA_CKPT_ROWS forces it), the pair takes the
// checkpointed path.  SA_CKPT_LIMIT_MB: direction bytes above which it is taken (default 60 % of the device memory);
// SA_CKPT_CHUNK_MB: direction bytes of one row chunk (default: 72 % of the device memory, 85 % of what is free).
bool want_checkpoints(sa_context *ctx, const sa_scoring *sc, uint64_t n, uint64_t m, bool traceback, uint64_t *chunk_rows)
{
    (void)traceback;                      // (a score-only call of that size runs the first pass alone)
    if (sc->mode != SA_GLOBAL || ctx->wide) return false;
    const char *eRows = std::getenv("SA_CKPT_ROWS");
    if (eRows || ctx->ckpt_rows > 0) {
        const long long r = eRows ? std::atoll(eRows) : ctx->ckpt_rows;
        if (r > 0 && (uint64_t)r < m) { *chunk_rows = (uint64_t)r; return true; }
        return false;
    }
    const char *eLimit = std::getenv("SA_CKPT_LIMIT_MB");
    const long double dirBytes = (long double)(n + 64) * (long double)(m + 512) / 4.0L;
    if (dirBytes < 4.0L * 1073741824.0L && !eLimit && ctx->ckpt_limit_mb <= 0) return false;      // (cudaMemGetInfo costs milliseconds)
    size_t freeB = 0, totalB = 0;
    if (cudaMemGetInfo(&freeB, &totalB) != cudaSuccess) { cudaGetLastError(); return false; }
    long double limit = 0.6L * (long double)totalB;
    if (ctx->ckpt_limit_mb > 0) limit = (long double)ctx->ckpt_limit_mb * 1048576.0L;
    if (eLimit) limit = (long double)std::atoll(eLimit) * 1048576.0L;
    if (dirBytes <= limit) return false;
    // chunks as tall as the memory allows, all of the same height: a chunk is one launch of the long-pair kernel, every
    // launch pays one sweep of the text width whatever its height (1 000 000 columns: 65 536 rows fill at 0.9 TCUPS,
    // 390 000 rows at 2.7), and the second pass re-fills all chunks but the last.  72 % of the device for one chunk's
    // directions leaves room for the row ring (15 GB at 1 M columns) and the rest.
    // (what is free now plus what the context already holds for directions: other users of the device count)
    long double chunkBytes = std::min(0.72L * (long double)totalB, 0.85L * ((long double)freeB + (long double)ctx->dirs.cap));
    if (ctx->ckpt_chunk_mb > 0) chunkBytes = (long double)ctx->ckpt_chunk_mb * 1048576.0L;
    if (const char *e = std::getenv("SA_CKPT_CHUNK_MB")) chunkBytes = (long double)std::atoll(e) * 1048576.0L;
    const uint64_t rowsMax = std::max<uint64_t>(1024, (uint64_t)(chunkBytes * 4.0L / (long double)(n + 64)));
    const uint64_t K = (m + rowsMax - 1) / rowsMax;
    *chunk_rows = (m + K - 1) / K;
    return K > 1;
}

// enqueue_long, or its checkpointed form when the direction matrix is too large for the device
int enqueue_long_auto(sa_context *ctx, const sa_scoring *sc, const uint8_t *d_text, uint64_t n, const uint8_t *d_pat, uint64_t m,
                      char *d_outT, char *d_outP, uint64_t cap, bool traceback, cudaStream_t st)
{
    uint64_t chunkRows = 0;
    if (want_checkpoints(ctx, sc, n, m, traceback, &chunkRows))
        return enqueue_long_checkpointed(ctx, sc, d_text, n, d_pat, m, d_outT, d_outP, cap, chunkRows, traceback, st);
    return enqueue_long(ctx, sc, d_text, n, d_pat, m, d_outT, d_outP, cap, traceback, st);
}

// members of a batch that the batch kernels can take (the device-side classifier applies the same rule)
bool batch_eligible(uint64_t n, uint64_t m) { return m <= BATCH_MAX_ROWS && n <= BATCH_MAX_TEXT; }

// routing of a SINGLE pair: small ones ride the batch kernels, everything else the strip kernel
bool use_batch_path(uint64_t n, uint64_t m)
{
    if (const char *e = std::getenv("SA_FORCE_PATH")) {
        if (!std::strcmp(e, "long")) return false;
        if (!std::strcmp(e, "batch")) return batch_eligible(n, m);
    }
    return m <= 384 && n <= 4096;
}

} // namespace

// =================================================================== C ABI
extern "C" {

const char *sa_version(void) { return "sa_b200 0.1 (sm_100a)"; }

// Page-locked host buffers for the callers of sa_align_batch / sa_align: the copies of a call run at PCIe speed only
// from pinned memory (pageable buffers go through the driver's bounce buffer at a fraction of it).
void *sa_host_alloc(uint64_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void sa_host_free(void *p) { if (p) cudaFreeHost(p); }
int sa_host_register(void *p, uint64_t bytes)
{
    if (!p || !bytes) return SA_ERR_ARGUMENT;
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return SA_ERR_MEMORY; }
    return SA_OK;
}
int sa_host_unregister(void *p)
{
    if (!p) return SA_ERR_ARGUMENT;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return SA_ERR_ARGUMENT; }
    return SA_OK;
}

const char *sa_status_string(int s)
{
    switch (s) {
    case SA_OK: return "ok";
    case SA_ERR_NO_DEVICE: return "no CUDA device";
    case SA_ERR_MEMORY: return "sequence is too long, not enough memory";
    case SA_ERR_COPY: return "could not copy to/from device memory";
    case SA_ERR_ARGUMENT: return "invalid argument";
    case SA_ERR_SCORE_RANGE: return "score matrix / gap outside the supported range";
    case SA_ERR_LAUNCH: return "kernel launch failed";
    case SA_ERR_CAPACITY: return "output buffer too small";
    }
    return "unknown";
}

int sa_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int sa_create(int device, sa_context **out)
{
    if (!out) return SA_ERR_ARGUMENT;
    *out = nullptr;
    if (device < 0 || device >= sa_device_count()) return SA_ERR_NO_DEVICE;
    sa_context *ctx = new (std::nothrow) sa_context();
    if (!ctx) return SA_ERR_MEMORY;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return SA_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return SA_ERR_NO_DEVICE; }
    ctx->sms = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SA_ERR_NO_DEVICE; }
    for (auto &e : ctx->ev) cudaEventCreate(&e);
    for (auto &e : ctx->evFill) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    // the fill kernels of a chunk outrank the traceback of the previous chunk (which runs on ctx->stream): when both
    // are pending the block scheduler places fill blocks first and the light traceback blocks take what is left
    int prLeast = 0, prGreatest = 0;
    cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest);
    for (auto &st : ctx->clsStream) cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prGreatest);
    cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->evSorted, cudaEventDisableTiming);
    for (auto &e : ctx->evJoin) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto &e : ctx->evTrace) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto &s : ctx->slot) {
        cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&s.in, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&s.packed, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&s.out, cudaEventDisableTiming);
    }
    if (const char *e = std::getenv("SA_DIRS_BUDGET_MB")) ctx->dirs_budget = ctx->dev_dirs_budget = (size_t)std::atoll(e) << 20;
    if (const char *e = std::getenv("SA_DEV_DIRS_BUDGET_MB")) ctx->dev_dirs_budget = (size_t)std::atoll(e) << 20;
    if (const char *e = std::getenv("SA_HOST_DIRS_BUDGET_MB")) ctx->host_dirs_budget = (size_t)std::atoll(e) << 20;
    // The small kernels that share the GPU with the fill must ask for the same (maximum) shared-memory carve-out:
    // an SM only changes its L1/shared split when it is empty, so with their default split the fill blocks of the
    // next chunk (74 KB each) could not join the traceback blocks of the previous one.
    cudaFuncSetAttribute(batch_classify_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(batch_scan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(batch_scatter_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(pack_local_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(pack_bases_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(compact_copy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (const char *e = std::getenv("SA_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)std::atoi(e));
    if (const char *e = std::getenv("SA_TB_BLOCKS_PER_SM")) ctx->tb_blocks_per_sm = std::max(1, std::atoi(e));
    *out = ctx;
    return SA_OK;
}

// Tuning knobs of a context by name (the SA_* environment variables of the same names are developer overrides read
// when the context is created or per call).  Sizes in MB; 0 restores the automatic choice where there is one.
namespace {
struct OptRef { const char *name; int kind; };      // kind 0: size_t bytes given in MB, 1: long long, 2: int >= 1
long long *opt_ll(sa_context *c, const char *n)
{
    if (!std::strcmp(n, "batch_min_chunks")) return &c->batch_min_chunks;
    if (!std::strcmp(n, "ckpt_rows")) return &c->ckpt_rows;
    if (!std::strcmp(n, "ckpt_limit_mb")) return &c->ckpt_limit_mb;
    if (!std::strcmp(n, "ckpt_chunk_mb")) return &c->ckpt_chunk_mb;
    return nullptr;
}
size_t *opt_mb(sa_context *c, const char *n)
{
    if (!std::strcmp(n, "dirs_budget_mb")) return &c->dirs_budget;
    if (!std::strcmp(n, "dev_dirs_budget_mb")) return &c->dev_dirs_budget;
    if (!std::strcmp(n, "host_dirs_budget_mb")) return &c->host_dirs_budget;
    return nullptr;
}
} // namespace

int sa_set_option(sa_context *ctx, const char *name, long long value)
{
    if (!ctx || !name || value < 0) return SA_ERR_ARGUMENT;
    if (size_t *p = opt_mb(ctx, name)) {
        if (value < 1) return SA_ERR_ARGUMENT;
        *p = (size_t)value << 20;
        return SA_OK;
    }
    if (long long *p = opt_ll(ctx, name)) { *p = value; return SA_OK; }
    if (!std::strcmp(name, "tb_blocks_per_sm")) {
        if (value < 1 || value > 16) return SA_ERR_ARGUMENT;
        ctx->tb_blocks_per_sm = (int)value;
        return SA_OK;
    }
    return SA_ERR_ARGUMENT;
}

int sa_get_option(const sa_context *ctx, const char *name, long long *value)
{
    if (!ctx || !name || !value) return SA_ERR_ARGUMENT;
    sa_context *c = const_cast<sa_context *>(ctx);
    if (const size_t *p = opt_mb(c, name)) { *value = (long long)(*p >> 20); return SA_OK; }
    if (const long long *p = opt_ll(c, name)) { *value = *p; return SA_OK; }
    if (!std::strcmp(name, "tb_blocks_per_sm")) { *value = ctx->tb_blocks_per_sm; return SA_OK; }
    return SA_ERR_ARGUMENT;
}

void sa_destroy(sa_context *ctx)
{
    if (!ctx) return;
    for (sa_context *w : ctx->workers) sa_destroy(w);
    ctx->workers.clear();
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (DevBuf *b : {&ctx->dS4, &ctx->dS, &ctx->dirs, &ctx->rowbuf, &ctx->fill, &ctx->misc, &ctx->dtext, &ctx->dpat, &ctx->doutT, &ctx->doutP, &ctx->sortbuf, &ctx->tbbuf, &ctx->snapbuf, &ctx->ckpt})
        b->release();
    ctx->pin.release();
    for (auto &s : ctx->slot) {
        for (DevBuf *b : {&s.text, &s.pattern, &s.toff, &s.poff, &s.results, &s.alnoff, &s.outT, &s.outP, &s.dirs, &s.fill, &s.order, &s.snap, &s.packT, &s.packP, &s.noff, &s.stats}) b->release();
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.done) cudaEventDestroy(s.done);
        if (s.in) cudaEventDestroy(s.in);
        if (s.packed) cudaEventDestroy(s.packed);
        if (s.out) cudaEventDestroy(s.out);
    }
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto &e : ctx->evpool) if (e) cudaEventDestroy(e);
    for (auto &e : ctx->evFill) if (e) cudaEventDestroy(e);
    for (auto &e : ctx->evTrace) if (e) cudaEventDestroy(e);
    for (auto &st : ctx->clsStream) if (st) cudaStreamDestroy(st);
    if (ctx->evFork) cudaEventDestroy(ctx->evFork);
    if (ctx->evSorted) cudaEventDestroy(ctx->evSorted);
    for (auto &e : ctx->evJoin) if (e) cudaEventDestroy(e);
    for (auto &b : ctx->clsSnap) b.release();
    for (auto &b : ctx->pdirs) b.release();
    for (auto &b : ctx->psort) b.release();
    ctx->packstate.release();
    ctx->errflag.release();
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int sa_last_timing(const sa_context *cctx, sa_timing *out)
{
    if (!cctx || !out) return SA_ERR_ARGUMENT;
    sa_context *ctx = const_cast<sa_context *>(cctx);
    if (ctx->timing_dirty) {
        // kernel times from the event triples (valid once the caller has synchronised the stream)
        double fill = 0, tb = 0;
        // (fill start, fill end, traceback start, traceback end) per chunk
        for (size_t i = 0; i + 3 < ctx->evused && i + 3 < ctx->evpool.size(); i += 4) {
            if (cudaEventQuery(ctx->evpool[i + 3]) != cudaSuccess) { cudaGetLastError(); return SA_ERR_LAUNCH; }
            fill += ev_us(ctx->evpool[i], ctx->evpool[i + 1]);
            tb += ev_us(ctx->evpool[i + 2], ctx->evpool[i + 3]);
        }
        ctx->timing.fill_us = fill;
        ctx->timing.traceback_us = tb;
        ctx->timing_dirty = false;
    }
    *out = ctx->timing;
    return SA_OK;
}

int sa_last_cuda_error(const sa_context *ctx) { return ctx ? ctx->last_cuda : 0; }

int sa_last_stats(sa_context *ctx, sa_stats *out)
{
    if (!ctx || !out) return SA_ERR_ARGUMENT;
    if (!ctx->have_stats) return SA_ERR_ARGUMENT;        // no alignment with strings yet (or the last call was fill-only)
    *out = ctx->last_stats;
    return SA_OK;
}

void *sa_context_stream(const sa_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

// --------------------------------------------------------------- single pair
static int align_single(sa_context *ctx, const sa_scoring *sc, const uint8_t *text, uint64_t n,
                        const uint8_t *pattern, uint64_t m, sa_result *result, char *outT, char *outP,
                        uint64_t cap, bool traceback, uint64_t *argmax)
{
    if (!ctx) return SA_ERR_ARGUMENT;
    if (!sc || !text || !pattern || !result || n == 0 || m == 0) return SA_ERR_ARGUMENT;
    if (n >= (1ull << 31) - 64 || m >= (1ull << 31) - 64) return SA_ERR_ARGUMENT;
    if (traceback && (!outT || !outP)) return SA_ERR_ARGUMENT;
    if (traceback && cap < n + m) return SA_ERR_CAPACITY;
    if (sc->alphabet_size < 2 || sc->alphabet_size > MAX_ALPHA) return SA_ERR_ARGUMENT;
    if (!residues_ok(text, n, sc->alphabet_size) || !residues_ok(pattern, m, sc->alphabet_size)) return SA_ERR_ARGUMENT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ERR_NO_DEVICE;
    cudaStream_t st = ctx->stream;
    reset_timing(ctx);
    ctx->timing.cells = (n + 1) * (m + 1);
    int rc = upload_scoring(ctx, sc, st);
    if (rc) return rc;

    const uint64_t slot = n + m;
    SA_TRY(ctx->dtext.reserve(n + 16), SA_ERR_MEMORY);
    SA_TRY(ctx->dpat.reserve(m + 16), SA_ERR_MEMORY);
    SA_TRY(ctx->doutT.reserve(slot + 16), SA_ERR_MEMORY);
    SA_TRY(ctx->doutP.reserve(slot + 16), SA_ERR_MEMORY);
    cudaEventRecord(ctx->ev[0], st);
    SA_TRY(cudaMemcpyAsync(ctx->dtext.p, text, n, cudaMemcpyHostToDevice, st), SA_ERR_COPY);
    SA_TRY(cudaMemcpyAsync(ctx->dpat.p, pattern, m, cudaMemcpyHostToDevice, st), SA_ERR_COPY);
    cudaEventRecord(ctx->ev[1], st);

    sa_result hres{};
    uint64_t hoff = 0, hargmax = 0;
    if (use_batch_path(n, m) && !ctx->wide) {
        // a one-pair batch through the batch kernels
        SA_TRY(ctx->misc.reserve(256), SA_ERR_MEMORY);
        int64_t hoffs[4] = {0, (int64_t)n, 0, (int64_t)m};
        int64_t *d_offs = ctx->misc.as<int64_t>();                 // [0,1]=text_off [2,3]=pattern_off
        sa_result *d_res = reinterpret_cast<sa_result *>(ctx->misc.as<char>() + 64);
        uint64_t *d_alnoff = reinterpret_cast<uint64_t *>(ctx->misc.as<char>() + 128);
        SA_TRY(cudaMemcpyAsync(d_offs, hoffs, sizeof hoffs, cudaMemcpyHostToDevice, st), SA_ERR_COPY);
        SA_TRY(cudaStreamSynchronize(st), SA_ERR_COPY);
        BatchClassTable T;
        if (!build_class_table((uint32_t)n, (uint32_t)m, fits_s16(sc, (uint32_t)n, (uint32_t)m), line16_ok(sc), &T)) return SA_ERR_ARGUMENT;
        SA_TRY(ctx->dirs.reserve(batch_dirs_bound(T, 1) * 4), SA_ERR_MEMORY);
        SA_TRY(ctx->fill.reserve(64), SA_ERR_MEMORY);
        SA_TRY(ctx->sortbuf.reserve(batch_sort_bytes(1)), SA_ERR_MEMORY);
        sa_batch b{1, ctx->dtext.as<uint8_t>(), d_offs, ctx->dpat.as<uint8_t>(), d_offs + 2};
        cudaEventRecord(ctx->ev[2], st);
        uint32_t *d_st = reinterpret_cast<uint32_t *>(ctx->misc.as<char>() + 192);
        rc = enqueue_batch(ctx, sc, &b, d_res, d_alnoff, d_st, ctx->doutT.as<char>(), ctx->doutP.as<char>(),
                           (uint32_t)n, (uint32_t)m, ctx->dirs.as<uint32_t>(), ctx->dirs.cap / 4, ctx->fill.p,
                           ctx->sortbuf.p, &ctx->snapbuf, st, 0, 1);
        if (rc) return rc;
        cudaEventRecord(ctx->ev[3], st);
        SA_TRY(cudaMemcpyAsync(&hres, d_res, sizeof hres, cudaMemcpyDeviceToHost, st), SA_ERR_COPY);
        SA_TRY(cudaMemcpyAsync(&hoff, d_alnoff, 8, cudaMemcpyDeviceToHost, st), SA_ERR_COPY);
        uint32_t hst[2] = {0, 0};
        SA_TRY(cudaMemcpyAsync(hst, d_st, 8, cudaMemcpyDeviceToHost, st), SA_ERR_COPY);
        SA_TRY(cudaStreamSynchronize(st), SA_ERR_LAUNCH);
        ctx->last_stats.identity = hst[0]; ctx->last_stats.gaps = hst[1]; ctx->have_stats = traceback;
        if (sc->mode == SA_LOCAL) {
            uint32_t ij[2] = {0, 0};
            cudaMemcpy(&ij[0], reinterpret_cast<uint32_t *>(ctx->fill.p) + 1, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(&ij[1], reinterpret_cast<uint32_t *>(ctx->fill.p) + 2, 4, cudaMemcpyDeviceToHost);
            hargmax = (uint64_t)ij[0] * (n + 1) + ij[1];
        }
    } else {
        cudaEventRecord(ctx->ev[2], st);
        rc = enqueue_long_auto(ctx, sc, ctx->dtext.as<uint8_t>(), n, ctx->dpat.as<uint8_t>(), m,
                               ctx->doutT.as<char>(), ctx->doutP.as<char>(), slot, traceback, st);
        if (rc) return rc;
        cudaEventRecord(ctx->ev[3], st);
        uint64_t hr[4]; int32_t hs = 0, hflag = 0;
        SA_TRY(cudaMemcpyAsync(hr, ctx->misc.p, sizeof hr, cudaMemcpyDeviceToHost, st), SA_ERR_COPY);
        SA_TRY(cudaMemcpyAsync(&hs, ctx->misc.as<char>() + 32, 4, cudaMemcpyDeviceToHost, st), SA_ERR_COPY);
        SA_TRY(cudaMemcpyAsync(&hflag, ctx->misc.as<char>() + 56, 4, cudaMemcpyDeviceToHost, st), SA_ERR_COPY);
        unsigned long long hst[2] = {0, 0};
        if (traceback) SA_TRY(cudaMemcpyAsync(hst, ctx->stats_src, 16, cudaMemcpyDeviceToHost, st), SA_ERR_COPY);
        SA_TRY(cudaStreamSynchronize(st), SA_ERR_LAUNCH);
        if (hflag) return SA_ERR_LAUNCH;          // a strip's watchdog fired (sa_tile.cuh): the launch gave up instead of hanging
        ctx->last_stats.identity = hst[0]; ctx->last_stats.gaps = hst[1]; ctx->have_stats = traceback;
        hres.score = hs; hres.aln_len = hr[0]; hres.start_text = hr[1]; hres.start_pattern = hr[2];
        hoff = slot - hr[0];
        hargmax = hr[3];
    }
    if (traceback && hres.aln_len) {
        SA_TRY(cudaMemcpyAsync(outT, ctx->doutT.as<char>() + hoff, hres.aln_len, cudaMemcpyDeviceToHost, st), SA_ERR_COPY);
        SA_TRY(cudaMemcpyAsync(outP, ctx->doutP.as<char>() + hoff, hres.aln_len, cudaMemcpyDeviceToHost, st), SA_ERR_COPY);
    }
    cudaEventRecord(ctx->ev[4], st);
    SA_TRY(cudaStreamSynchronize(st), SA_ERR_COPY);
    *result = hres;
    if (argmax) *argmax = hargmax;
    ctx->timing.h2d_us = ev_us(ctx->ev[0], ctx->ev[1]);
    ctx->timing.d2h_us = ev_us(ctx->ev[3], ctx->ev[4]);
    ctx->timing.total_us = ev_us(ctx->ev[0], ctx->ev[4]);
    {
        sa_timing k{};
        sa_last_timing(ctx, &k);     // folds the per-kernel event triples into fill_us / traceback_us
    }
    return SA_OK;
}

int sa_align(sa_context *ctx, const sa_scoring *sc, const uint8_t *text, uint64_t n, const uint8_t *pattern,
             uint64_t m, sa_result *result, char *outT, char *outP, uint64_t cap)
{
    return align_single(ctx, sc, text, n, pattern, m, result, outT, outP, cap, true, nullptr);
}

int sa_fill_only(sa_context *ctx, const sa_scoring *sc, const uint8_t *text, uint64_t n, const uint8_t *pattern,
                 uint64_t m, int32_t *score, uint64_t *argmax)
{
    sa_result r{};
    // long pairs skip the string emission; short ones go through the (cheap) batch traceback
    if (!ctx) return SA_ERR_ARGUMENT;
    const uint64_t cap = n + m;
    std::vector<char> t, p;
    char *pt = nullptr, *pp = nullptr;
    if (use_batch_path(n, m) && !scoring_is_wide(sc)) { t.resize(cap); p.resize(cap); pt = t.data(); pp = p.data(); }
    int rc = align_single(ctx, sc, text, n, pattern, m, &r, pt, pp, cap, pt != nullptr, argmax);
    if (rc) return rc;
    if (score) *score = r.score;
    return SA_OK;
}

// Device-resident single pair: sequences, output buffers (capacity n+m each) and the 4-word
// result {len, start_text, start_pattern, score} are device pointers; enqueued on `stream`.
int sa_align_device(sa_context *ctx, const sa_scoring *sc, const uint8_t *d_text, uint64_t n,
                    const uint8_t *d_pattern, uint64_t m, char *d_outT, char *d_outP, uint64_t *d_result4,
                    void *stream)
{
    if (!ctx || !sc || !d_text || !d_pattern || !d_outT || !d_outP || !d_result4 || n == 0 || m == 0) return SA_ERR_ARGUMENT;
    if (n >= (1ull << 31) - 64 || m >= (1ull << 31) - 64) return SA_ERR_ARGUMENT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ERR_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)stream;   // as given; NULL is the CUDA default stream
    int rc = upload_scoring(ctx, sc, st);
    if (rc) return rc;
    reset_timing(ctx);
    ctx->timing.cells = (n + 1) * (m + 1);
    rc = enqueue_long_auto(ctx, sc, d_text, n, d_pattern, m, d_outT, d_outP, n + m, true, st);
    if (rc) return rc;
    // misc = {len, start_text, start_pattern, argmax} then score at byte 32
    SA_TRY(cudaMemcpyAsync(d_result4, ctx->misc.p, 24, cudaMemcpyDeviceToDevice, st), SA_ERR_COPY);
    SA_TRY(cudaMemcpyAsync(d_result4 + 3, ctx->misc.as<char>() + 32, 4, cudaMemcpyDeviceToDevice, st), SA_ERR_COPY);
    return SA_OK;
}

// --------------------------------------------------------------- column slices of one pair (multi-GPU, config 5)
// Fill the slice [col0, col0+n) x all m rows of a GLOBAL alignment.  d_left_col / d_right_col hold 4*H(i, .)
// for i = 0..m (the form the kernels carry); d_left_col == NULL means the slice starts at the matrix border.
// The direction words stay in the context until the next fill; d_text / d_pattern must stay valid as well.
int sa_strip_begin(sa_context *ctx, const sa_scoring *sc, const uint8_t *d_text, uint64_t n, uint64_t col0,
                   uint64_t text_total, const uint8_t *d_pattern, uint64_t m, uint64_t chunk_rows_hint,
                   uint64_t *chunk_rows, void *stream)
{
    if (!ctx || !sc || !d_text || !d_pattern || n == 0 || m == 0) return SA_ERR_ARGUMENT;
    if (sc->mode != SA_GLOBAL) return SA_ERR_ARGUMENT;              // the arg-max of a local alignment is not sliced (yet)
    if (scoring_is_wide(sc)) return SA_ERR_SCORE_RANGE;
    if (col0 + n >= (1ull << 31) - 64 || m >= (1ull << 31) - 64 || text_total < col0 + n) return SA_ERR_ARGUMENT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ERR_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = upload_scoring(ctx, sc, st);
    if (rc) return rc;
    reset_timing(ctx);
    ctx->timing.cells = n * m;
    auto &S = ctx->strip;
    S.valid = false;
    // the strip height follows the rows that are in flight together: one chunk
    const uint64_t hint = chunk_rows_hint == 0 || chunk_rows_hint > m ? m : chunk_rows_hint;
    LongPlan P;
    rc = plan_long(ctx, sc, n, hint, &P, true, false, (int)std::min<uint64_t>(64, (text_total + n / 2) / n));
    if (rc) return rc;
    const uint64_t ROWS = 32ull * P.R;
    const uint64_t chunk = hint >= m ? m : (hint + ROWS - 1) / ROWS * ROWS;
    const uint32_t stripsTotal = (uint32_t)((m + ROWS - 1) / ROWS);
    SA_TRY(ctx->dirs.reserve((size_t)stripsTotal * P.strip_stride * 4), SA_ERR_MEMORY);
    SA_TRY(ctx->misc.reserve(128), SA_ERR_MEMORY);
    S.R = P.R; S.CB = P.CB; S.C = P.C; S.alpha = sc->alphabet_size; S.n_strips = stripsTotal; S.strip_stride = P.strip_stride;
    S.row_stride = P.row_stride; S.smem = P.smem; S.chunk = chunk;
    S.d_text = d_text; S.d_pat = d_pattern; S.n = n; S.m = m; S.col0 = col0; S.total = text_total; S.gap = sc->gap;
    std::memset(S.alphabet, 0, sizeof S.alphabet);
    std::memcpy(S.alphabet, sc->alphabet, sc->alphabet_size + 1);
    S.valid = true;
    if (chunk_rows) *chunk_rows = chunk;
    return SA_OK;
}

static int strip_launch(sa_context *ctx, uint64_t row0, uint64_t rows, const int32_t *d_left_col, int32_t *d_right_col,
                        const int32_t *d_top_row, int32_t *d_bottom_row, const unsigned long long *d_left64,
                        unsigned long long *d_right64, uint32_t xtag, int32_t *d_score, void *stream);

int sa_strip_fill_rows(sa_context *ctx, uint64_t row0, uint64_t rows, const int32_t *d_left_col, int32_t *d_right_col,
                       const int32_t *d_top_row, int32_t *d_bottom_row, int32_t *d_score, void *stream)
{
    if (!ctx || !ctx->strip.valid) return SA_ERR_ARGUMENT;
    if ((ctx->strip.col0 == 0) != (d_left_col == nullptr)) return SA_ERR_ARGUMENT;
    return strip_launch(ctx, row0, rows, d_left_col, d_right_col, d_top_row, d_bottom_row, nullptr, nullptr, 0, d_score, stream);
}

// ---- slices linked inside the launch: peer buffers through CUDA IPC ------------------------------------------
int sa_peer_alloc(sa_context *ctx, uint64_t bytes, void **dptr, unsigned char handle[64])
{
    if (!ctx || !dptr || !handle || bytes == 0) return SA_ERR_ARGUMENT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ERR_NO_DEVICE;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    SA_TRY(cudaMalloc(&p, bytes), SA_ERR_MEMORY);
    SA_TRY(cudaMemset(p, 0, bytes), SA_ERR_LAUNCH);
    cudaIpcMemHandle_t h;
    SA_TRY(cudaIpcGetMemHandle(&h, p), SA_ERR_LAUNCH);
    std::memcpy(handle, &h, 64);
    *dptr = p;
    return SA_OK;
}
int sa_peer_open(sa_context *ctx, const unsigned char handle[64], void **dptr)
{
    if (!ctx || !dptr || !handle) return SA_ERR_ARGUMENT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ERR_NO_DEVICE;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    SA_TRY(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess), SA_ERR_LAUNCH);
    return SA_OK;
}
int sa_peer_close(sa_context *ctx, void *dptr)
{
    if (!ctx || !dptr) return SA_ERR_ARGUMENT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ERR_NO_DEVICE;
    SA_TRY(cudaIpcCloseMemHandle(dptr), SA_ERR_LAUNCH);
    return SA_OK;
}
int sa_peer_free(sa_context *ctx, void *dptr)
{
    if (!ctx || !dptr) return SA_ERR_ARGUMENT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ERR_NO_DEVICE;
    SA_TRY(cudaFree(dptr), SA_ERR_LAUNCH);
    return SA_OK;
}

// The whole slice in one launch; the border column arrives in d_left_col64 (this GPU's memory, pattern_len+1 words
// {4H, tag}, written by the left neighbour's kernel; NULL for the first slice) and the right-most column is written
// into d_right_col64 (the right neighbour's buffer opened with sa_peer_open; NULL for the last slice).  `tag` must
// be the same on every rank and differ from call to call.
int sa_strip_fill_linked(sa_context *ctx, const uint64_t *d_left_col64, uint64_t *d_right_col64, uint32_t tag,
                         int32_t *d_score, void *stream)
{
    if (!ctx || !ctx->strip.valid || tag == 0) return SA_ERR_ARGUMENT;
    if ((ctx->strip.col0 == 0) != (d_left_col64 == nullptr)) return SA_ERR_ARGUMENT;
    return strip_launch(ctx, 0, ctx->strip.m, nullptr, nullptr, nullptr, nullptr,
                        reinterpret_cast<const unsigned long long *>(d_left_col64),
                        reinterpret_cast<unsigned long long *>(d_right_col64), tag, d_score, stream);
}

// SA_OK, or SA_ERR_LAUNCH when a strip of the last linked fill gave up waiting for its neighbour (synchronises).
int sa_strip_linked_status(sa_context *ctx, void *stream)
{
    if (!ctx || !ctx->strip.valid) return SA_ERR_ARGUMENT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ERR_NO_DEVICE;
    int flag = 0;
    SA_TRY(cudaMemcpyAsync(&flag, ctx->misc.as<char>() + 56, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream), SA_ERR_COPY);
    SA_TRY(cudaStreamSynchronize((cudaStream_t)stream), SA_ERR_LAUNCH);
    if (const char *path = std::getenv("SA_LONG_DBG")) {
        std::vector<unsigned long long> h((size_t)ctx->strip.dbg_strips * (16 + 128));      // 16 scalars + 64 {group, ns} stall events per strip
        if (!h.empty() && cudaMemcpy(h.data(), ctx->tbbuf.p, h.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
            char name[512];
            std::snprintf(name, sizeof name, "%s.dev%d", path, ctx->device);
            if (FILE *f = std::fopen(name, "wb")) { std::fwrite(h.data(), 8, h.size(), f); std::fclose(f); }
        }
    }
    return flag ? SA_ERR_LAUNCH : SA_OK;
}

static int strip_launch(sa_context *ctx, uint64_t row0, uint64_t rows, const int32_t *d_left_col, int32_t *d_right_col,
                        const int32_t *d_top_row, int32_t *d_bottom_row, const unsigned long long *d_left64,
                        unsigned long long *d_right64, uint32_t xtag, int32_t *d_score, void *stream)
{
    if (!ctx || !ctx->strip.valid || rows == 0) return SA_ERR_ARGUMENT;
    auto &S = ctx->strip;
    const uint64_t ROWS = 32ull * S.R;
    const bool last = row0 + rows == S.m;
    if (row0 + rows > S.m || row0 % ROWS != 0 || (!last && rows % ROWS != 0)) return SA_ERR_ARGUMENT;
    if ((row0 == 0) != (d_top_row == nullptr) || (!last && !d_bottom_row)) return SA_ERR_ARGUMENT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ERR_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t nStrips = (uint32_t)((rows + ROWS - 1) / ROWS);
    const bool linkedKernel = d_left64 || d_right64 || (std::getenv("SA_LONG_DBG") && !std::getenv("SA_LONG_DBG_PLAIN"));
    int occ = S.C ? tile_occupancy(S.R, S.C, false, linkedKernel, S.smem) : occupancy_long(S.R, false, S.smem, linkedKernel);
    if (occ < 1) return SA_ERR_LAUNCH;
    int perSm = std::min(occ, 2);
    if (S.C) perSm = tile_blocks_per_sm(ctx, nStrips, (int)std::min<uint64_t>(64, (S.total + S.n / 2) / S.n), occ);
    else if (const char *e = std::getenv("SA_LONG_BLOCKS_PER_SM")) { const int b = std::atoi(e); if (b >= 1) perSm = std::min(occ, b); }
    const uint64_t maxBlocks = (uint64_t)ctx->sms * perSm;
    const int grid = (int)std::min<uint64_t>(maxBlocks, (nStrips + LONG_WARPS - 1) / LONG_WARPS);
    const uint32_t ring = (uint32_t)std::min<uint64_t>(nStrips, (uint64_t)grid * LONG_WARPS + 1);
    const size_t rowEntries = (size_t)ring * S.row_stride;
    const bool fresh = rowEntries * 8 > ctx->rowbuf.cap;
    if (fresh) SA_TRY(cudaStreamSynchronize(st), SA_ERR_LAUNCH);       // an earlier chunk may still use the old ring
    SA_TRY(ctx->rowbuf.reserve(rowEntries * 8), SA_ERR_MEMORY);
    ctx->epoch = (ctx->epoch + 1) & 0x7ff;
    if (fresh || ctx->epoch == 0) {
        SA_TRY(cudaMemsetAsync(ctx->rowbuf.p, 0, ctx->rowbuf.cap, st), SA_ERR_LAUNCH);
        if (ctx->epoch == 0) ctx->epoch = 1;
    }
    LongArgs A{};
    A.text = S.d_text; A.n = (uint32_t)S.n; A.pattern = S.d_pat + row0; A.m = (uint32_t)rows;
    A.dirs = ctx->dirs.as<uint32_t>() + (size_t)(row0 / ROWS) * S.strip_stride; A.strip_stride = S.strip_stride;
    A.rowbuf = ctx->rowbuf.as<unsigned long long>(); A.ring = ring; A.row_stride = S.row_stride;
    A.S4 = ctx->dS4.as<int8_t>(); A.alpha = S.alpha; A.gap = S.gap;
    A.n_strips = nStrips; A.left_col = d_left_col ? d_left_col + row0 : nullptr;
    A.right_col = d_right_col ? d_right_col + row0 : nullptr; A.col0 = (uint32_t)S.col0;
    A.row_base = (uint32_t)row0; A.top_row = d_top_row; A.bottom_row = last ? nullptr : d_bottom_row;
    A.left_col64 = d_left64; A.right_col64 = d_right64; A.xtag = xtag;
    A.abort_flag = reinterpret_cast<int *>(ctx->misc.as<char>() + 56);
    SA_TRY(cudaMemsetAsync(A.abort_flag, 0, 4, st), SA_ERR_LAUNCH);
    if (std::getenv("SA_LONG_DBG")) {          // dev aid: per-strip timestamps, dumped by sa_strip_linked_status
        SA_TRY(ctx->tbbuf.reserve((size_t)nStrips * (128 + 1024) + 64), SA_ERR_MEMORY);
        SA_TRY(cudaMemsetAsync(ctx->tbbuf.p, 0, (size_t)nStrips * (128 + 1024), st), SA_ERR_LAUNCH);
        A.dbg = ctx->tbbuf.as<unsigned long long>();
        ctx->strip.dbg_strips = nStrips;
    }
    A.score = d_score ? d_score : reinterpret_cast<int32_t *>(ctx->misc.as<char>() + 32);
    A.tag_base = (uint32_t)ctx->epoch << 21;
    A.gmax = reinterpret_cast<int *>(ctx->misc.as<char>() + 48);
    cudaEvent_t e0 = next_event(ctx), e1 = next_event(ctx), e2 = next_event(ctx), e3 = next_event(ctx);
    cudaEventRecord(e0, st);
    {
        LongPlan P{};
        P.R = S.R; P.C = S.C; P.smem = S.smem;
        SA_TRY(launch_plan(P, A, false, grid, st, false), SA_ERR_LAUNCH);
    }
    ctx->timing.kernel_launches++;
    if (A.bottom_row) {     // the last strip left its bottom row in its ring row
        const unsigned long long *row = A.rowbuf + (size_t)((nStrips - 1) % ring) * S.row_stride;
        long_bottom_row_kernel<<<(unsigned)((S.n + 255) / 256), 256, 0, st>>>(row, A.bottom_row, (uint32_t)S.n);
        SA_TRY(cudaGetLastError(), SA_ERR_LAUNCH);
        ctx->timing.kernel_launches++;
    }
    cudaEventRecord(e1, st); cudaEventRecord(e2, st); cudaEventRecord(e3, st);
    ctx->timing_dirty = true;
    return SA_OK;
}

int sa_strip_fill(sa_context *ctx, const sa_scoring *sc, const uint8_t *d_text, uint64_t n, uint64_t col0,
                  uint64_t text_total, const uint8_t *d_pattern, uint64_t m, const int32_t *d_left_col,
                  int32_t *d_right_col, int32_t *d_score, void *stream)
{
    int rc = sa_strip_begin(ctx, sc, d_text, n, col0, text_total, d_pattern, m, 0, nullptr, stream);
    if (rc) return rc;
    return sa_strip_fill_rows(ctx, 0, m, d_left_col, d_right_col, nullptr, nullptr, d_score, stream);
}

// Follow the path through the slice filled last: it enters on the right edge at DP row start_row.  The piece is
// written right-aligned into d_outT / d_outP (capacity cap >= n + m); d_res4 = {len, exit_row, text index, pattern index}.
int sa_strip_traceback(sa_context *ctx, uint64_t start_row, char *d_outT, char *d_outP, uint64_t cap, uint64_t *d_res4,
                       void *stream)
{
    if (!ctx || !d_outT || !d_outP || !d_res4) return SA_ERR_ARGUMENT;
    const auto &S = ctx->strip;
    if (!S.valid || start_row > S.m || cap < S.n + S.m) return SA_ERR_ARGUMENT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ERR_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)stream;
    if (const char *e = std::getenv("SA_TB"); e && !std::strcmp(e, "serial")) {
        StripTraceArgs T{};
        T.text = S.d_text; T.n = (uint32_t)S.n; T.pattern = S.d_pat; T.m = (uint32_t)S.m;
        T.dirs = ctx->dirs.as<uint32_t>(); T.strip_stride = S.strip_stride;
        T.alpha = S.alpha; T.R = S.R; T.CB = S.CB; T.C = S.C; T.col0 = (uint32_t)S.col0; T.start_row = start_row;
        std::memcpy(T.alphabet, S.alphabet, sizeof T.alphabet);
        T.cap = cap; T.out_text = d_outT; T.out_pattern = d_outP; T.res = d_res4;
        strip_traceback_kernel<<<1, 32, 0, st>>>(T);
        SA_TRY(cudaGetLastError(), SA_ERR_LAUNCH);
        ctx->timing.kernel_launches++;
        return SA_OK;
    }
    LongPlan P{};
    P.R = S.R; P.CB = S.CB; P.C = S.C; P.NW = S.C ? tile_nwt(S.R, S.C) : S.R * S.CB / 16; P.n_strips = S.n_strips; P.strip_stride = S.strip_stride;
    return enqueue_parallel_traceback(ctx, P, S.n, S.m, S.d_text, S.d_pat, S.alpha, S.gap, S.alphabet, false, nullptr, nullptr,
                                      nullptr, nullptr, d_outT, d_outP, cap, d_res4, true, (long long)start_row, S.col0 > 0,
                                      (double)S.total / (double)S.m, st);
}

// --------------------------------------------------------------- batches
int sa_align_batch_device(sa_context *ctx, const sa_scoring *sc, const sa_batch *b, sa_batch_out *out,
                          uint32_t max_n, uint32_t max_m, void *stream)
{
    if (!ctx || !sc || !b || !out || !b->text || !b->pattern || !b->text_off || !b->pattern_off ||
        !out->results || !out->aln_off || !out->aligned_text || !out->aligned_pattern)
        return SA_ERR_ARGUMENT;
    if (b->n_pairs == 0) return SA_OK;
    if (b->n_pairs >= (1ull << 31)) return SA_ERR_ARGUMENT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ERR_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)stream;   // as given; NULL is the CUDA default stream
    if (scoring_is_wide(sc)) return SA_ERR_SCORE_RANGE;      // wide matrices only through the single-pair kernels (host entry points)
    int rc = upload_scoring(ctx, sc, st);
    if (rc) return rc;
    BatchClassTable T;
    if (max_n > BATCH_MAX_TEXT || max_m > BATCH_MAX_ROWS || !build_class_table(max_n, max_m, fits_s16(sc, max_n, max_m), line16_ok(sc), &T))
        return SA_ERR_ARGUMENT;
    // Chunks are software-pipelined: sort+fill of chunk c+1 runs on the caller's stream while the
    // (latency-bound) traceback of chunk c runs on the context's stream; two buffer sets.
    const double perPair = (double)batch_dirs_bound(T, 1 << 20) / (double)(1 << 20) * 4.0;      // bytes per pair
    uint64_t chunk = std::max<uint64_t>(32, (uint64_t)((double)ctx->dev_dirs_budget / perPair));
    const char *ps = std::getenv("SA_BATCH_PIPELINE");
    const bool pipeline = !(ps && ps[0] == '0') && b->n_pairs >= 8192;
    // at least 4 chunks to overlap; small batches 2 -- every chunk costs ~0.15 ms of binning and kernel tails
    // (125 000 pairs: 3.51 / 3.55 / 3.62 / 3.79 ms with 2 / 3 / 4 / 6 chunks; 250 000: 7.18 / 6.96 / 6.85 / 6.86)
    uint64_t minChunks = b->n_pairs >= 200000 ? 4 : 2;
    if (ctx->batch_min_chunks >= 1) minChunks = (uint64_t)ctx->batch_min_chunks;
    if (const char *e = std::getenv("SA_BATCH_MIN_CHUNKS")) { const int v = std::atoi(e); if (v >= 1) minChunks = (uint64_t)v; }
    if (pipeline) chunk = std::min<uint64_t>(chunk, (b->n_pairs + minChunks - 1) / minChunks);
    chunk = std::min<uint64_t>(chunk, b->n_pairs);
    chunk = (b->n_pairs + (b->n_pairs + chunk - 1) / chunk - 1) / ((b->n_pairs + chunk - 1) / chunk);      // equal chunks, no stub at the end
    for (int k = 0; k < (pipeline ? 2 : 1); ++k) {
        SA_TRY(ctx->pdirs[k].reserve(batch_dirs_bound(T, chunk) * 4), SA_ERR_MEMORY);
        SA_TRY(ctx->psort[k].reserve(batch_sort_bytes(chunk)), SA_ERR_MEMORY);
    }
    SA_TRY(ctx->fill.reserve(b->n_pairs * 12 + 64), SA_ERR_MEMORY);
    reset_timing(ctx);
    cudaEventRecord(ctx->ev[2], st);
    int c = 0;
    for (uint64_t first = 0; first < b->n_pairs; first += chunk, ++c) {
        const uint32_t count = (uint32_t)std::min<uint64_t>(chunk, b->n_pairs - first);
        const int k = pipeline ? (c & 1) : 0;
        if (pipeline && c >= 2) cudaStreamWaitEvent(st, ctx->evTrace[k], 0);      // buffer set k is free again
        rc = enqueue_batch(ctx, sc, b, out->results, out->aln_off, out->stats, out->aligned_text, out->aligned_pattern,
                           max_n, max_m, ctx->pdirs[k].as<uint32_t>(), ctx->pdirs[k].cap / 4, ctx->fill.p, ctx->psort[k].p,
                           &ctx->snapbuf, st, (uint32_t)first, count, pipeline ? ctx->stream : nullptr,
                           pipeline ? ctx->evFill[k] : nullptr, pipeline && first + chunk < b->n_pairs);
        if (rc) return rc;
        if (pipeline) cudaEventRecord(ctx->evTrace[k], ctx->stream);
    }
    if (pipeline) {
        // join: the caller's stream continues only after the last tracebacks
        cudaStreamWaitEvent(st, ctx->evTrace[0], 0);
        if (c >= 2) cudaStreamWaitEvent(st, ctx->evTrace[1], 0);
    }
    cudaEventRecord(ctx->ev[3], st);
    return SA_OK;
}

int sa_align_batch(sa_context *ctx, const sa_scoring *sc, const sa_batch *b, sa_batch_out *out)
{
    if (!ctx || !sc || !b || !out || !b->text || !b->pattern || !b->text_off || !b->pattern_off ||
        !out->results || !out->aln_off || !out->aligned_text || !out->aligned_pattern)
        return SA_ERR_ARGUMENT;
    const uint64_t N = b->n_pairs;
    if (N == 0) return SA_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SA_ERR_NO_DEVICE;
    const int64_t *to = b->text_off, *po = b->pattern_off;
    if ((uint64_t)(to[N] - to[0] + po[N] - po[0]) > out->arena_capacity) return SA_ERR_CAPACITY;

    // pairs too long for the batch kernels (all of them with a wide score matrix) are aligned one by one through sa_align
    const bool wideScores = scoring_is_wide(sc);
    uint32_t max_n = 0, max_m = 0;
    uint64_t cells = 0;
    std::vector<uint64_t> longPairs;
    {
        // one pass over the offsets (cell count, longest members, members for the long-pair kernels); big batches split it
        // over a few host threads -- at 1 M pairs the single-threaded pass was ~1.5 ms in front of the first copy
        struct Part { uint64_t cells = 0; uint32_t max_n = 0, max_m = 0; bool bad = false; std::vector<uint64_t> longs; };
        const unsigned nT = N >= 262144 ? std::max(1u, std::min(8u, std::thread::hardware_concurrency())) : 1u;
        std::vector<Part> part(nT);
        auto scan = [&](unsigned t) {
            Part &P = part[t];
            for (uint64_t p = N * t / nT, e = N * (t + 1) / nT; p < e; ++p) {
                const uint64_t n = (uint64_t)(to[p + 1] - to[p]), m = (uint64_t)(po[p + 1] - po[p]);
                if (n == 0 || m == 0) { P.bad = true; return; }
                P.cells += (n + 1) * (m + 1);
                if (wideScores || !batch_eligible(n, m)) { P.longs.push_back(p); continue; }
                P.max_n = std::max<uint32_t>(P.max_n, (uint32_t)n);
                P.max_m = std::max<uint32_t>(P.max_m, (uint32_t)m);
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nT; ++t) th.emplace_back(scan, t);
        scan(0);
        for (auto &t : th) t.join();
        for (const Part &P : part) {
            if (P.bad) return SA_ERR_ARGUMENT;
            cells += P.cells; max_n = std::max(max_n, P.max_n); max_m = std::max(max_m, P.max_m);
            longPairs.insert(longPairs.end(), P.longs.begin(), P.longs.end());
        }
    }
    int rc = upload_scoring(ctx, sc, ctx->stream);
    if (rc) return rc;
    reset_timing(ctx);
    sa_timing tm{};
    tm.cells = cells;
    uint64_t d2hBytes = 0;

    if (max_m > 0) {
        BatchClassTable T;
        if (!build_class_table(max_n, max_m, fits_s16(sc, max_n, max_m), line16_ok(sc), &T)) return SA_ERR_ARGUMENT;
        SA_TRY(ctx->errflag.reserve(16), SA_ERR_MEMORY);
        SA_TRY(cudaMemsetAsync(ctx->errflag.p, 0, 4, ctx->stream), SA_ERR_LAUNCH);
        auto validate = [&](const void *d_text, uint64_t tbytes, const void *d_pat, uint64_t pbytes, cudaStream_t vst) {
            int *flag = ctx->errflag.as<int>();
            validate_residues_kernel<<<(unsigned)std::min<uint64_t>(4 * ctx->sms, (tbytes / 4096) + 1), 256, 0, vst>>>((const uint8_t *)d_text, tbytes, (unsigned)sc->alphabet_size, flag);
            validate_residues_kernel<<<(unsigned)std::min<uint64_t>(4 * ctx->sms, (pbytes / 4096) + 1), 256, 0, vst>>>((const uint8_t *)d_pat, pbytes, (unsigned)sc->alphabet_size, flag);
            ctx->timing.kernel_launches += 2;
        };
        // chunk size: bounded by the direction budget and by ~1/8 of the batch for copy/compute overlap
        const double perPair = (double)batch_dirs_bound(T, 1 << 20) / (double)(1 << 20) * 4.0;
        uint64_t chunk = std::max<uint64_t>(32, (uint64_t)((double)ctx->host_dirs_budget / NSLOT / perPair));
        uint64_t nChunks = 8;
        if (const char *e = std::getenv("SA_HOST_CHUNKS")) nChunks = (uint64_t)std::max(1, std::atoi(e));
        chunk = std::min<uint64_t>(chunk, std::max<uint64_t>(4096, (N + nChunks - 1) / nChunks));
        cudaEventRecord(ctx->ev[0], ctx->stream);
        for (auto &s : ctx->slot) cudaStreamWaitEvent(s.stream, ctx->ev[0], 0);
        // Staged pipeline (default; SA_HOST_PIPELINE=slots selects the older one below): the shape of the device-resident
        // pipeline with the copies on their own streams.  All fills run on ONE stream, all tracebacks on ctx->stream
        // (one block per SM next to the fill blocks of the following chunk), host->device copies on a third stream and
        // device->host copies on a fourth; everything is ordered by events, the host only waits at the end.
        const char *hp = std::getenv("SA_HOST_PIPELINE");
        const bool staged = !(hp && std::strcmp(hp, "slots") == 0) && N >= 8192;
        bool packedOut = false;       // the strings came back packed: aln_off already counts from the start of the arenas
        if (staged) {
            const cudaStream_t stIn = ctx->slot[0].stream, stOut = ctx->slot[1].stream, stFill = ctx->slot[2].stream;
            int NIO = 4;          // I/O buffer sets in flight (copy in | fill | traceback + pack | copy out)
            if (const char *e = std::getenv("SA_HOST_IOSETS")) NIO = std::max(3, std::min(NIO_MAX, std::atoi(e)));
            // two direction sets of the device pipeline's size (the slot pipeline's third of 8 GB made 15 chunks of 1 M
            // pairs, whose fills add up to 29.4 ms against 25.4 ms in 5-8 chunks)
            chunk = std::max<uint64_t>(32, (uint64_t)((double)ctx->dirs_budget / perPair));
            chunk = std::min<uint64_t>(chunk, std::max<uint64_t>(4096, (N + nChunks - 1) / nChunks));
            chunk = std::min<uint64_t>(chunk, N);
            const uint64_t nc = (N + chunk - 1) / chunk;
            chunk = (N + nc - 1) / nc;                                   // equal chunks
            // Chunk boundaries.  Nothing overlaps the first chunk's copy in or the last chunk's traceback and copy out, so
            // big batches start and end with small chunks: 1/32, 2/32, 3/32 of the batch, eighths in the middle, then
            // 3/32, 2/32, 1/32 (SA_HOST_CHUNKS selects equal chunks instead).
            std::vector<uint64_t> bounds{0};
            if (!std::getenv("SA_HOST_CHUNKS") && N >= 65536 && chunk >= N / 8) {
                // (fewer, larger chunks for smaller batches -- a chunk's fill should stay above ~1 ms, below that the class
                // kernels' tails and the launch gaps show: 125 k pairs 5.95 -> 5.3 ms, 250 k pairs 10.1 -> 9.3 ms)
                std::vector<int> w{1, 2, 3, 4, 4, 4, 4, 4, 3, 2, 1};
                if (N < 200000) w = {2, 3, 3, 2};
                else if (N < 400000) w = {1, 2, 3, 3, 3, 2, 1};
                if (const char *e = std::getenv("SA_HOST_SCHEDULE")) {          // e.g. "1,2,4,5,5,5,5,3,2": parts of the batch
                    std::vector<int> u;
                    for (const char *q = e; *q;) { const int v = std::atoi(q); if (v > 0) u.push_back(v); while (*q && *q != ',') ++q; if (*q) ++q; }
                    if (!u.empty()) w = u;
                }
                uint64_t acc = 0, tot = 0;
                for (int k : w) tot += k;
                for (int k : w) { acc += k; bounds.push_back(acc == tot ? N : N * acc / tot); }
                chunk = 0;
                for (size_t k = 1; k < bounds.size(); ++k) chunk = std::max(chunk, bounds[k] - bounds[k - 1]);
            } else {
                for (uint64_t first = chunk; first < N; first += chunk) bounds.push_back(first);
                bounds.push_back(N);
            }
            const uint64_t nChunk = bounds.size() - 1;
            uint64_t maxT = 0, maxP = 0;
            for (uint64_t k = 0; k < nChunk; ++k) {
                maxT = std::max<uint64_t>(maxT, (uint64_t)(to[bounds[k + 1]] - to[bounds[k]]));
                maxP = std::max<uint64_t>(maxP, (uint64_t)(po[bounds[k + 1]] - po[bounds[k]]));
            }
            // buffers at their final size before anything is in flight (growing one frees it, which waits for the device)
            for (int q = 0; q < NIO; ++q) { Slot &s = ctx->slot[q];
                SA_TRY(s.text.reserve(maxT + 16), SA_ERR_MEMORY);
                SA_TRY(s.pattern.reserve(maxP + 16), SA_ERR_MEMORY);
                SA_TRY(s.toff.reserve((chunk + 1) * 8), SA_ERR_MEMORY);
                SA_TRY(s.poff.reserve((chunk + 1) * 8), SA_ERR_MEMORY);
                SA_TRY(s.results.reserve(chunk * sizeof(sa_result)), SA_ERR_MEMORY);
                SA_TRY(s.alnoff.reserve(chunk * 8), SA_ERR_MEMORY);
                SA_TRY(s.outT.reserve(maxT + maxP + 16), SA_ERR_MEMORY);
                SA_TRY(s.outP.reserve(maxT + maxP + 16), SA_ERR_MEMORY);
                SA_TRY(s.fill.reserve(chunk * 12 + 64), SA_ERR_MEMORY);
                if (out->stats) SA_TRY(s.stats.reserve(chunk * 8 + 64), SA_ERR_MEMORY);
            }
            for (int k = 0; k < 2; ++k) {
                SA_TRY(ctx->pdirs[k].reserve(batch_dirs_bound(T, chunk) * 4), SA_ERR_MEMORY);
                SA_TRY(ctx->psort[k].reserve(batch_sort_bytes(chunk)), SA_ERR_MEMORY);
            }
            // strings packed on the device before the copy (SA_HOST_PACK=0: whole slots as they are).  Long members write
            // into their slots afterwards, which the packed layout would overlap: such batches stay unpacked.
            const char *pe = std::getenv("SA_HOST_PACK");
            const bool pack = longPairs.empty() && !(pe && pe[0] == '0');
            packedOut = pack;
            if (pack) {
                for (int q = 0; q < NIO; ++q) { Slot &s = ctx->slot[q];
                    SA_TRY(s.packT.reserve(maxT + maxP + 16), SA_ERR_MEMORY);
                    SA_TRY(s.packP.reserve(maxT + maxP + 16), SA_ERR_MEMORY);
                    SA_TRY(s.noff.reserve((chunk + chunk / PACK_BLOCK + 4) * 8), SA_ERR_MEMORY);
                }
                SA_TRY(ctx->pin.reserve(256), SA_ERR_MEMORY);
                SA_TRY(ctx->packstate.reserve(64), SA_ERR_MEMORY);
                SA_TRY(cudaMemsetAsync(ctx->packstate.as<unsigned long long>(), 0, 8, ctx->stream), SA_ERR_LAUNCH);   // running total
            }
            uint64_t hostBase = 0;
            auto drain = [&](uint64_t k) -> int {          // device->host copies of chunk k, sized by its packed total
                Slot &s = ctx->slot[k % NIO];
                const uint64_t first = bounds[k], count = bounds[k + 1] - first;
                SA_TRY(cudaEventSynchronize(s.packed), SA_ERR_LAUNCH);
                const uint64_t total = reinterpret_cast<volatile unsigned long long *>(ctx->pin.p)[k % NIO];
                if (hostBase + total > out->arena_capacity) return SA_ERR_CAPACITY;
                SA_TRY(cudaMemcpyAsync(out->results + first, s.results.p, count * sizeof(sa_result), cudaMemcpyDeviceToHost, stOut), SA_ERR_COPY);
                SA_TRY(cudaMemcpyAsync(out->aln_off + first, s.alnoff.p, count * 8, cudaMemcpyDeviceToHost, stOut), SA_ERR_COPY);
                if (out->stats) { SA_TRY(cudaMemcpyAsync(out->stats + 2 * first, s.stats.p, count * 8, cudaMemcpyDeviceToHost, stOut), SA_ERR_COPY); d2hBytes += count * 8; }
                if (total) {
                    SA_TRY(cudaMemcpyAsync(out->aligned_text + hostBase, s.packT.p, total, cudaMemcpyDeviceToHost, stOut), SA_ERR_COPY);
                    SA_TRY(cudaMemcpyAsync(out->aligned_pattern + hostBase, s.packP.p, total, cudaMemcpyDeviceToHost, stOut), SA_ERR_COPY);
                }
                SA_TRY(cudaEventRecord(s.out, stOut), SA_ERR_LAUNCH);
                hostBase += total;
                d2hBytes += count * (sizeof(sa_result) + 8) + 2 * total;
                return SA_OK;
            };
            uint64_t c = 0;
            for (; c < nChunk; ++c) {
                Slot &s = ctx->slot[c % NIO];
                const int d = (int)(c & 1);
                const uint64_t first = bounds[c], count = bounds[c + 1] - first;
                const int64_t tb = to[first], pb = po[first];
                const uint64_t tbytes = (uint64_t)(to[first + count] - tb), pbytes = (uint64_t)(po[first + count] - pb);
                const uint64_t arena = tbytes + pbytes;
                if (c >= (uint64_t)NIO) cudaStreamWaitEvent(stIn, s.out, 0);          // the set's previous chunk has drained
                SA_TRY(cudaMemcpyAsync(s.text.p, b->text + tb, tbytes, cudaMemcpyHostToDevice, stIn), SA_ERR_COPY);
                SA_TRY(cudaMemcpyAsync(s.pattern.p, b->pattern + pb, pbytes, cudaMemcpyHostToDevice, stIn), SA_ERR_COPY);
                SA_TRY(cudaMemcpyAsync(s.toff.p, to + first, (count + 1) * 8, cudaMemcpyHostToDevice, stIn), SA_ERR_COPY);
                SA_TRY(cudaMemcpyAsync(s.poff.p, po + first, (count + 1) * 8, cudaMemcpyHostToDevice, stIn), SA_ERR_COPY);
                validate(s.text.p, tbytes, s.pattern.p, pbytes, stIn);
                SA_TRY(cudaEventRecord(s.in, stIn), SA_ERR_LAUNCH);
                cudaStreamWaitEvent(stFill, s.in, 0);
                if (c >= 2) cudaStreamWaitEvent(stFill, ctx->evTrace[d], 0);  // direction set d is free again
                sa_batch cb{count, s.text.as<uint8_t>() - tb, s.toff.as<int64_t>(), s.pattern.as<uint8_t>() - pb, s.poff.as<int64_t>()};
                char *oT = s.outT.as<char>() - (tb + pb), *oP = s.outP.as<char>() - (tb + pb);
                rc = enqueue_batch(ctx, sc, &cb, s.results.as<sa_result>(), s.alnoff.as<uint64_t>(), out->stats ? s.stats.as<uint32_t>() : nullptr,
                                   oT, oP, max_n, max_m, ctx->pdirs[d].as<uint32_t>(), ctx->pdirs[d].cap / 4, s.fill.p, ctx->psort[d].p, &ctx->snapbuf,
                                   stFill, 0, (uint32_t)count, ctx->stream, ctx->evFill[d], /*tbShare=*/c + 1 < nChunk);
                if (rc) return rc;
                SA_TRY(cudaEventRecord(ctx->evTrace[d], ctx->stream), SA_ERR_LAUNCH);
                if (!pack) {
                    cudaStreamWaitEvent(stOut, ctx->evTrace[d], 0);
                    SA_TRY(cudaMemcpyAsync(out->results + first, s.results.p, count * sizeof(sa_result), cudaMemcpyDeviceToHost, stOut), SA_ERR_COPY);
                    SA_TRY(cudaMemcpyAsync(out->aln_off + first, s.alnoff.p, count * 8, cudaMemcpyDeviceToHost, stOut), SA_ERR_COPY);
                    if (out->stats) { SA_TRY(cudaMemcpyAsync(out->stats + 2 * first, s.stats.p, count * 8, cudaMemcpyDeviceToHost, stOut), SA_ERR_COPY); d2hBytes += count * 8; }
                    SA_TRY(cudaMemcpyAsync(out->aligned_text + (tb - to[0]) + (pb - po[0]), s.outT.p, arena, cudaMemcpyDeviceToHost, stOut), SA_ERR_COPY);
                    SA_TRY(cudaMemcpyAsync(out->aligned_pattern + (tb - to[0]) + (pb - po[0]), s.outP.p, arena, cudaMemcpyDeviceToHost, stOut), SA_ERR_COPY);
                    SA_TRY(cudaEventRecord(s.out, stOut), SA_ERR_LAUNCH);
                    d2hBytes += count * (sizeof(sa_result) + 8) + 2 * arena;
                    continue;
                }
                // pack the chunk's strings (behind its traceback, next to the following fill), then let the host -- two
                // chunks behind the enqueue front, so that the fills never wait for it -- copy exactly the packed bytes
                CompactArgs K{};
                K.results = s.results.as<sa_result>(); K.aln_off = s.alnoff.as<uint64_t>(); K.count = (uint32_t)count;
                K.srcT = oT; K.srcP = oP; K.dstT = s.packT.as<char>(); K.dstP = s.packP.as<char>();
                K.noff = s.noff.as<unsigned long long>(); K.running = ctx->packstate.as<unsigned long long>();
                K.host_total = reinterpret_cast<volatile unsigned long long *>(ctx->pin.p) + (c % NIO);
                K.block_base = K.noff + chunk;
                pack_local_kernel<<<(unsigned)((count + PACK_BLOCK - 1) / PACK_BLOCK), 128, 0, ctx->stream>>>(K);
                pack_bases_kernel<<<1, 128, 0, ctx->stream>>>(K);
                compact_copy_kernel<<<(unsigned)((count * 32 + 127) / 128), 128, 0, ctx->stream>>>(K);
                SA_TRY(cudaGetLastError(), SA_ERR_LAUNCH);
                ctx->timing.kernel_launches += 3;
                SA_TRY(cudaEventRecord(s.packed, ctx->stream), SA_ERR_LAUNCH);
                if (c >= 2) { rc = drain(c - 2); if (rc) return rc; }
            }
            if (pack) for (uint64_t k = c >= 2 ? c - 2 : 0; k < c; ++k) { rc = drain(k); if (rc) return rc; }
            cudaStreamWaitEvent(ctx->stream, ctx->slot[(c - 1) % NIO].out, 0);      // stOut is in order: the last chunk's drain ends it
        }
        int si = 0;
        for (uint64_t first = 0; !staged && first < N; first += chunk, si = (si + 1) % NSLOT) {
            Slot &s = ctx->slot[si];
            const uint64_t count = std::min<uint64_t>(chunk, N - first);
            const int64_t tb = to[first], pb = po[first];
            const uint64_t tbytes = (uint64_t)(to[first + count] - tb), pbytes = (uint64_t)(po[first + count] - pb);
            const uint64_t arena = tbytes + pbytes;
            // wait until this slot's previous chunk has fully drained
            SA_TRY(cudaEventSynchronize(s.done), SA_ERR_LAUNCH);
            SA_TRY(s.text.reserve(tbytes + 16), SA_ERR_MEMORY);
            SA_TRY(s.pattern.reserve(pbytes + 16), SA_ERR_MEMORY);
            SA_TRY(s.toff.reserve((count + 1) * 8), SA_ERR_MEMORY);
            SA_TRY(s.poff.reserve((count + 1) * 8), SA_ERR_MEMORY);
            SA_TRY(s.results.reserve(count * sizeof(sa_result)), SA_ERR_MEMORY);
            SA_TRY(s.alnoff.reserve(count * 8), SA_ERR_MEMORY);
            SA_TRY(s.outT.reserve(arena + 16), SA_ERR_MEMORY);
            SA_TRY(s.outP.reserve(arena + 16), SA_ERR_MEMORY);
            SA_TRY(s.dirs.reserve(batch_dirs_bound(T, count) * 4), SA_ERR_MEMORY);
            SA_TRY(s.fill.reserve(count * 12 + 64), SA_ERR_MEMORY);
            SA_TRY(s.order.reserve(batch_sort_bytes(count)), SA_ERR_MEMORY);
            if (out->stats) SA_TRY(s.stats.reserve(count * 8 + 64), SA_ERR_MEMORY);
            SA_TRY(cudaMemcpyAsync(s.text.p, b->text + tb, tbytes, cudaMemcpyHostToDevice, s.stream), SA_ERR_COPY);
            SA_TRY(cudaMemcpyAsync(s.pattern.p, b->pattern + pb, pbytes, cudaMemcpyHostToDevice, s.stream), SA_ERR_COPY);
            SA_TRY(cudaMemcpyAsync(s.toff.p, to + first, (count + 1) * 8, cudaMemcpyHostToDevice, s.stream), SA_ERR_COPY);
            SA_TRY(cudaMemcpyAsync(s.poff.p, po + first, (count + 1) * 8, cudaMemcpyHostToDevice, s.stream), SA_ERR_COPY);
            validate(s.text.p, tbytes, s.pattern.p, pbytes, s.stream);
            // device-side offsets are relative to the chunk: rebase with a tiny kernel-free trick --
            // the kernels subtract nothing, so pass pointers shifted by the chunk base instead.
            sa_batch cb{count, s.text.as<uint8_t>() - tb, s.toff.as<int64_t>(), s.pattern.as<uint8_t>() - pb, s.poff.as<int64_t>()};
            // output slots are addressed by text_off+pattern_off as well: shift the arenas likewise
            char *oT = s.outT.as<char>() - (tb + pb), *oP = s.outP.as<char>() - (tb + pb);
            // members too long for the batch kernels are skipped by the device-side classifier
            // (and aligned one by one below)
            rc = enqueue_batch(ctx, sc, &cb, s.results.as<sa_result>(), s.alnoff.as<uint64_t>(), out->stats ? s.stats.as<uint32_t>() : nullptr,
                               oT, oP, max_n, max_m, s.dirs.as<uint32_t>(), s.dirs.cap / 4, s.fill.p, s.order.p, &s.snap, s.stream, 0, (uint32_t)count,
                               nullptr, nullptr, /*tbShare=*/false);      // (measured: sharing does not pay with three slots in flight)
            if (rc) return rc;
            SA_TRY(cudaMemcpyAsync(out->results + first, s.results.p, count * sizeof(sa_result), cudaMemcpyDeviceToHost, s.stream), SA_ERR_COPY);
            SA_TRY(cudaMemcpyAsync(out->aln_off + first, s.alnoff.p, count * 8, cudaMemcpyDeviceToHost, s.stream), SA_ERR_COPY);
            if (out->stats) { SA_TRY(cudaMemcpyAsync(out->stats + 2 * first, s.stats.p, count * 8, cudaMemcpyDeviceToHost, s.stream), SA_ERR_COPY); d2hBytes += count * 8; }
            SA_TRY(cudaMemcpyAsync(out->aligned_text + (tb - to[0]) + (pb - po[0]), s.outT.p, arena, cudaMemcpyDeviceToHost, s.stream), SA_ERR_COPY);
            SA_TRY(cudaMemcpyAsync(out->aligned_pattern + (tb - to[0]) + (pb - po[0]), s.outP.p, arena, cudaMemcpyDeviceToHost, s.stream), SA_ERR_COPY);
            SA_TRY(cudaEventRecord(s.done, s.stream), SA_ERR_LAUNCH);
            d2hBytes += count * (sizeof(sa_result) + 8) + 2 * arena;
        }
        for (auto &s : ctx->slot) SA_TRY(cudaStreamSynchronize(s.stream), SA_ERR_LAUNCH);
        cudaEventRecord(ctx->ev[4], ctx->stream);
        SA_TRY(cudaStreamSynchronize(ctx->stream), SA_ERR_LAUNCH);
        {
            int bad = 0;
            SA_TRY(cudaMemcpy(&bad, ctx->errflag.p, 4, cudaMemcpyDeviceToHost), SA_ERR_COPY);
            if (bad) return SA_ERR_ARGUMENT;       // a residue >= alphabet_size somewhere in the batch
        }
        tm.total_us = ev_us(ctx->ev[0], ctx->ev[4]);
        {
            sa_timing k{};
            sa_last_timing(ctx, &k);
            tm.fill_us = k.fill_us; tm.traceback_us = k.traceback_us; tm.kernel_launches = k.kernel_launches;
        }
        // unpacked aln_off values are absolute (text_off+pattern_off based); make them relative to the arenas.  Packed
        // ones already are: compact_copy_kernel rewrote them to positions that start at 0.
        const uint64_t base0 = (uint64_t)(to[0] + po[0]);
        if (base0 && !packedOut) for (uint64_t p = 0; p < N; ++p) out->aln_off[p] -= base0;
    }
    // Long members: one by one through sa_align -- but several at a time when there are several: a pair of a few
    // thousand residues fills a fraction of the GPU (its cooperative launch has a handful of blocks), so up to
    // SA_LONG_WORKERS (default 8, memory permitting) sub-contexts run them side by side, one host thread each.
    auto align_member = [&](sa_context *c, uint64_t p, sa_timing *k) -> int {
        const uint64_t n = (uint64_t)(to[p + 1] - to[p]), m = (uint64_t)(po[p + 1] - po[p]);
        const uint64_t slot = (uint64_t)(to[p] - to[0]) + (uint64_t)(po[p] - po[0]);
        const int r = sa_align(c, sc, b->text + to[p], n, b->pattern + po[p], m, &out->results[p],
                               out->aligned_text + slot, out->aligned_pattern + slot, n + m);
        if (r) return r;
        out->aln_off[p] = slot;
        if (out->stats) {
            sa_stats st{};
            sa_last_stats(c, &st);
            out->stats[2 * p] = (uint32_t)st.identity; out->stats[2 * p + 1] = (uint32_t)st.gaps;
        }
        sa_last_timing(c, k);
        return SA_OK;
    };
    size_t nWorkers = 1;
    if (longPairs.size() >= 2) {
        nWorkers = 8;
        if (const char *e = std::getenv("SA_LONG_WORKERS")) nWorkers = (size_t)std::max(1, std::atoi(e));
        uint64_t worst = 0;                                     // direction bytes of the largest member
        for (uint64_t p : longPairs) worst = std::max<uint64_t>(worst, (uint64_t)(to[p + 1] - to[p] + 64) * (uint64_t)(po[p + 1] - po[p] + 512) / 4);
        size_t freeB = 0, totalB = 0;
        cudaMemGetInfo(&freeB, &totalB);
        nWorkers = std::min<size_t>({nWorkers, longPairs.size(), std::max<size_t>(1, (size_t)(0.6 * (double)freeB / (double)(worst + (64u << 20))))});
    }
    if (nWorkers <= 1) {
        for (uint64_t p : longPairs) {
            sa_timing k{};
            rc = align_member(ctx, p, &k);
            if (rc) return rc;
            tm.total_us += k.total_us; tm.fill_us += k.fill_us; tm.traceback_us += k.traceback_us;
            tm.kernel_launches += k.kernel_launches;
        }
    } else {
        while (ctx->workers.size() < nWorkers) {
            sa_context *w = nullptr;
            if (sa_create(ctx->device, &w) != SA_OK) break;
            ctx->workers.push_back(w);
        }
        nWorkers = std::min(nWorkers, ctx->workers.size());
        if (nWorkers == 0) return SA_ERR_MEMORY;
        std::atomic<size_t> next(0);
        std::atomic<int> firstErr(0);
        std::mutex mu;
        const auto t0 = std::chrono::steady_clock::now();
        auto work = [&](size_t w) {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= longPairs.size() || firstErr.load()) break;
                sa_timing k{};
                const int r = align_member(ctx->workers[w], longPairs[i], &k);
                if (r) { int z = 0; firstErr.compare_exchange_strong(z, r); break; }
                std::lock_guard<std::mutex> g(mu);
                tm.fill_us += k.fill_us; tm.traceback_us += k.traceback_us; tm.kernel_launches += k.kernel_launches;
            }
        };
        std::vector<std::thread> th;
        for (size_t w = 0; w < nWorkers; ++w) th.emplace_back(work, w);
        for (auto &t : th) t.join();
        if (firstErr.load()) return firstErr.load();
        tm.total_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    }
    tm.h2d_bytes = (uint64_t)(to[N] - to[0]) + (uint64_t)(po[N] - po[0]) + 16 * (N + 1);
    tm.d2h_bytes = d2hBytes;
    ctx->timing = tm;
    ctx->timing_dirty = false;
    return SA_OK;
}

// --------------------------------------------------------------- multi-GPU dispatcher (one process)
namespace {
std::mutex g_multi_mu;
sa_context *g_multi_ctx[64] = {};
sa_timing g_multi_timing[8] = {};
}

int sa_options_from_env(sa_options *o)
{
    if (!o) return SA_ERR_ARGUMENT;
    std::memset(o, 0, sizeof *o);
    o->n_devices = 1;
    const char *e = std::getenv("SA_DEVICES");
    if (!e || !*e) return SA_OK;
    std::vector<int> v;
    for (const char *q = e; *q;) {
        char *end = nullptr;
        const long d = std::strtol(q, &end, 10);
        if (end == q) return SA_ERR_ARGUMENT;
        v.push_back((int)d);
        q = end;
        if (*q == ',') ++q; else if (*q) return SA_ERR_ARGUMENT;
    }
    if (v.size() == 1 && std::strchr(e, ',') == nullptr && v[0] >= 1) {          // a count
        const int cnt = v[0];
        v.clear();
        for (int d = 0; d < cnt; ++d) v.push_back(d);
    }
    if (v.empty() || v.size() > 8) return SA_ERR_ARGUMENT;
    o->n_devices = (int32_t)v.size();
    for (size_t k = 0; k < v.size(); ++k) o->devices[k] = v[k];
    return SA_OK;
}

int sa_multi_last_timing(int k, sa_timing *out)
{
    if (k < 0 || k >= 8 || !out) return SA_ERR_ARGUMENT;
    std::lock_guard<std::mutex> g(g_multi_mu);
    *out = g_multi_timing[k];
    return SA_OK;
}

int sa_align_batch_multi(const sa_options *opt, const sa_scoring *sc, const sa_batch *b, sa_batch_out *out)
{
    if (!sc || !b || !out || !b->text_off || !b->pattern_off || !out->results || !out->aln_off) return SA_ERR_ARGUMENT;
    sa_options one{};
    one.n_devices = 1;
    if (!opt || opt->n_devices == 0) opt = &one;
    const int D = opt->n_devices;
    if (D < 1 || D > 8) return SA_ERR_ARGUMENT;
    const int have = sa_device_count();
    if (have < 1) return SA_ERR_NO_DEVICE;
    for (int k = 0; k < D; ++k) {
        if (opt->devices[k] < 0 || opt->devices[k] >= have || opt->devices[k] >= 64) return SA_ERR_NO_DEVICE;
        for (int j = 0; j < k; ++j) if (opt->devices[j] == opt->devices[k]) return SA_ERR_ARGUMENT;
    }
    const uint64_t N = b->n_pairs;
    if (N == 0) return SA_OK;
    std::lock_guard<std::mutex> g(g_multi_mu);
    for (int k = 0; k < D; ++k) {
        sa_context *&c = g_multi_ctx[opt->devices[k]];
        if (!c) { const int rc = sa_create(opt->devices[k], &c); if (rc) { c = nullptr; return rc; } }
    }
    const int64_t *to = b->text_off, *po = b->pattern_off;
    if ((uint64_t)(to[N] - to[0] + po[N] - po[0]) > out->arena_capacity) return SA_ERR_CAPACITY;
    std::vector<uint64_t> first((size_t)D + 1);
    int rc = sa_partition_batch(to, po, N, D, first.data());
    if (rc) return rc;
    std::vector<int> status((size_t)D, SA_OK);
    auto work = [&](int k) {
        const uint64_t f = first[k], cnt = first[k + 1] - f;
        g_multi_timing[k] = sa_timing{};
        if (cnt == 0) return;
        // the device's range as a batch of its own: offsets stay absolute (sa_align_batch rebases), the output arenas
        // start at the range's own slot base, and aln_off comes back relative to that base
        const uint64_t base = (uint64_t)(to[f] - to[0]) + (uint64_t)(po[f] - po[0]);
        sa_batch sb{cnt, b->text, to + f, b->pattern, po + f};
        sa_batch_out so{out->results + f, out->aln_off + f, out->aligned_text + base, out->aligned_pattern + base,
                        (uint64_t)(to[f + cnt] - to[f]) + (uint64_t)(po[f + cnt] - po[f]), out->stats ? out->stats + 2 * f : nullptr};
        sa_context *c = g_multi_ctx[opt->devices[k]];
        status[k] = sa_align_batch(c, sc, &sb, &so);
        if (status[k] == SA_OK) {
            if (base) for (uint64_t p = 0; p < cnt; ++p) so.aln_off[p] += base;
            sa_last_timing(c, &g_multi_timing[k]);
        }
    };
    if (D == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int k = 1; k < D; ++k) th.emplace_back(work, k);
        work(0);
        for (auto &t : th) t.join();
    }
    for (int k = 0; k < D; ++k) if (status[k] != SA_OK) return status[k];
    return SA_OK;
}

int sa_partition_batch(const int64_t *to, const int64_t *po, uint64_t N, int world, uint64_t *first)
{
    if (!to || !po || !first || world < 1) return SA_ERR_ARGUMENT;
    // prefix of cells; rank r takes the pairs whose cumulative cell count falls in its 1/world share
    long double total = 0;
    for (uint64_t p = 0; p < N; ++p) total += (long double)(to[p + 1] - to[p] + 1) * (long double)(po[p + 1] - po[p] + 1);
    first[0] = 0;
    long double acc = 0;
    uint64_t p = 0;
    for (int r = 1; r < world; ++r) {
        const long double target = total * r / world;
        while (p < N && acc + (long double)(to[p + 1] - to[p] + 1) * (long double)(po[p + 1] - po[p] + 1) / 2 <= target) {
            acc += (long double)(to[p + 1] - to[p] + 1) * (long double)(po[p + 1] - po[p] + 1);
            ++p;
        }
        first[r] = p;
    }
    first[world] = N;
    return SA_OK;
}

} // extern "C"

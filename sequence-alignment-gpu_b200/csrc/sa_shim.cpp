// sa_shim.cpp -- C++ side of the drop-in boundary: SequenceAlignment::alignSequenceGPU with the
// reference's ownership and error conventions, implemented on the C ABI (include/sa_b200.h).
//   reference: alignSequenceGPU.cu:463-653 (entry), :402-403 (new char[] outputs),
//              :541-546 / :588-594 (stdout messages + return 1), :613-626 (BENCHMARK return value).
// Thread safety: the reference keeps all state per call; here the calls share cached contexts (streams, growing
// workspaces), so every entry point holds one lock for its whole duration -- concurrent callers are serialised, not
// corrupted.  Devices: SA_DEVICES ("0,1,2,3" or a count) selects the GPUs of alignSequenceGPUBatch; single pairs run
// on the first of them (the reference hard-codes device 0, alignSequenceGPU.cu:476).
#include "../../include/SequenceAlignment.hpp"
#include "../../include/sa_b200.h"

#include <algorithm>
#include <cstring>
#include <iostream>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

namespace {

std::mutex g_mu;                 // held for the whole of every entry point
sa_context *g_ctx = nullptr;     // single-pair context, on the first device of SA_DEVICES
int g_ctx_device = -1;

sa_context *context_locked()
{
    sa_options o;
    if (sa_options_from_env(&o) != SA_OK) { o.n_devices = 1; o.devices[0] = 0; }
    if (g_ctx && g_ctx_device != o.devices[0]) { sa_destroy(g_ctx); g_ctx = nullptr; }
    if (!g_ctx) {
        if (sa_create(o.devices[0], &g_ctx) != SA_OK) g_ctx = nullptr;
        g_ctx_device = o.devices[0];
    }
    return g_ctx;
}

sa_scoring scoringOf(const SequenceAlignment::Request &rq)
{
    sa_scoring sc;
    sc.mode = rq.alignmentType == SequenceAlignment::GLOBAL ? SA_GLOBAL : SA_LOCAL;
    sc.alphabet_size = rq.alphabetSize;
    sc.score_matrix = rq.scoreMatrix;
    sc.gap = rq.gapPenalty;
    sc.alphabet = rq.alphabet;
    return sc;
}

// The reference knows two failures and prints both to stdout: MEM_ERROR (:541-546) and the copy error (:588-594).
// Everything else this library can refuse (bad arguments, score range, no device, launch failure) has no reference
// message: it goes to stderr with the library's own text.  The return value is 1 either way (mainDriver.cu:22).
uint64_t fail(int status)
{
    if (status == SA_ERR_MEMORY || status == SA_ERR_CAPACITY) std::cout << SequenceAlignment::MEM_ERROR;
    else if (status == SA_ERR_COPY) std::cout << "error: could not copy from device memory\n";
    else std::cerr << "error: " << sa_status_string(status) << "\n";
    return 1;
}

// page-locked staging of the batch entry, kept between calls (grown on demand)
struct Pinned {
    void *p = nullptr; uint64_t cap = 0;
    void *reserve(uint64_t bytes)
    {
        if (bytes <= cap) return p;
        if (p) sa_host_free(p);
        cap = bytes + bytes / 4 + 4096;
        p = sa_host_alloc(cap);
        if (!p) cap = 0;
        return p;
    }
};
Pinned g_pin[8];     // text, pattern, text offsets, pattern offsets, results, aln_off, aligned text, aligned pattern

} // namespace

uint64_t SequenceAlignment::alignSequenceGPU(const Request &rq, Response *rs)
{
    // anything but GLOBAL / LOCAL: both reference paths silently do nothing (alignSequenceCPU.cpp:318-328)
    if (rq.alignmentType != GLOBAL && rq.alignmentType != LOCAL) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    sa_context *ctx = context_locked();
    if (!ctx) return fail(SA_ERR_NO_DEVICE);
    // the reference allocates 2*text (alignSequenceGPU.cu:402-403), enough only when text >= pattern
    const uint64_t cap = std::max<uint64_t>(2 * rq.textNumBytes, rq.textNumBytes + rq.patternNumBytes);
    char *outT = new (std::nothrow) char[cap];
    char *outP = new (std::nothrow) char[cap];
    if (!outT || !outP) { delete[] outT; delete[] outP; return fail(SA_ERR_MEMORY); }
    const sa_scoring sc = scoringOf(rq);
    sa_result r;
    const int st = sa_align(ctx, &sc, reinterpret_cast<const uint8_t *>(rq.textBytes), rq.textNumBytes,
                            reinterpret_cast<const uint8_t *>(rq.patternBytes), rq.patternNumBytes, &r, outT, outP, cap);
    if (st != SA_OK) { delete[] outT; delete[] outP; return fail(st); }
    delete[] rs->alignedTextBytes;        // the reference leaks these when a Response is reused
    delete[] rs->alignedPatternBytes;
    rs->alignedTextBytes = outT;
    rs->alignedPatternBytes = outP;
    rs->numAlignmentBytes = r.aln_len;
    rs->startInAlignedText = r.start_text;
    rs->startInAlignedPattern = r.start_pattern;
    rs->score = r.score;
#ifdef BENCHMARK
    // what the reference times under its BENCHMARK switch: the kernels and the device-to-host copies, no allocation,
    // no host-to-device copy, no traceback (alignSequenceGPU.cu:555-558, 613-626)
    sa_timing t;
    sa_last_timing(ctx, &t);
    return std::max<uint64_t>(1, (uint64_t)(t.fill_us + t.d2h_us));
#else
    return 0;
#endif
}

uint64_t SequenceAlignment::alignSequenceGPUFillMicros(const Request &rq, Response *rs)
{
    std::lock_guard<std::mutex> lk(g_mu);
    sa_context *ctx = context_locked();
    if (!ctx) return fail(SA_ERR_NO_DEVICE);
    const sa_scoring sc = scoringOf(rq);
    int32_t score = 0;
    uint64_t arg = 0;
    const int st = sa_fill_only(ctx, &sc, reinterpret_cast<const uint8_t *>(rq.textBytes), rq.textNumBytes,
                                reinterpret_cast<const uint8_t *>(rq.patternBytes), rq.patternNumBytes, &score, &arg);
    if (st != SA_OK) return fail(st);
    rs->score = score;
    sa_timing t;
    sa_last_timing(ctx, &t);
    return std::max<uint64_t>(1, (uint64_t)t.fill_us);
}

uint64_t SequenceAlignment::alignSequenceGPUBatch(const Request *rq, Response *rs, uint64_t n)
{
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    sa_options opt;
    if (sa_options_from_env(&opt) != SA_OK) return fail(SA_ERR_ARGUMENT);
    // pack the requests into the CSR layout of sa_align_batch, in page-locked staging buffers: the copies of the
    // batch path run at link speed only from pinned memory (pageable staging was 5x slower)
    int64_t *toff = static_cast<int64_t *>(g_pin[2].reserve((n + 1) * 8));
    int64_t *poff = static_cast<int64_t *>(g_pin[3].reserve((n + 1) * 8));
    if (!toff || !poff) return fail(SA_ERR_MEMORY);
    toff[0] = poff[0] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        toff[i + 1] = toff[i] + (int64_t)rq[i].textNumBytes;
        poff[i + 1] = poff[i] + (int64_t)rq[i].patternNumBytes;
    }
    const uint64_t arena = (uint64_t)(toff[n] + poff[n]);
    uint8_t *text = static_cast<uint8_t *>(g_pin[0].reserve((uint64_t)toff[n] + 16));
    uint8_t *pat = static_cast<uint8_t *>(g_pin[1].reserve((uint64_t)poff[n] + 16));
    sa_result *res = static_cast<sa_result *>(g_pin[4].reserve(n * sizeof(sa_result)));
    uint64_t *off = static_cast<uint64_t *>(g_pin[5].reserve(n * 8));
    char *aT = static_cast<char *>(g_pin[6].reserve(arena + 16));
    char *aP = static_cast<char *>(g_pin[7].reserve(arena + 16));
    if (!text || !pat || !res || !off || !aT || !aP) return fail(SA_ERR_MEMORY);
    // (a few host threads for big batches: 1 M requests are 600 MB of small copies)
    const unsigned nT = n >= 65536 ? std::max(1u, std::min(8u, std::thread::hardware_concurrency())) : 1u;
    auto packRange = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            std::memcpy(text + toff[i], rq[i].textBytes, rq[i].textNumBytes);
            std::memcpy(pat + poff[i], rq[i].patternBytes, rq[i].patternNumBytes);
        }
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nT; ++t) th.emplace_back(packRange, n * t / nT, n * (t + 1) / nT);
        packRange(0, n / nT);
        for (auto &t : th) t.join();
    }
    const sa_scoring sc = scoringOf(rq[0]);
    sa_batch b{n, text, toff, pat, poff};
    sa_batch_out o{res, off, aT, aP, arena};
    const int st = sa_align_batch_multi(&opt, &sc, &b, &o);
    if (st != SA_OK) return fail(st);
    bool oom = false;
    auto unpackRange = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            const uint64_t cap = std::max<uint64_t>(2 * rq[i].textNumBytes, rq[i].textNumBytes + rq[i].patternNumBytes);
            char *t = new (std::nothrow) char[cap], *p = new (std::nothrow) char[cap];
            if (!t || !p) { delete[] t; delete[] p; oom = true; return; }
            std::memcpy(t, aT + off[i], res[i].aln_len);
            std::memcpy(p, aP + off[i], res[i].aln_len);
            delete[] rs[i].alignedTextBytes;
            delete[] rs[i].alignedPatternBytes;
            rs[i].alignedTextBytes = t;
            rs[i].alignedPatternBytes = p;
            rs[i].numAlignmentBytes = res[i].aln_len;
            rs[i].startInAlignedText = res[i].start_text;
            rs[i].startInAlignedPattern = res[i].start_pattern;
            rs[i].score = res[i].score;
        }
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nT; ++t) th.emplace_back(unpackRange, n * t / nT, n * (t + 1) / nT);
        unpackRange(0, n / nT);
        for (auto &t : th) t.join();
    }
    return oom ? fail(SA_ERR_MEMORY) : 0;
}

// sa_shim.cpp -- C++ side of the drop-in boundary: SequenceAlignment::alignSequenceGPU with the
// reference's ownership and error conventions, implemented on the C ABI (include/sa_b200.h).
//   reference: alignSequenceGPU.cu:463-653 (entry), :402-403 (new char[] outputs),
//              :541-546 / :588-594 (stdout messages + return 1), :613-626 (BENCHMARK return value).
#include "../../include/SequenceAlignment.hpp"
#include "../../include/sa_b200.h"

#include <algorithm>
#include <iostream>
#include <mutex>
#include <new>
#include <vector>

namespace {

sa_context *g_ctx = nullptr;
std::mutex g_mu;

sa_context *context()
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx && sa_create(0, &g_ctx) != SA_OK) g_ctx = nullptr;   // device 0, like alignSequenceGPU.cu:476
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

uint64_t fail(int status)
{
    if (status == SA_ERR_MEMORY || status == SA_ERR_CAPACITY) std::cout << SequenceAlignment::MEM_ERROR;
    else std::cout << "error: could not copy from device memory\n";
    return 1;
}

} // namespace

uint64_t SequenceAlignment::alignSequenceGPU(const Request &rq, Response *rs)
{
    // anything but GLOBAL / LOCAL: both reference paths silently do nothing (alignSequenceCPU.cpp:318-328)
    if (rq.alignmentType != GLOBAL && rq.alignmentType != LOCAL) return 0;
    sa_context *ctx = context();
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
    sa_timing t;
    sa_last_timing(ctx, &t);
    return (uint64_t)(t.fill_us + t.d2h_us);
#else
    return 0;
#endif
}

uint64_t SequenceAlignment::alignSequenceGPUFillMicros(const Request &rq, Response *rs)
{
    sa_context *ctx = context();
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
    sa_context *ctx = context();
    if (!ctx) return fail(SA_ERR_NO_DEVICE);
    // pack the requests into the CSR layout of sa_align_batch
    std::vector<int64_t> toff(n + 1, 0), poff(n + 1, 0);
    for (uint64_t i = 0; i < n; ++i) {
        toff[i + 1] = toff[i] + (int64_t)rq[i].textNumBytes;
        poff[i + 1] = poff[i] + (int64_t)rq[i].patternNumBytes;
    }
    std::vector<uint8_t> text(toff[n]), pat(poff[n]);
    for (uint64_t i = 0; i < n; ++i) {
        std::copy_n(rq[i].textBytes, rq[i].textNumBytes, text.begin() + toff[i]);
        std::copy_n(rq[i].patternBytes, rq[i].patternNumBytes, pat.begin() + poff[i]);
    }
    const uint64_t arena = (uint64_t)(toff[n] + poff[n]);
    std::vector<sa_result> res(n);
    std::vector<uint64_t> off(n);
    std::vector<char> aT(arena), aP(arena);
    const sa_scoring sc = scoringOf(rq[0]);
    sa_batch b{n, text.data(), toff.data(), pat.data(), poff.data()};
    sa_batch_out o{res.data(), off.data(), aT.data(), aP.data(), arena};
    const int st = sa_align_batch(ctx, &sc, &b, &o);
    if (st != SA_OK) return fail(st);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t cap = std::max<uint64_t>(2 * rq[i].textNumBytes, rq[i].textNumBytes + rq[i].patternNumBytes);
        char *t = new (std::nothrow) char[cap], *p = new (std::nothrow) char[cap];
        if (!t || !p) { delete[] t; delete[] p; return fail(SA_ERR_MEMORY); }
        std::copy_n(aT.data() + off[i], res[i].aln_len, t);
        std::copy_n(aP.data() + off[i], res[i].aln_len, p);
        delete[] rs[i].alignedTextBytes;
        delete[] rs[i].alignedPatternBytes;
        rs[i].alignedTextBytes = t;
        rs[i].alignedPatternBytes = p;
        rs[i].numAlignmentBytes = res[i].aln_len;
        rs[i].startInAlignedText = res[i].start_text;
        rs[i].startInAlignedPattern = res[i].start_pattern;
        rs[i].score = res[i].score;
    }
    return 0;
}

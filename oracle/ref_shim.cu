// ref_shim.cu -- TEST INFRASTRUCTURE ONLY.
// Thin extern "C" wrapper around the UNMODIFIED reference, compiled from the
// sources where they lie under /root/reference (never copied into this repo):
//   nvcc -I/root/reference oracle/ref_shim.cu -> oracle/_ref/libsa_ref*.so   (see oracle/Makefile)
// The reference is a unity build: SequenceAlignment.hpp pulls in utilities.cpp,
// alignSequenceCPU.cpp and alignSequenceGPU.cu (SequenceAlignment.hpp:138-140),
// so this one translation unit contains the whole reference program.
// Used to (a) pin oracle/sa_oracle.c, (b) generate tests/golden/*, and
// (c) serve as bench.py's cpu_baseline / --impl reference (kind "reference").
#include "SequenceAlignment.hpp"

#include <atomic>
#include <cstring>
#include <sstream>
#include <thread>
#include <vector>

namespace {

void fillRequest(SequenceAlignment::Request &rq, int mode, int alphabetSize, const int *matrix,
                 int gap, const char *text, uint64_t n, const char *pattern, uint64_t m)
{
    const bool protein = alphabetSize == (int)SequenceAlignment::NUM_PROTEIN_CHARS;
    rq.deviceType = SequenceAlignment::programArgs::CPU;
    rq.sequenceType = protein ? SequenceAlignment::programArgs::PROTEIN
                              : SequenceAlignment::programArgs::DNA;
    rq.alignmentType = mode == 0 ? SequenceAlignment::programArgs::GLOBAL
                                 : SequenceAlignment::programArgs::LOCAL;
    rq.alphabet = protein ? SequenceAlignment::PROTEIN_ALPHABET : SequenceAlignment::DNA_ALPHABET;
    rq.alphabetSize = alphabetSize;
    rq.gapPenalty = gap;
    rq.textNumBytes = n;
    rq.patternNumBytes = m;
    rq.textBytes = new char[n];      // ~Request delete[]s these (SequenceAlignment.hpp:92-98)
    rq.patternBytes = new char[m];
    std::memcpy(rq.textBytes, text, n);
    std::memcpy(rq.patternBytes, pattern, m);
    std::memcpy(rq.scoreMatrix, matrix, sizeof(int) * alphabetSize * alphabetSize);
}

} // namespace

extern "C" {

// Runs SequenceAlignment::alignSequenceCPU (alignSequenceCPU.cpp:287).
// outT/outP must hold 2*n bytes (the reference's own capacity, :306-307);
// callers keep text >= pattern like parseArguments does (utilities.cpp:225-230).
int ref_align_cpu(int mode, int alphabetSize, const int *matrix, int gap,
                  const char *text, uint64_t n, const char *pattern, uint64_t m,
                  int *score, uint64_t *alnLen, uint64_t *startText, uint64_t *startPattern,
                  char *outT, char *outP)
{
    SequenceAlignment::Request rq;
    SequenceAlignment::Response rs;
    fillRequest(rq, mode, alphabetSize, matrix, gap, text, n, pattern, m);
    const uint64_t err = SequenceAlignment::alignSequenceCPU(rq, &rs);
    if (err) return (int)err;
    *score = rs.score;
    *alnLen = rs.numAlignmentBytes;
    *startText = rs.startInAlignedText;
    *startPattern = rs.startInAlignedPattern;
    if (outT) std::memcpy(outT, rs.alignedTextBytes, rs.numAlignmentBytes);
    if (outP) std::memcpy(outP, rs.alignedPatternBytes, rs.numAlignmentBytes);
    return 0;
}

// A batch of independent pairs (CSR layout) through alignSequenceCPU, one pair per task on
// `nthreads` host threads (the reference itself is single-threaded; pairs are independent).
// This is bench.py's cpu_baseline / --impl reference leg for the batch workload.
// Returns the number of failed pairs.  checksum = sum over pairs of (score + alnLen).
int ref_align_cpu_batch(int mode, int alphabetSize, const int *matrix, int gap,
                        const char *text, const int64_t *toff, const char *pattern, const int64_t *poff,
                        uint64_t nPairs, int nthreads, int *scores, uint64_t *alnLens)
{
    std::atomic<uint64_t> next(0);
    std::atomic<int> failed(0);
    auto worker = [&]() {
        for (;;) {
            const uint64_t p = next.fetch_add(1);
            if (p >= nPairs) break;
            int score = 0; uint64_t len = 0, st = 0, sp = 0;
            const uint64_t n = (uint64_t)(toff[p + 1] - toff[p]), m = (uint64_t)(poff[p + 1] - poff[p]);
            if (ref_align_cpu(mode, alphabetSize, matrix, gap, text + toff[p], n, pattern + poff[p], m,
                              &score, &len, &st, &sp, nullptr, nullptr) != 0) { failed++; continue; }
            if (scores) scores[p] = score;
            if (alnLens) alnLens[p] = len;
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < (nthreads > 1 ? nthreads : 1); ++t) th.emplace_back(worker);
    for (auto &t : th) t.join();
    return failed.load();
}

// The same loop as a CHECKER (tests/, bench.py "verified"): every Response field and both strings of the listed pairs,
// computed by the unmodified alignSequenceCPU, against a result set laid out like sa_align_batch's (results[p] =
// {int32 score, pad, u64 aln_len, u64 start_text, u64 start_pattern}; strings at aligned_*[aln_off[p] ..)).
// idx == nullptr checks pairs 0..nIdx-1.  Returns the number of mismatching pairs, *firstBad the smallest one or -1.
// Pairs whose pattern is longer than their text are aligned with the operands as given; the reference's 2*text output
// capacity (alignSequenceCPU.cpp:306-307) holds for the mutate.py-style batches this is used on.
uint64_t ref_check_batch(int mode, int alphabetSize, const int *matrix, int gap,
                         const char *text, const int64_t *toff, const char *pattern, const int64_t *poff,
                         const uint64_t *idx, uint64_t nIdx, const void *results, const uint64_t *alnOff,
                         const char *alnT, const char *alnP, int nthreads, int64_t *firstBad)
{
    struct Res { int32_t score; int32_t pad; uint64_t len, st, sp; };
    const Res *R = static_cast<const Res *>(results);
    std::atomic<uint64_t> next(0), bad(0);
    std::atomic<int64_t> first(-1);
    auto worker = [&]() {
        for (;;) {
            const uint64_t k = next.fetch_add(1);
            if (k >= nIdx) break;
            const uint64_t p = idx ? idx[k] : k;
            const uint64_t n = (uint64_t)(toff[p + 1] - toff[p]), m = (uint64_t)(poff[p + 1] - poff[p]);
            SequenceAlignment::Request rq;
            SequenceAlignment::Response rs;
            fillRequest(rq, mode, alphabetSize, matrix, gap, text + toff[p], n, pattern + poff[p], m);
            bool ok = SequenceAlignment::alignSequenceCPU(rq, &rs) == 0;
            ok = ok && R[p].score == rs.score && R[p].len == rs.numAlignmentBytes && R[p].st == rs.startInAlignedText &&
                 R[p].sp == rs.startInAlignedPattern &&
                 std::memcmp(alnT + alnOff[p], rs.alignedTextBytes, rs.numAlignmentBytes) == 0 &&
                 std::memcmp(alnP + alnOff[p], rs.alignedPatternBytes, rs.numAlignmentBytes) == 0;
            if (!ok) {
                bad++;
                int64_t cur = first.load();
                while ((cur < 0 || (int64_t)p < cur) && !first.compare_exchange_weak(cur, (int64_t)p)) {}
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < (nthreads > 1 ? nthreads : 1); ++t) th.emplace_back(worker);
    for (auto &t : th) t.join();
    if (firstBad) *firstBad = first.load();
    return bad.load();
}

// Fill only (file-local fillMatrixNW / fillMatrixSW, alignSequenceCPU.cpp:203,116):
// what tests/benchmarks.cu:150-157 times as the CPU "MCUPS".  M = (m+1)*(n+1) bytes.
int ref_fill_cpu(int mode, int alphabetSize, const int *matrix, int gap,
                 const char *text, uint64_t n, const char *pattern, uint64_t m,
                 char *M, int *score, uint64_t *argmax)
{
    SequenceAlignment::Request rq;
    fillRequest(rq, mode, alphabetSize, matrix, gap, text, n, pattern, m);
    if (mode == 0) {
        *score = fillMatrixNW(M, m + 1, n + 1, rq);
        *argmax = 0;
    } else {
        auto r = fillMatrixSW(M, m + 1, n + 1, rq);
        *score = r.first;
        *argmax = r.second;
    }
    return 0;
}

// readSequenceFile -> validateAndTransform (utilities.cpp:65,31): file -> alphabet indices.
// Returns number of residues, or -1.
int64_t ref_read_sequence(const char *fname, int alphabetSize, char *out, uint64_t cap)
{
    SequenceAlignment::Request rq;
    const bool protein = alphabetSize == (int)SequenceAlignment::NUM_PROTEIN_CHARS;
    rq.alphabet = protein ? SequenceAlignment::PROTEIN_ALPHABET : SequenceAlignment::DNA_ALPHABET;
    rq.alphabetSize = alphabetSize;
    rq.textNumBytes = 0;
    rq.patternNumBytes = 0;
    if (readSequenceFile(fname, &rq) != 0 || rq.textNumBytes == 0) return -1;
    if (rq.textNumBytes > cap) return -(int64_t)rq.textNumBytes;
    std::memcpy(out, rq.textBytes, rq.textNumBytes);
    return (int64_t)rq.textNumBytes;
}

// validateAndTransform on an in-memory string (utilities.cpp:31). In place; returns count.
int ref_validate_and_transform(char *buf, uint64_t len, int alphabetSize)
{
    const bool protein = alphabetSize == (int)SequenceAlignment::NUM_PROTEIN_CHARS;
    std::string s(buf, buf + len);
    const int nRead = validateAndTransform(
        s, protein ? SequenceAlignment::PROTEIN_ALPHABET : SequenceAlignment::DNA_ALPHABET, alphabetSize);
    std::memcpy(buf, s.data(), nRead > 0 ? nRead : 0);
    return nRead;
}

// parseScoreMatrixFile (utilities.cpp:106). Returns 0 / -1 like the reference.
int ref_parse_score_matrix(const char *fname, int alphabetSize, int *out)
{
    return parseScoreMatrixFile(fname, alphabetSize, out);
}

// prettyAlignmentPrint (utilities.cpp:253) into a caller buffer; returns bytes needed.
uint64_t ref_pretty_print(const char *alnT, const char *alnP, uint64_t len, uint64_t startT,
                          uint64_t startP, int score, char *out, uint64_t cap)
{
    SequenceAlignment::Response rs;
    rs.alignedTextBytes = new char[len ? len : 1];
    rs.alignedPatternBytes = new char[len ? len : 1];
    std::memcpy(rs.alignedTextBytes, alnT, len);
    std::memcpy(rs.alignedPatternBytes, alnP, len);
    rs.numAlignmentBytes = len;
    rs.startInAlignedText = startT;
    rs.startInAlignedPattern = startP;
    rs.score = score;
    std::ostringstream os;
    prettyAlignmentPrint(rs, os);
    const std::string s = os.str();
    if (out && cap) std::memcpy(out, s.data(), s.size() < cap ? s.size() : cap);
    return s.size();
}

// The reference's own GPU path, SequenceAlignment::alignSequenceGPU (alignSequenceGPU.cu:463), unmodified: the
// secondary baseline "reference kernels on the same B200" of benchmarks.py.  Built normally it
// returns 0 and fills the response (CPU traceback included); built with -DBENCHMARK (libsa_refgpu_bench.so) it
// returns the microseconds of fill + device-to-host copy like tests/benchmarks.cu uses it (:613-626).
uint64_t ref_align_gpu(int mode, int alphabetSize, const int *matrix, int gap,
                       const char *text, uint64_t n, const char *pattern, uint64_t m,
                       int *score, uint64_t *alnLen, uint64_t *startText, uint64_t *startPattern,
                       char *outT, char *outP)
{
    SequenceAlignment::Request rq;
    SequenceAlignment::Response rs;
    fillRequest(rq, mode, alphabetSize, matrix, gap, text, n, pattern, m);
    rq.deviceType = SequenceAlignment::programArgs::GPU;
    const uint64_t r = SequenceAlignment::alignSequenceGPU(rq, &rs);
#ifdef BENCHMARK
    return r;
#else
    if (r) return r;
    if (score) *score = rs.score;
    if (alnLen) *alnLen = rs.numAlignmentBytes;
    if (startText) *startText = rs.startInAlignedText;
    if (startPattern) *startPattern = rs.startInAlignedPattern;
    if (outT) std::memcpy(outT, rs.alignedTextBytes, rs.numAlignmentBytes);
    if (outP) std::memcpy(outP, rs.alignedPatternBytes, rs.numAlignmentBytes);
    return 0;
#endif
}

} // extern "C"

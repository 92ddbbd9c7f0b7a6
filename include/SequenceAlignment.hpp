// SequenceAlignment.hpp -- drop-in surface for the B200 build.
//
// Declares the same public types and entry points as the reference's header
// (/root/reference/SequenceAlignment.hpp:10-131) so that callers written against
// it -- mainDriver.cu:17-24, tests/tests.cu, tests/benchmarks.cu -- compile and
// link unchanged against libsa_b200.so.  Differences on purpose:
//   * no unity-build #includes (the reference pulls utilities.cpp,
//     alignSequenceCPU.cpp and alignSequenceGPU.cu in at :138-140); the GPU entry
//     point lives in libsa_b200.so (csrc/sa_shim.cpp -> include/sa_b200.h);
//   * alignSequenceGPUBatch is new (the reference loops over single calls,
//     tests/benchmarks.cu:318-322).
// Field order and types of Request / Response are ABI-relevant and match the
// reference exactly; do not reorder.
#pragma once

#include <cstdint>
#include <string>

namespace SequenceAlignment
{
    // Same enumerators, same order (values 0..8) as the reference's programArgs.
    enum programArgs { CPU, GPU, DNA, PROTEIN, GLOBAL, LOCAL, SEMI_GLOBAL, SCORE_MATRIX, GAP_PENALTY };

    const unsigned int NUM_DNA_CHARS = 4;
    const unsigned int NUM_PROTEIN_CHARS = 23;
    // letter -> index order is fixed by the score-matrix files; the last char is the gap.
    const char DNA_ALPHABET[] = {'A', 'T', 'C', 'G', '-'};
    const char PROTEIN_ALPHABET[] = {'A', 'R', 'N', 'D', 'C', 'Q', 'E', 'G', 'H', 'I', 'L', 'K', 'M',
                                     'F', 'P', 'S', 'T', 'W', 'Y', 'V', 'B', 'Z', 'X', '-'};
    const short DEFAULT_GAP_PENALTY = 5;
    const std::string MEM_ERROR = "error: sequence is too long, not enough memory\n";

    struct Request
    {
        programArgs deviceType;
        programArgs sequenceType;
        programArgs alignmentType;
        char *textBytes = nullptr;        // alphabet indices, one per byte
        uint64_t textNumBytes;
        char *patternBytes = nullptr;
        uint64_t patternNumBytes;
        const char *alphabet;
        int alphabetSize;
        int scoreMatrix[NUM_PROTEIN_CHARS * NUM_PROTEIN_CHARS];   // row-major, stride alphabetSize
        int gapPenalty;

        ~Request()
        {
            delete[] textBytes;
            delete[] patternBytes;
            textBytes = patternBytes = nullptr;
        }
    };

    struct Response
    {
        char *alignedTextBytes = nullptr;     // new char[]-allocated by the aligner, owned here
        char *alignedPatternBytes = nullptr;
        uint64_t numAlignmentBytes;
        uint64_t startInAlignedText;
        uint64_t startInAlignedPattern;
        int score;

        ~Response()
        {
            delete[] alignedTextBytes;
            delete[] alignedPatternBytes;
            alignedTextBytes = alignedPatternBytes = nullptr;
        }
    };

    enum DIRECTION { LEFT, DIAG, TOP, STOP };

    // 0 on success; 1 after printing MEM_ERROR (or the copy error) to stdout, like the reference.
    // Built with -DBENCHMARK the shim returns elapsed microseconds of fill + D2H instead
    // (alignSequenceGPU.cu:613-626) -- see sa_shim.cpp.
    uint64_t alignSequenceGPU(const Request &, Response *);

    // New: n independent requests (same scoring scheme) in one device batch.
    uint64_t alignSequenceGPUBatch(const Request *, Response *, uint64_t n);

    // BENCHMARK-mode twin of alignSequenceGPU: fill only, returns elapsed microseconds.
    uint64_t alignSequenceGPUFillMicros(const Request &, Response *);
}

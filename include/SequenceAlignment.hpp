// SequenceAlignment.hpp -- drop-in surface of the B200 build.
//
// Declares the public types, constants and entry points of the reference's header
// (/root/reference/SequenceAlignment.hpp:10-131) plus the front-end functions of its utilities.cpp
// (:10, :19, :31, :65, :106, :131, :253), so that a caller written like mainDriver.cu:4-27 -- parseArguments,
// alignSequenceGPU, prettyAlignmentPrint -- compiles against this header and links against libsa_b200.so alone.
// What the library implements:
//   * alignSequenceGPU                      the hot path (csrc/sa_shim.cpp -> include/sa_b200.h -> CUDA);
//   * alignSequenceGPUBatch                 new: n requests in one device batch, sharded over SA_DEVICES GPUs;
//   * parseArguments, readSequenceFile, validateAndTransform, parseScoreMatrixFile, indexOfLetter, getScore,
//     prettyAlignmentPrint                  restated over the C front end (csrc/sa_utilities.cpp, csrc/sa_frontend.cpp).
// What it does NOT implement: alignSequenceCPU, traceBackNW, traceBackSW.  They are declared so that reference
// callers compile, but this is a GPU library with no CPU fallback: a program that also wants `-c` keeps linking the
// reference's own alignSequenceCPU.cpp next to the library (INTEGRATION.md shows the reference tree built that way).
// Differences on purpose: no unity-build #includes (the reference pulls its .cpp/.cu files in at :138-140).
// Field order and types of Request / Response are ABI-relevant and match the reference exactly; do not reorder.
#pragma once

#include <cstdint>
#include <iostream>
#include <string>
#include <unordered_map>

namespace SequenceAlignment
{
    // Same enumerators, same order (values 0..8) as the reference's programArgs.
    enum programArgs { CPU, GPU, DNA, PROTEIN, GLOBAL, LOCAL, SEMI_GLOBAL, SCORE_MATRIX, GAP_PENALTY };

    // command-line flags (SequenceAlignment.hpp:23-32)
    const std::unordered_map<std::string, programArgs> argumentMap = {
        {"--cpu", CPU}, {"-c", CPU}, {"--gpu", GPU}, {"-g", GPU}, {"--dna", DNA}, {"-d", DNA},
        {"--protein", PROTEIN}, {"-p", PROTEIN}, {"--global", GLOBAL}, {"--local", LOCAL},
        {"--score-matrix", SCORE_MATRIX}, {"-s", SCORE_MATRIX}, {"--gap-penalty", GAP_PENALTY},
    };

    // user messages, byte for byte (SequenceAlignment.hpp:35-50): the reference's tests compare them
    const std::string USAGE =
        "Usage: ./alignSequence [-d|-p] [-c|-g] [--global|--local] [-s <file>] [--gap-penalty <int>] <file> <file>\n"
        "       -d, --dna             - align dna sequences (default)\n"
        "       -p, --protein         - align protein sequence\n"
        "       -c, --cpu             - use cpu device (default)\n"
        "       -g, --gpu             - use gpu device\n"
        "       --global              - use global alignment (default)\n"
        "       --local               - use local alignment\n"
        "       -s, --score-matrix    - next argument is a score matrix file\n"
        "       --gap-penalty         - next argument is a gap open penalty (default 5)\n";
    const std::string SEQ_NOT_READ_ERROR = "error: text sequence or pattern sequence not read\n";
    const std::string MEM_ERROR = "error: sequence is too long, not enough memory\n";
    const std::string SCORE_MATRIX_NOT_READ_ERROR = "error: matrix scores not read. Only integer scores accepted (int)\n";
    const std::string GAP_PENALTY_NOT_READ_ERROR = "error: gap penalty not read. Only integer scores accepted (int)\n";

    const unsigned int NUM_DNA_CHARS = 4;
    const unsigned int NUM_PROTEIN_CHARS = 23;
    // letter -> index order is fixed by the score-matrix files; the last char is the gap.
    const char DNA_ALPHABET[] = {'A', 'T', 'C', 'G', '-'};
    const char PROTEIN_ALPHABET[] = {'A', 'R', 'N', 'D', 'C', 'Q', 'E', 'G', 'H', 'I', 'L', 'K', 'M',
                                     'F', 'P', 'S', 'T', 'W', 'Y', 'V', 'B', 'Z', 'X', '-'};

    // defaults (SequenceAlignment.hpp:61-68)
    const programArgs DEFAULT_DEVICE = CPU;
    const programArgs DEFAULT_SEQUENCE = DNA;
    const programArgs DEFAULT_ALIGNMENT_TYPE = GLOBAL;
    static const char *DEFAULT_ALPHABET = DNA_ALPHABET;
    const int DEFAULT_ALPHABET_SIZE = NUM_DNA_CHARS;
    const short DEFAULT_GAP_PENALTY = 5;
    const std::string DEFAULT_DNA_SCORE_MATRIX_FILE = "scoreMatrices/dna/blast.txt";
    const std::string DEFAULT_PROTEIN_SCORE_MATRIX_FILE = "scoreMatrices/protein/blosum50.txt";

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
    // The -DBENCHMARK build of the library (libsa_b200_bench.so) returns elapsed microseconds of fill + D2H instead
    // (alignSequenceGPU.cu:613-626) -- see sa_shim.cpp.
    uint64_t alignSequenceGPU(const Request &, Response *);

    // New: n independent requests (same scoring scheme) in one device batch.  SA_DEVICES ("0,1,2,3" or a count)
    // selects the GPUs the batch is sharded over, one host thread per device; default: device 0.
    uint64_t alignSequenceGPUBatch(const Request *, Response *, uint64_t n);

    // BENCHMARK-mode twin of alignSequenceGPU: fill only, returns elapsed microseconds.
    uint64_t alignSequenceGPUFillMicros(const Request &, Response *);

    // Declared for source compatibility only -- NOT in libsa_b200.so (no CPU fallback, see the banner).
    uint64_t alignSequenceCPU(const Request &, Response *);
    void traceBackNW(const char *, const uint64_t, const uint64_t, const Request &, Response *);
    void traceBackSW(const char *, const uint64_t, const uint64_t, const uint64_t, const Request &, Response *);
}

// front end of the reference's utilities.cpp (global namespace there too), implemented in csrc/sa_utilities.cpp
char indexOfLetter(const char letter, const char *alphabet, const int alphabetSize);
int getScore(char char1, char char2, const char *alphabet, const int alphabetSize, const int *scoreMatrix);
int validateAndTransform(std::string &sequence, const char *alphabet, const int alphabetSize);
int readSequenceFile(const std::string fname, SequenceAlignment::Request *request);
int parseScoreMatrixFile(const std::string &fname, const int alphabetSize, int *buffer);
int parseArguments(int argc, const char *argv[], SequenceAlignment::Request *request);
void prettyAlignmentPrint(SequenceAlignment::Response &response, std::ostream &stream);

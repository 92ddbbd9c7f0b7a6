// dropin_main.cpp -- a caller written against include/SequenceAlignment.hpp the way the reference's mainDriver.cu:4-27 is
// written against its header, linked against libsa_b200.so ONLY (tests/test_dropin.py builds it with g++).
//
//   dropin_main <reference CLI arguments>      parseArguments -> alignSequenceGPU -> prettyAlignmentPrint
//                                              (the `-g` path of ./alignSequence; `-c` is refused: no CPU fallback)
//   dropin_main --batch <case file>            alignSequenceGPUBatch on the requests of a binary case file, one line
//                                              per pair on stdout:  score len startText startPattern text pattern
//   dropin_main --twice <reference CLI args>   the same Response object through alignSequenceGPU twice (buffer reuse)
//   dropin_main --return <reference CLI args>  prints what alignSequenceGPU returns: 0, or with the -DBENCHMARK build of
//                                              the library the microseconds of fill + D2H (tests/benchmarks.cu:171-175)
#include "SequenceAlignment.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

namespace SA = SequenceAlignment;

static int run_cli(int argc, const char *argv[], int repeats)
{
    SA::Request request;
    SA::Response response;
    if (parseArguments(argc, argv, &request)) return 1;
    if (request.deviceType != SA::programArgs::GPU) {
        std::cerr << "dropin_main: this build has no CPU path, use -g\n";
        return 2;
    }
    for (int k = 0; k < repeats; ++k)
        if (SA::alignSequenceGPU(request, &response)) return 1;
    prettyAlignmentPrint(response, std::cout);
    return 0;
}

// case file: int32 mode(0 global / 1 local), alpha, gap, n ; alpha*alpha int32 matrix ; n x {uint32 textLen, patternLen} ;
// then the residues of every pair, text then pattern
static int run_batch(const char *path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f.good()) return 3;
    int32_t hdr[4];
    f.read(reinterpret_cast<char *>(hdr), sizeof hdr);
    const int mode = hdr[0], alpha = hdr[1], gap = hdr[2];
    const uint64_t n = (uint64_t)hdr[3];
    std::vector<int32_t> mat((size_t)alpha * alpha);
    f.read(reinterpret_cast<char *>(mat.data()), mat.size() * 4);
    std::vector<uint32_t> lens(2 * n);
    f.read(reinterpret_cast<char *>(lens.data()), lens.size() * 4);
    std::vector<SA::Request> rq(n);
    std::vector<SA::Response> rs(n);
    for (uint64_t i = 0; i < n; ++i) {
        SA::Request &r = rq[i];
        r.deviceType = SA::GPU;
        r.sequenceType = alpha == (int)SA::NUM_DNA_CHARS ? SA::DNA : SA::PROTEIN;
        r.alignmentType = mode ? SA::LOCAL : SA::GLOBAL;
        r.alphabet = alpha == (int)SA::NUM_DNA_CHARS ? SA::DNA_ALPHABET : SA::PROTEIN_ALPHABET;
        r.alphabetSize = alpha;
        r.gapPenalty = gap;
        std::memcpy(r.scoreMatrix, mat.data(), mat.size() * 4);
        r.textNumBytes = lens[2 * i];
        r.patternNumBytes = lens[2 * i + 1];
        r.textBytes = new char[r.textNumBytes];
        r.patternBytes = new char[r.patternNumBytes];
        f.read(r.textBytes, (std::streamsize)r.textNumBytes);
        f.read(r.patternBytes, (std::streamsize)r.patternNumBytes);
    }
    if (!f.good()) return 3;
    if (SA::alignSequenceGPUBatch(rq.data(), rs.data(), n)) return 1;
    for (uint64_t i = 0; i < n; ++i) {
        std::printf("%d %llu %llu %llu ", rs[i].score, (unsigned long long)rs[i].numAlignmentBytes,
                    (unsigned long long)rs[i].startInAlignedText, (unsigned long long)rs[i].startInAlignedPattern);
        std::fwrite(rs[i].alignedTextBytes, 1, rs[i].numAlignmentBytes, stdout);
        std::fputc(' ', stdout);
        std::fwrite(rs[i].alignedPatternBytes, 1, rs[i].numAlignmentBytes, stdout);
        std::fputc('\n', stdout);
    }
    return 0;
}

int main(int argc, const char *argv[])
{
    if (argc >= 3 && !std::strcmp(argv[1], "--batch")) return run_batch(argv[2]);
    if (argc >= 2 && !std::strcmp(argv[1], "--twice")) return run_cli(argc - 1, argv + 1, 2);
    if (argc >= 2 && !std::strcmp(argv[1], "--return")) {
        SA::Request request;
        SA::Response response;
        if (parseArguments(argc - 1, argv + 1, &request)) return 1;
        const uint64_t r = SA::alignSequenceGPU(request, &response);
        std::printf("%llu %d\n", (unsigned long long)r, response.score);
        return 0;
    }
    return run_cli(argc, argv, 1);
}

// sa_utilities.cpp -- the reference's front-end functions (utilities.cpp:10-315) with their own signatures, restated over
// the C front end of this library (sa_frontend.cpp), so that a mainDriver-style caller links against libsa_b200.so alone.
// Return values and stderr messages are the reference's (its tests compare them, tests/tests.cu:67-114).
#include "../../include/SequenceAlignment.hpp"
#include "../../include/sa_b200.h"

#include <cstring>
#include <fstream>
#include <new>
#include <vector>

namespace SA = SequenceAlignment;

// utilities.cpp:10-16: position of the letter in the alphabet, -1 when it is not there
char indexOfLetter(const char letter, const char *alphabet, const int alphabetSize)
{
    const void *hit = alphabetSize > 0 ? std::memchr(alphabet, letter, (size_t)alphabetSize) : nullptr;
    return hit ? (char)(static_cast<const char *>(hit) - alphabet) : (char)-1;
}

// utilities.cpp:19-25
int getScore(char char1, char char2, const char *alphabet, const int alphabetSize, const int *scoreMatrix)
{
    return scoreMatrix[indexOfLetter(char1, alphabet, alphabetSize) * alphabetSize + indexOfLetter(char2, alphabet, alphabetSize)];
}

// utilities.cpp:31-63: in place; the number of residues, or 0 after "'X' letter not in alphabet." on stderr
int validateAndTransform(std::string &sequence, const char *alphabet, const int alphabetSize)
{
    char bad = 0;
    const int64_t n = sa_validate_and_transform(sequence.empty() ? nullptr : &sequence[0], sequence.size(), alphabet, alphabetSize, &bad);
    if (n == 0 && bad) std::cerr << "'" << bad << "'" << " letter not in alphabet." << std::endl;
    return (int)n;
}

// utilities.cpp:65-104: the first file read becomes the text, the second the pattern
int readSequenceFile(const std::string fname, SA::Request *request)
{
    std::ifstream f(fname, std::ios::binary);
    if (!f.good()) {
        std::cerr << fname << " file does not exist" << std::endl;
        return -1;
    }
    std::string contents((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const int n = validateAndTransform(contents, request->alphabet, request->alphabetSize);
    if (n <= 0) return 0;
    const bool asText = request->textNumBytes == 0;
    if (!asText && request->patternNumBytes != 0) return 0;        // both already read: a third file is ignored
    char *buf = new (std::nothrow) char[n];
    if (!buf) {
        std::cerr << SA::MEM_ERROR;
        return -1;
    }
    std::memcpy(buf, contents.data(), (size_t)n);
    if (asText) { request->textBytes = buf; request->textNumBytes = (uint64_t)n; }
    else { request->patternBytes = buf; request->patternNumBytes = (uint64_t)n; }
    return 0;
}

// utilities.cpp:106-129: -1 for a token that is not an integer; a missing file prints a message and returns 0
int parseScoreMatrixFile(const std::string &fname, const int alphabetSize, int *buffer)
{
    {
        std::ifstream probe(fname);
        if (!probe.good()) {
            std::cerr << fname << " file does not exist" << std::endl;
            return 0;
        }
    }
    std::vector<int32_t> m((size_t)alphabetSize * alphabetSize);
    if (sa_parse_score_matrix_file(fname.c_str(), alphabetSize, m.data()) != SA_OK) {
        // the reference has stored the scores it could read before it gives up: do the same
        std::ifstream f(fname);
        int v;
        for (int i = 0; i < alphabetSize * alphabetSize && (f >> v); ++i) buffer[i] = v;
        return -1;
    }
    for (size_t i = 0; i < m.size(); ++i) buffer[i] = m[i];
    return 0;
}

// utilities.cpp:131-251.  Flags come from argumentMap; "--gap-penalty" / "-s" take the next non-flag argument; every
// other non-flag argument is a sequence file.  Afterwards the longer sequence becomes the text and, unless -s was read,
// the default matrix of the sequence type is loaded.
int parseArguments(int argc, const char *argv[], SA::Request *request)
{
    if (argc == 1) {
        std::cerr << SA::USAGE;
        return 1;
    }
    request->deviceType = SA::DEFAULT_DEVICE;
    request->sequenceType = SA::DEFAULT_SEQUENCE;
    request->alignmentType = SA::DEFAULT_ALIGNMENT_TYPE;
    request->alphabet = SA::DEFAULT_ALPHABET;
    request->alphabetSize = SA::DEFAULT_ALPHABET_SIZE;
    request->gapPenalty = SA::DEFAULT_GAP_PENALTY;
    request->textNumBytes = 0;
    request->patternNumBytes = 0;

    bool wantGap = false, wantMatrix = false, haveMatrix = false;
    for (int i = 1; i < argc; ++i) {
        const auto flag = SA::argumentMap.find(argv[i]);
        if (flag != SA::argumentMap.end()) {
            switch (flag->second) {
            case SA::CPU: case SA::GPU: request->deviceType = flag->second; break;
            case SA::DNA: case SA::PROTEIN: request->sequenceType = flag->second; break;
            case SA::GLOBAL: case SA::LOCAL: case SA::SEMI_GLOBAL: request->alignmentType = flag->second; break;
            case SA::SCORE_MATRIX: wantMatrix = true; haveMatrix = false; break;
            case SA::GAP_PENALTY: wantGap = true; break;
            }
            const bool dna = request->sequenceType == SA::DNA;
            request->alphabet = dna ? SA::DNA_ALPHABET : SA::PROTEIN_ALPHABET;
            request->alphabetSize = dna ? (int)SA::NUM_DNA_CHARS : (int)SA::NUM_PROTEIN_CHARS;
        } else if (wantGap) {              // (the gap penalty wins when both flags are pending, like the reference)
            try {
                request->gapPenalty = std::stoi(argv[i]);
            } catch (...) {
                std::cerr << SA::GAP_PENALTY_NOT_READ_ERROR;
                return 1;
            }
            wantGap = false;
        } else if (wantMatrix) {
            if (parseScoreMatrixFile(argv[i], request->alphabetSize, request->scoreMatrix) == -1) {
                std::cerr << SA::SCORE_MATRIX_NOT_READ_ERROR;
                return 1;
            }
            wantMatrix = false;
            haveMatrix = true;
        } else if (readSequenceFile(argv[i], request) == -1) {
            std::cerr << SA::SEQ_NOT_READ_ERROR;
            return 1;
        }
    }
    if (request->textNumBytes == 0 || request->patternNumBytes == 0) {
        std::cerr << SA::SEQ_NOT_READ_ERROR << SA::USAGE;
        return 1;
    }
    if (request->textNumBytes < request->patternNumBytes) {
        std::swap(request->textBytes, request->patternBytes);
        std::swap(request->textNumBytes, request->patternNumBytes);
    }
    if (!haveMatrix)
        parseScoreMatrixFile(request->sequenceType == SA::DNA ? SA::DEFAULT_DNA_SCORE_MATRIX_FILE : SA::DEFAULT_PROTEIN_SCORE_MATRIX_FILE,
                             request->alphabetSize, request->scoreMatrix);
    return 0;
}

// utilities.cpp:253-315
void prettyAlignmentPrint(SA::Response &response, std::ostream &stream)
{
    const uint64_t need = sa_pretty_print(response.alignedTextBytes, response.alignedPatternBytes, response.numAlignmentBytes,
                                          response.startInAlignedText, response.startInAlignedPattern, response.score,
                                          nullptr, 0, nullptr, nullptr);
    std::string out(need, '\0');
    if (need) sa_pretty_print(response.alignedTextBytes, response.alignedPatternBytes, response.numAlignmentBytes,
                              response.startInAlignedText, response.startInAlignedPattern, response.score, &out[0], need, nullptr, nullptr);
    stream << out;
}

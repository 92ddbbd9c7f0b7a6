// sa_frontend.cpp -- the callers' side of the hot path: sequence files and score-matrix files -> the byte / int
// buffers of a Request.  Host C++ behind the same C ABI; semantics are those of the
// reference's utilities.cpp, restated (not copied):
//
//   validateAndTransform (utilities.cpp:31-63)   a byte stream becomes alphabet indices, one per byte:
//       * a '>' met outside a header starts a header that lasts up to the next '\n' (FASTA);
//       * bytes above 'Z' are shifted down by 32 (that is how lower case is folded), then everything outside
//         'A'..'Z' is dropped -- digits, blanks, line ends, '*', '-', bytes >= 0x80 (negative as signed char);
//       * a letter that is not in the alphabet is an error: the reference prints it and returns 0.
//   readSequenceFile (utilities.cpp:65-104)       whole file -> validateAndTransform.
//   parseScoreMatrixFile (utilities.cpp:106-129)  alpha*alpha whitespace-separated integers, row-major with stride
//       alpha; a token that is not an integer returns -1; a MISSING FILE returns 0 and leaves the buffer as it was
//       (the reference's silent failure; sa_parse_score_matrix_file reports it as SA_ERR_ARGUMENT instead and the
//       C++ mirror keeps the reference's return value).
//
// New surface for batches (the reference has none): sa_read_fasta_batch turns a multi-record FASTA file into the
// CSR arrays sa_align_batch takes, one record per sequence, in one pass with a 256-entry table per alphabet.
#include "../../include/sa_b200.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

// byte -> alphabet index, 0xFE = dropped, 0xFF = letter outside the alphabet
void build_table(const char *alphabet, int alpha, unsigned char (&tab)[256])
{
    for (int b = 0; b < 256; ++b) {
        const int sc = (b >= 128) ? b - 256 : b;              // the reference works on (signed) char
        const int up = sc > 90 ? sc - 32 : sc;
        if (up < 65 || up > 90) { tab[b] = 0xFE; continue; }
        tab[b] = 0xFF;
        for (int k = 0; k < alpha; ++k)
            if (alphabet[k] == (char)up) { tab[b] = (unsigned char)k; break; }
    }
}

bool read_whole_file(const char *path, std::vector<unsigned char> &buf)
{
    FILE *f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf.resize(sz > 0 ? (size_t)sz : 0);
    const size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
    std::fclose(f);
    buf.resize(got);
    return true;
}

} // namespace

extern "C" {

// In place: buf[0..return) holds the alphabet indices.  Returns the number of residues, 0 when a letter is not in
// the alphabet (*bad_letter receives it, upper-cased) -- exactly the reference's return value.
int64_t sa_validate_and_transform(char *buf, uint64_t len, const char *alphabet, int alphabet_size, char *bad_letter)
{
    if (!buf || !alphabet || alphabet_size < 1) return 0;
    unsigned char tab[256];
    build_table(alphabet, alphabet_size, tab);
    bool header = false;
    uint64_t n = 0;
    for (uint64_t i = 0; i < len; ++i) {
        const unsigned char b = (unsigned char)buf[i];
        if (!header && b == '>') { header = true; continue; }
        if (header) { if (b == '\n') header = false; continue; }
        const unsigned char t = tab[b];
        if (t == 0xFE) continue;
        if (t == 0xFF) {
            if (bad_letter) { const int sc = b >= 128 ? b - 256 : b; *bad_letter = (char)(sc > 90 ? sc - 32 : sc); }
            return 0;
        }
        buf[n++] = (char)t;
    }
    return (int64_t)n;
}

void sa_free(void *p) { std::free(p); }

// Whole file -> alphabet indices in a malloc'ed buffer (*out, release with sa_free).  SA_ERR_ARGUMENT: the file
// cannot be read, holds no residue or holds a letter outside the alphabet (*bad_letter, 0 otherwise).
int sa_read_sequence_file(const char *path, const char *alphabet, int alphabet_size, uint8_t **out, uint64_t *n,
                          char *bad_letter)
{
    if (!path || !alphabet || !out || !n) return SA_ERR_ARGUMENT;
    if (bad_letter) *bad_letter = 0;
    *out = nullptr; *n = 0;
    std::vector<unsigned char> buf;
    if (!read_whole_file(path, buf)) return SA_ERR_ARGUMENT;
    const int64_t k = sa_validate_and_transform(reinterpret_cast<char *>(buf.data()), buf.size(), alphabet, alphabet_size, bad_letter);
    if (k <= 0) return SA_ERR_ARGUMENT;
    uint8_t *p = static_cast<uint8_t *>(std::malloc((size_t)k));
    if (!p) return SA_ERR_MEMORY;
    std::memcpy(p, buf.data(), (size_t)k);
    *out = p; *n = (uint64_t)k;
    return SA_OK;
}

// alpha*alpha integers, row-major.  SA_OK, SA_ERR_ARGUMENT (file missing / not enough integers).
int sa_parse_score_matrix_file(const char *path, int alphabet_size, int32_t *matrix)
{
    if (!path || !matrix || alphabet_size < 1) return SA_ERR_ARGUMENT;
    FILE *f = std::fopen(path, "r");
    if (!f) return SA_ERR_ARGUMENT;
    int rc = SA_OK;
    for (int i = 0; i < alphabet_size * alphabet_size; ++i) {
        int v;
        if (std::fscanf(f, "%d", &v) != 1) { rc = SA_ERR_ARGUMENT; break; }
        matrix[i] = v;
    }
    std::fclose(f);
    return rc;
}

// Multi-record FASTA -> CSR: record r is (*residues)[(*offsets)[r] .. (*offsets)[r+1]).  Text before the first '>'
// counts as a record of its own when it holds residues (so a plain one-sequence file works too).  Both arrays are
// malloc'ed (sa_free).  Header handling, case folding and dropped bytes follow validateAndTransform.
int sa_read_fasta_batch(const char *path, const char *alphabet, int alphabet_size, uint8_t **residues,
                        int64_t **offsets, uint64_t *n_records, char *bad_letter)
{
    if (!path || !alphabet || !residues || !offsets || !n_records) return SA_ERR_ARGUMENT;
    if (bad_letter) *bad_letter = 0;
    *residues = nullptr; *offsets = nullptr; *n_records = 0;
    std::vector<unsigned char> buf;
    if (!read_whole_file(path, buf)) return SA_ERR_ARGUMENT;
    unsigned char tab[256];
    build_table(alphabet, alphabet_size, tab);
    std::vector<int64_t> off;
    off.push_back(0);
    bool header = false;
    uint64_t n = 0;
    for (size_t i = 0; i < buf.size(); ++i) {
        const unsigned char b = buf[i];
        if (!header && b == '>') {
            header = true;
            if ((int64_t)n > off.back()) off.push_back((int64_t)n);      // close the record in progress
            continue;
        }
        if (header) { if (b == '\n') header = false; continue; }
        const unsigned char t = tab[b];
        if (t == 0xFE) continue;
        if (t == 0xFF) {
            if (bad_letter) { const int sc = b >= 128 ? b - 256 : b; *bad_letter = (char)(sc > 90 ? sc - 32 : sc); }
            return SA_ERR_ARGUMENT;
        }
        buf[n++] = t;
    }
    if ((int64_t)n > off.back()) off.push_back((int64_t)n);
    const size_t recs = off.size() - 1;
    uint8_t *r = static_cast<uint8_t *>(std::malloc(n ? n : 1));
    int64_t *o = static_cast<int64_t *>(std::malloc(off.size() * sizeof(int64_t)));
    if (!r || !o) { std::free(r); std::free(o); return SA_ERR_MEMORY; }
    std::memcpy(r, buf.data(), n);
    std::memcpy(o, off.data(), off.size() * sizeof(int64_t));
    *residues = r; *offsets = o; *n_records = recs;
    return SA_OK;
}

// ---- the step after the path: prettyAlignmentPrint (utilities.cpp:253-315), restated ---------------------------
// Writes the reference's report -- blocks of 50 columns with 1-based positions and a match line ('|' identical,
// ' ' gap, '.' mismatch), then "# Length / # Identity / # Gaps / # Score" -- into out (capacity cap) and returns the
// number of bytes the full report needs (call with cap = 0 to size the buffer).  identity / gaps (may be NULL)
// receive the two counts.  An empty alignment prints nothing, like the reference.
uint64_t sa_pretty_print(const char *aligned_text, const char *aligned_pattern, uint64_t len, uint64_t start_text,
                         uint64_t start_pattern, int32_t score, char *out, uint64_t cap, uint64_t *identity,
                         uint64_t *gaps)
{
    if (identity) *identity = 0;
    if (gaps) *gaps = 0;
    if (len == 0 || !aligned_text || !aligned_pattern) return 0;
    std::string s;
    s.reserve((size_t)len * 4 + 256);
    const unsigned CHARS_PER_LINE = 50;
    int width = 0;
    int maxI = (int)(len + (start_text > start_pattern ? start_text : start_pattern));      // int, like the reference
    do { maxI /= 10; ++width; } while (maxI != 0);
    auto padded = [&](const std::string &v) { if ((int)v.size() < width) s.append((size_t)width - v.size(), ' '); s += v; };
    uint64_t nId = 0, nGap = 0;
    for (uint64_t i = 0; i < len; i += CHARS_PER_LINE) {
        const uint64_t end = (i + CHARS_PER_LINE < len) ? i + CHARS_PER_LINE : len;
        padded(std::to_string(i + 1 + start_text)); s += ' ';
        s.append(aligned_text + i, (size_t)(end - i));
        s += "   "; s += std::to_string(end + start_pattern); s += " \n";
        padded(" "); s += ' ';
        for (uint64_t j = i; j < end; ++j) {
            if (aligned_text[j] == aligned_pattern[j]) { s += '|'; ++nId; }
            else if (aligned_text[j] == '-' || aligned_pattern[j] == '-') { s += ' '; ++nGap; }
            else s += '.';
        }
        s += '\n';
        padded(std::to_string(i + 1)); s += ' ';
        s.append(aligned_pattern + i, (size_t)(end - i));
        s += "   "; s += std::to_string(end); s += "\n\n";
    }
    char num[64];
    s += "# Length: \t" + std::to_string(len) + "\n";
    std::snprintf(num, sizeof num, "%.3g", (double)nId / ((double)len * 1.0) * 100);
    s += "# Identity: \t" + std::to_string(nId) + "/" + std::to_string(len) + " (" + num + "%)\n";
    std::snprintf(num, sizeof num, "%.3g", (double)nGap / ((double)len * 1.0) * 100);
    s += "# Gaps: \t" + std::to_string(nGap) + "/" + std::to_string(len) + " (" + num + "%)\n";
    s += "# Score: \t" + std::to_string(score) + "\n";
    if (identity) *identity = nId;
    if (gaps) *gaps = nGap;
    if (out && cap) std::memcpy(out, s.data(), s.size() < cap ? s.size() : (size_t)cap);
    return s.size();
}

} // extern "C"

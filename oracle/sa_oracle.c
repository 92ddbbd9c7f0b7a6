/*
 * sa_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's
 * pairwise alignment algorithm (robertszafa/sequence-alignment-gpu,
 * alignSequenceCPU.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path
 * (sequence-alignment-gpu_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (a) every golden vector the reference's own tests hold for this path
 *       (tests/tests.cu:116-368, committed as tests/golden/reference_goldens.json),
 *   (b) the unmodified reference compiled from /root/reference into
 *       oracle/_ref/libsa_ref.so (oracle/Makefile), on the data/ sweep and on
 *       randomised inputs, field by field.
 *
 * Conventions (same as the reference):
 *   text    t[0..n-1]  -> columns, numCols = n+1
 *   pattern p[0..m-1]  -> rows,    numRows = m+1
 *   sequences are alphabet indices (one per byte), score matrix is row-major with
 *   stride alphabetSize and is indexed [pattern][text] (alignSequenceCPU.cpp:172,256),
 *   gap is a positive magnitude that is subtracted.
 *   DIRECTION { LEFT=0, DIAG=1, TOP=2, STOP=3 } (SequenceAlignment.hpp:122).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { SA_LEFT = 0, SA_DIAG = 1, SA_TOP = 2, SA_STOP = 3 };

typedef struct {
    int32_t  score;
    uint64_t aln_len;
    uint64_t start_text;
    uint64_t start_pattern;
} sa_oracle_result;

static inline int imax(int a, int b) { return a > b ? a : b; }

/* One cell of the recurrence; follows alignSequenceCPU.cpp:170-192 (SW) and
 * :254-273 (NW): diagonal only on a strict win, gap tie goes LEFT. */
static inline int cell(int left, int top, int diag, int s, int g, char *dir)
{
    const int fromLeft = left - g;
    const int fromTop  = top - g;
    const int fromDiag = diag + s;
    const int withGap  = imax(fromLeft, fromTop);
    const int best     = imax(fromDiag, withGap);
    if (fromDiag > withGap)        *dir = SA_DIAG;
    else if (fromLeft >= fromTop)  *dir = SA_LEFT;
    else                           *dir = SA_TOP;
    return best;
}

/* NW fill, alignSequenceCPU.cpp:203-284.  M is (m+1)*(n+1) bytes. Returns H(m,n). */
int sa_oracle_fill_nw(char *M, uint64_t m, uint64_t n, const uint8_t *text,
                      const uint8_t *pattern, const int32_t *S, int alpha, int g)
{
    const uint64_t cols = n + 1;
    int *prev = (int *)malloc(sizeof(int) * cols);
    int *cur  = (int *)malloc(sizeof(int) * cols);
    if (!prev || !cur) { free(prev); free(cur); return 0; }
    for (uint64_t j = 0; j < cols; ++j) { cur[j] = (int)j * -g; M[j] = SA_LEFT; }   /* :232-236 */
    for (uint64_t i = 1; i <= m; ++i) {
        int *tmp = prev; prev = cur; cur = tmp;
        char *row = M + i * cols;
        cur[0] = (int)i * -g;                                                        /* :247 */
        row[0] = SA_TOP;                                                             /* :248 */
        const int32_t *Srow = S + (int)pattern[i - 1] * alpha;                       /* :256 */
        for (uint64_t j = 1; j <= n; ++j)
            cur[j] = cell(cur[j - 1], prev[j], prev[j - 1], Srow[text[j - 1]], g, &row[j]);
    }
    const int score = cur[cols - 1];                                                 /* :279 */
    free(prev); free(cur);
    return score;
}

/* SW fill, alignSequenceCPU.cpp:116-201.  Returns best score, *argmax = first
 * (row-major) cell attaining it (strict '>' update, :191-192); 0 if none > 0. */
int sa_oracle_fill_sw(char *M, uint64_t m, uint64_t n, const uint8_t *text,
                      const uint8_t *pattern, const int32_t *S, int alpha, int g,
                      uint64_t *argmax)
{
    const uint64_t cols = n + 1;
    int *prev = (int *)malloc(sizeof(int) * cols);
    int *cur  = (int *)malloc(sizeof(int) * cols);
    if (!prev || !cur) { free(prev); free(cur); *argmax = 0; return 0; }
    for (uint64_t j = 0; j < cols; ++j) { cur[j] = 0; M[j] = SA_STOP; }              /* :145-149 */
    int best = 0; uint64_t bestIJ = 0;
    for (uint64_t i = 1; i <= m; ++i) {
        int *tmp = prev; prev = cur; cur = tmp;
        char *row = M + i * cols;
        cur[0] = 0; row[0] = SA_STOP;                                                /* :163-164 */
        const int32_t *Srow = S + (int)pattern[i - 1] * alpha;                       /* :172 */
        for (uint64_t j = 1; j <= n; ++j) {
            char d;
            const int b = cell(cur[j - 1], prev[j], prev[j - 1], Srow[text[j - 1]], g, &d);
            row[j] = b > 0 ? d : SA_STOP;                                            /* :189 */
            cur[j] = imax(0, b);                                                     /* :190 */
            if (cur[j] > best) { best = cur[j]; bestIJ = i * cols + j; }             /* :191-192 */
        }
    }
    free(prev); free(cur);
    *argmax = bestIJ;
    return best;
}

static void reverse_bytes(char *b, uint64_t len)
{
    for (uint64_t a = 0, z = len; a + 1 < z; ++a, --z) { char t = b[a]; b[a] = b[z - 1]; b[z - 1] = t; }
}

/* NW traceback, alignSequenceCPU.cpp:64-114 (border override :78-81, clamped
 * indices :100-101). */
void sa_oracle_traceback_nw(const char *M, uint64_t m, uint64_t n, const uint8_t *text,
                            const uint8_t *pattern, const char *alphabet, int alpha,
                            char *outT, char *outP, sa_oracle_result *res)
{
    const uint64_t cols = n + 1;
    uint64_t curr = (m + 1) * cols - 1;
    int ti = (int)n - 1, pi = (int)m - 1;
    uint64_t len = 0;
    const char GAP = alphabet[alpha];
    while (curr > 0) {
        char dir = M[curr];
        if (curr % cols == 0)   dir = SA_TOP;
        else if (curr < cols)   dir = SA_LEFT;
        const int takeT = (dir == SA_DIAG || dir == SA_LEFT);
        const int takeP = (dir == SA_DIAG || dir == SA_TOP);
        outT[len] = takeT ? alphabet[text[ti]] : GAP;
        outP[len] = takeP ? alphabet[pattern[pi]] : GAP;
        ++len;
        ti = imax(0, ti - takeT);
        pi = imax(0, pi - takeP);
        curr -= (dir == SA_LEFT) ? 1 : (dir == SA_DIAG) ? cols + 1 : cols;
    }
    res->aln_len = len;
    res->start_text = (uint64_t)(int64_t)ti;
    res->start_pattern = (uint64_t)(int64_t)pi;
    reverse_bytes(outT, len);
    reverse_bytes(outP, len);
}

/* SW traceback, alignSequenceCPU.cpp:10-62.  Note the break-before-update on
 * reaching row 0 / column 0 (:45-46) and the int(-1) -> uint64 wrap when the
 * best score is 0 (:13-14,:56-57). */
void sa_oracle_traceback_sw(const char *M, uint64_t start, uint64_t m, uint64_t n,
                            const uint8_t *text, const uint8_t *pattern,
                            const char *alphabet, int alpha,
                            char *outT, char *outP, sa_oracle_result *res)
{
    (void)m;
    const uint64_t cols = n + 1;
    int ti = (int)(start % cols) - 1;
    int pi = (int)(start / cols) - 1;
    uint64_t len = 0, curr = start;
    const char GAP = alphabet[alpha];
    while (M[curr] != SA_STOP) {
        const char dir = M[curr];
        const int takeT = (dir == SA_DIAG || dir == SA_LEFT);
        const int takeP = (dir == SA_DIAG || dir == SA_TOP);
        outT[len] = takeT ? alphabet[text[ti]] : GAP;
        outP[len] = takeP ? alphabet[pattern[pi]] : GAP;
        ++len;
        curr -= (dir == SA_LEFT) ? 1 : (dir == SA_DIAG) ? cols + 1 : cols;
        if (curr % cols == 0 || curr < cols) break;
        ti = imax(0, ti - takeT);
        pi = imax(0, pi - takeP);
    }
    res->aln_len = len;
    res->start_text = (uint64_t)(int64_t)ti;
    res->start_pattern = (uint64_t)(int64_t)pi;
    reverse_bytes(outT, len);
    reverse_bytes(outP, len);
}

/* alignSequenceCPU, alignSequenceCPU.cpp:287-333.  mode 0 = global, 1 = local.
 * outT/outP need capacity >= m+n.  If dirs_out != NULL it receives the full
 * (m+1)*(n+1) byte direction matrix (debug aid for the parity tests).
 * Returns 0 on success, 1 on allocation failure (reference: MEM_ERROR + 1). */
int sa_oracle_align(int mode, int alpha, const int32_t *S, int g, const char *alphabet,
                    const uint8_t *text, uint64_t n, const uint8_t *pattern, uint64_t m,
                    char *outT, char *outP, sa_oracle_result *res, char *dirs_out)
{
    const uint64_t cells = (m + 1) * (n + 1);
    char *M = dirs_out ? dirs_out : (char *)malloc(cells);
    if (!M) return 1;
    if (mode == 0) {
        res->score = sa_oracle_fill_nw(M, m, n, text, pattern, S, alpha, g);
        sa_oracle_traceback_nw(M, m, n, text, pattern, alphabet, alpha, outT, outP, res);
    } else {
        uint64_t start = 0;
        res->score = sa_oracle_fill_sw(M, m, n, text, pattern, S, alpha, g, &start);
        sa_oracle_traceback_sw(M, start, m, n, text, pattern, alphabet, alpha, outT, outP, res);
    }
    if (!dirs_out) free(M);
    return 0;
}

/* Score-only variants with two rolling rows (O(n) memory) for sizes where the
 * reference's 1 B/cell matrix cannot be allocated.
 * Same recurrence as above; validated against sa_oracle_align in tests. */
int sa_oracle_score_only(int mode, int alpha, const int32_t *S, int g,
                         const uint8_t *text, uint64_t n, const uint8_t *pattern, uint64_t m,
                         int32_t *score, uint64_t *argmax)
{
    const uint64_t cols = n + 1;
    int *prev = (int *)malloc(sizeof(int) * cols);
    int *cur  = (int *)malloc(sizeof(int) * cols);
    if (!prev || !cur) { free(prev); free(cur); return 1; }
    int best = 0; uint64_t bestIJ = 0;
    for (uint64_t j = 0; j < cols; ++j) cur[j] = mode == 0 ? (int)j * -g : 0;
    for (uint64_t i = 1; i <= m; ++i) {
        int *tmp = prev; prev = cur; cur = tmp;
        cur[0] = mode == 0 ? (int)i * -g : 0;
        const int32_t *Srow = S + (int)pattern[i - 1] * alpha;
        for (uint64_t j = 1; j <= n; ++j) {
            char d;
            int b = cell(cur[j - 1], prev[j], prev[j - 1], Srow[text[j - 1]], g, &d);
            if (mode != 0) {
                b = imax(0, b);
                if (b > best) { best = b; bestIJ = i * cols + j; }
            }
            cur[j] = b;
        }
    }
    *score = mode == 0 ? cur[cols - 1] : best;
    if (argmax) *argmax = bestIJ;
    free(prev); free(cur);
    return 0;
}

/* Re-score an emitted alignment (size-independent property used at the full
 * BASELINE sizes): sum of substitution scores and gap penalties over the
 * aligned columns must reproduce the reported score. letters are ASCII. */
int64_t sa_oracle_rescore(const char *alnT, const char *alnP, uint64_t len,
                          const char *alphabet, int alpha, const int32_t *S, int g)
{
    int8_t idx[256];
    memset(idx, -1, sizeof idx);
    for (int a = 0; a < alpha; ++a) idx[(unsigned char)alphabet[a]] = (int8_t)a;
    const char GAP = alphabet[alpha];
    int64_t total = 0;
    for (uint64_t k = 0; k < len; ++k) {
        if (alnT[k] == GAP || alnP[k] == GAP) total -= g;
        else total += S[idx[(unsigned char)alnP[k]] * alpha + idx[(unsigned char)alnT[k]]];
    }
    return total;
}

/* ---- batch checker (tests/, bench.py "verified"): aligns the listed pairs of a CSR batch with sa_oracle_align on
 * `nthreads` host threads and compares EVERY Response field and both strings (alignSequenceCPU.cpp:287-333 is the
 * truth) with a result set laid out like sa_align_batch's: results[p] = {int32 score, pad, u64 aln_len, u64
 * start_text, u64 start_pattern}, strings of pair p at aligned_*[aln_off[p] .. + aln_len).  idx == NULL checks pairs
 * 0..n_idx-1.  Returns the number of mismatching pairs; *first_bad (may be NULL) receives the smallest such pair
 * index or -1. */
typedef struct {
    int mode, alpha, g; const int32_t *S; const char *alphabet;
    const uint8_t *text; const int64_t *toff; const uint8_t *pattern; const int64_t *poff;
    const uint64_t *idx; uint64_t n_idx;
    const sa_oracle_result *results; const uint64_t *aln_off; const char *alnT; const char *alnP;
    uint64_t next; uint64_t bad; int64_t first_bad;
    pthread_mutex_t mu;
} check_job;

static void *check_worker(void *arg)
{
    check_job *J = (check_job *)arg;
    char *oT = NULL, *oP = NULL; uint64_t cap = 0;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        const uint64_t k0 = J->next; J->next += 16;
        pthread_mutex_unlock(&J->mu);
        if (k0 >= J->n_idx) break;
        for (uint64_t k = k0; k < k0 + 16 && k < J->n_idx; ++k) {
            const uint64_t p = J->idx ? J->idx[k] : k;
            const uint64_t n = (uint64_t)(J->toff[p + 1] - J->toff[p]), m = (uint64_t)(J->poff[p + 1] - J->poff[p]);
            if (n + m + 1 > cap) { cap = 2 * (n + m) + 64; free(oT); free(oP); oT = (char *)malloc(cap); oP = (char *)malloc(cap); }
            sa_oracle_result r; memset(&r, 0, sizeof r);
            int ok = oT && oP && sa_oracle_align(J->mode, J->alpha, J->S, J->g, J->alphabet, J->text + J->toff[p], n,
                                                 J->pattern + J->poff[p], m, oT, oP, &r, NULL) == 0;
            const sa_oracle_result *g = &J->results[p];
            ok = ok && g->score == r.score && g->aln_len == r.aln_len && g->start_text == r.start_text &&
                 g->start_pattern == r.start_pattern &&
                 memcmp(J->alnT + J->aln_off[p], oT, r.aln_len) == 0 && memcmp(J->alnP + J->aln_off[p], oP, r.aln_len) == 0;
            if (!ok) {
                pthread_mutex_lock(&J->mu);
                J->bad++;
                if (J->first_bad < 0 || (int64_t)p < J->first_bad) J->first_bad = (int64_t)p;
                pthread_mutex_unlock(&J->mu);
            }
        }
    }
    free(oT); free(oP);
    return NULL;
}

uint64_t sa_oracle_check_batch(int mode, int alpha, const int32_t *S, int g, const char *alphabet,
                               const uint8_t *text, const int64_t *toff, const uint8_t *pattern, const int64_t *poff,
                               const uint64_t *idx, uint64_t n_idx, const void *results, const uint64_t *aln_off,
                               const char *alnT, const char *alnP, int nthreads, int64_t *first_bad)
{
    check_job J;
    memset(&J, 0, sizeof J);
    J.mode = mode; J.alpha = alpha; J.g = g; J.S = S; J.alphabet = alphabet;
    J.text = text; J.toff = toff; J.pattern = pattern; J.poff = poff; J.idx = idx; J.n_idx = n_idx;
    J.results = (const sa_oracle_result *)results; J.aln_off = aln_off; J.alnT = alnT; J.alnP = alnP;
    J.first_bad = -1;
    pthread_mutex_init(&J.mu, NULL);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads - 1; ++t) if (pthread_create(&th[started], NULL, check_worker, &J) == 0) ++started;
    check_worker(&J);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&J.mu);
    if (first_bad) *first_bad = J.first_bad;
    return J.bad;
}

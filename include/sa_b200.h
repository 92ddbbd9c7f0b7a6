/*
 * sa_b200.h -- C ABI of the B200-native pairwise alignment hot path.
 *
 * This is the drop-in boundary for the `-g` device path of
 * robertszafa/sequence-alignment-gpu: every entry point below replaces a piece
 * of /root/reference/alignSequenceGPU.cu (cited per function).  Plain pointers
 * and sizes only; no C++/torch types.  The functions never print and never
 * throw; they return SA_OK (0) or a negative sa_status.  There is NO CPU
 * fallback: without a CUDA device every compute entry returns SA_ERR_NO_DEVICE.
 *
 * Conventions (identical to the reference's Request, SequenceAlignment.hpp:71-99):
 *   - sequences are alphabet indices 0..alphabet_size-1, one per byte
 *     (utilities.cpp:47-52), NOT ASCII;
 *   - score_matrix is row-major with stride alphabet_size and is indexed
 *     [pattern letter][text letter] (alignSequenceCPU.cpp:172,256);
 *   - gap is a positive magnitude that is subtracted (linear gap penalty);
 *   - alphabet has alphabet_size+1 chars, the last one is the gap char '-'
 *     (SequenceAlignment.hpp:56-58);
 *   - results are bit-identical to the reference's alignSequenceCPU
 *     (alignSequenceCPU.cpp:287-333): score, numAlignmentBytes, both start
 *     indices (incl. the 2^64-1 wrap for a zero-score local alignment) and
 *     both aligned strings (forward order, no terminator).
 */
#ifndef SA_B200_H
#define SA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SA_OK = 0,
    SA_ERR_NO_DEVICE = -1,   /* no CUDA device / driver */
    SA_ERR_MEMORY = -2,      /* device or pinned allocation failed (reference: MEM_ERROR, alignSequenceGPU.cu:541-546) */
    SA_ERR_COPY = -3,        /* host<->device copy failed (reference: "could not copy from device memory", :588-594) */
    SA_ERR_ARGUMENT = -4,    /* null pointer, empty sequence, alphabet_size > 32, ...; from the HOST-buffer entry points
                              * (sa_align, sa_fill_only, sa_align_batch) also a residue >= alphabet_size.  The
                              * device-resident entry points do not scan their inputs: there such a residue is
                              * read as letter alphabet_size-1 (clamped, never out of bounds). */
    SA_ERR_SCORE_RANGE = -5, /* |score| > 4064 or gap > 2^24; also: a matrix beyond +-31 given to the device-resident
                              * batch / slice entry points (such matrices run through the single-pair kernels only) */
    SA_ERR_LAUNCH = -6,      /* kernel launch / execution error */
    SA_ERR_CAPACITY = -7     /* caller-provided output buffer too small */
} sa_status;

enum { SA_GLOBAL = 0, SA_LOCAL = 1 };   /* programArgs::GLOBAL / LOCAL, SequenceAlignment.hpp:17 */

/* Scoring scheme: the non-sequence half of Request (SequenceAlignment.hpp:87-91). */
typedef struct {
    int32_t        mode;           /* SA_GLOBAL (Needleman-Wunsch) or SA_LOCAL (Smith-Waterman) */
    int32_t        alphabet_size;  /* 4 (DNA) or 23 (protein); any 2..32 is accepted */
    const int32_t *score_matrix;   /* alphabet_size^2 ints, [pattern][text] */
    int32_t        gap;            /* >= 0 */
    const char    *alphabet;       /* alphabet_size+1 chars */
} sa_scoring;

/* One alignment result: Response (SequenceAlignment.hpp:101-120) as a POD. */
typedef struct {
    int32_t  score;
    uint64_t aln_len;         /* numAlignmentBytes */
    uint64_t start_text;      /* startInAlignedText */
    uint64_t start_pattern;   /* startInAlignedPattern */
} sa_result;

/* Timing of the last call on a context, microseconds, from CUDA events on the
 * context's stream.  Serves the reference's `#define BENCHMARK` behaviour
 * (alignSequenceGPU.cu:555-558,613-626: return elapsed us of fill + D2H). */
typedef struct {
    double h2d_us;        /* host->device copies */
    double fill_us;       /* DP fill kernels */
    double traceback_us;  /* device traceback + string emission */
    double d2h_us;        /* device->host copies of results */
    double total_us;      /* first event to last event */
    uint64_t cells;       /* sum over pairs of (pattern_len+1)*(text_len+1) (tests/benchmarks.cu:85) */
    uint32_t kernel_launches;
    uint64_t h2d_bytes;   /* sa_align_batch: bytes copied host->device ... */
    uint64_t d2h_bytes;   /* ... and device->host by the last call (0 from the other entry points) */
} sa_timing;

typedef struct sa_context sa_context;

/* ---- lifecycle ------------------------------------------------------------ */

/* Number of usable CUDA devices (0 when none). */
int sa_device_count(void);

/* A context is NOT thread-safe: it owns streams, events and growing workspaces that its calls reuse.  Use one
 * context per host thread (several contexts on one device are fine; the tests run up to eight).
 * Creates a context bound to `device` (the reference hard-codes device 0,
 * alignSequenceGPU.cu:476).  Owns a stream, pinned staging and a growing
 * device workspace, so repeated calls do not pay cudaMalloc/cudaMallocHost per
 * call like initMemory does (alignSequenceGPU.cu:362-461). */
int sa_create(int device, sa_context **out);
void sa_destroy(sa_context *ctx);
/* Tuning knobs of a context by name (the reference has none: its launch shape is compiled in, alignSequenceGPU.cu:7-12).
 * Sizes in MB.  "dev_dirs_budget_mb" / "dirs_budget_mb" / "host_dirs_budget_mb": direction words per chunk of the
 * device-resident batch / of the staged host batch / of the slot pipeline; "batch_min_chunks": floor of the chunk count
 * of a device batch (0 = automatic); "tb_blocks_per_sm": traceback blocks next to the following chunk's fill;
 * "ckpt_rows" / "ckpt_limit_mb" / "ckpt_chunk_mb": checkpointed traceback of global alignments -- rows per chunk,
 * direction bytes above which it is taken, direction bytes of one chunk (0 = automatic).  Unknown names and values
 * out of range return SA_ERR_ARGUMENT.  The SA_* environment variables remain as developer overrides. */
int sa_set_option(sa_context *ctx, const char *name, long long value);
int sa_get_option(const sa_context *ctx, const char *name, long long *value);

const char *sa_status_string(int status);
int sa_last_timing(const sa_context *ctx, sa_timing *out);
/* cudaError_t of the last failing CUDA call on this context (0 if none): the reference
 * swallows these (no cudaGetLastError after launches, alignSequenceGPU.cu:575-611). */
int sa_last_cuda_error(const sa_context *ctx);
/* The context's own (non-blocking) cudaStream_t, used by the host-buffer entry points. */
void *sa_context_stream(const sa_context *ctx);

/* ---- single pair: replaces alignSequenceGPU (alignSequenceGPU.cu:463-653) ---
 * HOST buffers in, HOST buffers out.  aligned_text / aligned_pattern must hold
 * at least text_len + pattern_len bytes each (capacity given in out_capacity).
 * Any pair size: short pairs go through the batch kernel, long ones through
 * the persistent wavefront kernel with the direction matrix in HBM. */
int sa_align(sa_context *ctx, const sa_scoring *scoring,
             const uint8_t *text, uint64_t text_len,
             const uint8_t *pattern, uint64_t pattern_len,
             sa_result *result, char *aligned_text, char *aligned_pattern,
             uint64_t out_capacity);

/* Fill only, no traceback (what the reference times under BENCHMARK).  Returns
 * the score (and for SA_LOCAL the row-major-first arg-max as i*(n+1)+j). */
int sa_fill_only(sa_context *ctx, const sa_scoring *scoring,
                 const uint8_t *text, uint64_t text_len,
                 const uint8_t *pattern, uint64_t pattern_len,
                 int32_t *score, uint64_t *argmax);

/* DEVICE-resident single pair through the long-pair kernels: d_text / d_pattern hold the
 * residues, d_aligned_* have capacity text_len+pattern_len and receive the strings
 * right-aligned (they end at the buffer end), d_result4 receives
 * {aln_len, start_text, start_pattern, score}.  Enqueued on `stream`, no synchronisation. */
int sa_align_device(sa_context *ctx, const sa_scoring *scoring,
                    const uint8_t *d_text, uint64_t text_len,
                    const uint8_t *d_pattern, uint64_t pattern_len,
                    char *d_aligned_text, char *d_aligned_pattern, uint64_t *d_result4,
                    void *stream);

/* ---- one GLOBAL alignment split into column slices, one slice per GPU (BASELINE config 5).
 * The reference's alignSequenceGPU cannot hold such a pair (alignSequenceGPU.cu:410-416 caps the
 * direction matrix at one device); the slices replace that single-device matrix.
 *
 * sa_strip_fill: fill DP columns col0+1 .. col0+text_len for all pattern_len rows.  d_left_col /
 * d_right_col are DEVICE arrays of pattern_len+1 int32 holding 4*H(i, col0) / 4*H(i, col0+text_len),
 * i = 0..pattern_len (the scaled form the kernels carry); d_left_col is NULL exactly when col0 == 0.
 * text_total_len is the length of the whole text (it only steers the traceback's search band).
 * d_score (device, may be NULL) receives H(pattern_len, col0+text_len).  The slice's directions stay
 * in the context; d_text / d_pattern must stay valid until sa_strip_traceback has run.
 *
 * sa_strip_traceback: the path enters the slice on its right edge at DP row start_row
 * (pattern_len for the last slice) and is followed with the reference's rules
 * (alignSequenceCPU.cpp:64-114) to the slice's left edge, or to the origin in the first slice.
 * The piece is written right-aligned into d_aligned_* (capacity cap >= text_len + pattern_len);
 * d_result4 = {piece_len, row where the path leaves the slice, text index, pattern index}.
 * The whole alignment is the concatenation of the pieces in slice order.  Both calls are enqueued
 * on `stream` without synchronisation. */
int sa_strip_fill(sa_context *ctx, const sa_scoring *scoring,
                  const uint8_t *d_text, uint64_t text_len, uint64_t col0, uint64_t text_total_len,
                  const uint8_t *d_pattern, uint64_t pattern_len,
                  const int32_t *d_left_col, int32_t *d_right_col, int32_t *d_score, void *stream);
/* The same fill in ROW CHUNKS, so that neighbouring GPUs overlap: as soon as a GPU has filled rows
 * (row0, row0+rows] of its slice it can hand that part of its right-most column over and the next GPU
 * starts on those rows while this one continues below.
 *   sa_strip_begin      plans the slice (strip height chosen for `chunk_rows_hint` rows in flight; 0 = all rows),
 *                       reserves the direction words and returns the chunk height to use in *chunk_rows
 *                       (a multiple of the strip height).
 *   sa_strip_fill_rows  fills rows row0+1 .. row0+rows; row0 must be a multiple of *chunk_rows' strip height and so
 *                       must rows unless the chunk ends at pattern_len.  d_left_col / d_right_col are the WHOLE
 *                       pattern_len+1 columns (the call touches rows row0..row0+rows); d_top_row (text_len int32,
 *                       NULL when row0 == 0) is the d_bottom_row the previous chunk produced; d_bottom_row may be
 *                       NULL for the last chunk.
 * sa_strip_fill is begin + one chunk covering all rows. */
int sa_strip_begin(sa_context *ctx, const sa_scoring *scoring,
                   const uint8_t *d_text, uint64_t text_len, uint64_t col0, uint64_t text_total_len,
                   const uint8_t *d_pattern, uint64_t pattern_len,
                   uint64_t chunk_rows_hint, uint64_t *chunk_rows, void *stream);
int sa_strip_fill_rows(sa_context *ctx, uint64_t row0, uint64_t rows,
                       const int32_t *d_left_col, int32_t *d_right_col,
                       const int32_t *d_top_row, int32_t *d_bottom_row, int32_t *d_score, void *stream);
/* Slices LINKED inside the launch (the pipelined multi-GPU wavefront): every GPU launches its slice at once; a
 * strip of slice k+1 starts as soon as the same strip of slice k has written its right-most column into GPU k+1's
 * border buffer -- plain 8-byte {4H, tag} stores over NVLink into memory mapped with CUDA IPC, polled locally.
 *   sa_peer_alloc   a border buffer in this GPU's memory (>= 8*(pattern_len+1) bytes, zeroed) and its IPC handle
 *   sa_peer_open    maps the RIGHT neighbour's buffer (handle from its sa_peer_alloc) into this process
 *   sa_strip_fill_linked  after sa_strip_begin: the whole slice in one launch.  d_left_col64 = this rank's own border
 *                   buffer (NULL for the first slice), d_right_col64 = the mapped buffer of the right neighbour (NULL
 *                   for the last slice); tag != 0 identifies the call (same on all ranks, different every call).
 *   sa_strip_linked_status  synchronises and reports SA_ERR_LAUNCH if a strip gave up waiting (~10 s) for a
 *                   neighbour that never delivered. */
int sa_peer_alloc(sa_context *ctx, uint64_t bytes, void **dptr, unsigned char handle[64]);
int sa_peer_open(sa_context *ctx, const unsigned char handle[64], void **dptr);
int sa_peer_close(sa_context *ctx, void *dptr);
int sa_peer_free(sa_context *ctx, void *dptr);
int sa_strip_fill_linked(sa_context *ctx, const uint64_t *d_left_col64, uint64_t *d_right_col64, uint32_t tag,
                         int32_t *d_score, void *stream);
int sa_strip_linked_status(sa_context *ctx, void *stream);
int sa_strip_traceback(sa_context *ctx, uint64_t start_row,
                       char *d_aligned_text, char *d_aligned_pattern, uint64_t cap,
                       uint64_t *d_result4, void *stream);

/* ---- front end: files -> Request buffers (host code; replaces utilities.cpp:31-129 for callers of this library).
 * sa_validate_and_transform restates validateAndTransform (utilities.cpp:31-63) in place: FASTA header lines are
 * skipped, lower case is folded, everything outside A-Z is dropped, the residues become alphabet indices; returns
 * their number, or 0 when a letter is not in the alphabet (*bad_letter, may be NULL).
 * sa_read_sequence_file = readSequenceFile (:65-104): *out is malloc'ed (sa_free).  sa_parse_score_matrix_file =
 * parseScoreMatrixFile (:106-129) but a missing file is an error instead of a silent success.
 * sa_read_fasta_batch (new): every record of a multi-FASTA file becomes one sequence of a CSR batch. */
int64_t sa_validate_and_transform(char *buf, uint64_t len, const char *alphabet, int alphabet_size, char *bad_letter);
int sa_read_sequence_file(const char *path, const char *alphabet, int alphabet_size, uint8_t **out, uint64_t *n,
                          char *bad_letter);
int sa_parse_score_matrix_file(const char *path, int alphabet_size, int32_t *matrix);
int sa_read_fasta_batch(const char *path, const char *alphabet, int alphabet_size, uint8_t **residues,
                        int64_t **offsets, uint64_t *n_records, char *bad_letter);
void sa_free(void *p);

/* Identity / gap counts of the alignment the last sa_align call on this context produced, counted
 * on the device during string emission (prettyAlignmentPrint's "# Identity" and "# Gaps", utilities.cpp:262-283):
 * identity = columns with the same letter in both strings, gaps = columns with a gap character in either. */
typedef struct { uint64_t identity; uint64_t gaps; } sa_stats;
int sa_last_stats(sa_context *ctx, sa_stats *out);

/* The report of prettyAlignmentPrint (utilities.cpp:253-315), byte for byte: returns the size of the full report and
 * writes at most cap bytes of it into out; *identity / *gaps (may be NULL) receive the two counts. */
uint64_t sa_pretty_print(const char *aligned_text, const char *aligned_pattern, uint64_t len, uint64_t start_text,
                         uint64_t start_pattern, int32_t score, char *out, uint64_t cap, uint64_t *identity,
                         uint64_t *gaps);

/* ---- batch of independent pairs (new surface; the reference's "batch" is a
 * loop of single calls, tests/benchmarks.cu:318-322) --------------------------
 * CSR layout: pair p's text is text[text_off[p] .. text_off[p+1]) and likewise
 * for pattern.  Aligned strings of pair p are written to
 *   aligned_text   [aln_off[p] .. aln_off[p] + results[p].aln_len)
 * where aln_off is an OUTPUT array of n_pairs entries and the two output
 * arenas must hold text_off[n_pairs] + pattern_off[n_pairs] bytes each.
 * Where a pair's strings land inside the arenas is the library's choice --
 * read aln_off: big host batches come back packed (pair p+1 right behind pair
 * p; only the used bytes cross PCIe), other calls leave every pair in its own
 * slot of text_len + pattern_len bytes. */
typedef struct {
    uint64_t       n_pairs;
    const uint8_t *text;         const int64_t *text_off;     /* n_pairs+1 */
    const uint8_t *pattern;      const int64_t *pattern_off;  /* n_pairs+1 */
} sa_batch;

typedef struct {
    sa_result *results;          /* n_pairs */
    uint64_t  *aln_off;          /* n_pairs, offsets into the two arenas */
    char      *aligned_text;     /* arena */
    char      *aligned_pattern;  /* arena */
    uint64_t   arena_capacity;   /* bytes available in each arena */
    uint32_t  *stats;            /* optional (NULL = not wanted): 2 per pair, {identity, gaps} -- the counts of
                                  * prettyAlignmentPrint (utilities.cpp:262-283), taken on the device while the
                                  * strings are emitted */
} sa_batch_out;

/* HOST buffers; copies in, aligns on the device, copies results out, pipelined
 * in chunks over the context's streams. */
int sa_align_batch(sa_context *ctx, const sa_scoring *scoring,
                   const sa_batch *batch, sa_batch_out *out);

/* DEVICE-resident variant: every pointer in `batch` and `out` is a device
 * pointer (e.g. torch tensors' data_ptr()), work is enqueued on `stream`
 * (a cudaStream_t used as given; NULL is the CUDA default stream) and the call
 * returns without synchronising.  max_text_len / max_pattern_len bound the pair sizes (they
 * size the direction workspace).  A pair with an empty side, a text longer than max_text_len or a
 * pattern longer than max_pattern_len is not aligned: its result is the sentinel
 * {score INT32_MIN, aln_len 0, starts 0} and aln_off 0. */
int sa_align_batch_device(sa_context *ctx, const sa_scoring *scoring,
                          const sa_batch *batch, sa_batch_out *out,
                          uint32_t max_text_len, uint32_t max_pattern_len,
                          void *stream);

/* ---- the multi-GPU dispatcher behind the batch entry (one process, one host thread per device) ----
 * sa_align_batch over several GPUs of this host: the batch is cut into contiguous cell-balanced ranges
 * (sa_partition_batch), every device aligns its range from / into the caller's HOST buffers through its own context
 * and streams, no collective.  The reference hard-codes device 0 (alignSequenceGPU.cu:476).
 *   options == NULL or n_devices == 0: device 0 only.  Contexts are created on first use and cached per device.
 * Results are those of sa_align_batch on one device; where a pair's strings land is again reported in aln_off
 * (every device packs its range behind its own base offset).  Thread-safe (calls are serialised). */
typedef struct {
    int32_t n_devices;            /* 1..8 */
    int32_t devices[8];           /* CUDA device ordinals */
} sa_options;
int sa_align_batch_multi(const sa_options *options, const sa_scoring *scoring,
                         const sa_batch *batch, sa_batch_out *out);
/* the devices SA_DEVICES names ("0,2,3", or a count "4" = devices 0..3); device 0 when it is unset */
int sa_options_from_env(sa_options *options);
/* per-device timing of the last sa_align_batch_multi call: device k of the option list */
int sa_multi_last_timing(int k, sa_timing *out);

/* ---- multi-GPU helpers ------------------------------------------------------
 * Deterministic cell-balanced split of a batch over `world` ranks (each rank
 * aligns pairs [first[r], first[r+1]) ): no data-path collective is needed. */
int sa_partition_batch(const int64_t *text_off, const int64_t *pattern_off,
                       uint64_t n_pairs, int world, uint64_t *first /* world+1 */);

/* ---- page-locked host buffers ----------------------------------------------
 * The reference pins its own direction matrix per call (cudaMallocHost of rows*cols bytes,
 * alignSequenceGPU.cu:541-546); here the caller's input and output buffers are what crosses PCIe, and the copies
 * of sa_align / sa_align_batch run at link speed only from page-locked memory (measured: 300 000 pairs in 59 ms from
 * pageable numpy arrays against 11 ms from pinned ones).  Allocate the buffers of a call with sa_host_alloc, or pin
 * existing ones once with sa_host_register and reuse them.  sa_host_alloc returns NULL on failure. */
void *sa_host_alloc(uint64_t bytes);
void  sa_host_free(void *p);
int   sa_host_register(void *p, uint64_t bytes);
int   sa_host_unregister(void *p);

/* Library / build info. */
const char *sa_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SA_B200_H */

/* ntgpu.h — C ABI of libntgpu, the B200-native (sm_100a) implementation of needletail's
 * per-record hot path.  This header IS the drop-in boundary: the reference (a Rust crate,
 * /root/reference, v0.7.3) exports no FFI of its own (Cargo.toml:17-19 builds a cdylib only
 * for PyO3), so every entry point below names the Rust item it replaces (file:line) and is
 * shaped so a `cc`+`bindgen` shim can re-expose the reference's public surface
 * (src/lib.rs:56-57: parse_fastx_file / parse_fastx_reader / parse_fastx_stdin / FastxReader /
 * Sequence) on top of it — see INTEGRATION.md for the binding a maintainer would add.
 *
 * Conventions
 *  - plain C types only; no CUDA / torch types cross the boundary (streams and device
 *    pointers travel as void* / uint64_t).
 *  - every function returns an ntg_status (0 = ok).  Parse errors use the reference's
 *    ParseErrorKind numbering (src/errors.rs:28-43); misuse is NTG_EINVAL, never UB.
 *  - one ntg_ctx = one CUDA device + its streams; not thread-safe per context (same contract
 *    as the reference's `&mut self` reader, src/parser/utils.rs:119-130), movable across threads.
 *  - batch layout: `seqs` is the concatenation of n sequences, `offs` has n+1 entries
 *    (offs[i]..offs[i+1] is sequence i).  Per-item outputs are likewise CSR: item_offs[n+1].
 *  - result objects (ntg_records, ntg_items, ...) own pinned host arrays that stay valid until
 *    the matching *_free call ("valid until next next()" in the reference).
 *  - there is NO CPU fallback: if no CUDA device is present ntg_create fails with NTG_ECUDA.
 */
#ifndef NTGPU_H
#define NTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTG_ABI_VERSION 3

typedef enum ntg_status {
    NTG_OK = 0,
    /* == needletail::errors::ParseErrorKind (src/errors.rs:28-43) */
    NTG_EIO = 1,
    NTG_EUNKNOWN_FORMAT = 2,
    NTG_EINVALID_START = 3,
    NTG_EINVALID_SEPARATOR = 4,
    NTG_EUNEQUAL_LENGTHS = 5,
    NTG_EUNEXPECTED_END = 6,
    NTG_EEMPTY_FILE = 7,
    /* library-level */
    NTG_EINVAL = 16,   /* bad argument (k == 0, k > 64, bit path with k > 32, m > k, null ptr ...) */
    NTG_ECUDA = 17,    /* CUDA runtime failure; text in ntg_last_error() */
    NTG_ENCCL = 18,    /* NCCL failure / NCCL not loadable */
    NTG_ENOMEM = 19,
    NTG_EUNSUPPORTED = 20
} ntg_status;

typedef enum ntg_format { NTG_FMT_NONE = 0, NTG_FMT_FASTA = 1, NTG_FMT_FASTQ = 2 } ntg_format;          /* parser/utils.rs:75-88 */
typedef enum ntg_line_ending { NTG_LE_NONE = 0, NTG_LE_UNIX = 1, NTG_LE_WINDOWS = 2 } ntg_line_ending;  /* parser/utils.rs:91-117 */

typedef struct ntg_ctx ntg_ctx;

/* ---- context --------------------------------------------------------------------------- */
int ntg_abi_version(void);
int ntg_device_count(int* count);
int ntg_create(int device, ntg_ctx** out);
void ntg_destroy(ntg_ctx* ctx);
const char* ntg_last_error(const ntg_ctx* ctx);       /* never NULL */
int ntg_device_info(ntg_ctx* ctx, int* sm_count, size_t* total_mem, int* cc_major, int* cc_minor);
int ntg_sync(ntg_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t ntg_launch_count(const ntg_ctx* ctx);

/* The record scanner and the Sequence batch calls keep their device scratch and the pinned buffers of freed result tables between
   calls (grow-only, sized by the largest window / batch seen so far: cudaMalloc / cudaMallocHost per call cost more than the work).
   ntg_release_scratch hands all of it back to the driver; the next call allocates again.  ntg_destroy releases it too. */
int ntg_release_scratch(ntg_ctx* ctx);

/* pinned host memory for streaming feeds (cudaMemcpyAsync needs it to overlap) */
int ntg_alloc_pinned(size_t bytes, void** out);
int ntg_free_pinned(void* p);
/* device buffers for callers that keep inputs resident in HBM */
int ntg_device_alloc(ntg_ctx* ctx, size_t bytes, uint64_t* dptr);
int ntg_device_free(ntg_ctx* ctx, uint64_t dptr);
int ntg_memcpy_h2d(ntg_ctx* ctx, uint64_t dptr, const void* host, size_t bytes);
int ntg_memcpy_d2h(ntg_ctx* ctx, void* host, uint64_t dptr, size_t bytes);

/* CUDA-event timing on the context's compute stream (the stream every kernel below is
 * launched on).  slot in [0,64). */
int ntg_event_record(ntg_ctx* ctx, int slot);
int ntg_event_elapsed_ms(ntg_ctx* ctx, int slot_start, int slot_stop, float* ms);

/* ---- (1) FASTX record scanner ----------------------------------------------------------
 * replaces: parse_fastx_reader / get_fastx_reader (src/parser/mod.rs:85-150,37-46; format sniff,
 * decompression stays on the host side of the boundary), fastq::Reader::{find,validate,
 * check_end,next} (src/parser/fastq.rs:155-187,240-285,337-356,388-449), fasta::Reader::{_find,
 * find,next} (src/parser/fasta.rs:220-243,200-216,291-367) and the SequenceRecord accessors
 * id/raw_seq/qual/all/num_bases/position (src/parser/record.rs:57-154).
 * Input: the whole (decompressed) byte stream.  Output: one row per record delivered before the
 * first error, exactly the records the reference's `next()` loop yields. */
typedef struct ntg_record {
    uint64_t start;            /* offset of '@' / '>' == position().byte()            */
    uint64_t id_b, id_e;       /* id()      = bytes[id_b..id_e)                        */
    uint64_t seq_b, seq_e;     /* raw_seq() = bytes[seq_b..seq_e)                      */
    uint64_t qual_b, qual_e;   /* qual()    = bytes[qual_b..qual_e)   (0,0 for FASTA)  */
    uint64_t all_e;            /* all()     = bytes[start..all_e)                      */
    uint64_t num_bases;        /* num_bases()                                          */
    uint64_t line;             /* start_line_number() == position().line()             */
} ntg_record;

typedef struct ntg_parse_error {   /* needletail::errors::ParseError (src/errors.rs:46-56) */
    int32_t kind;              /* ntg_status in 1..7, or 0                              */
    int32_t format;            /* ntg_format                                            */
    uint64_t line;             /* ErrorPosition.line                                    */
    uint64_t record_index;     /* index of the record that failed                       */
    int32_t has_id;            /* ErrorPosition.id is Some                              */
    char id[236];              /* first space-delimited token of the id, NUL-terminated */
} ntg_parse_error;

typedef struct ntg_records {
    int32_t format;            /* ntg_format                                            */
    int32_t line_ending;       /* FastxReader::line_ending() after the first record     */
    uint64_t n_records;
    const ntg_record* records; /* pinned host, n_records rows                           */
    ntg_parse_error error;     /* kind == 0 when the stream ended cleanly               */
    uint64_t final_line, final_byte;  /* FastxReader::position() after the last next()  */
    void* _priv;
} ntg_records;

int ntg_parse_fastx(ntg_ctx* ctx, const uint8_t* bytes, size_t n, ntg_records** out);      /* n < 4 GiB: one window */
void ntg_records_free(ntg_records* r);
/* The incremental form behind `FastxReader::next()` over an `R: Read` of any length (src/parser/mod.rs:85-87, refill loop
 * src/parser/fastq.rs:312-384, src/parser/fasta.rs:291-346): one WINDOW of the stream per call (n < 4 GiB).
 *   format  : NTG_FMT_NONE on the first window (sniff), afterwards the format the first window reported
 *   at_eof  : 0 = the stream continues behind this window: only records complete inside it are delivered (FASTQ: four
 *             newlines; FASTA: followed by another '>' line start) and no end-of-stream rule runs; 1 = last window
 *   consumed: offset of the first byte not covered by a delivered record — the next window starts there (0 with no
 *             error: not even one complete record, pass a larger window)
 * Offsets, lines and record indices in *out are relative to the window; out->final_line is the (relative, 1-based) line at
 * `consumed`, so the caller carries  line_base += final_line - 1,  record_base += n_records,  byte_base += consumed. */
int ntg_parse_fastx_chunk(ntg_ctx* ctx, const uint8_t* bytes, size_t n, int format, int at_eof, ntg_records** out, uint64_t* consumed);

/* Record writers for filtered output: SequenceRecord::write / write_fasta / write_fastq (src/parser/record.rs:158-247) over a
 * record table.  The text of every record i with keep[i] != 0 (keep == NULL: all), in table order, with the given line ending
 * (NTG_LE_UNIX / NTG_LE_WINDOWS: the `forced_line_ending` of the reference): '>' id EOL raw_seq EOL, or '@' id EOL raw_seq EOL
 * '+' EOL qual EOL.  `bytes` is the buffer the table indexes.  *out_len = bytes needed (also when out_cap is too small:
 * NTG_EINVAL then, nothing written). */
int ntg_write_records(ntg_ctx* ctx, const uint8_t* bytes, size_t n, int format, const ntg_record* records, size_t n_records,
                      const uint8_t* keep, int line_ending, uint8_t* out, size_t out_cap, size_t* out_len);

/* ---- (2) Sequence trait, batch form -----------------------------------------------------
 * Every call works on a batch of sequences so that one FFI crossing amortises over many
 * records (a per-k-mer FFI call would dominate).  */

/* sequence::normalize / Sequence::normalize (src/sequence.rs:19-62,226-232).
 * out must hold offs[n] bytes; out_offs n+1 entries; changed[i] = 0 means the reference returns
 * None / Cow::Borrowed for sequence i (out then holds a copy of the input). */
int ntg_normalize(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, int allow_iupac,
                  uint8_t* out, uint64_t* out_offs, uint8_t* changed);
/* Sequence::strip_returns (src/sequence.rs:165-191) */
int ntg_strip_returns(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n,
                      uint8_t* out, uint64_t* out_offs, uint8_t* changed);
/* sequence::complement + Sequence::reverse_complement (src/sequence.rs:67-105,202-208);
 * same offsets in and out */
int ntg_reverse_complement(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, uint8_t* out);
/* QualitySequence::quality_mask (src/sequence.rs:280-297).  qual_offs = the offsets of the quals batch: NTG_EINVAL unless they
 * equal offs (the reference only masks records whose lengths were validated as equal); NULL = the caller vouches for that. */
int ntg_quality_mask(ntg_ctx* ctx, const uint8_t* seqs, const uint8_t* quals, const uint64_t* offs, const uint64_t* qual_offs,
                     size_t n, uint8_t score, uint8_t* out);

typedef struct ntg_items {
    uint64_t n_seqs;
    uint64_t n_items;
    const uint64_t* item_offs; /* n_seqs+1 : items of sequence i are [item_offs[i], item_offs[i+1]) */
    const uint32_t* pos;       /* Item.0 : position of the k-mer within its sequence               */
    const uint8_t* was_rc;     /* Item.2 (canonical_kmers / bit_kmers) ; NULL for minimizers       */
    const uint64_t* val_lo;    /* 2-bit pack (first base most significant), low 64 bits            */
    const uint64_t* val_hi;    /* bits 64..127 (k > 32, canonical_kmers only) else NULL            */
    void* _priv;
} ntg_items;
void ntg_items_free(ntg_items* it);

/* Sequence::canonical_kmers / kmer::CanonicalKmers (src/sequence.rs:237-239, src/kmer.rs:73-129),
 * is_good_base (src/kmer.rs:6-8).  rc may be NULL (then it is the reverse complement of each
 * sequence, as at every reference call site) or a batch with the same offsets.  1 <= k <= 64.
 * The chosen slice is seqs[pos..pos+k) when !was_rc, else rc[len-pos-k..len-pos); val_* is its
 * 2-bit pack (bases compared as raw bytes exactly like the reference; ties => was_rc = 1). */
int ntg_canonical_kmers(ntg_ctx* ctx, const uint8_t* seqs, const uint8_t* rc, const uint64_t* offs, size_t n,
                        uint32_t k, ntg_items** out);
/* Sequence::kmers / kmer::Kmers (src/sequence.rs:245-247, src/kmer.rs:13-41): every window of k bytes, whatever the bytes are.
 * The item is the input slice seqs[pos..pos+k) itself: only item_offs and pos are filled (was_rc, val_lo, val_hi are NULL). */
int ntg_kmers(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, uint32_t k, ntg_items** out);
/* Sequence::bit_kmers / bitkmer::BitNuclKmer (+ canonical) (src/sequence.rs:250-252,
 * src/bitkmer.rs:26-143).  1 <= k <= 32.  ties => (kmer, false). */
int ntg_bit_kmers(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, uint32_t k, int canonical,
                  ntg_items** out);
/* for each item (pos, kmer, _) of bit_kmers(k, false): (pos, bitkmer::minimizer(kmer, m).0)
 * (src/bitkmer.rs:146-162, RC taken at width k).  1 <= m <= k <= 32. */
int ntg_bit_minimizers(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, uint32_t k, uint32_t m,
                       ntg_items** out);
/* element-wise BitKmer helpers on arrays of u64 (src/bitkmer.rs:112-162) */
int ntg_bitkmer_reverse_complement(ntg_ctx* ctx, const uint64_t* in, size_t n, uint32_t k, uint64_t* out);
int ntg_bitkmer_canonical(ntg_ctx* ctx, const uint64_t* in, size_t n, uint32_t k, uint64_t* out, uint8_t* was_rc);
int ntg_bitkmer_minimizer(ntg_ctx* ctx, const uint64_t* in, size_t n, uint32_t k, uint32_t m, uint64_t* out);

/* ---- (3) the fused hot path: scan + normalize + canonical k-mers + minimizers -> tallies ---
 * One pass over the FASTX bytes computes what the reference's README / bench loop computes
 * (src/lib.rs:16-36, benches/benchmark.rs:32-41,55-63), per record:
 *     norm = rec.normalize(iupac); rc = norm.reverse_complement();
 *     for (pos, kmer, was_rc) in norm.canonical_kmers(k, &rc) { tallies }
 *     for (pos, bk, _) in norm.bit_kmers(k, false) { bitkmer::minimizer(bk, m) }   // m > 0, k <= 32
 * Records after the first parse error are not tallied (iterator semantics); the error is
 * reported in *err and the call still returns NTG_OK (the status only reports library failures). */
typedef struct ntg_tally_config {
    uint32_t k;                /* 1..64                                                  */
    uint32_t m;                /* 0 = no minimizers; else 1..k and k <= 32               */
    uint32_t allow_iupac;      /* normalize(iupac) flag (does not change the tallies)    */
    uint32_t has_query;        /* count canonical k-mers equal to `query` (lib.rs:31-35) */
    uint8_t query[64];         /* k ASCII bases ACGT                                     */
    uint32_t flags;            /* NTG_TALLY_* bits, 0 for normal use                     */
    uint32_t qmask_score;      /* != 0 (FASTQ): QualitySequence::quality_mask(score) (src/sequence.rs:280-297) is applied to every
                                  record before the loop above: a base whose quality byte is < score counts as 'N'.  Fused into
                                  the record-owned short-read kernel (the quality line is resident beside its sequence line: no
                                  extra DRAM traffic); other inputs are masked in a device copy first.  Not for stream sessions. */
} ntg_tally_config;
/* diagnostic: FASTQ tiles wait for the look-back instead of starting on the locally inferred line phase (same results) */
#define NTG_TALLY_NO_SPECULATION 1u
/* enqueue/collect form only, after ntg_comm_init: the tallies of all ranks are summed by ONE ncclAllReduce enqueued on the
 * compute stream right behind the kernel (k_finalize writes the NCCL send buffer; no host round trip in between).  collect
 * then returns the job-wide tallies on every rank.  If any rank's shard needed the host (parse error replay, exact path)
 * collect returns that rank's LOCAL tallies with NTG_RESERVED_NOT_REDUCED set in reserved[0]: reduce them with
 * ntg_comm_allreduce_tallies. */
#define NTG_TALLY_ALLREDUCE 2u
/* diagnostic: short-read FASTQ goes through the general tile kernel (fused::k_fused) instead of the record-owned fast path */
#define NTG_TALLY_NO_FASTPATH 4u
#define NTG_RESERVED_NOT_REDUCED (1ull << 32)
/* reserved[1] bit: the record-owned short-read FASTQ kernel (fastq_warp.cuh) produced the tallies (whole-buffer entry points) */
#define NTG_RESERVED_FAST_PATH (1ull << 32)

typedef struct ntg_tallies {
    uint64_t n_records;
    uint64_t n_bases;          /* sum of num_bases()                                     */
    uint64_t n_kmers;          /* canonical_kmers items                                  */
    uint64_t n_not_rc;         /* items with was_rc == false (benchmark.rs:37-39)        */
    uint64_t kmer_sum_lo;      /* wrapping sum of val_lo of every canonical k-mer        */
    uint64_t kmer_sum_hi;      /* wrapping sum of val_hi (k > 32)                        */
    uint64_t n_query;
    uint64_t n_minimizers;     /* bit_kmers(k,false) items                               */
    uint64_t minimizer_sum;    /* wrapping sum of bitkmer::minimizer(kmer, m).0          */
    uint64_t reserved[7];      /* reserved[0]: 0 = single-pass fused kernel produced these; else bitmask of why the exact
                                  record-table path re-ran (1 parse error, 2 newline-dense tile, 4 whitespace run > halo);
                                  reserved[1]: non-zero when a speculated FASTQ line phase was wrong and the call re-ran without speculation */
} ntg_tallies;

/* host bytes of ANY size: streamed through three 64 MiB device segments, H2D copies overlapped with the kernel of the
 * previous segment (the end-to-end path; pinned caller memory copies at PCIe speed).  Device memory use is bounded. */
int ntg_tally_fastx(ntg_ctx* ctx, const uint8_t* bytes, size_t n, const ntg_tally_config* cfg,
                    ntg_tallies* out, ntg_parse_error* err);
/* bytes already resident in HBM (dptr 16-byte aligned) */
int ntg_tally_fastx_device(ntg_ctx* ctx, uint64_t dptr, size_t n, const ntg_tally_config* cfg,
                           ntg_tallies* out, ntg_parse_error* err);
/* asynchronous form for benchmarking: enqueue only; collect synchronises and finalises.
 * `fused_kernel_ms` (may be NULL) receives the CUDA-event duration of the fused kernel alone. */
int ntg_tally_fastx_device_enqueue(ntg_ctx* ctx, uint64_t dptr, size_t n, const ntg_tally_config* cfg);
int ntg_tally_fastx_device_collect(ntg_ctx* ctx, ntg_tallies* out, ntg_parse_error* err, float* fused_kernel_ms);

/* ---- (3b) streaming session: the tally path over an `R: Read` ---------------------------------
 * replaces: parse_fastx_reader<R: Read + Send>(reader) (src/parser/mod.rs:85-150) + the per-record loop, for streams of
 * unknown length (stdin, sockets, decompressors).  Pieces of any size are staged in pinned host buffers and go to the device
 * as 64 MiB segments while the next one fills; the look-back state carries across kernel launches.  Iterator semantics are
 * kept without re-reading the stream: a parse error ends the stream in front of the failing record (the launch that met it is
 * replayed truncated while its segment is still resident) and is reported by finish with the reference's kind / line / id.
 * Inputs that need the exact record-table path (newline-dense tiles, whitespace runs > 128 B inside sequences, records
 * longer than a segment that fail) make finish return NTG_EUNSUPPORTED: use ntg_tally_fastx on the whole input.
 * One session per context at a time; not thread-safe (the `&mut self` contract of FastxReader, parser/utils.rs:119-130). */
typedef struct ntg_stream ntg_stream;
int ntg_stream_open(ntg_ctx* ctx, const ntg_tally_config* cfg, ntg_stream** out);
int ntg_stream_feed(ntg_stream* s, const uint8_t* bytes, size_t n);          /* copies into the staging buffer */
/* zero-copy producers (read(2), inflate): write up to *avail bytes at *ptr (pinned), then commit what was written */
int ntg_stream_acquire(ntg_stream* s, uint8_t** ptr, size_t* avail);
int ntg_stream_commit(ntg_stream* s, size_t n);
/* one more piece of a gzip stream (magic 1f 8b; src/parser/mod.rs:96-108): multi-member like flate2::MultiGzDecoder.
 * threads > 1 and a BGZF file (member sizes in the gzip extra field): members are inflated in parallel, in place. */
int ntg_stream_feed_gz(ntg_stream* s, const uint8_t* gz, size_t n, int threads);
/* threads == NTG_GZ_DEVICE (0) on a session that has been fed nothing else, and a BGZF file: the members are inflated ON THE
 * DEVICE (one thread per member, RFC 1951 decoder in needletail_b200/csrc/inflate.cuh).  Only the compressed bytes cross PCIe;
 * the text is written straight into the device segment the fused kernel reads and never exists on the host.  Lengths are
 * checked against ISIZE, CRC-32 is not recomputed.  A non-BGZF gzip stream falls back to the sequential host inflate. */
#define NTG_GZ_DEVICE 0
/* the device-side DEFLATE decoder on its own: a whole BGZF blob, host to host.  *out_len = decompressed size (also when
 * out_cap is too small: NTG_EINVAL then).  Corrupt members: NTG_EIO. */
int ntg_inflate_bgzf(ntg_ctx* ctx, const uint8_t* gz, size_t n, uint8_t* out, size_t out_cap, size_t* out_len);
int ntg_stream_finish(ntg_stream* s, ntg_tallies* out, ntg_parse_error* err);
uint64_t ntg_stream_bytes(const ntg_stream* s);                              /* decompressed bytes fed so far */
void ntg_stream_close(ntg_stream* s);
/* parse_fastx_file(path) (src/parser/mod.rs:160-165) for the tally path: reads the file piecewise into the staging buffers,
 * inflating gzip on the way (`threads` workers for BGZF).  bzip2 / xz / zstd files: NTG_EUNSUPPORTED (zlib only). */
int ntg_tally_fastx_file(ntg_ctx* ctx, const char* path, const ntg_tally_config* cfg, int threads,
                         ntg_tallies* out, ntg_parse_error* err);

/* ---- (3c) k-mer spectrum: count per distinct canonical k-mer -------------------------------------
 * The consumer of `canonical_kmers` (the README loop counts ONE k-mer: src/lib.rs:31-35).  Definition: the multiset
 * of items of `rec.normalize(false).canonical_kmers(k, &rc)` over every record delivered before the first parse error, keyed by
 * the 2-bit pack of the canonical k-mer (== bit_kmers(k, true)).  k <= 14: dense histogram of 4^k u32 counters (`capacity`
 * ignored); 15 <= k <= 32: open-addressing hash table of `capacity` slots (rounded up to a power of two; NTG_ENOMEM when full).
 * Counts saturate nowhere: they wrap at 2^32. */
typedef struct ntg_spectrum ntg_spectrum;
int ntg_spectrum_create(ntg_ctx* ctx, uint32_t k, uint64_t capacity, ntg_spectrum** out);
void ntg_spectrum_destroy(ntg_spectrum* sp);
int ntg_spectrum_clear(ntg_spectrum* sp);
/* add every canonical k-mer of a FASTX input (host bytes of any size / bytes resident in HBM, 16-byte aligned).  tallies (may be
 * NULL) receives the counting pass's tallies (n_kmers == k-mers added).  A parse error ends the input in front of the failing
 * record (err->kind != 0; use ntg_tally_fastx for its exact kind / line). */
int ntg_spectrum_add_fastx(ntg_spectrum* sp, const uint8_t* bytes, size_t n, ntg_tallies* tallies, ntg_parse_error* err);
int ntg_spectrum_add_fastx_device(ntg_spectrum* sp, uint64_t dptr, size_t n, ntg_tallies* tallies, ntg_parse_error* err);
/* count of one k-mer given as k ASCII bases ACGT (its canonical form is looked up) — the README example */
int ntg_spectrum_count(ntg_spectrum* sp, const uint8_t* kmer, uint64_t* count);
/* the distinct k-mers and their counts, in arbitrary order; *n_distinct is the full number also when cap is smaller */
int ntg_spectrum_export(ntg_spectrum* sp, uint64_t* keys, uint32_t* counts, uint64_t cap, uint64_t* n_distinct);
/* count-of-counts: hist[c] = number of distinct k-mers seen c times (c >= n_bins - 1 collected in the last bin) */
int ntg_spectrum_histogram(ntg_spectrum* sp, uint64_t* hist, uint32_t n_bins);
/* multi-GPU (after ntg_comm_init, same k / capacity on every rank).  Dense: ONE ncclAllReduce over the whole histogram, every rank
 * ends with the job-wide counts.  Hash: entries travel to their owner rank (hash of the key) with grouped ncclSend / ncclRecv:
 * every rank ends with the job-wide counts of the keys it owns (reduce-scatter by k-mer hash). */
int ntg_spectrum_reduce(ntg_spectrum* sp);
uint64_t ntg_spectrum_kmers(const ntg_spectrum* sp);             /* k-mers added on this rank so far */

/* ---- (4) synthetic inputs (DESIGN.md "Synthetic inputs"; same bytes as oracle/synth.hpp) ---- */
int ntg_synth_fastq_device(ntg_ctx* ctx, uint64_t dptr, uint64_t seed, uint64_t rec0, uint64_t nrec,
                           uint32_t read_len, uint32_t n_thresh);
int ntg_synth_fasta_device(ntg_ctx* ctx, uint64_t dptr, uint64_t seed, uint64_t rec0, uint64_t nrec,
                           uint32_t read_len, uint32_t n_thresh);

/* ---- (5) multi-GPU: one context per rank, tallies reduced with one ncclAllReduce ---------- */
#define NTG_NCCL_ID_BYTES 128
int ntg_comm_unique_id(uint8_t id[NTG_NCCL_ID_BYTES]);                 /* rank 0, then broadcast by the launcher */
int ntg_comm_init(ntg_ctx* ctx, int n_ranks, int rank, const uint8_t id[NTG_NCCL_ID_BYTES]);
int ntg_comm_allreduce_tallies(ntg_ctx* ctx, ntg_tallies* inout);      /* ncclUint64 / ncclSum over NVLink */
int ntg_comm_destroy(ntg_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* NTGPU_H */

/*
 * tobias_b200.h -- C-ABI of libtobias_b200.so: the B200 (sm_100a) implementation of the TOBIAS per-base
 * signal hot path.  Plain C types only; the caller owns every host pointer; all device memory is
 * owned by the context; a context is bound to one GPU and must be used from one thread at a time
 * (one process per GPU).  Every entry point returns TB_OK (0) or a negative error code; the message is
 * available from tb_last_error().  No exceptions cross this boundary.  There is no CPU fallback: if no
 * CUDA device is usable tb_init fails.
 *
 * The reference (loosolab/TOBIAS 0.17.3) has no FFI of its own for this path: its boundary is the
 * Cython / worker-function layer.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes binding.
 *
 * Coordinates: contig-local 0-based half-open [start, end) as in BED / pysam, int64.
 * Nucleotide codes follow the reference: A=0 T=1 C=2 G=3 (complement = code ^ 1), anything else = N
 * (tobias/utils/sequences.pyx:398-420).
 */
#ifndef TOBIAS_B200_H
#define TOBIAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tb_ctx tb_ctx;

enum {
    TB_OK = 0,
    TB_ERR_CUDA = -1,      /* a CUDA runtime call or kernel failed */
    TB_ERR_ARG = -2,       /* invalid argument */
    TB_ERR_STATE = -3,     /* call order violated (e.g. no genome / reads / model uploaded) */
    TB_ERR_NOMEM = -4,     /* device or host allocation failed */
    TB_ERR_NCCL = -5       /* NCCL call failed / library built without NCCL */
};

/* read flag bits of tb_reads_upload */
enum {
    TB_READ_REVERSE = 1,   /* is_reverse                                     (ngs.pyx:83)          */
    TB_READ_5P_MATCH = 2,  /* 5'-most CIGAR op is M            (atacorrect_functions.py:156-158)  */
    TB_READ_SKIP = 4       /* is_unmapped or is_duplicate: dropped by ReadList.from_bam (ngs.pyx:153) */
};

/* score kinds of tb_score_regions (tobias/tools/score_bigwig.py:67-84) */
enum { TB_SCORE_FOOTPRINT = 0, TB_SCORE_SUM = 1, TB_SCORE_MEAN = 2, TB_SCORE_NONE = 3, TB_SCORE_FOS = 4 };

/* ------------------------------------------------------------------------------------------------ context */
int tb_abi_version(void);
int tb_init(int device, tb_ctx **out);
void tb_destroy(tb_ctx *ctx);
const char *tb_last_error(const tb_ctx *ctx);   /* ctx may be NULL: message of the last failed tb_init */
/* sm_count, bytes free / total on the bound device; 1 if this build is the GPU-less kernel emulator used by CI */
int tb_device_info(tb_ctx *ctx, int *sm_count, int64_t *free_bytes, int64_t *total_bytes, int *is_emulator);
/* elapsed device time (ms, CUDA events on the context's stream) of the most recent *_run / compute call */
int tb_last_kernel_ms(tb_ctx *ctx, float *ms);
/* per-stage device times (ms) of the most recent multi-kernel call; tb_correct_run: { K1 pileup, K3a score, K3b fit,
 * K3c correct }, tb_bias_counts: { scatter, scan+accumulate } (summed over chunks) */
int tb_stage_ms(tb_ctx *ctx, float *out, int n);
/* CUDA-event stopwatch on the context's stream spanning any sequence of calls (device time line, host gaps included) */
int tb_timer_begin(tb_ctx *ctx);
int tb_timer_end(tb_ctx *ctx, float *ms);
/* page-locked host memory for the caller's input / output buffers (full-speed, truly asynchronous copies) */
int tb_host_alloc(tb_ctx *ctx, int64_t bytes, void **out);
int tb_host_free(tb_ctx *ctx, void *p);
/* number of kernel launches issued by this context so far */
int tb_launch_count(tb_ctx *ctx, int64_t *n);
/* Kernel-variant / chunk-size switches for comparison runs and tests (every setting yields the same results; production code never
 * calls this): "k3a_sum_form", "k3a_no_tmem", "k3a_tmem_groups", "k3c_fast", "k2_tile", "k2_site_cap", "k2_chunk", "k2_no_hist",
 * "k2_smem_chunk", "k2_no_smem_hist", "k3c_chunks", "k4_chunks", "k4_form", "k4_threads", "k3b_persistent", "k3b_tile", "k3b_cap", "k3b_threads", "host_trace".  value < 0 restores the default.  (These
 * replace the environment variables the library read in its first version.) */
int tb_set_option(tb_ctx *ctx, const char *key, int64_t value);

/* ------------------------------------------------------------------------------------------------ genome
 * Replaces GenomicSequence.from_fasta (tobias/utils/sequences.pyx:385-428): the genome is kept resident in
 * HBM, 2 bits per base + 1 N-mask bit per base; ASCII -> 2-bit packing runs on the device. */
int tb_genome_create(tb_ctx *ctx, int n_contigs, const int64_t *lengths);
/* upload ASCII bases [offset, offset+n) of one contig; offset must be a multiple of 32 */
int tb_genome_upload(tb_ctx *ctx, int contig, int64_t offset, const uint8_t *ascii, int64_t n);
/* asynchronous form of tb_genome_upload: copies and packing are enqueued on the context's copy stream and the call returns;
 * `ascii` (page-locked: tb_host_alloc) must stay untouched until tb_genome_wait.  Every entry point that reads the genome waits on
 * the device for the contigs it needs, so tb_count_reads / tb_bias_counts on the first contigs overlap with the transfer of the rest */
int tb_genome_upload_async(tb_ctx *ctx, int contig, int64_t offset, const uint8_t *ascii, int64_t n);
/* blocks the host until every asynchronous upload has completed */
/* A contig that is already in the device layout (seq2: (n + 15) / 16 words, base k in bits 2 * (k & 15) of word k >> 4, A0 T1 C2 G3; nmask:
 * (n + 31) / 32 words, bit (k & 31) of word k >> 5 set for N; bits past the end set).  0.375 bytes per base cross the link instead of 1.
 * Host packer: tobias_b200/io/hostpack.c (cached next to the FASTA as <fasta>.tb2bit).  async = 1: like tb_genome_upload_async. */
int tb_genome_upload_packed(tb_ctx *ctx, int contig, const uint32_t *seq2, const uint32_t *nmask, int64_t n_bases, int async);
/* The same with the N mask given as runs [run_start[r], run_end[r]) of contig positions (assembly gaps are few and long: a human
 * genome then crosses the bus as 0.25 bytes per base); the device builds the mask words. */
int tb_genome_upload_packed_runs(tb_ctx *ctx, int contig, const uint32_t *seq2, int64_t n_runs, const int64_t *run_start,
                                 const int64_t *run_end, int64_t n_bases, int async);
int tb_genome_wait(tb_ctx *ctx);
/* debugging / tests: nucleotide codes (0..4) of contig[start, start+n) */
int tb_genome_fetch(tb_ctx *ctx, int contig, int64_t start, int64_t n, uint8_t *codes_out);

/* ------------------------------------------------------------------------------------------------ reads
 * Replaces the list of OneRead objects (tobias/utils/ngs.pyx:55-89,150-160) by five columns, sorted by
 * (contig, ref_start).  clip5 = soft-clipped bases at the read's 5' end (ngs.pyx:98): query_alignment_start
 * for forward reads, query_length - query_alignment_end for reverse reads.  Replaces any earlier read set. */
int tb_reads_upload(tb_ctx *ctx, int64_t n, const int32_t *contig, const int32_t *ref_start,
                    const int32_t *ref_end, const int32_t *clip5, const uint8_t *flags);
/* Compact columns of records grouped by contig in the genome's contig order (what a coordinate-sorted BAM gives): the records of
 * contig c are [rec_off[c], rec_off[c + 1]) (n_contigs + 1 offsets, rec_off[0] = 0, rec_off[n_contigs] = n); reference start int32,
 * reference length (ref_end - ref_start) and 5' clip as uint16 -- 9 bytes per record instead of 17.  Same flag bits. */
int tb_reads_upload_compact(tb_ctx *ctx, int64_t n, const int64_t *rec_off, const int32_t *ref_start, const uint16_t *ref_len,
                            const uint16_t *clip5, const uint8_t *flags);

/* ------------------------------------------------------------------------------------------------ K1
 * count_reads (tobias/tools/atacorrect_functions.py:80-105): number of (read, region) pairs with the read
 * overlapping the region and start < cutsite <= end.  Regions must be sorted by (contig, start). */
int tb_count_reads(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start,
                   const int64_t *end, int shift_fwd, int shift_rev, int64_t *count_out);

/* Per-strand cut-site pileup of each region: ReadList.signal (tobias/utils/ngs.pyx:186-199) after the filter
 * of atacorrect_functions.py:237 (start < cutsite < end, read overlaps the region).  out_fwd / out_rev receive
 * the concatenated per-region arrays (sum of region lengths), uint32.  Regions sorted by (contig, start). */
int tb_pileup(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start,
              const int64_t *end, int shift_fwd, int shift_rev, uint32_t *out_fwd, uint32_t *out_rev);
/* ReadList.signal(region) (tobias/utils/ngs.pyx:186-199) over the resident records: values[cut - start - 1] += 1 for every record with
 * start < cut <= end -- no overlap test and the upper bound inclusive, unlike the filtered pileup above.  Same output layout. */
int tb_read_signal(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start, const int64_t *end,
                   int shift_fwd, int shift_rev, uint32_t *out_fwd, uint32_t *out_rev);

/* ------------------------------------------------------------------------------------------------ K2
 * bias_estimation (tobias/tools/atacorrect_functions.py:109-181) over all input regions at once, i.e. the
 * SequenceMatrix.add_sequence / add_background calls of sequences.pyx:63-105,178-223 with per-(region,
 * strand, cutsite) multiplicity capped at `cap`.  Regions sorted by (contig, start), non-overlapping.
 * is_dwm=1: counts_out is int64 [2 strands][2 (bias, background)][5][5][L][L]  (Sm, Sn, m, n)
 * is_dwm=0: counts_out is int64 [2 strands][2 (bias, background)][5][L]
 * scalars_out: int64 [5] = { no_reads, no_bias_fwd, no_bg_fwd, no_bias_rev, no_bg_rev }.
 * The counts are integers, so sums over shards are exact in any order (tb_allreduce_i64). */
int tb_bias_counts(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start,
                   const int64_t *end, int k_flank, int bg_shift, int shift_fwd, int shift_rev, int cap,
                   int is_dwm, int64_t *counts_out, int64_t *scalars_out);

/* AtacBias.join / SequenceMatrix.add_counts (atacorrect_functions.py:54-60, sequences.pyx:35-42) across
 * the ranks of a one-process-per-GPU job: in-place sum of a host buffer over NCCL.  `nccl_unique_id` is the
 * 128-byte ncclUniqueId created by rank 0 (tb_nccl_unique_id) and distributed by the host launcher. */
int tb_nccl_unique_id(uint8_t id_out[128]);
int tb_nccl_init(tb_ctx *ctx, const uint8_t nccl_unique_id[128], int rank, int world_size);
int tb_allreduce_i64(tb_ctx *ctx, int64_t *buf, int64_t n);
int tb_allreduce_f64(tb_ctx *ctx, double *buf, int64_t n);

/* ------------------------------------------------------------------------------------------------ K3
 * Score tables of one strand (0 = forward, 1 = reverse) as produced by prepare_mat
 * (tobias/utils/sequences.pyx:107-117,226-257), all log2, row order A,T,C,G,N:
 *   is_dwm=1: bias_pwm_log[5][L], bg_pwm_log[5][L], bias_dwm_log[5][5][L][L], bg_dwm_log[5][5][L][L]
 *   is_dwm=0: bias_pwm_log holds the pssm[5][L]; the other three pointers are ignored. */
int tb_set_bias_model(tb_ctx *ctx, int strand, int is_dwm, int L, const double *bias_pwm_log,
                      const double *bg_pwm_log, const double *bias_dwm_log, const double *bg_dwm_log);

/* What tb_set_bias_model stored for a strand: *is_dwm (-1: no model yet), *L, and *product_form = 1 when DWM scoring
 * (sequences.pyx:263-342) runs as products of powers of two of the table rows, 0 when the tables are so wide (a column
 * product could pass 2^480) that the kernel evaluating exp2 / log2 per column is used instead.  Any pointer may be NULL. */
int tb_bias_model_info(tb_ctx *ctx, int strand, int *is_dwm, int *L, int *product_form);

/* SequenceMatrix.score_sequence (sequences.pyx:123-151,263-342) of one strand over contig[start, end),
 * with the reference's conventions: NaN in the first/last L//2 positions and for windows holding N; the
 * reverse strand is the score of the reverse complement, reversed (atacorrect_functions.py:258-260). */
int tb_score_sequence(tb_ctx *ctx, int strand, int contig, int64_t start, int64_t end, double *out);

/* bias_correction (tobias/tools/atacorrect_functions.py:191-410) over all OUTPUT regions [start, end) at once
 * (each is extended internally by k_flank + window/2, which must stay inside its contig).
 * Window fits: one thread per window, rounds over compacted work lists.  lm_exact_order = 1: MINPACK's own
 * arithmetic and summation order -- bit-faithful to scipy's lmdif on identical inputs (PWM mode: tracks bit-identical to the
 * reference).  0: the outer-iteration head uses the five inner products of the Jacobian columns (one pass instead of four):
 * same iteration and decisions, parameters ~1e-8 apart, the scatter a 1e-13 change of the bias input causes in the reference
 * itself (what the DWM bias input carries anyway).  2 / 3 select the first-generation warp-per-window kernels (comparison runs).
 * Device-side only; results stay in HBM until tb_correct_download.  Regions sorted by (contig, start), and
 * non-overlapping before extension. */
int tb_correct_run(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start,
                   const int64_t *end, int k_flank, int window, double correction_factor,
                   int shift_fwd, int shift_rev, int lm_exact_order);
/* Tracks of the last tb_correct_run, trimmed to the output regions and concatenated (sum of region lengths).
 * tracks[t][s]: t = 0 uncorrected, 1 bias, 2 expected, 3 corrected; split_strands=0: s=0 is forward+reverse and
 * s=1 must be NULL; split_strands=1: s=0 forward, s=1 reverse.  NULL pointers are skipped (--track-off).
 * as_f64=0: float32 (what the bigWig writer stores); as_f64=1: float64 (what the reference queues).
 * prepost_out (may be NULL): double [2 strands][3][5][L] = verification PWM counts
 * (atacorrect_functions.py:358-373): [0] pre counts, [1] post counts, [2] post neg_counts. */
int tb_correct_download(tb_ctx *ctx, int split_strands, int as_f64, void *tracks[4][2], double *prepost_out);
/* Asynchronous form: _begin enqueues the packing and the device-to-host copies on a second stream, ordered after the
 * work already submitted; kernels launched afterwards (tb_signal_from_corrected, tb_score_run) overlap with the copies.
 * The host buffers (pinned for real overlap) are complete when _wait returns.  tb_correct_download = _begin + _wait. */
int tb_correct_download_begin(tb_ctx *ctx, int split_strands, int as_f64, void *tracks[4][2], double *prepost_out);
int tb_correct_download_wait(tb_ctx *ctx);
/* Streaming form: names the destination of the NEXT tb_correct_run's tracks (same arguments as tb_correct_download; tracks = NULL
 * withdraws it).  That run then computes its last stage in chunks of output regions and packs / copies every finished chunk to the
 * host on the second stream while the next one computes -- what the reference's writer processes do with the queues the workers fill
 * (atacorrect.py:379-424).  The run returns as before (compute stream drained); the host buffers are complete after
 * tb_correct_download_wait.  A sink serves one run. */
int tb_correct_set_sink(tb_ctx *ctx, int split_strands, int as_f64, void *tracks[4][2], double *prepost_out);
/* per-window fit diagnostics of the last tb_correct_run (tests): double [2 strands][n_windows_total][6] =
 * { mode (0 fit, 1 fallback, 2 empty window), a, b (bias_min for mode 1), max(prediction), MINPACK info, nfev };
 * n may be queried with out=NULL. */
int tb_correct_fit_trace(tb_ctx *ctx, int64_t *n_windows_total, double *out);

/* ------------------------------------------------------------------------------------------------ K4
 * calculate_scores (tobias/tools/score_bigwig.py:33-98) over all regions at once.  `signal` holds the
 * concatenated float32 bigWig values of the regions ALREADY extended by `flank` on both sides (NaN allowed:
 * nan_to_num is applied), region r occupying ext_len[r] values.  Output: concatenated scores trimmed by `flank`
 * (ext_len[r] - 2*flank values per region).  use_min / use_max switch --min-limit / --max-limit.
 * tb_score_regions = upload + run + download; the three pieces are exposed for resident-in-HBM timing. */
typedef struct tb_score_params {
    int kind;                 /* TB_SCORE_* */
    int absolute;             /* --absolute */
    int use_min, use_max;     /* --min-limit / --max-limit given */
    double min_limit, max_limit;
    int window;               /* --window (sum / mean) */
    int smooth;               /* --smooth (<=1: off) */
    int flank_min, flank_max, fp_min, fp_max;
    int flank;                /* region_flank: trimmed from both ends of each region */
    int as_f64;               /* output element type: 0 float32, 1 float64 */
} tb_score_params;

int tb_signal_upload(tb_ctx *ctx, const float *signal, int64_t n_regions, const int64_t *ext_len);
int tb_score_run(tb_ctx *ctx, const tb_score_params *p);
int tb_score_download(tb_ctx *ctx, void *out);
/* Streaming form: destination of the NEXT tb_score_run's output (element type as in its tb_score_params; NULL withdraws it).  The footprint
 * scan then runs in chunks of regions whose scores are copied out on the second stream while the next chunk computes; complete after
 * tb_score_sink_wait (or tb_correct_download_wait: both wait for the second stream). */
int tb_score_set_sink(tb_ctx *ctx, void *out);
int tb_score_sink_wait(tb_ctx *ctx);
/* Scores of the last tb_score_run at `n` positions of its concatenated output space (region offsets = prefix sums of the
 * trimmed region lengths), as doubles, without downloading the track: the reads BINDetect's scan_and_score makes from the
 * footprint bigWig -- the middle of every TFBS and the random background positions of every region
 * (tobias/tools/bindetect_functions.py:296-330).  An index outside the track is an error (TB_ERR_ARG). */
int tb_score_gather(tb_ctx *ctx, int64_t n, const int64_t *index, double *out);
int tb_score_regions(tb_ctx *ctx, const float *signal, int64_t n_regions, const int64_t *ext_len,
                     const tb_score_params *p, void *out);
/* Chain ATACorrect -> ScoreBigwig on the device: use the corrected track of the last tb_correct_run (rounded to
 * float32 as a bigWig would store it, zero outside the output regions) as the signal of the same regions
 * extended by p->flank. */
int tb_signal_from_corrected(tb_ctx *ctx, int flank);
/* The same for ANY scoring regions (sorted by contig and start; each extended by `flank` internally, which must stay inside its contig):
 * every position reads the corrected track where an output region of the last tb_correct_run covers it and 0 elsewhere -- what
 * `TOBIAS ScoreBigwig --signal <prefix>_corrected.bw --regions ...` sees (score_bigwig.py:136-170 merges the peaks but does not split them at
 * 50 kb as ATACorrect does, atacorrect.py:233; pass the merged regions to reproduce it across the split pieces). */
int tb_signal_from_corrected_regions(tb_ctx *ctx, int flank, int64_t n_regions, const int32_t *contig, const int64_t *start,
                                     const int64_t *end);

/* Gathers of the two downstream tools that read the tracks this path writes, over the signal of the last tb_signal_upload
 * (one uploaded region = one BED region / one site window; NaN reads as 0 like OneRegion.get_signal, tobias/utils/regions.py:172):
 * tb_signal_summary: ScoreBed's score functions (tobias/tools/score_bed.py:30-54) per region -- op 0 start, 1 mid (index int(len / 2)),
 *   2 end, 3 min, 4 max, 5 mean, 6 sum; out: double [n_regions].
 * tb_signal_aggregate: PlotAggregate's column means over the rows of a signal matrix (tobias/tools/plot_aggregate.py:236-257): all regions
 *   of one width; reverse[r] != 0 reads row r back to front (minus-strand sites), keep[r] == 0 leaves it out (outlier removal), both may be
 *   NULL; log_transform: sign(v) log2(|v| + 1); out: double [width]. */
int tb_signal_summary(tb_ctx *ctx, int op, double *out);
int tb_signal_aggregate(tb_ctx *ctx, const uint8_t *reverse, const uint8_t *keep, int log_transform, double *out);

/* Array-level mirror of SequenceMatrix.add_sequence / add_background (tobias/utils/sequences.pyx:63-105 PWM, :178-223 DWM) for a batch of
 * explicit k-mers: kmers int64 [n][L] codes (A0 T1 C2 G3 N4), amounts double [n].  The increments are ADDED to the caller's arrays: counts
 * [5][L] (PWM) or [5][5][L][L] (DWM); neg_counts [5][L] receives the non-positive amounts of a PWM add_sequence (may be NULL otherwise);
 * *total += the amounts that were added (no_bias / no_bg).  A k-mer holding N is skipped, except by the PWM add_background (:100-102). */
int tb_add_kmers(tb_ctx *ctx, int L, int is_dwm, int is_background, int64_t n, const int64_t *kmers, const double *amounts,
                 double *counts, double *neg_counts, double *total);

/* fast_rolling_math(arr, w, operation) (tobias/utils/signals.pyx:102-177) on one float64 array.  is_mean: 0 "sum", 1 "mean" (NaN
 * flanks, C-float accumulator), 2 "max", 3 "min" (C-float extremum, every position defined, index 0 never inspected -- the reference's
 * `start_i + j > 0`), 4 "prod". */
int tb_rolling(tb_ctx *ctx, const double *arr, int64_t n, int w, int is_mean, double *out);

#ifdef __cplusplus
}
#endif
#endif /* TOBIAS_B200_H */

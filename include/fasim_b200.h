/*
 * fasim_b200.h — C ABI of the B200-native Fasim-LongTarget hot path (libfasim_b200.so).
 *
 * Drop-in boundary.  The reference has no plugin API; its seams are the process
 * surface (`fasim -f1 .. -f2 .. -O ..`, the `-TFOsorted` file) and a handful of internal C/C++
 * entry points.  Every function below names the reference interface it replaces.  Plain C types
 * only; no exceptions cross the boundary; every call returns LTG_OK (0) or a negative error code
 * and `ltg_last_error()` gives the message.  A context owns one GPU and is not thread-safe; use
 * one context per GPU (and per host thread).  There is no CPU fallback: if no CUDA device / kernel
 * image is usable the calls fail with LTG_ERR_CUDA.
 */
#ifndef FASIM_B200_H
#define FASIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTG_OK 0
#define LTG_ERR_ARG (-1)      /* bad argument / bad parameter value (reference: exit(1) in rules.h:281-284) */
#define LTG_ERR_CUDA (-2)     /* CUDA runtime error, no device, kernel image missing */
#define LTG_ERR_IO (-3)       /* file could not be read / written */
#define LTG_ERR_LIMIT (-4)    /* input outside the supported envelope (see DESIGN.md "limits") */
#define LTG_ERR_STATE (-5)    /* call sequence error (e.g. no query loaded) */

typedef struct ltg_context ltg_context;

/* struct para — fastsim.h:22-45; defaults of initEnv — Fasim-LongTarget.cpp:284-303 */
typedef struct ltg_params {
    int32_t rule;          /* -r : 0 = all rules, else 1..6 (para) / 1..18 (anti)             */
    int32_t cut_length;    /* -c : segment length, default 5000                                */
    int32_t strand;        /* -t : 0 both, >0 parallel only, <0 anti-parallel only             */
    int32_t overlap;       /* -o : segment overlap, default 100                                */
    int32_t nt_min;        /* -ni: default 20                                                  */
    int32_t nt_max;        /* -na: default 100000                                              */
    int32_t min_identity;  /* -i : default 60   (parsed with atoi in the reference)             */
    int32_t min_stability; /* -S : default 1    (parsed with atoi in the reference)             */
    int32_t penalty_t;     /* -pt: default -1000                                               */
    int32_t penalty_c;     /* -pc: default 0                                                   */
    int32_t c_distance;    /* -ds: default 15                                                  */
    int32_t c_length;      /* -lg: default 50                                                  */
} ltg_params;

/* struct triplex — sim.h:20-45 (one reported triplex).  Strings live in the result's text pool. */
typedef struct ltg_triplex {
    int32_t stari, endi;        /* QueryStart / QueryEnd (1-based, on the lncRNA)                 */
    int32_t starj, endj;        /* StartInSeq / EndInSeq (1-based, on the DNA record)             */
    int32_t reverse;            /* Para: +1 parallel, -1 anti-parallel                            */
    int32_t strand;             /* 0 / 1                                                          */
    int32_t rule;
    int32_t nt;
    float score, identity, tri_score;
    int32_t middle, center, motif;   /* MidPoint / Center / Class (filled by ltg_cluster)          */
    int64_t genomestart, genomeend;
    int64_t tfo_off, tts_off;   /* offsets of the NUL-terminated "TFO sequence" / "TTS sequence"  */
    int64_t chr_off;            /* offset of the NUL-terminated chromosome tag                    */
    int32_t record;             /* index of the DNA record this triplex came from                 */
    int32_t pad_;
} ltg_triplex;

/* Result of scanning DNA against the loaded lncRNA; owned by the library, freed by ltg_result_free. */
typedef struct ltg_result {
    int64_t n_triplex;
    ltg_triplex* triplex;
    int64_t text_bytes;
    char* text;
    /* work counters for GCUPS / Mbp/s */
    int64_t n_segments;         /* non-homopolymer segments scanned                                */
    int64_t n_tasks;            /* (segment, rule, strand, orientation) tasks                      */
    int64_t scan_cells;         /* sum over tasks of m * n_seg (one pass; unpadded m)              */
    int64_t dna_bases;          /* bases of DNA covered                                            */
    int64_t n_peaks;            /* candidate windows (peaks)                                       */
    int64_t window_cells;       /* forward window cells actually computed (m * cut per alignment)  */
    int64_t n_literal_tasks;    /* tasks re-run through the literal striped emulation (Q4 guard)   */
    int64_t n_literal_windows;
    double gpu_ms_scan;         /* CUDA-event time of the scan + peak kernels                      */
    double gpu_ms_window;       /* CUDA-event time of the window kernels                           */
    int64_t gpu_launches;       /* kernels launched for this result                                */
    double gpu_ms_scan_kernel;  /* CUDA-event time of the k_scan launches alone (roofline numerator) */
    int64_t n_scan_launches;    /* number of k_scan launches                                       */
    int64_t h2d_bytes;          /* bytes copied host -> device for this result (DNA, descriptors)   */
    int64_t d2h_bytes;          /* bytes copied device -> host (per-peak records, strings, counters) */
    int64_t n_q4_probed;        /* task pairs swept a second time to decide whether the Q4 emulation is needed */
} ltg_result;

/* ---- context ---------------------------------------------------------------------------------- */
int ltg_create(int device, ltg_context** out);                /* replaces: process start (main, Fasim-LongTarget.cpp:78) */
void ltg_destroy(ltg_context* ctx);
const char* ltg_last_error(void);
void ltg_default_params(ltg_params* p);                        /* initEnv defaults — Fasim-LongTarget.cpp:284-303 */
int ltg_set_params(ltg_context* ctx, const ltg_params* p);

/* lowercase_variant != 0 selects the behaviour of the OLDER driver that ships next to the canonical one (fasim-LongTarget.cpp +
 * fastSim.h): per-peak window loop of fastSim.h:194-226 (no start clamp, acceptance only on equality, no best-candidate
 * tracking, the last window's alignment is always converted) and no per-task identity / stability filter (:311-313).  The
 * `fasim --compat lowercase` command line adds that variant's output naming (fasim-LongTarget.cpp:883).                    */
int ltg_set_compat(ltg_context* ctx, int lowercase_variant);

/* replaces the -F switch (paraList.doFastSim = false, Fasim-LongTarget.cpp:360-362): with on != 0 every task is aligned by
 * SIM() — sim.h:410-1143, Huang & Miller's k best non-intersecting local alignments — instead of fastSIM(); the scan calls
 * then return SIM's rows (dispatch at Fasim-LongTarget.cpp:419-426 and its 15 sibling call sites).                          */
int ltg_set_sim_mode(ltg_context* ctx, int on);

/* replaces readRna + the per-call RNA preparation (TranslateBase ssw_cpp.cpp:323, cg_str stats.h:306,
 * ssw_init/qP_byte sswNew.cpp:1274/176, init_work stats.h:386): uploads the lncRNA once and builds the
 * device-resident packed query profiles for every task pair.                                          */
int ltg_set_query(ltg_context* ctx, const char* name, const char* rna, int64_t len);

/* replaces LongTarget() for one DNA record — Fasim-LongTarget.cpp:379-598 (cutSequence, same_seq, the
 * 48-task loop with calc_score_once + fastSIM, final filter) and the coordinate fix-up of main() :141-149.
 * `dna` is host memory; the call includes the H2D copy of the record and the D2H copy of the hits.     */
int ltg_scan_record(ltg_context* ctx, const char* dna, int64_t len, const char* chr, int64_t record_start,
                    ltg_result** out);

/* Several records in one call (the multi-record -f1 file of main(), Fasim-LongTarget.cpp:133-163): the segments of all
 * records share the device batches, which keeps the GPU busy when the records are short.  The result equals the
 * concatenation of ltg_scan_record over the records; ltg_triplex.record is the index into the arrays.                  */
int ltg_scan_records(ltg_context* ctx, int64_t n_records, const char* const* dna, const int64_t* len, const char* const* chr,
                     const int64_t* record_start, ltg_result** out);

/* ltg_scan_records with every record given by a pointer that is either host memory (dna_on_device = 0) or device memory
 * (1: the records are already resident in HBM — bench.py's device-resident figure for the multi-record example sets).   */
int ltg_scan_records_at(ltg_context* ctx, int64_t n_records, const void* const* dna, int dna_on_device, const int64_t* len,
                        const char* const* chr, const int64_t* record_start, ltg_result** out);

/* Same computation with the DNA already resident in HBM (device pointer to `len` ASCII bytes). Used by
 * bench.py for the device-resident throughput figure.                                                  */
int ltg_scan_device(ltg_context* ctx, const void* d_dna, int64_t len, const char* chr, int64_t record_start,
                    ltg_result** out);

/* One shard of a record for the one-process-per-GPU deployment: scans the segments
 * [first_segment, first_segment + n_segments) of cutSequence's enumeration (fastsim.h:71-90) of a record of `record_len`
 * bases.  `dna` points at the first byte of segment `first_segment` (host memory, or device memory when dna_on_device),
 * `len` bytes are readable.  Coordinates in the result are those of the whole record, so appending the shards' results in
 * shard order (ltg_result_append) reproduces ltg_scan_record of the whole record exactly.                              */
int ltg_scan_shard(ltg_context* ctx, const void* dna, int dna_on_device, int64_t len, const char* chr, int64_t record_start,
                   int64_t record_len, int64_t first_segment, int64_t n_segments, ltg_result** out);

/* ltg_scan_shard for DNA kept 2-bit packed (chromosome-scale stores).  `packed` holds bases in UCSC .2bit
 * coding (4 per byte, the first in the two high bits, T0 C1 A2 G3) in host memory or — the packed genome store resident in HBM,
 * 0.25 B/base — in device memory; the shard's first base has index `first_base` in it, `len` bases are readable.  N runs come as
 * sorted (start, size) blocks RELATIVE to first_base (the nBlock arrays of a .2bit record).  The bases are expanded on the
 * device, the segment descriptors of cutSequence (fastsim.h:71-90) and the same_seq flags (Fasim-LongTarget.cpp:873) are
 * produced there too.  Replaces readDna + cutSequence for such inputs; results equal ltg_scan_shard on the expanded text.     */
int ltg_scan_packed(ltg_context* ctx, const void* packed, int packed_on_device, int64_t first_base, int64_t len,
                    const uint32_t* n_start, const uint32_t* n_size, int32_t n_blocks, const char* chr, int64_t record_start,
                    int64_t record_len, int64_t first_segment, int64_t n_segments, ltg_result** out);

/* concatenates src into dst (dst may be empty); triplex order is preserved (main() :150-163)          */
int ltg_result_append(ltg_result* dst, const ltg_result* src);
int ltg_result_new(ltg_result** out);
void ltg_result_free(ltg_result* r);

/* replaces cluster_triplex + sort — Fasim-LongTarget.cpp:600-691, 812-813 (in place)                    */
int ltg_cluster(ltg_result* r, const ltg_params* p);
/* replaces printResult's -TFOsorted writer — Fasim-LongTarget.cpp:797-829; `r` must be clustered.       */
int ltg_write_tfosorted(const ltg_result* r, const char* path);
/* replaces print_cluster (the -TFOclass{1,2}-<ds>-<lg> bedGraph files) — Fasim-LongTarget.cpp:694-795   */
int ltg_write_tfoclass(const ltg_result* r, const ltg_params* p, const char* sorted_path, const char* chr,
                       int64_t record_start, int64_t dna_size, const char* rna_name);

/* replaces main() — Fasim-LongTarget.cpp:78-172: same flags (-f1 -f2 -O -r -c -m -t -i -S -ni -na -pc -pt
 * -o -ds -lg -cn -d), same output file names, plus `--device N`.                                        */
int ltg_main(int argc, char* const* argv);

/* ---- function-level seams (used by the parity tests; same meaning as the reference calls) ------------- */
/* calc_score_once — stats.h:879 — for the task (para, strand, rule) of one segment (<= cut_length bases) */
/* Aligner::preAlign — ssw_cpp.cpp:388 — per-column maxima (ssw_pre_align, sswNew.cpp:1309) and peaks      */
typedef struct ltg_task_probe {
    int32_t para, strand, rule;   /* in  */
    int32_t max_score;            /* out: calc_score_once                                                  */
    int32_t threshold;            /* out: (int)(max_score * 0.8) — Fasim-LongTarget.cpp:413                */
    int32_t n_peaks;              /* out                                                                   */
    int32_t literal;              /* out: 1 if the task went through the literal striped emulation         */
    int32_t pad_;
} ltg_task_probe;
/* Runs the scan stage for `n_tasks` tasks of one segment.  colmax (may be NULL) receives n_tasks*seg_len
 * ints; peak_score/peak_pos (may be NULL) receive up to peak_cap entries per task (row-major).           */
int ltg_probe_segment(ltg_context* ctx, const char* seg, int32_t seg_len, ltg_task_probe* tasks, int32_t n_tasks,
                      int32_t* colmax, int32_t* peak_score, int32_t* peak_pos, int32_t peak_cap);

/* Aligner::Align — ssw_cpp.cpp:599 (ssw_align sswNew.cpp:1446: forward SW, reverse SW, banded_sw
 * traceback) for `n` windows given as translated DNA text.  out6 = n rows of {sw_score, ref_begin,
 * ref_end, query_begin, query_end, n_cigar}; cigar = n rows of cigar_cap BAM-encoded ops.              */
int ltg_probe_align(ltg_context* ctx, const char* const* windows, const int32_t* window_len, int32_t n,
                    int32_t* out6, uint32_t* cigar, int32_t cigar_cap);

/* window-stage work counters accumulated over the context's life: out[2*round + retry] = windows planned in forward
 * round 0..3 (retry 1 = second, wider sweep after an inconclusive pruned one), out[8] = reverse sweeps,
 * out[10 + k] = the DP cells of the same entries, out[20..23] = alignments handed on by traceback tiers 1..4.
 * out[24..29] (only with LTG_FILTER_STATS=1) = alignments: live, nt bound below the cut, nt ok, identity ok, stability ok,
 * all three ok.  `out` holds 30 entries.                                                                        */
void ltg_debug_stats(ltg_context* ctx, int64_t* out30, int reset);

/* device-timing helpers for bench.py: opaque cudaStream_t of the context */
void* ltg_stream(ltg_context* ctx);
int ltg_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FASIM_B200_H */

/*
 * simmr_b200.h — C ABI of libsimmr_b200.so, the B200-native replacement for
 * the read-simulation hot path of genomicsoup/simmr.
 *
 * The reference is a single Rust binary with no FFI; the boundary replaced is
 * the call pair in simmr/src/main.rs:180-206
 *     simulate::simulate_pe_reads / simulate_long_reads   (simulate.rs:110, :323)
 *     fastq::write_to_fastq                                (fastq.rs:14)
 * together with the plugin traits those consume
 *     trait ErrorProfile      (error_profiles/base.rs:6-32)
 *     trait AbundanceProfile  (abundance_profiles/base.rs:10-69)
 * Per-read virtual dispatch cannot cross to a GPU, so profiles cross the
 * boundary as plain-data descriptors handed over once.
 * INTEGRATION.md shows the Rust `extern "C"` block a maintainer would add.
 *
 * Conventions: every function returns 0 on success or a negative SIMMR_E_*;
 * simmr_last_error() gives the message.  The caller owns every input buffer;
 * the library owns device memory and the pinned output ring; nothing outlives
 * the context.  A context is not thread-safe.  There is no CPU fallback: any
 * compute entry point fails with SIMMR_E_CUDA when no sm_100 device is usable.
 *
 * All citations are relative to the reference tree (/root/reference/).
 */
#ifndef SIMMR_B200_H
#define SIMMR_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SIMMR_OK            0
#define SIMMR_E_INVALID    -1   /* bad argument                                 */
#define SIMMR_E_CUDA       -2   /* CUDA runtime / no device                      */
#define SIMMR_E_IO         -3   /* file could not be read or written            */
#define SIMMR_E_PARSE      -4   /* malformed FASTA / genome file / profile blob */
#define SIMMR_E_STATE      -5   /* genomes or profile not set                   */
#define SIMMR_E_TOO_SMALL  -6   /* contig <= minimum_genome_size (main.rs:124)  */
#define SIMMR_E_SINK       -7   /* sink callback returned non-zero              */
#define SIMMR_E_LIMIT      -8   /* header/read too large for the device tiles   */

/* cli.rs:61-70 ErrorProfile; CUSTOM_LONG is the is_long blob path (custom_long.rs) */
typedef enum {
    SIMMR_EP_PERFECT_SHORT = 0, SIMMR_EP_MINIMAL_SHORT = 1, SIMMR_EP_PERFECT_LONG = 2,
    SIMMR_EP_MINIMAL_LONG = 3, SIMMR_EP_CUSTOM_SHORT = 4, SIMMR_EP_CUSTOM_LONG = 5
} simmr_eprofile_kind;
/* cli.rs:75-84 AbundanceProfile */
typedef enum { SIMMR_AP_EXACT = 0, SIMMR_AP_UNIFORM = 1, SIMMR_AP_CUSTOM = 2 } simmr_aprofile_kind;

/* start of a long read drawn in [0,RL) as simulate.rs:484 does, instead of [0,size-RL) */
#define SIMMR_FLAG_COMPAT_LONG_START 1u

typedef struct simmr_ctx simmr_ctx;

/* Flat genome set == the device layout (genome.rs:15-41 Genome/Seq):
 * contigs concatenated, one offset table, ids as string pools. */
typedef struct {
    uint32_t n_genomes, n_contigs;
    const uint8_t  *bases;          /* Seq.seq of every contig, concatenated      */
    const uint64_t *contig_off;     /* [n_contigs+1] byte offsets into bases      */
    const uint64_t *contig_size;    /* [n_contigs]   Seq.size (differs from the
                                       byte length only with --contiguous)        */
    const uint32_t *genome_contig;  /* [n_genomes+1] first contig of each genome  */
    const char     *contig_ids;     /* Seq.id pool                                */
    const uint32_t *contig_id_off;  /* [n_contigs+1]                              */
    const char     *genome_ids;     /* Genome.uuid pool                           */
    const uint32_t *genome_id_off;  /* [n_genomes+1]                              */
} simmr_genome_set;

/* cli.rs:229-301 determine_error_profile, as data */
typedef struct {
    int32_t  kind;           /* simmr_eprofile_kind                              */
    uint16_t read_length;    /* --read-length  (cli.rs:124-130)                  */
    uint16_t insert_size;    /* --insert-size  (cli.rs:141-147)                  */
    uint8_t  mean_phred;     /* --mean-phred-score (cli.rs:150-156)              */
    uint8_t  pad_[3];
    uint32_t flags;          /* SIMMR_FLAG_*                                     */
    const uint8_t *blob;     /* --custom-profile bytes (encoding.rs:268-281)     */
    uint64_t blob_len;
} simmr_eprofile_desc;

/* cli.rs:306-320 determine_abundance_profile, as data */
typedef struct {
    int32_t kind;            /* simmr_aprofile_kind                              */
    int32_t size_adjusted;   /* --size-adjusted                                  */
    const double *abundances;/* custom: one per genome (main.rs:165-175)         */
} simmr_aprofile_desc;

typedef struct {
    uint64_t seed;           /* --seed; has_seed==0 -> from entropy              */
    int32_t  has_seed;
    uint32_t read_id_base;   /* first read id (simulate.rs:85-89 starts at 0)    */
    const char *header_format;   /* --read-header-format; NULL = cli.rs:196 default */
    uint64_t unit_begin, unit_end;  /* shard of the global unit index space (unit = pair (PE) or long read).
                                       Without SIMMR_RUN_EXPLICIT_RANGE, 0,0 = all; with it the range is taken
                                       literally, so an empty shard that starts at 0 stays empty */
    uint64_t batch_units;    /* units per device batch; 0 = library default      */
    uint32_t flags;          /* SIMMR_RUN_*                                      */
    uint32_t pad_;
} simmr_run_opts;
#define SIMMR_RUN_EXPLICIT_RANGE 1u

/* One chunk of FASTQ bytes, in file order.  `data` is pinned host memory owned
 * by the library, valid until the callback returns.  Non-zero return aborts.
 * The callback runs on a writer thread owned by the library (one per run), never
 * concurrently with itself; the launching thread keeps the GPU busy meanwhile. */
typedef int (*simmr_sink_fn)(void *user, const uint8_t *data, uint64_t len,
                             uint64_t first_unit, uint64_t n_units);
/* A sink that returns SIMMR_SINK_RETAINED keeps `data` after it returns (e.g. for an I/O thread of its own) and gives
 * it back with simmr_chunk_release(data), from any thread.  A context owns SIMMR_RING_BUFFERS pinned buffers; the run
 * waits while all of them are retained, and every retained chunk must be released before the run's entry point returns
 * (the entry point waits for that). */
#define SIMMR_SINK_RETAINED 2
#define SIMMR_RING_BUFFERS  8
void simmr_chunk_release(const uint8_t *data);

typedef struct {
    uint64_t units, records, bases, bytes;  /* produced by the last run          */
    uint64_t clamped;          /* windows clipped at a contig end */
    uint64_t algorithmic_bytes;/* sum over records of 3L + H + 5 */
    double   device_ms;        /* CUDA-event time of all kernels of the run       */
    double   format_ms;        /* ... of the fused gather/format kernel only      */
    uint64_t kernel_launches;  /* kernels launched by the run                     */
    uint64_t h2d_bytes, d2h_bytes;
} simmr_stats;

/* Replay draws: the semantic draws of simulate.rs, recorded
 * per unit / per base, consumed instead of Philox.  Host arrays. */
typedef struct {
    uint64_t n_units;
    const uint32_t *genome;      /* genome index                                  */
    const uint32_t *contig;      /* simulate.rs:182 / :375 index within genome    */
    const uint16_t *read_length; /* simulate.rs:211 / :358                        */
    const uint16_t *insert_size; /* simulate.rs:212                               */
    const uint64_t *start;       /* simulate.rs:233 fwd_start / :484 read_start   */
    const uint64_t *redraw;      /* simulate.rs:245 new_rev_end / :489 read_end   */
    const uint64_t *base_off;    /* [n_units+1] first per-base slot of each unit  */
    const float    *phred_z;     /* StandardNormal deviate per base               */
    const uint32_t *phred_cand;  /* custom: WeightedAliasIndex candidate          */
    const double   *phred_u;     /* custom: uniform in [0, weight_sum)            */
    const uint32_t *phred_inbin; /* custom: raw draw for Uniform::new_inclusive   */
    const float    *mut_u;       /* minimal_short.rs:123 gen::<f32>()             */
    const uint8_t  *mut_k;       /* minimal_short.rs:125 choose() index           */
} simmr_draws;

/* ---- context ------------------------------------------------------------ */
int  simmr_ctx_create(int device, simmr_ctx **out);
void simmr_ctx_destroy(simmr_ctx *ctx);
const char *simmr_last_error(const simmr_ctx *ctx);   /* ctx may be NULL */
int  simmr_device_info(const simmr_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor,
                       uint64_t *hbm_bytes);

/* ---- inputs ------------------------------------------------------------- */
/* replaces holding Vec<Genome> across simulate (main.rs:37-110): copies the
 * set into device-resident buffers once */
int  simmr_genomes_load(simmr_ctx *ctx, const simmr_genome_set *set);
/* replaces Box<dyn ErrorProfile> (cli.rs:229-301) */
int  simmr_error_profile_set(simmr_ctx *ctx, const simmr_eprofile_desc *desc);
/* ErrorProfile::minimum_genome_size / is_long_read (base.rs:29-31), host only */
int  simmr_eprofile_minimum_genome_size(const simmr_eprofile_desc *desc, uint16_t *out);
int  simmr_eprofile_is_long_read(const simmr_eprofile_desc *desc, int *out);
/* AbundanceProfile::determine_abundances + adjust_for_size (base.rs:27-68;
 * uniform.rs:18-95, exact.rs:17-35, custom.rs:20-96), host only, bit-identical */
int  simmr_abundance_counts(const simmr_aprofile_desc *desc, uint64_t total_reads, uint32_t n_genomes,
                            const uint64_t *genome_sizes, uint64_t *out_reads, double *out_abundance);

/* ---- the hot path ------------------------------------------------------- */
/* replaces simulate_pe_reads / simulate_long_reads + write_to_fastq.
 * reads_per_genome[i] is the .0 of determine_abundances for genome i.
 * sink == NULL: device-resident run (FASTQ bytes stay in HBM staging). */
int  simmr_simulate(simmr_ctx *ctx, const uint64_t *reads_per_genome, const simmr_run_opts *opts,
                    simmr_sink_fn sink, void *user);
/* The same run for a consumer on the device: no copy to the host, every batch's FASTQ bytes are handed over where the
 * format kernel left them in HBM (`device_data`, valid until the callback returns; the kernels of the batch have
 * completed).  Non-zero return aborts.  The callback runs on the calling thread. */
typedef int (*simmr_device_sink_fn)(void *user, const uint8_t *device_data, uint64_t len,
                                    uint64_t first_unit, uint64_t n_units);
int  simmr_simulate_device(simmr_ctx *ctx, const uint64_t *reads_per_genome, const simmr_run_opts *opts,
                           simmr_device_sink_fn sink, void *user);
/* same output path, draws supplied instead of Philox */
int  simmr_replay(simmr_ctx *ctx, const simmr_draws *draws, const simmr_run_opts *opts,
                  simmr_sink_fn sink, void *user);
uint64_t simmr_total_units(const simmr_ctx *ctx, const uint64_t *reads_per_genome);
/* FASTQ bytes and records the unit range of `opts` will produce (needs a seed): runs the
 * geometry kernel only.  This is the per-shard count of the host-side exchange that gives
 * every GPU its offset in the single output file (the reference appends genome after
 * genome, main.rs:201-206). */
int  simmr_measure(simmr_ctx *ctx, const uint64_t *reads_per_genome, const simmr_run_opts *opts,
                   uint64_t *bytes, uint64_t *records);
int  simmr_stats_get(const simmr_ctx *ctx, simmr_stats *out);

/* ---- the same path over several GPUs of one box ----- */
/* One context per device, genomes and profile replicated.  A run is cut into batches of consecutive units; batch j is
 * produced by device j mod n on that device's own launch thread, and ONE writer thread hands the batches to the sink in
 * file order — the bytes are those of the one-GPU run (the reference appends genome after genome, main.rs:201-206).
 * No device collective.  devices: CUDA ordinals, n >= 1. */
typedef struct simmr_multi simmr_multi;
int  simmr_multi_create(const int *devices, int n_devices, simmr_multi **out);
void simmr_multi_destroy(simmr_multi *m);
int  simmr_multi_genomes_load(simmr_multi *m, const simmr_genome_set *set);
int  simmr_multi_error_profile_set(simmr_multi *m, const simmr_eprofile_desc *desc);
uint64_t simmr_multi_total_units(const simmr_multi *m, const uint64_t *reads_per_genome);
int  simmr_multi_simulate(simmr_multi *m, const uint64_t *reads_per_genome, const simmr_run_opts *opts,
                          simmr_sink_fn sink, void *user);
/* sums over the devices; device_ms / format_ms are the slowest device's */
int  simmr_multi_stats_get(const simmr_multi *m, simmr_stats *out);
const char *simmr_multi_last_error(const simmr_multi *m);

/* ---- host-side I/O of the same path (genome.rs, files.rs, encoding.rs) -- */
typedef struct simmr_genome_builder simmr_genome_builder;
int  simmr_builder_create(simmr_genome_builder **out);
void simmr_builder_destroy(simmr_genome_builder *b);
/* Genome::from_fasta (genome.rs:89-162): uuid NULL -> random (util.rs:124-129) */
int  simmr_builder_add_fasta(simmr_genome_builder *b, const char *path, const char *uuid, int contiguous);
/* the same for n files at once, read, decompressed (gzip, bzip2, xz by their magic bytes, as needletail does) and normalised
 * in parallel; uuids may be NULL or hold NULL entries; threads <= 0: one per core, at most 32 */
int  simmr_builder_add_fastas(simmr_genome_builder *b, const char *const *paths, const char *const *uuids, uint32_t n,
                              int contiguous, int threads);
/* main.rs:117-162: drop contigs with size <= min_size, then empty genomes.
 * kept[i] (optional, one per genome added) says whether genome i survived. */
int  simmr_builder_filter(simmr_genome_builder *b, uint16_t min_size, uint8_t *kept);
int  simmr_builder_view(const simmr_genome_builder *b, simmr_genome_set *out);
uint64_t simmr_builder_genome_size(const simmr_genome_builder *b, uint32_t genome);  /* Genome.size, pre-filter */
const char *simmr_builder_error(const simmr_genome_builder *b);

/* files.rs:56-86 parse_genome_file.  Returns the record count (>=0) and writes
 * "path\tuuid\tabundance\n" per record ("" / "nan" when absent). */
int64_t simmr_parse_genome_file(const char *path, char *out, uint64_t cap);
/* files.rs:100-134 write_metadata */
int  simmr_write_metadata(const char *output, uint32_t n, const char *const *uuid,
                          const char *const *path, const uint64_t *num_reads, const double *abundance);

/* ---- output side: what the reference's readme lists as to-do (readme.md:209-211) and fastq.rs:14-124 writes as one
 * plain interleaved file.  Host-side helpers for a sink; the bytes a sink receives stay plain interleaved FASTQ. ---- */
/* One gzip member (RFC 1952) holding src[0, n): members may be concatenated, so the slices of a chunk can be compressed
 * on as many threads as the caller has and written in order.  Returns the member's bytes, or a negative SIMMR_E_*
 * (SIMMR_E_LIMIT: cap too small; simmr_gzip_bound gives a sufficient cap). */
int64_t  simmr_gzip_member(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, int level);
uint64_t simmr_gzip_bound(uint64_t n);
/* Splits interleaved paired-end FASTQ (whole records, mate 1 first) into its two mates: r1 and r2 need n bytes of room
 * each.  Writes the byte counts; SIMMR_E_PARSE if src does not hold an even number of four-line records. */
int  simmr_fastq_split_mates(const uint8_t *src, uint64_t n, uint8_t *r1, uint64_t *n1, uint8_t *r2, uint64_t *n2);

/* ---- introspection used by the parity tests (host only) ----------------- */
/* kind 3 (short reads) of a mutating profile: the (npos << cb) quality words are followed by the 256 words of the
 * substitution-gap table and the (1 << cb) words of the quality-given-substitution table */
int  simmr_debug_phred_table(const simmr_eprofile_desc *desc, uint32_t *kind, uint32_t *cb, uint32_t *npos,
                             uint32_t *const_q, uint32_t *entries, uint64_t cap);
int  simmr_debug_len_table(const simmr_eprofile_desc *desc, int which, uint32_t *kind, uint32_t *cb,
                           uint32_t *entries, uint32_t *lo, uint32_t *hi, uint64_t cap);
/* custom-long alt-k-mer tables, flattened: ref, keep_thr, row, cb (n_ref words each), then entries and alts */
int  simmr_debug_err_table(const simmr_eprofile_desc *desc, uint32_t *k, uint32_t *n_ref, uint32_t *words, uint64_t cap);
void simmr_debug_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
const char *simmr_version(void);

#ifdef __cplusplus
}
#endif
#endif

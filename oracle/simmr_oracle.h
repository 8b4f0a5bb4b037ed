/*
 * simmr_oracle.h — C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the read-simulation
 * path of genomicsoup/simmr (Rust).  It exists to check the CUDA path; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  Nothing under simmr_b200/ links or calls it.
 *
 * PARITY STATUS.  The reference is Rust; no rustc/cargo exists in this image,
 * so the reference itself cannot be run here (no oracle/_ref).  The oracle is
 * pinned against every known-answer vector the reference's own tests hold for
 * this path (tests/test_oracle_golden.py lists them with file:line).  What the
 * reference's tests do NOT pin — the rand/rand_distr draw streams, the bincode
 * layout, needletail normalisation of exotic bytes — is restated from the
 * crates' published algorithms and is "parity unpinned" (see DESIGN.md §3).
 *
 * All citations are relative to /root/reference/.
 */
#ifndef SIMMR_ORACLE_H
#define SIMMR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- profile kinds (simmr/src/cli.rs:61-70; custom-long per custom_long.rs) */
enum { ORC_PERFECT_SHORT = 0, ORC_MINIMAL_SHORT = 1, ORC_PERFECT_LONG = 2,
       ORC_MINIMAL_LONG = 3, ORC_CUSTOM_SHORT = 4, ORC_CUSTOM_LONG = 5 };
/* abundance kinds (simmr/src/cli.rs:75-84) */
enum { ORC_AB_EXACT = 0, ORC_AB_UNIFORM = 1, ORC_AB_CUSTOM = 2 };
/* flags */
enum { ORC_FLAG_COMPAT_LONG_START = 1 };  /* simulate.rs:484 start in [0,RL) */

typedef struct {
    uint32_t n_genomes, n_contigs;
    const uint8_t  *bases;          /* all contigs, concatenated            */
    const uint64_t *contig_off;     /* [n_contigs+1] physical offsets       */
    const uint64_t *contig_size;    /* [n_contigs] logical Seq.size         */
    const uint32_t *genome_contig;  /* [n_genomes+1] first contig of genome */
    const char     *contig_ids;     /* concatenated Seq.id bytes            */
    const uint32_t *contig_id_off;  /* [n_contigs+1]                        */
    const char     *genome_ids;     /* concatenated Genome.uuid bytes       */
    const uint32_t *genome_id_off;  /* [n_genomes+1]                        */
} orc_genomes;

typedef struct {
    int32_t  kind;
    uint16_t read_length;
    uint16_t insert_size;
    uint8_t  mean_phred;
    uint8_t  pad_[3];
    uint32_t flags;
    const uint8_t *blob;   /* custom profiles: bincode ErrorModelParams */
    uint64_t blob_len;
} orc_profile;

typedef struct {
    uint64_t seed;
    const char *header_format;  /* NULL -> cli.rs:196 default */
    uint64_t unit_begin, unit_end;  /* half-open; 0,0 = everything */
    uint32_t read_id_base;
} orc_opts;

/* Replay draws.  Per-unit arrays of length n_units; per-base
 * arrays are concatenated, base_off[u] giving the first per-base slot of unit
 * u (PE: mate 1 slots then mate 2 slots, RL each; long: end-start slots). */
typedef struct {
    uint64_t n_units;
    const uint32_t *genome;      /* genome index                               */
    const uint32_t *contig;      /* contig index within the genome             */
    const uint16_t *read_length; /* RL                                         */
    const uint16_t *insert_size; /* IS (PE only)                               */
    const uint64_t *start;       /* fwd_start (PE) / read_start (long)         */
    const uint64_t *redraw;      /* new_rev_end (PE case 4a) / read_end redraw */
    const uint64_t *base_off;    /* [n_units+1]                                */
    const float    *phred_z;     /* standard-normal deviates (minimal/perfect-long) */
    const uint32_t *phred_cand;  /* custom: alias candidate                    */
    const double   *phred_u;     /* custom: uniform in [0, weight_sum)         */
    const uint32_t *phred_inbin; /* custom: raw 32-bit draw for the in-bin uniform */
    const float    *mut_u;       /* gen::<f32>() in [0,1)                      */
    const uint8_t  *mut_k;       /* choose() index 0..2                        */
} orc_draws;

typedef struct {
    uint64_t records, bases, bytes, units, clamped, mutated;
} orc_stats;

/* -------- primitives (simmr/src/util.rs, shared/src/util.rs) -------- */
uint8_t  orc_complement(uint8_t n);
void     orc_reverse_complement(const uint8_t *in, uint64_t n, uint8_t *out);
void     orc_encode_quality_scores(const uint8_t *in, uint64_t n, uint8_t *out);
float    orc_phred_to_probability(uint8_t q);
float    orc_phred_to_accuracy(uint8_t q);
uint8_t  orc_probability_to_phred(float p);
uint8_t  orc_accuracy_to_phred(float acc);
int      orc_bytes_to_string(const uint8_t *in, uint64_t n, char *out);  /* returns len or 0 */
double   orc_gaussian(double x, const double *xs, uint64_t n, double bw);
int      orc_kmer_encode(const uint8_t *kmer, uint32_t k, uint32_t esize, uint32_t *out);
int      orc_kmer_decode(uint32_t code, uint32_t k, uint32_t esize, int skip_n, uint8_t *out);
/* k-mer splice (custom_long.rs:187-232); alt choice i for the j-th matched
 * window is choice[j] (replay) */
int64_t  orc_simulate_errors(const uint8_t *seq, uint64_t n, uint32_t k,
                             const uint32_t *ref_kmers, const uint32_t *alt_off,
                             const uint32_t *alt_kmers, uint32_t n_ref,
                             const uint32_t *choice, uint64_t n_choice,
                             uint8_t *out, uint64_t out_cap);
void     orc_format_f64(double v, char *out, int cap);

/* -------- Philox4x32-10 (Salmon et al. 2011, Random123) -------- */
void     orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* -------- abundance (abundance_profiles/{uniform,exact,custom}.rs) -------- */
int      orc_abundance(int kind, int size_adjusted, uint64_t total_reads, uint32_t n,
                       const double *abund_in, const uint64_t *genome_sizes,
                       uint64_t *out_reads, double *out_abund);

/* -------- profile-level scalars (error_profiles/ *.rs) -------- */
int      orc_profile_min_genome_size(const orc_profile *p, uint16_t *out);
int      orc_profile_is_long(const orc_profile *p, int *out);
/* quantised alias tables the native mode samples from */
int      orc_profile_phred_table(const orc_profile *p, uint32_t *kind, uint32_t *cb, uint32_t *npos,
                                 uint32_t *const_q, uint32_t *entries, uint64_t cap);
int      orc_profile_err_table(const orc_profile *p, uint32_t *k, uint32_t *n_ref, uint32_t *words, uint64_t cap);
int      orc_profile_len_table(const orc_profile *p, int which /*0=RL,1=IS*/,
                               uint32_t *kind, uint32_t *cb, uint32_t *entries,
                               uint32_t *lo, uint32_t *hi, uint64_t cap);

/* -------- FASTA / genome file / metadata (genome.rs, files.rs) -------- */
typedef struct orc_fasta orc_fasta;
int      orc_fasta_load(const char *path, int contiguous, orc_fasta **out, char *err, int errcap);
void     orc_fasta_free(orc_fasta *);
uint32_t orc_fasta_num_seqs(const orc_fasta *);
uint64_t orc_fasta_size(const orc_fasta *);
uint64_t orc_fasta_seq_size(const orc_fasta *, uint32_t i);
uint64_t orc_fasta_seq_len(const orc_fasta *, uint32_t i);     /* physical bytes */
const uint8_t *orc_fasta_seq(const orc_fasta *, uint32_t i);
const char *orc_fasta_seq_id(const orc_fasta *, uint32_t i, uint32_t *len);
/* genome file: returns number of records or <0; fields written as
 * "path\tuuid\tabundance\n" lines ("" / "nan" when absent) */
int64_t  orc_parse_genome_file(const char *path, char *out, uint64_t cap);
int64_t  orc_metadata_tsv(uint32_t n, const char *const *uuid, const char *const *path,
                          const uint64_t *num_reads, const double *abundance,
                          char *out, uint64_t cap);

/* -------- custom-profile blob (shared/src/encoding.rs:82-117) -------- */
int      orc_blob_summary(const uint8_t *blob, uint64_t len, uint64_t *bin_size,
                          uint64_t *n_pos, uint64_t *kmer_size, uint64_t *n_kmers,
                          double *is_mean, double *rl_mean, int *has_is_bins, int *is_long);

/* -------- drivers -------- */
/* units for a run: PE = sum floor(reads/2), long = sum reads */
uint64_t orc_total_units(const orc_profile *p, const uint64_t *reads_per_genome, uint32_t n);
/* native (Philox) mode: FASTQ bytes of units [unit_begin, unit_end).  Call with
 * out==NULL to get the size.  Returns bytes written/needed or <0. */
int64_t  orc_simulate(const orc_genomes *g, const orc_profile *p,
                      const uint64_t *reads_per_genome, const orc_opts *o,
                      uint8_t *out, uint64_t cap, orc_stats *st);
/* replay mode: reference arithmetic on supplied draws */
int64_t  orc_replay(const orc_genomes *g, const orc_profile *p, const orc_draws *d,
                    const orc_opts *o, uint8_t *out, uint64_t cap, orc_stats *st);
/* native geometry only (tests): per unit genome, contig(global), RL, IS, start, lo */
int      orc_native_geometry(const orc_genomes *g, const orc_profile *p,
                             const uint64_t *reads_per_genome, const orc_opts *o,
                             uint32_t *genome, uint32_t *contig, uint32_t *rl, uint32_t *is_,
                             uint64_t *start, uint64_t *lo_or_end);

/* -------- reference-arithmetic samplers (statistical tests only) -------- */
/* draws from the reference's formulas with an independent CPU generator */
void     orc_ref_minimal_phred(uint8_t mean_phred, uint64_t n, uint64_t seed, uint8_t *out);
void     orc_ref_perfect_long_phred(uint64_t n, uint64_t seed, uint8_t *out);
void     orc_ref_normal_len(double mean, double sd, uint64_t n, uint64_t seed, uint16_t *out);
void     orc_ref_gamma_len(uint64_t n, uint64_t seed, uint16_t *out);
void     orc_ref_mutate(const uint8_t *seq, const uint8_t *q, uint64_t n, uint64_t seed, uint8_t *out);
int      orc_ref_custom_phred(const uint8_t *blob, uint64_t blob_len, uint32_t pos,
                              uint64_t n, uint64_t seed, uint8_t *out);
int      orc_ref_custom_len(const uint8_t *blob, uint64_t blob_len, int which,
                            uint64_t n, uint64_t seed, uint16_t *out);

/* CPU baseline of bench.py (paired-end): mode 0 = native draw schedule, mode 1 = the reference's own stream plumbing
 * (ChaCha12 re-seeded per pair, ziggurat normal + f32 powf per base; ref_stream.inc), `threads` std::threads, buffers
 * touched before the clock. */
int      orc_bench_pe(const orc_genomes *g, const orc_profile *p, const uint64_t *reads_per_genome, uint64_t seed,
                      uint64_t unit_begin, uint64_t n_units, int mode, int threads, double *seconds,
                      uint64_t *bases, uint64_t *records, uint64_t *bytes);
int64_t  orc_native_errors(const orc_profile *p, uint64_t seed, uint64_t unit, const uint8_t *seq, uint64_t n,
                           uint8_t *out, uint64_t cap, uint32_t *choices, uint64_t choice_cap, uint64_t *n_choices);
void     orc_chacha12_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]);
void     orc_ref_ziggurat(uint64_t seed, uint64_t n, double *out);
int64_t  orc_ref_stream(const orc_genomes *g, const orc_profile *p, uint64_t n_pairs, uint64_t seed,
                        uint8_t *out, uint64_t cap, orc_stats *st);

const char *orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif

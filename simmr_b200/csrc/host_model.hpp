/* host_model.hpp — host-side model of the simmr read-simulation path:
 * profile descriptors -> sampling tables, custom-profile blob, abundance
 * counts, FASTA / genome-file ingest, per-contig header templates.
 * Citations are relative to the reference tree (/root/reference/). */
#pragma once
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/simmr_b200.h"
#include "simmr_types.h"

namespace simmr {

/* shared/src/encoding.rs:82-117 */
struct BinSet {
    uint64_t num_bins = 0, bin_width = 0;
    std::vector<double> density;
    std::vector<uint32_t> lo, hi;
    void range_of(size_t j, uint32_t &l, uint32_t &h) const;
};
struct ProfileBlob {
    uint64_t bin_size = 0, kmer_size = 0;
    uint8_t bit_encoding = 0;
    std::vector<BinSet> quality;
    std::vector<uint32_t> ref_kmer;              /* probabilities: flattened */
    std::vector<uint32_t> alt_off, alt_kmer;
    std::vector<float> alt_prob;
    double insert_mean = 0, insert_std = 0, length_mean = 0, length_std = 0;
    bool has_insert_bins = false, is_long = false;
    BinSet insert_bins, length_bins;
};
bool decode_blob(const uint8_t *p, uint64_t n, ProfileBlob &out);

struct uint2_host { uint32_t x, y; };

struct LenModel {
    uint32_t kind = LEN_CONST, value = 0, cb = 0, vmax = 0;
    std::vector<uint32_t> entries, lo, hi;
};

/* k-mer error injection of custom-long profiles (custom_long.rs:187-232), DESIGN.md §5.  Per distinct reference k-mer
 * (3-bit codes, ascending): the chance that the sampled alternate is the k-mer itself as a 32-bit threshold, and an alias row
 * (thr20<<12 | alias12 per column) over the other alternates; `entries` and `alts` share one row offset. */
struct ErrModel {
    bool on = false;
    uint32_t k = 0;
    std::vector<uint32_t> ref, keep_thr, row, cb, entries, alts;
    /* device look-up of windows of A,C,G,T only: 4^k words pairs {keep_thr, row index + 1 (0: k-mer not observed)} by 2-bit code */
    std::vector<uint2_host> direct;
};

struct ProfileModel {
    int kind = 0;
    bool is_long = false, mutates = false;
    uint16_t min_size = 0;
    uint32_t flags = 0;
    /* native sampling tables */
    uint32_t phred_kind = PT_CONST;     /* PT_* */
    uint8_t const_q = 0;
    uint32_t cb = 0, npos = 0;
    std::vector<uint32_t> phred_entries;   /* PT_T64 / PT_T32: quality bytes are raw phreds; PT_Q32: characters as printed (+33) */
    /* PT_Q32 of a mutating profile: bases until the next substitution, quality given a substitution */
    bool has_sub = false;
    std::vector<uint32_t> gap_entries, qsub_entries;
    LenModel rl, is;
    ErrModel err;                        /* custom-long blobs with an alt-k-mer table */
    GammaConst gamma{};
    uint32_t max_read_len = 0;          /* upper bound on a read's length    */
    uint32_t nominal_read_len = 0;      /* typical length (sizes the shared-memory stage) */
    /* replay arithmetic */
    uint32_t replay_kind = RP_CONST;
    float mean_phred_f = 0.f, acc20 = 0.f;
    float acc_lut[256];                 /* util.rs:96-98 per quality         */
    float pl_thr[64]; uint32_t pl_qmax = 0;   /* perfect_long.rs:66-77 as thresholds on accuracy */
    std::vector<double> r_odds;         /* rand_distr WeightedAliasIndex<f64> per position */
    std::vector<uint32_t> r_alias, r_off, r_lo, r_hi;
    std::vector<double> r_sum;
};
int build_profile(const simmr_eprofile_desc *d, ProfileModel &m, std::string &err);

/* abundance_profiles/{uniform,exact,custom}.rs */
int abundance_counts(const simmr_aprofile_desc *d, uint64_t total_reads, uint32_t n, const uint64_t *sizes,
                     uint64_t *out_reads, double *out_abund, std::string &err);

/* growable byte buffer that does not zero what it grows by (a 4 GB genome set is filled by parallel copies) */
class ByteBuf {
    uint8_t *p_ = nullptr; size_t n_ = 0, cap_ = 0;
public:
    ByteBuf() = default;
    ByteBuf(const ByteBuf &) = delete;
    ByteBuf &operator=(const ByteBuf &) = delete;
    ~ByteBuf() { delete[] p_; }
    size_t size() const { return n_; }
    uint8_t *data() { return p_; }
    const uint8_t *data() const { return p_; }
    void grow(size_t n) {             /* size becomes n >= size(); new bytes are uninitialised */
        if (n > cap_) {
            const size_t cap = std::max(n, cap_ + cap_ / 2);
            uint8_t *q = new uint8_t[cap];
            if (n_) memcpy(q, p_, n_);
            delete[] p_; p_ = q; cap_ = cap;
        }
        n_ = n;
    }
    void shrink(size_t n) { if (n < n_) n_ = n; }
};

/* genome.rs:15-41 flattened */
struct GenomeStore {
    ByteBuf bases;
    std::vector<uint64_t> contig_off{0}, contig_size;
    std::vector<uint32_t> genome_contig{0};
    std::string contig_ids, genome_ids;
    std::vector<uint32_t> contig_id_off{0}, genome_id_off{0};
    std::vector<uint64_t> genome_size;         /* Genome.size, never updated by the filter (main.rs:158-161) */
    std::vector<std::string> paths;
    std::string err;
    int add_fasta(const char *path, const char *uuid, bool contiguous);
    /* n files at once: read, decompressed and normalised in parallel (threads == 0: one per core, at most 32) */
    int add_fastas(const char *const *path, const char *const *uuid, uint32_t n, bool contiguous, unsigned threads);
    void filter(uint16_t min_size, uint8_t *kept);
    void view(simmr_genome_set &v) const;
};

bool utf8_ok(const uint8_t *s, size_t n);
std::string rust_f64(double v);
int64_t parse_genome_file(const char *path, std::string &out, std::string &err);
std::string metadata_tsv(uint32_t n, const char *const *uuid, const char *const *path,
                         const uint64_t *reads, const double *abund);

/* fastq.rs:34-57 compiled once per contig */
struct HeaderPlan {
    std::vector<HdrTok> toks;
    std::vector<ContigHdr> contigs;
    std::string pool;
    uint32_t hdr_cap = 0;       /* upper bound on any header's bytes */
    uint32_t hdr_typ = 0;       /* a header's bytes when positions are printed with the digits of its contig's size (sizes a warp's stage) */
};
extern const char *const DEFAULT_HEADER_FORMAT;   /* cli.rs:196 */
int compile_headers(const simmr_genome_set &g, const char *fmt, HeaderPlan &plan, std::string &err);

void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* message behind simmr_last_error(NULL): one slot per host thread, shared by every entry point */
std::string &last_error_slot();

}  // namespace simmr

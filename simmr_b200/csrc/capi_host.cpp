/* capi_host.cpp — the host-only entry points of include/simmr_b200.h: genome
 * ingest (genome.rs, main.rs:117-162), genome file and metadata (files.rs),
 * abundance counts (abundance_profiles/), profile scalars (error_profiles/)
 * and the table introspection the parity tests use.  No CUDA calls here. */
#include "host_model.hpp"

#include <zlib.h>

#include <cstring>
#include <fstream>

using namespace simmr;

namespace { int flunk(int code, const std::string &m) { last_error_slot() = m; return code; } }

struct simmr_genome_builder { GenomeStore store; };

extern "C" {

/* ---- output side helpers (no reference equivalent: fastq.rs writes one plain file) ---- */
uint64_t simmr_gzip_bound(uint64_t n) { return n + n / 1000 + 64 + 18; }
int64_t simmr_gzip_member(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, int level) {
    if ((!src && n) || !dst) return flunk(SIMMR_E_INVALID, "null buffer");
    if (n > 0xfffffff0ull) return flunk(SIMMR_E_LIMIT, "a gzip member holds less than 4 GiB here: compress the chunk in slices");
    z_stream z{};
    if (deflateInit2(&z, level < 0 ? 1 : level > 9 ? 9 : level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return flunk(SIMMR_E_IO, "zlib: deflateInit2 failed");
    z.next_in = const_cast<Bytef *>(src); z.avail_in = uInt(n);
    z.next_out = dst; z.avail_out = uInt(cap > 0xfffffff0ull ? 0xfffffff0ull : cap);
    const int rc = deflate(&z, Z_FINISH);
    const uint64_t out = z.total_out;
    deflateEnd(&z);
    if (rc != Z_STREAM_END) return flunk(SIMMR_E_LIMIT, "gzip output buffer too small");
    return int64_t(out);
}
int simmr_fastq_split_mates(const uint8_t *src, uint64_t n, uint8_t *r1, uint64_t *n1, uint8_t *r2, uint64_t *n2) {
    if ((!src && n) || !r1 || !r2 || !n1 || !n2) return flunk(SIMMR_E_INVALID, "null buffer");
    uint64_t at = 0, a = 0, b = 0, rec = 0;
    while (at < n) {                                  /* a record is four lines (fastq.rs:61-69) */
        uint64_t end = at;
        for (int line = 0; line < 4; ++line) {
            const void *nl = memchr(src + end, '\n', n - end);
            if (!nl) return flunk(SIMMR_E_PARSE, "FASTQ chunk ends inside a record");
            end = uint64_t(static_cast<const uint8_t *>(nl) - src) + 1;
        }
        if (rec & 1) { memcpy(r2 + b, src + at, end - at); b += end - at; } else { memcpy(r1 + a, src + at, end - at); a += end - at; }
        at = end; ++rec;
    }
    if (rec & 1) return flunk(SIMMR_E_PARSE, "FASTQ chunk holds an odd number of records");
    *n1 = a; *n2 = b;
    return SIMMR_OK;
}

int simmr_builder_create(simmr_genome_builder **out) { if (!out) return SIMMR_E_INVALID; *out = new simmr_genome_builder; return SIMMR_OK; }
void simmr_builder_destroy(simmr_genome_builder *b) { delete b; }
const char *simmr_builder_error(const simmr_genome_builder *b) { return b ? b->store.err.c_str() : ""; }
int simmr_builder_add_fasta(simmr_genome_builder *b, const char *path, const char *uuid, int contiguous) {
    if (!b || !path) return SIMMR_E_INVALID;
    return b->store.add_fasta(path, uuid, contiguous != 0);
}
int simmr_builder_add_fastas(simmr_genome_builder *b, const char *const *paths, const char *const *uuids, uint32_t n, int contiguous, int threads) {
    if (!b || (!paths && n)) return SIMMR_E_INVALID;
    return b->store.add_fastas(paths, uuids, n, contiguous != 0, threads > 0 ? unsigned(threads) : 0u);
}
int simmr_builder_filter(simmr_genome_builder *b, uint16_t min_size, uint8_t *kept) {
    if (!b) return SIMMR_E_INVALID;
    b->store.filter(min_size, kept);
    return SIMMR_OK;
}
int simmr_builder_view(const simmr_genome_builder *b, simmr_genome_set *out) {
    if (!b || !out) return SIMMR_E_INVALID;
    b->store.view(*out);
    return SIMMR_OK;
}
uint64_t simmr_builder_genome_size(const simmr_genome_builder *b, uint32_t g) { return b && g < b->store.genome_size.size() ? b->store.genome_size[g] : 0; }

int64_t simmr_parse_genome_file(const char *path, char *out, uint64_t cap) {
    std::string text, err;
    const int64_t n = parse_genome_file(path, text, err);
    if (n < 0) return flunk(int(n), err);
    if (text.size() + 1 > cap) return flunk(SIMMR_E_INVALID, "buffer too small");
    memcpy(out, text.c_str(), text.size() + 1);
    return n;
}
int simmr_write_metadata(const char *output, uint32_t n, const char *const *uuid, const char *const *path,
                         const uint64_t *num_reads, const double *abundance) {
    std::remove(output);                                          /* files.rs:106-108 */
    std::ofstream f(output, std::ios::binary);
    if (!f) return flunk(SIMMR_E_IO, std::string("cannot write ") + output);
    f << metadata_tsv(n, uuid, path, num_reads, abundance);
    return f.good() ? SIMMR_OK : flunk(SIMMR_E_IO, std::string("write failed: ") + output);
}

int simmr_abundance_counts(const simmr_aprofile_desc *d, uint64_t total_reads, uint32_t n, const uint64_t *sizes,
                           uint64_t *out_reads, double *out_abund) {
    if (!d || !out_reads || !out_abund) return flunk(SIMMR_E_INVALID, "null argument");
    std::string err;
    const int rc = abundance_counts(d, total_reads, n, sizes, out_reads, out_abund, err);
    return rc ? flunk(rc, err) : rc;
}

int simmr_eprofile_minimum_genome_size(const simmr_eprofile_desc *d, uint16_t *out) {
    ProfileModel m; std::string err;
    const int rc = build_profile(d, m, err);
    if (rc) return flunk(rc, err);
    *out = m.min_size; return SIMMR_OK;
}
int simmr_eprofile_is_long_read(const simmr_eprofile_desc *d, int *out) {
    ProfileModel m; std::string err;
    const int rc = build_profile(d, m, err);
    if (rc) return flunk(rc, err);
    *out = m.is_long; return SIMMR_OK;
}

int simmr_debug_phred_table(const simmr_eprofile_desc *d, uint32_t *kind, uint32_t *cb, uint32_t *npos, uint32_t *const_q, uint32_t *entries, uint64_t cap) {
    ProfileModel m; std::string err;
    const int rc = build_profile(d, m, err);
    if (rc) return flunk(rc, err);
    *kind = m.phred_kind; *cb = m.cb; *npos = m.npos; *const_q = m.const_q;
    /* PT_Q32 of a mutating profile: the quality table is followed by gap[256] and qsub[1 << cb] */
    std::vector<uint32_t> all = m.phred_entries;
    if (m.has_sub) { all.insert(all.end(), m.gap_entries.begin(), m.gap_entries.end()); all.insert(all.end(), m.qsub_entries.begin(), m.qsub_entries.end()); }
    if (all.size() > cap) return flunk(SIMMR_E_INVALID, "buffer too small");
    if (!all.empty()) memcpy(entries, all.data(), all.size() * 4);
    return int(all.size());
}
int simmr_debug_len_table(const simmr_eprofile_desc *d, int which, uint32_t *kind, uint32_t *cb, uint32_t *entries, uint32_t *lo, uint32_t *hi, uint64_t cap) {
    ProfileModel m; std::string err;
    const int rc = build_profile(d, m, err);
    if (rc) return flunk(rc, err);
    const LenModel &t = which ? m.is : m.rl;
    *kind = t.kind; *cb = t.cb;
    if (t.kind == LEN_CONST) { lo[0] = hi[0] = t.value; return 0; }
    if (t.kind == LEN_GAMMA) return 0;
    if (t.entries.size() > cap) return flunk(SIMMR_E_INVALID, "buffer too small");
    memcpy(entries, t.entries.data(), t.entries.size() * 4);
    memcpy(lo, t.lo.data(), t.lo.size() * 4); memcpy(hi, t.hi.data(), t.hi.size() * 4);
    return int(t.entries.size());
}
int simmr_debug_err_table(const simmr_eprofile_desc *d, uint32_t *k, uint32_t *n_ref, uint32_t *words, uint64_t cap) {
    ProfileModel m; std::string err;
    const int rc = build_profile(d, m, err);
    if (rc) return flunk(rc, err);
    const ErrModel &e = m.err;
    *k = e.on ? e.k : 0; *n_ref = uint32_t(e.ref.size());
    std::vector<uint32_t> all;
    for (const std::vector<uint32_t> *v : {&e.ref, &e.keep_thr, &e.row, &e.cb, &e.entries, &e.alts}) all.insert(all.end(), v->begin(), v->end());
    if (all.size() > cap) return flunk(SIMMR_E_INVALID, "buffer too small");
    if (!all.empty()) memcpy(words, all.data(), all.size() * 4);
    return int(all.size());
}
void simmr_debug_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }

}  // extern "C"

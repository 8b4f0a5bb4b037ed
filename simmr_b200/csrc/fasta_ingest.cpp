/* fasta_ingest.cpp — FASTA assemblies -> the flat genome set (GenomeStore), in parallel.
 *
 * Replaces Genome::from_fasta (simmr/src/genome.rs:89-162), which reads one file at a time through needletail 0.4.1:
 * record.id() is the whole header line, record.normalize(false) upper-cases a,c,g,t,n, turns u/U into T and . ~ into -,
 * drops white space, keeps A C G T N - and turns everything else (IUPAC codes included) into N; gzip, bzip2 and xz input
 * is detected by its magic bytes.  Here: every file of a call is read and normalised on its own thread, a large file is
 * cut at line starts into chunks that are normalised in parallel, and the pieces are copied into the flat buffer in
 * parallel too.  gzip goes through zlib; bzip2 and xz through libbz2 / liblzma opened at run time (the image carries the
 * shared objects but no headers), so a machine without them only loses those two formats. */
#include "host_model.hpp"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>

namespace simmr {
namespace {

struct Norm {          /* needletail sequence.rs normalize(seq, false); 0 = dropped */
    uint8_t map[256];
    Norm() {
        for (int i = 0; i < 256; ++i) map[i] = 'N';
        for (char c : {'A', 'C', 'G', 'T', 'N', '-'}) map[uint8_t(c)] = uint8_t(c);
        map['a'] = 'A'; map['c'] = 'C'; map['g'] = 'G'; map['t'] = 'T'; map['n'] = 'N';
        map['u'] = map['U'] = 'T'; map['.'] = map['~'] = '-';
        map[' '] = map['\t'] = map['\r'] = map['\n'] = 0;
    }
};
const Norm kNorm;

/* ---- whole file into memory, decompressed --------------------------------- */
bool read_file(const char *path, std::vector<uint8_t> &raw, std::string &err) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { err = std::string("cannot open ") + path; return false; }
    struct stat st;
    if (fstat(fd, &st) == 0 && st.st_size > 0) raw.reserve(size_t(st.st_size));
    std::vector<uint8_t> buf(size_t(4) << 20);
    for (;;) {
        const ssize_t got = read(fd, buf.data(), buf.size());
        if (got < 0) { err = std::string("read error in ") + path; close(fd); return false; }
        if (!got) break;
        raw.insert(raw.end(), buf.begin(), buf.begin() + got);
    }
    close(fd);
    return true;
}
bool gunzip(const std::vector<uint8_t> &in, std::vector<uint8_t> &out, std::string &err) {   /* concatenated members included */
    z_stream z{};
    if (inflateInit2(&z, 15 + 32) != Z_OK) { err = "zlib: inflateInit2 failed"; return false; }
    out.clear(); out.resize(std::max<size_t>(in.size() * 4, 1 << 16));
    z.next_in = const_cast<Bytef *>(in.data());
    size_t left = in.size(), done = 0, members = 0;
    for (;;) {
        if (!z.avail_in && left) { z.avail_in = uInt(std::min<size_t>(left, 1u << 30)); left -= z.avail_in; }
        if (done == out.size()) out.resize(out.size() * 2);
        const size_t room = std::min<size_t>(out.size() - done, 1u << 30);
        z.next_out = out.data() + done; z.avail_out = uInt(room);
        const int rc = inflate(&z, Z_NO_FLUSH);
        done += room - z.avail_out;
        if (rc == Z_STREAM_END) {
            ++members;
            if (!z.avail_in && !left) break;
            inflateReset(&z);                                       /* another member follows (next_in / avail_in stay) */
        } else if (rc == Z_BUF_ERROR && !z.avail_in && !left && z.avail_out) {
            inflateEnd(&z); err = "gzip data ends early"; return false;
        } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
            if (members) break;                                     /* bytes after the last member that are no gzip member: ignored, as gzip does */
            inflateEnd(&z); err = "gzip data is corrupt"; return false;
        }
    }
    inflateEnd(&z);
    out.resize(done);
    return true;
}
/* libbz2's public stream structure and entry points (bzlib.h 1.0.x) */
struct BzStream {
    char *next_in; unsigned avail_in, total_in_lo32, total_in_hi32;
    char *next_out; unsigned avail_out, total_out_lo32, total_out_hi32;
    void *state; void *(*bzalloc)(void *, int, int); void (*bzfree)(void *, void *); void *opaque;
};
bool bunzip2(const std::vector<uint8_t> &in, std::vector<uint8_t> &out, std::string &err) {
    void *lib = dlopen("libbz2.so.1.0", RTLD_NOW);
    if (!lib) lib = dlopen("libbz2.so.1", RTLD_NOW);
    if (!lib) { err = "bzip2 input needs libbz2.so.1.0 on this machine"; return false; }
    auto init = reinterpret_cast<int (*)(BzStream *, int, int)>(dlsym(lib, "BZ2_bzDecompressInit"));
    auto step = reinterpret_cast<int (*)(BzStream *)>(dlsym(lib, "BZ2_bzDecompress"));
    auto end = reinterpret_cast<int (*)(BzStream *)>(dlsym(lib, "BZ2_bzDecompressEnd"));
    if (!init || !step || !end) { err = "libbz2 lacks the decompression entry points"; return false; }
    out.clear(); out.resize(std::max<size_t>(in.size() * 5, 1 << 16));
    size_t done = 0, at = 0;
    while (at < in.size()) {                                         /* one iteration per stream of a multi-stream file */
        BzStream s{};
        if (init(&s, 0, 0) != 0) { err = "libbz2: init failed"; return false; }
        s.next_in = reinterpret_cast<char *>(const_cast<uint8_t *>(in.data() + at));
        size_t left = in.size() - at;
        s.avail_in = unsigned(std::min<size_t>(left, 1u << 30)); left -= s.avail_in;
        for (;;) {
            if (done == out.size()) out.resize(out.size() * 2);
            const size_t room = std::min<size_t>(out.size() - done, 1u << 30);
            s.next_out = reinterpret_cast<char *>(out.data() + done); s.avail_out = unsigned(room);
            const int rc = step(&s);
            done += room - s.avail_out;
            if (rc == 4 /* BZ_STREAM_END */) break;
            if (rc != 0 /* BZ_OK */) { end(&s); err = "bzip2 data is corrupt"; return false; }
            if (!s.avail_in) {
                if (!left) { if (s.avail_out) { end(&s); err = "bzip2 data ends early"; return false; } continue; }
                s.avail_in = unsigned(std::min<size_t>(left, 1u << 30)); left -= s.avail_in;
            }
        }
        at = size_t(reinterpret_cast<uint8_t *>(s.next_in) - in.data());
        end(&s);
    }
    out.resize(done);
    return true;
}
bool unxz(const std::vector<uint8_t> &in, std::vector<uint8_t> &out, std::string &err) {
    void *lib = dlopen("liblzma.so.5", RTLD_NOW);
    if (!lib) { err = "xz input needs liblzma.so.5 on this machine"; return false; }
    /* lzma_ret lzma_stream_buffer_decode(uint64_t *memlimit, uint32_t flags, const lzma_allocator *, const uint8_t *in,
     *                                     size_t *in_pos, size_t in_size, uint8_t *out, size_t *out_pos, size_t out_size) */
    auto decode = reinterpret_cast<int (*)(uint64_t *, uint32_t, const void *, const uint8_t *, size_t *, size_t, uint8_t *, size_t *, size_t)>(
        dlsym(lib, "lzma_stream_buffer_decode"));
    if (!decode) { err = "liblzma lacks lzma_stream_buffer_decode"; return false; }
    for (size_t cap = std::max<size_t>(in.size() * 6, 1 << 16);; cap *= 2) {
        out.clear(); out.resize(cap);
        uint64_t memlimit = UINT64_MAX; size_t in_pos = 0, out_pos = 0;
        const int rc = decode(&memlimit, 0x08 /* LZMA_CONCATENATED */, nullptr, in.data(), &in_pos, in.size(), out.data(), &out_pos, cap);
        if (rc == 0 /* LZMA_OK */) { out.resize(out_pos); return true; }
        if (rc != 10 /* LZMA_BUF_ERROR: output buffer too small */ || cap > std::max<size_t>(in.size() * 2000, size_t(1) << 32)) { err = "xz data is corrupt"; return false; }
    }
}
bool slurp(const char *path, std::vector<uint8_t> &data, std::string &err) {
    std::vector<uint8_t> raw;
    if (!read_file(path, raw, err)) return false;
    bool ok = true;
    if (raw.size() >= 2 && raw[0] == 0x1f && raw[1] == 0x8b) ok = gunzip(raw, data, err);
    else if (raw.size() >= 3 && raw[0] == 'B' && raw[1] == 'Z' && raw[2] == 'h') ok = bunzip2(raw, data, err);
    else if (raw.size() >= 6 && raw[0] == 0xFD && !memcmp(raw.data() + 1, "7zXZ", 4)) ok = unxz(raw, data, err);
    else data.swap(raw);
    if (!ok) err = std::string(path) + ": " + err;
    return ok;
}

/* ---- one chunk of a file: lines [b, e) where b is a line start -------------- */
struct Header { std::string id; size_t at; };        /* `at`: normalised bytes of the chunk before this record */
struct Chunk { size_t b = 0, e = 0; std::vector<uint8_t> seq; std::vector<Header> headers; };
void scan_chunk(const std::vector<uint8_t> &data, Chunk &c) {
    const uint8_t *p = data.data();
    c.seq.resize(c.e - c.b);
    uint8_t *out = c.seq.data();
    size_t k = 0;
    for (size_t i = c.b; i < c.e;) {
        const uint8_t *nl = static_cast<const uint8_t *>(memchr(p + i, '\n', c.e - i));
        const size_t eol = nl ? size_t(nl - p) : c.e;
        if (p[i] == '>') {
            size_t idend = eol; if (idend > i && p[idend - 1] == '\r') --idend;
            c.headers.push_back(Header{std::string(reinterpret_cast<const char *>(p + i + 1), idend - i - 1), k});
        } else {
            for (size_t j = i; j < eol; ++j) { const uint8_t v = kNorm.map[p[j]]; out[k] = v; k += v != 0; }
        }
        i = eol + 1;
    }
    c.seq.resize(k);
}

struct Parsed {        /* one file: records as spans of the chunks' normalised bytes, in order */
    std::vector<Chunk> chunks;
    struct Rec { std::string id; uint64_t len; };
    std::vector<Rec> recs;
    uint64_t total = 0;
    std::string err;
    int rc = SIMMR_OK;
};
void parse_file(const char *path, unsigned inner_threads, Parsed &out) {
    std::vector<uint8_t> data;
    if (!slurp(path, data, out.err)) { out.rc = SIMMR_E_IO; return; }
    size_t i = 0; const size_t n = data.size();
    while (i < n && (data[i] == '\n' || data[i] == '\r')) ++i;
    if (i < n && data[i] != '>') { out.err = std::string(path) + ": not FASTA (a record must start with '>')"; out.rc = SIMMR_E_PARSE; return; }
    /* chunks of about 32 MB (SIMMR_FASTA_CHUNK bytes, for the tests) that start on a line start */
    const char *env = getenv("SIMMR_FASTA_CHUNK");
    const size_t chunk = env && atoll(env) > 0 ? size_t(atoll(env)) : size_t(32) << 20;
    const size_t want = std::max<size_t>(1, std::min<size_t>(inner_threads, (n - i + chunk - 1) / chunk));
    std::vector<size_t> cut{i};
    for (size_t c = 1; c < want; ++c) {
        size_t at = i + (n - i) * c / want;
        const uint8_t *nl = static_cast<const uint8_t *>(memchr(data.data() + at, '\n', n - at));
        at = nl ? size_t(nl - data.data()) + 1 : n;
        if (at > cut.back() && at < n) cut.push_back(at);
    }
    cut.push_back(n);
    out.chunks.resize(cut.size() - 1);
    for (size_t c = 0; c + 1 < cut.size(); ++c) { out.chunks[c].b = cut[c]; out.chunks[c].e = cut[c + 1]; }
    if (out.chunks.size() == 1) scan_chunk(data, out.chunks[0]);
    else {
        std::vector<std::thread> th;
        for (Chunk &c : out.chunks) th.emplace_back([&data, &c] { scan_chunk(data, c); });
        for (auto &t : th) t.join();
    }
    /* records: a header opens one; bytes before a chunk's first header belong to the record still open */
    for (const Chunk &c : out.chunks) {
        size_t prev = 0;
        for (const Header &h : c.headers) {
            if (!out.recs.empty()) out.recs.back().len += h.at - prev;
            out.recs.push_back(Parsed::Rec{h.id, 0});
            prev = h.at;
        }
        if (!out.recs.empty()) out.recs.back().len += c.seq.size() - prev;
        out.total += c.seq.size();
    }
}

std::string random_uuid() {   /* util.rs:124-129: top 64 bits of a v4 UUID, `{:x}` */
    std::random_device rd;
    uint64_t v = (uint64_t(rd()) << 32) | rd();
    v = (v & ~0xF000ull) | 0x4000ull;
    char b[32]; snprintf(b, sizeof b, "%llx", (unsigned long long)v);
    return b;
}
}  // namespace

int GenomeStore::add_fastas(const char *const *file, const char *const *uuid, uint32_t n, bool contiguous, unsigned threads) {
    if (!threads) threads = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    std::vector<Parsed> parsed(n);
    {   /* files in parallel; what is left of the thread budget goes to the chunks inside a file */
        const unsigned outer = std::max(1u, std::min<unsigned>(threads, n)), inner = std::max(1u, threads / outer);
        std::atomic<uint32_t> next{0};
        std::vector<std::thread> th;
        for (unsigned t = 0; t < outer; ++t) th.emplace_back([&] { for (uint32_t f; (f = next++) < n;) parse_file(file[f], inner, parsed[f]); });
        for (auto &t : th) t.join();
    }
    for (uint32_t f = 0; f < n; ++f) if (parsed[f].rc) { err = parsed[f].err; return parsed[f].rc; }
    /* layout, then the copies in parallel */
    struct Piece { const std::vector<uint8_t> *src; uint64_t from, len, at; };
    std::vector<Piece> pieces;
    std::vector<uint64_t> joints;                                    /* where the 'N' after each contig of a --contiguous genome goes */
    uint64_t end = bases.size();
    for (uint32_t f = 0; f < n; ++f) {
        Parsed &p = parsed[f];
        if (contiguous) {   /* genome.rs:117-136: one Seq "whole genome", every contig followed by 'N', size without the Ns */
            size_t ci = 0; uint64_t used = 0;
            for (const Parsed::Rec &r : p.recs) {
                for (uint64_t left = r.len; left;) {                 /* the record's bytes, chunk by chunk */
                    while (used == p.chunks[ci].seq.size()) { ++ci; used = 0; }
                    const uint64_t take = std::min<uint64_t>(left, p.chunks[ci].seq.size() - used);
                    pieces.push_back(Piece{&p.chunks[ci].seq, used, take, end});
                    end += take; used += take; left -= take;
                }
                joints.push_back(end++);
            }
            contig_off.push_back(end); contig_size.push_back(p.total);
            contig_ids += "whole genome"; contig_id_off.push_back(uint32_t(contig_ids.size()));
        } else {
            uint64_t at = end;
            for (const Chunk &c : p.chunks) { pieces.push_back(Piece{&c.seq, 0, c.seq.size(), end}); end += c.seq.size(); }
            for (const Parsed::Rec &r : p.recs) {
                at += r.len;
                contig_off.push_back(at); contig_size.push_back(r.len);
                contig_ids += r.id; contig_id_off.push_back(uint32_t(contig_ids.size()));
            }
        }
        genome_contig.push_back(uint32_t(contig_size.size()));
        genome_ids += uuid && uuid[f] ? std::string(uuid[f]) : random_uuid();
        genome_id_off.push_back(uint32_t(genome_ids.size()));
        genome_size.push_back(p.total);
        paths.push_back(file[f]);
    }
    bases.grow(end);
    uint8_t *dst = bases.data();
    for (uint64_t j : joints) dst[j] = 'N';
    {
        std::atomic<size_t> next{0};
        std::vector<std::thread> th;
        for (unsigned t = 0; t < std::max(1u, std::min<unsigned>(threads, unsigned(pieces.size()))); ++t) th.emplace_back([&] {
            for (size_t i; (i = next++) < pieces.size();) if (pieces[i].len) memcpy(dst + pieces[i].at, pieces[i].src->data() + pieces[i].from, pieces[i].len);
        });
        for (auto &t : th) t.join();
    }
    return SIMMR_OK;
}

int GenomeStore::add_fasta(const char *path, const char *uuid, bool contiguous) { return add_fastas(&path, &uuid, 1, contiguous, 0); }

/* main.rs:117-162: contigs with size <= min_size are dropped, then genomes left without contigs; in place (nothing moves
 * when nothing is dropped) */
void GenomeStore::filter(uint16_t min_size, uint8_t *kept) {
    const uint32_t ng = uint32_t(genome_contig.size()) - 1;
    std::vector<uint64_t> n_off{0}, n_size, n_gsize; std::vector<uint32_t> n_gc{0}, n_cid{0}, n_gid{0};
    std::string n_cids, n_gids; std::vector<std::string> n_paths;
    uint8_t *p = bases.data();
    uint64_t w = 0;
    for (uint32_t g = 0; g < ng; ++g) {
        const size_t before = n_size.size();
        for (uint32_t c = genome_contig[g]; c < genome_contig[g + 1]; ++c) {
            if (contig_size[c] <= min_size) continue;                 /* main.rs:124 */
            const uint64_t len = contig_off[c + 1] - contig_off[c];
            if (w != contig_off[c]) memmove(p + w, p + contig_off[c], len);
            w += len;
            n_off.push_back(w); n_size.push_back(contig_size[c]);
            n_cids.append(contig_ids, contig_id_off[c], contig_id_off[c + 1] - contig_id_off[c]);
            n_cid.push_back(uint32_t(n_cids.size()));
        }
        const bool keep = n_size.size() > before;                     /* main.rs:144-156 */
        if (kept) kept[g] = keep;
        if (!keep) continue;
        n_gc.push_back(uint32_t(n_size.size()));
        n_gids.append(genome_ids, genome_id_off[g], genome_id_off[g + 1] - genome_id_off[g]);
        n_gid.push_back(uint32_t(n_gids.size()));
        n_gsize.push_back(genome_size[g]);                            /* Genome.size is not updated by the filter (main.rs:158-161) */
        n_paths.push_back(paths[g]);
    }
    bases.shrink(w);
    contig_off.swap(n_off); contig_size.swap(n_size); genome_contig.swap(n_gc); contig_ids.swap(n_cids); contig_id_off.swap(n_cid);
    genome_ids.swap(n_gids); genome_id_off.swap(n_gid); genome_size.swap(n_gsize); paths.swap(n_paths);
}

void GenomeStore::view(simmr_genome_set &v) const {
    v.n_genomes = uint32_t(genome_contig.size()) - 1; v.n_contigs = uint32_t(contig_size.size());
    v.bases = bases.data(); v.contig_off = contig_off.data(); v.contig_size = contig_size.data();
    v.genome_contig = genome_contig.data(); v.contig_ids = contig_ids.data(); v.contig_id_off = contig_id_off.data();
    v.genome_ids = genome_ids.data(); v.genome_id_off = genome_id_off.data();
}

}  // namespace simmr

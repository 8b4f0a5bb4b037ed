/* simmr_main.cpp — the `simmr` command line on top of libsimmr_b200.so.
 *
 * Mirrors the reference binary (simmr/src/main.rs:20-266, flags of cli.rs:93-220): same
 * flag names, defaults and kebab-case enum values, same outputs (<output> and
 * <output>.tsv), errors to stderr and exit code 1.  Everything goes through the C ABI of
 * include/simmr_b200.h; nothing here computes reads.  Added: --gpus N (the run's batches are
 * dealt to N devices of this box through simmr_multi, byte-identical output), --output-parts P
 * (P part files <output>.000 ... written by P I/O threads: one file takes about 2.5 GB/s,
 * whatever feeds it), --gzip [--gzip-level L] (the output as concatenated gzip members, the slices of a chunk compressed
 * on all host threads), --split-mates (two files, <stem>_R1<ext> and <stem>_R2<ext>, instead of the interleaved one:
 * readme.md:209-211 lists that as to-do), --compat-long-start (simulate.rs:484 start draw) and the `custom-long`
 * profile value. */
#include "../../include/simmr_b200.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Args {
    std::vector<std::string> genome;                 /* cli.rs:95-100  */
    std::string genome_file, output, custom_profile; /* :103-112, :178-183 */
    uint64_t num_reads = 1000;                       /* :115-121 */
    uint16_t read_length = 150, insert_size = 150;   /* :124-130, :141-147 */
    double read_length_std = 10.0;                   /* :132-138 (unused by the profiles, as in the reference) */
    uint8_t mean_phred = 30;                         /* :150-156 */
    std::string error_profile = "perfect-short";     /* :159-166 */
    std::string abundance_profile = "uniform";       /* :169-176 */
    std::string header_format;                       /* :194-200, empty = default */
    bool has_seed = false; uint64_t seed = 0;        /* :202-203 */
    bool size_adjusted = false, contiguous = false;  /* :206-219 */
    bool compat_long_start = false;
    int gpus = 1, parts = 1, gzip_level = 1;
    bool gzip = false, split_mates = false;
    uint64_t batch_units = 0;                        /* 0: library default (about 256 MiB of FASTQ per chunk) */
};

[[noreturn]] void die(const std::string &m) { fprintf(stderr, "ERROR %s\n", m.c_str()); exit(1); }
void info(const std::string &m) { fprintf(stderr, " INFO %s\n", m.c_str()); }
void warn(const std::string &m) { fprintf(stderr, " WARN %s\n", m.c_str()); }

void usage() {
    fputs("simmr (B200) — metagenomic read simulator\n\n"
          "USAGE: simmr [OPTIONS] --output <OUTPUT> <--genome <GENOME>...|--genome-file <GENOME_FILE>>\n\n"
          "  --genome <path>            FASTA assembly (repeatable)\n"
          "  --genome-file <path>       list of paths, or TSV with path|filepath, uuid|id|genome_id, abundance\n"
          "  --output <path>            FASTQ output; metadata goes to <output>.tsv\n"
          "  --num-reads <n>            [1000] total reads (uniform/custom) or reads per genome (exact)\n"
          "  --read-length <n>          [150]\n  --read-length-std <x>      [10]\n  --insert-size <n>          [150]\n"
          "  --mean-phred-score <n>     [30]\n"
          "  --error-profile <p>        [perfect-short] minimal-short, minimal-long, perfect-short, perfect-long, custom-short, custom-long\n"
          "  --abundance-profile <p>    [uniform] exact, uniform, custom\n"
          "  --custom-profile <path>    simmrd profile blob\n  --with-ani <n>             accepted, not implemented (as in the reference)\n"
          "  --read-header-format <s>   placeholders {:read_id:} {:genome_id:} {:sequence_id:} {:start_position:} {:end_position:} {:reverse_complement:} {:pair:}\n"
          "  --seed <n>\n  --size-adjusted\n  --contiguous\n  --gpus <n>                 [1] devices to share the run\n"
          "  --output-parts <n>         [1] write <output>.000 ... <output>.(n-1) instead of <output>: chunk j of the run goes to part j mod n\n"
          "  --gzip                     write gzip (concatenated members, compressed on all host threads); --gzip-level <1-9> [1]\n"
          "  --split-mates              paired-end reads into <stem>_R1<ext> and <stem>_R2<ext> instead of one interleaved file\n"
          "  --batch-units <n>          units (pairs / long reads) per chunk; default: about 256 MiB of FASTQ\n"
          "  --compat-long-start\n", stderr);
}

Args parse(int argc, char **argv) {
    Args a;
    auto need = [&](int &i) -> std::string { if (i + 1 >= argc) die(std::string("missing value for ") + argv[i]); return argv[++i]; };
    auto num = [&](const std::string &s, const char *flag, double lo, double hi) {
        char *end = nullptr; const double v = strtod(s.c_str(), &end);
        if (s.empty() || *end || v < lo || v > hi || v != std::floor(v)) die(std::string("invalid value '") + s + "' for " + flag);
        return v;
    };
    for (int i = 1; i < argc; ++i) {
        std::string f = argv[i], val; bool has_val = false;
        const size_t eq = f.find('=');
        if (f.rfind("--", 0) == 0 && eq != std::string::npos) { val = f.substr(eq + 1); f = f.substr(0, eq); has_val = true; }
        auto v = [&]() { return has_val ? val : need(i); };
        if (f == "--genome") a.genome.push_back(v());
        else if (f == "--genome-file") a.genome_file = v();
        else if (f == "--output") a.output = v();
        else if (f == "--num-reads") a.num_reads = uint64_t(num(v(), "--num-reads", 0, 1.8e19));
        else if (f == "--read-length") a.read_length = uint16_t(num(v(), "--read-length", 0, 65535));
        else if (f == "--read-length-std") a.read_length_std = strtod(v().c_str(), nullptr);
        else if (f == "--insert-size") a.insert_size = uint16_t(num(v(), "--insert-size", 0, 65535));
        else if (f == "--mean-phred-score") a.mean_phred = uint8_t(num(v(), "--mean-phred-score", 0, 255));
        else if (f == "--error-profile") a.error_profile = v();
        else if (f == "--abundance-profile") a.abundance_profile = v();
        else if (f == "--custom-profile") a.custom_profile = v();
        else if (f == "--with-ani") (void)v();
        else if (f == "--read-header-format") a.header_format = v();
        else if (f == "--seed") { a.seed = strtoull(v().c_str(), nullptr, 10); a.has_seed = true; }
        else if (f == "--size-adjusted") a.size_adjusted = true;
        else if (f == "--contiguous") a.contiguous = true;
        else if (f == "--compat-long-start") a.compat_long_start = true;
        else if (f == "--gpus") a.gpus = int(num(v(), "--gpus", 1, 64));
        else if (f == "--output-parts") a.parts = int(num(v(), "--output-parts", 1, 64));
        else if (f == "--gzip") a.gzip = true;
        else if (f == "--gzip-level") a.gzip_level = int(num(v(), "--gzip-level", 1, 9));
        else if (f == "--split-mates") a.split_mates = true;
        else if (f == "--batch-units") a.batch_units = uint64_t(num(v(), "--batch-units", 1, 2147483647));
        else if (f == "-h" || f == "--help") { usage(); exit(0); }
        else die("unexpected argument '" + f + "'");
    }
    if (a.output.empty()) die("the following required arguments were not provided: --output <OUTPUT>");
    if (a.genome.empty() && a.genome_file.empty()) die("the following required arguments were not provided: <--genome <GENOME>|--genome-file <GENOME_FILE>>");
    if (!a.genome.empty() && !a.genome_file.empty()) die("the argument '--genome' cannot be used with '--genome-file'");   /* cli.rs:88-92 */
    return a;
}

/* The run's chunks arrive in file order on the library's writer thread.  One part: written right there.  Several parts:
 * chunk j goes to part j mod P, each part has its own I/O thread (a single file takes about 2.5 GB/s whatever feeds it),
 * and the chunk's pinned buffer is kept (SIMMR_SINK_RETAINED) until that thread has written it. */
struct PartWriter {
    struct Part {
        int fd = -1, fd2 = -1; std::thread th; std::mutex m; std::condition_variable cv;
        std::deque<std::pair<const uint8_t *, uint64_t>> q; bool closed = false;
    };
    std::vector<Part> parts;
    std::atomic<bool> failed{false};
    uint64_t next = 0;
    bool gz = false, split = false; int level = 1;
    unsigned workers = std::max(1u, std::thread::hardware_concurrency());
    /* [p, p + n) to fd: as it is, or as gzip members — one per slice of about 4 MiB, the slices compressed in parallel */
    bool emit(int fd, const uint8_t *p, uint64_t n) const {
        if (!gz) return write_all(fd, p, n);
        const uint64_t nsl = std::max<uint64_t>(1, std::min<uint64_t>(workers / std::max<size_t>(1, parts.size()) + 1, (n + (4u << 20) - 1) / (4u << 20)));
        std::vector<std::vector<uint8_t>> out(nsl);
        std::vector<int64_t> got(nsl, 0);
        std::vector<std::thread> th;
        for (uint64_t k = 0; k < nsl; ++k) th.emplace_back([&, k] {
            const uint64_t lo = n * k / nsl, hi = n * (k + 1) / nsl;
            out[k].resize(simmr_gzip_bound(hi - lo));
            got[k] = simmr_gzip_member(p + lo, hi - lo, out[k].data(), out[k].size(), level);
        });
        for (auto &t : th) t.join();
        for (uint64_t k = 0; k < nsl; ++k) if (got[k] < 0 || !write_all(fd, out[k].data(), uint64_t(got[k]))) return false;
        return true;
    }
    /* one chunk of the run to its part: split into mates and / or compressed on the way */
    bool process(Part &p, const uint8_t *data, uint64_t len) const {
        if (!split) return emit(p.fd, data, len);
        std::vector<uint8_t> r1(len), r2(len);
        uint64_t n1 = 0, n2 = 0;
        if (simmr_fastq_split_mates(data, len, r1.data(), &n1, r2.data(), &n2)) return false;
        return emit(p.fd, r1.data(), n1) && emit(p.fd2, r2.data(), n2);
    }
    bool threaded() const { return parts.size() > 1 || gz || split; }      /* chunks are kept and handled by the parts' own threads */
    static bool write_all(int fd, const uint8_t *p, uint64_t n) {
        while (n) { const ssize_t w = write(fd, p, n); if (w <= 0) return false; p += w; n -= uint64_t(w); }
        return true;
    }
    explicit PartWriter(size_t n) : parts(n) {}
    void start() {
        if (!threaded()) return;
        for (Part &p : parts) p.th = std::thread([this, &p] {
            for (;;) {
                std::pair<const uint8_t *, uint64_t> c;
                { std::unique_lock<std::mutex> l(p.m); p.cv.wait(l, [&] { return !p.q.empty() || p.closed; }); if (p.q.empty()) return; c = p.q.front(); p.q.pop_front(); }
                if (!failed && !process(p, c.first, c.second)) failed = true;
                simmr_chunk_release(c.first);
            }
        });
    }
    void finish() {
        for (Part &p : parts) { { std::lock_guard<std::mutex> l(p.m); p.closed = true; } p.cv.notify_all(); if (p.th.joinable()) p.th.join(); if (p.fd >= 0) close(p.fd); if (p.fd2 >= 0) close(p.fd2); }
    }
    static int call(void *user, const uint8_t *data, uint64_t len, uint64_t, uint64_t) {
        PartWriter *w = static_cast<PartWriter *>(user);
        if (w->failed) return 1;
        if (!w->threaded()) { if (!write_all(w->parts[0].fd, data, len)) { w->failed = true; return 1; } return 0; }
        Part &p = w->parts[w->next++ % w->parts.size()];
        { std::lock_guard<std::mutex> l(p.m); p.q.emplace_back(data, len); }
        p.cv.notify_all();
        return SIMMR_SINK_RETAINED;
    }
};

}  // namespace

int main(int argc, char **argv) {
    const Args args = parse(argc, argv);

    /* cli.rs:229-301 determine_error_profile */
    simmr_eprofile_desc ep{};
    ep.read_length = args.read_length; ep.insert_size = args.insert_size; ep.mean_phred = args.mean_phred;
    ep.flags = args.compat_long_start ? SIMMR_FLAG_COMPAT_LONG_START : 0;
    std::string blob;
    const std::string &e = args.error_profile;
    if (e == "perfect-short") ep.kind = SIMMR_EP_PERFECT_SHORT;
    else if (e == "minimal-short") ep.kind = SIMMR_EP_MINIMAL_SHORT;
    else if (e == "perfect-long") ep.kind = SIMMR_EP_PERFECT_LONG;
    else if (e == "minimal-long") ep.kind = SIMMR_EP_MINIMAL_LONG;
    else if (e == "custom-short" || e == "custom-long") {
        ep.kind = e == "custom-short" ? SIMMR_EP_CUSTOM_SHORT : SIMMR_EP_CUSTOM_LONG;
        if (args.custom_profile.empty()) die("a custom error profile needs --custom-profile");
        std::ifstream f(args.custom_profile, std::ios::binary);
        if (!f) { fprintf(stderr, "Error parsing custom error profile: cannot open %s\n", args.custom_profile.c_str()); return 1; }
        std::stringstream ss; ss << f.rdbuf(); blob = ss.str();
        ep.blob = reinterpret_cast<const uint8_t *>(blob.data()); ep.blob_len = blob.size();
    } else die("invalid value '" + e + "' for '--error-profile <ERROR_PROFILE>'");
    int is_long = 0; uint16_t min_size = 0;
    if (simmr_eprofile_is_long_read(&ep, &is_long) || simmr_eprofile_minimum_genome_size(&ep, &min_size)) {
        fprintf(stderr, "Error parsing custom error profile: %s\n", simmr_last_error(nullptr)); return 1;
    }
    if (ep.kind == SIMMR_EP_CUSTOM_SHORT && is_long)    /* main.rs:30-33 */
        die("You specified a custom short-read error profile but the provided error profile is for long reads");
    int ap_kind;
    if (args.abundance_profile == "exact") ap_kind = SIMMR_AP_EXACT;
    else if (args.abundance_profile == "uniform") ap_kind = SIMMR_AP_UNIFORM;
    else if (args.abundance_profile == "custom") ap_kind = SIMMR_AP_CUSTOM;
    else die("invalid value '" + args.abundance_profile + "' for '--abundance-profile <ABUNDANCE_PROFILE>'");

    info("Loading genomes");
    std::vector<std::string> paths, uuids; std::vector<bool> has_uuid; std::vector<double> abund;
    if (!args.genome_file.empty()) {                    /* main.rs:38-93 */
        std::vector<char> buf(size_t(64) << 20);
        const int64_t n = simmr_parse_genome_file(args.genome_file.c_str(), buf.data(), buf.size());
        if (n < 0) die(std::string("Failed to read genome file: ") + simmr_last_error(nullptr));
        std::istringstream in(buf.data()); std::string line;
        while (std::getline(in, line)) {
            const size_t t1 = line.find('\t'), t2 = line.find('\t', t1 + 1);
            paths.push_back(line.substr(0, t1)); uuids.push_back(line.substr(t1 + 1, t2 - t1 - 1)); has_uuid.push_back(!uuids.back().empty());
            const std::string ab = line.substr(t2 + 1);
            abund.push_back(ab == "nan" ? NAN : strtod(ab.c_str(), nullptr));
        }
        for (const auto &p : paths) { struct stat st; if (stat(p.c_str(), &st)) die("Genome (" + p + ") does not exist"); }
    } else {
        for (const auto &g : args.genome) { paths.push_back(g); uuids.emplace_back(); has_uuid.push_back(false); abund.push_back(NAN); }
    }
    simmr_genome_builder *gb = nullptr;
    if (simmr_builder_create(&gb)) die("out of memory");
    for (size_t i = 0; i < paths.size(); ++i)
        if (ap_kind == SIMMR_AP_CUSTOM && std::isnan(abund[i]))    /* main.rs:80-86 */
            die("You used a custom abundance profile but didn't provide abundances for genome " + paths[i]);
    {   /* every assembly is read, decompressed and normalised in parallel (genome.rs:89-162 does them one by one) */
        std::vector<const char *> pp(paths.size()), up(paths.size());
        for (size_t i = 0; i < paths.size(); ++i) { pp[i] = paths[i].c_str(); up[i] = has_uuid[i] ? uuids[i].c_str() : nullptr; }
        if (simmr_builder_add_fastas(gb, pp.data(), up.data(), uint32_t(paths.size()), args.contiguous, 0))
            die(std::string("Failed to parse a genome: ") + simmr_builder_error(gb));
    }
    std::vector<uint64_t> all_sizes(paths.size());
    for (size_t i = 0; i < paths.size(); ++i) all_sizes[i] = simmr_builder_genome_size(gb, uint32_t(i));

    info("Ensuring genomes meet minimum sequence length requirements for simulation");
    std::vector<uint8_t> kept(paths.size(), 1);
    if (!args.contiguous) simmr_builder_filter(gb, min_size, kept.data());   /* main.rs:117-162 */
    std::vector<std::string> k_paths; std::vector<double> k_abund; std::vector<uint64_t> k_sizes;
    for (size_t i = 0; i < paths.size(); ++i) {
        if (kept[i]) { k_paths.push_back(paths[i]); k_abund.push_back(abund[i]); k_sizes.push_back(all_sizes[i]); }
        else warn("Removing " + paths[i] + " from simulation, it doesn't have usable sequences");
    }
    simmr_genome_set gs{};
    simmr_builder_view(gb, &gs);
    if (!gs.n_genomes) die("no genome has a sequence longer than the minimum size " + std::to_string(min_size));

    /* abundance_profiles/: reads per genome, bit-identical rounding */
    simmr_aprofile_desc ap{ap_kind, args.size_adjusted ? 1 : 0, ap_kind == SIMMR_AP_CUSTOM ? k_abund.data() : nullptr};
    std::vector<uint64_t> reads(gs.n_genomes); std::vector<double> ab_out(gs.n_genomes);
    if (simmr_abundance_counts(&ap, args.num_reads, gs.n_genomes, k_sizes.data(), reads.data(), ab_out.data())) die(simmr_last_error(nullptr));

    if (ep.kind == SIMMR_EP_CUSTOM_LONG) {
        uint32_t ek = 0, en = 0;
        std::vector<uint32_t> words(size_t(1) << 24);
        if (simmr_debug_err_table(&ep, &ek, &en, words.data(), words.size()) >= 0 && ek == 0)
            warn("the custom long-read profile has no alt-k-mer table: reads carry its quality scores but no sequencing errors");
    }
    info(is_long ? "Simulating long reads" : "Simulating short reads");
    info("Writing simulated reads to " + args.output);
    unlink(args.output.c_str());                        /* main.rs:191-198 */
    unlink((args.output + ".tsv").c_str());
    if (args.split_mates && is_long) die("--split-mates needs a paired-end error profile");
    PartWriter writer{size_t(args.parts)};
    writer.gz = args.gzip; writer.level = args.gzip_level; writer.split = args.split_mates;
    auto mate_path = [&](const std::string &path, const char *tag) {      /* reads.fq.gz -> reads_R1.fq.gz */
        size_t cut = path.size();
        for (const char *ext : {".gz", ".fastq", ".fq"}) { const size_t n = strlen(ext); if (cut >= n && path.compare(cut - n, n, ext) == 0) cut -= n; }
        const size_t slash = path.rfind('/');
        if (slash != std::string::npos && cut <= slash) cut = path.size();
        return path.substr(0, cut) + tag + path.substr(cut);
    };
    for (int p = 0; p < args.parts; ++p) {
        char suffix[16]; snprintf(suffix, sizeof suffix, ".%03d", p);
        const std::string path = args.parts > 1 ? args.output + suffix : args.output;
        const std::string first = args.split_mates ? mate_path(path, "_R1") : path;
        writer.parts[p].fd = open(first.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (writer.parts[p].fd < 0) die("Failed to write reads to the output file: cannot create " + first);
        if (args.split_mates) {
            const std::string second = mate_path(path, "_R2");
            writer.parts[p].fd2 = open(second.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
            if (writer.parts[p].fd2 < 0) die("Failed to write reads to the output file: cannot create " + second);
        }
    }

    simmr_run_opts opts{};
    opts.has_seed = 1;
    if (args.has_seed) opts.seed = args.seed;
    else { std::ifstream r("/dev/urandom", std::ios::binary); r.read(reinterpret_cast<char *>(&opts.seed), 8); }   /* from_entropy, simulate.rs:174 */
    opts.header_format = args.header_format.empty() ? nullptr : args.header_format.c_str();
    opts.batch_units = args.batch_units;

    /* one context and launch thread per GPU inside the library, batches dealt round-robin, one writer thread calling
     * PartWriter::call in file order (genomes replicated, no device collective) */
    std::vector<int> devices(args.gpus);
    for (int k = 0; k < args.gpus; ++k) devices[k] = k;
    simmr_multi *multi = nullptr;
    if (simmr_multi_create(devices.data(), args.gpus, &multi)) die(simmr_multi_last_error(nullptr));
    if (simmr_multi_genomes_load(multi, &gs) || simmr_multi_error_profile_set(multi, &ep)) die(simmr_multi_last_error(multi));
    writer.start();
    const auto t_run = std::chrono::steady_clock::now();
    const int run_rc = simmr_multi_simulate(multi, reads.data(), &opts, &PartWriter::call, &writer);
    writer.finish();
    {
        simmr_stats rs{};
        if (!run_rc && !simmr_multi_stats_get(multi, &rs)) {
            const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_run).count();
            char line[200]; snprintf(line, sizeof line, "Wrote %llu FASTQ bytes (%llu records, %llu bases) in %.3f s (%.2f GB/s)", (unsigned long long)rs.bytes,
                                     (unsigned long long)rs.records, (unsigned long long)rs.bases, secs, double(rs.bytes) / secs / 1e9);
            info(line);
        }
    }
    if (writer.failed) die("Failed to write reads to the output file");
    if (run_rc) die(simmr_multi_last_error(multi));
    simmr_multi_destroy(multi);

    info("Writing simulation metadata to " + args.output + ".tsv");
    std::vector<std::string> ids(gs.n_genomes);
    std::vector<const char *> idp(gs.n_genomes), pp(gs.n_genomes);
    for (uint32_t g = 0; g < gs.n_genomes; ++g) {
        ids[g].assign(gs.genome_ids + gs.genome_id_off[g], gs.genome_id_off[g + 1] - gs.genome_id_off[g]);
        idp[g] = ids[g].c_str(); pp[g] = k_paths[g].c_str();
    }
    if (simmr_write_metadata((args.output + ".tsv").c_str(), gs.n_genomes, idp.data(), pp.data(), reads.data(), ab_out.data()))
        fprintf(stderr, "ERROR Failed to write metadata file: %s\n", simmr_last_error(nullptr));
    simmr_builder_destroy(gb);
    return 0;
}

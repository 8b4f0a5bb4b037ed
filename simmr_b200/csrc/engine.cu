/* engine.cu — context, device residency and the batch pipeline behind the
 * compute entry points of include/simmr_b200.h.
 *
 * Replaces the simulate + write phase of the reference's run_main
 * (simmr/src/main.rs:180-206): simulate::simulate_pe_reads /
 * simulate_long_reads (simulate.rs:110-150, 323-406) and
 * fastq::write_to_fastq (fastq.rs:14-124).  There is no CPU fallback. */
#include "device_code.cuh"
#include "host_model.hpp"

#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

using namespace simmr;

namespace {

struct DevMem {    /* owning device allocation */
    void *p = nullptr; size_t bytes = 0;
    DevMem() = default;
    DevMem(const DevMem &) = delete;
    DevMem &operator=(const DevMem &) = delete;
    ~DevMem() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    cudaError_t reserve(size_t n) {
        if (n <= bytes) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, n ? n : 1);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    template <class T> cudaError_t upload(const T *src, size_t count, cudaStream_t s) {
        cudaError_t e = reserve(count * sizeof(T) + 16);
        if (e != cudaSuccess) return e;
        return count ? cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s) : cudaSuccess;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
};
/* NUMA placement, best effort (a container may forbid either call; then nothing changes): pinned buffers are allocated
 * with the calling thread's memory policy preferring the GPU's node, and the threads the library owns run on that node's
 * cores — eight GPUs copying 50 GB/s each into one socket's memory was the end-to-end limit of round 1. */
int numa_node_of_device(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return -1;
    for (char *c = bus; *c; ++c) if (*c >= 'A' && *c <= 'F') *c = char(*c - 'A' + 'a');
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}
struct PreferNode {          /* scope: memory of this thread comes from `node` if possible (MPOL_PREFERRED = 1) */
    bool on = false;
    explicit PreferNode(int node) {
        if (node < 0 || node >= 1024) return;
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        on = syscall(SYS_set_mempolicy, 1, mask, 1024ul + 1) == 0;
    }
    ~PreferNode() { if (on) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul); }
};
void run_on_node(int node) {  /* the calling (library-owned) thread onto the cores of `node` */
    if (node < 0) return;
    FILE *f = fopen(("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist").c_str(), "r");
    if (!f) return;
    cpu_set_t set; CPU_ZERO(&set);
    int a, b, any = 0;
    while (fscanf(f, "%d", &a) == 1) {
        b = a;
        int c = fgetc(f);
        if (c == '-') { if (fscanf(f, "%d", &b) != 1) break; c = fgetc(f); }
        for (int i = a; i <= b && i < CPU_SETSIZE; ++i) { CPU_SET(i, &set); any = 1; }
        if (c != ',') break;
    }
    fclose(f);
    if (any) sched_setaffinity(0, sizeof set, &set);
}
struct PinnedMem {
    void *p = nullptr; size_t bytes = 0;
    ~PinnedMem() { if (p) cudaFreeHost(p); }
    cudaError_t reserve(size_t n, int node = -1) {
        if (n <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        PreferNode prefer(node);
        cudaError_t e = cudaHostAlloc(&p, n, cudaHostAllocDefault);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
};
}  // namespace

struct simmr_ctx {
    int device = 0, numa_node = -1;
    cudaDeviceProp prop{};
    cudaStream_t compute = nullptr, copy = nullptr, prep = nullptr;   /* format kernels; D2H of finished batches; geometry + offsets of the next batch */
    std::string err;
    /* genomes */
    bool have_genomes = false, any_impure = true;   /* any_impure: some 32-byte block of `bases` holds a byte other than A,C,G,T */
    simmr_genome_set host_view{};                 /* host copies of the small tables */
    std::vector<uint64_t> h_contig_off, h_contig_size;
    std::vector<uint32_t> h_genome_contig, h_contig_id_off, h_genome_id_off;
    std::string h_contig_ids, h_genome_ids;
    DevMem d_bases, d_impure, d_packed, d_contig_off, d_contig_size, d_genome_contig;
    /* headers */
    std::string header_fmt; bool have_headers = false;
    HeaderPlan plan;
    uint32_t hdr_fast = 0, hdr_fields = 0, hdr_kinds = 0;     /* decimal fields when every contig has the same ones */
    DevMem d_hdr;                                  /* header templates, one fixed-stride block per (contig, mate) */
    uint32_t hdr_stride = 0;
    /* profile */
    bool have_profile = false;
    ProfileModel model;
    DevMem d_phred, d_gap, d_qsub, d_rl_e, d_rl_lo, d_rl_hi, d_is_e, d_is_lo, d_is_hi;
    DevMem d_acc, d_plthr, d_rodds, d_ralias, d_roff, d_rlo, d_rhi;
    DevMem d_err_direct, d_err_ref, d_err_keep, d_err_row, d_err_cb, d_err_entries, d_err_alts, d_err_direct2, d_err_ent4;     /* k-mer error tables */
    DevMem d_epacked, d_ebases;                    /* edited reads of a batch (k_kmer_errors) */
    /* run scratch */
    DevMem d_unit_start, d_counters, d_out[2];
    DevMem d_sizes[2], d_offsets[2], d_block_sums[2], d_desc[2], d_desc_ext[2];   /* per-batch scratch, two sets: batch b+1 is prepared while batch b is formatted */
    DevMem d_seg_perm[2];                          /* k-mer error injection: the reads of every segment ordered by size (k_segment_sort) */
    PinnedMem h_out[SIMMR_RING_BUFFERS], h_total;                   /* pinned ring: batches wait here for the writer thread */
    PinnedMem h_in[2];                                              /* genome upload: pinned bounce buffers (simmr_genomes_load) */
    cudaEvent_t ev[24] = {};                       /* created with the context (Events below) */
    DevMem d_draw[13];
    simmr_stats stats{};
    bool dbg_guard = false, dbg_sync = false, dbg_serial = false;   /* SIMMR_DEBUG_GUARD / SIMMR_DEBUG_SYNC / SIMMR_PREP_SERIAL, read per run */
};

struct simmr_multi {
    std::vector<simmr_ctx *> ctx;
    std::vector<int> devices;
    std::string err;
    simmr_stats stats{};
};

namespace {

int fail(simmr_ctx *c, int code, const std::string &m) {
    last_error_slot() = m;
    if (c) c->err = m;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(ctx, SIMMR_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

/* how the short-read kernel samples qualities: a single position's table is copied once per lane (no bank conflicts) */
int quality_mode(const ProfileModel &m) {
    if (m.phred_kind != PT_Q32) return QM_CONST;
    return m.npos == 1 && m.cb <= 8 ? QM_LANE : QM_POS;
}
size_t lanes_smem(const ProfileModel &m, uint32_t warps, uint32_t stage_cap) {
    return SM_QTAB + qtab_bytes(quality_mode(m), m.cb, m.npos) + size_t(warps) * warp_region_bytes(stage_cap);
}

/* kernel instantiations: the word path is specialised per quality mode, replay is generic */
using FormatKernel = void (*)(RunParams, ReplayDev);
FormatKernel lanes_kernel(bool replay, int qm) {
    if (replay) return k_format_lanes<true, QM_CONST>;
    switch (qm) {
    case QM_LANE: return k_format_lanes<false, QM_LANE>;
    case QM_POS: return k_format_lanes<false, QM_POS>;
    default: return k_format_lanes<false, QM_CONST>;
    }
}
FormatKernel direct_kernel(bool replay) { return replay ? k_format_direct<true> : k_format_direct<false>; }
using LongKernel = void (*)(RunParams);
LongKernel long_kernel(int qm, uint32_t cb) {
    if (qm == QM_LANE) return cb == 7 ? k_format_long<QM_LANE, 7> : k_format_long<QM_LANE, 0>;
    return qm == QM_POS ? k_format_long<QM_POS, 0> : k_format_long<QM_CONST, 0>;
}

struct Plan {      /* launch geometry for one run */
    bool tiles = false;            /* warp-private stages in shared memory (short reads) vs one CTA per unit */
    bool lanes = false;            /* long reads, native mode: one lane per run (k_format_long) */
    uint32_t stage_cap = 0, max_unit_bytes = 0, threads = 0;
    size_t smem = 0;
    int grid = 0;
};

int make_plan(simmr_ctx *ctx, bool replay, Plan &pl) {
    const ProfileModel &m = ctx->model;
    const uint64_t per_rec = uint64_t(ctx->plan.hdr_cap) + 2ull * m.max_read_len + 5;
    const uint64_t unit_bytes = per_rec * (m.is_long ? 1 : 2);
    if (unit_bytes > 0xfff00000ull) return fail(ctx, SIMMR_E_LIMIT, "a unit's FASTQ bytes exceed 4 GiB");
    pl.max_unit_bytes = uint32_t(unit_bytes);
    const size_t smem_max = ctx->prop.sharedMemPerBlockOptin;
    pl.tiles = !m.is_long;
    if (pl.tiles) {
        /* One CTA per SM, its warps' stages share what the tables leave of shared memory.  A stage holds up to WARP_UNITS
         * units, each at its global offset modulo 16 (a slot is at most the unit's bytes + 30); k_segment_sort fills a
         * tile with as many units as fit.  Fewer warps when a single unit needs more than a sixteenth. */
        const uint64_t need = (unit_bytes + 30 + 15) & ~15ull;
        const uint64_t tables = SM_QTAB + qtab_bytes(quality_mode(m), m.cb, m.npos);
        auto cap_for = [&](uint32_t w) -> int64_t { return tables >= smem_max ? -1 : int64_t(((smem_max - tables) / w) & ~size_t(15)) - 64; };
        uint32_t warps = FORMAT_THREADS / 32;
        if (const char *e = getenv("SIMMR_WARPS")) warps = uint32_t(std::max(1, std::min(FORMAT_THREADS / 32, atoi(e))));
        while (warps > 1 && cap_for(warps) < int64_t(need)) --warps;
        /* A stage should hold a whole tile — WARP_UNITS units of typical size (nominal read length, header positions with
         * the digits of the contigs' sizes) and 5 % more: a tile that does not fit is cut in two, and its second part runs
         * with most lanes idle.  Large per-position quality tables (NovaSeq: 38 KB) are paid for with warps, not with tiles. */
        const uint64_t typ_unit = (uint64_t(ctx->plan.hdr_typ) + 2ull * m.nominal_read_len + 5) * 2;
        const int64_t want = int64_t(std::min<uint64_t>(WARP_UNITS * need, WARP_UNITS * typ_unit * 21 / 20 + 32));
        while (warps > 8 && cap_for(warps) < want) --warps;
        if (cap_for(warps) < int64_t(need)) pl.tiles = false;   /* a single unit does not fit: direct path */
        else {
            const uint64_t cap = std::min<uint64_t>(uint64_t(cap_for(warps)), WARP_UNITS * need);
            pl.stage_cap = uint32_t(cap); pl.threads = warps * 32; pl.smem = lanes_smem(m, warps, uint32_t(cap));
        }
    }
    int occ = 0;
    const size_t long_smem = SM_QTAB + qtab_bytes(quality_mode(m), m.cb, m.npos) + long_extra_smem();   /* tables, a genome-word ring and two line tiles per warp, an event list per lane */
    if (!pl.tiles && !replay && m.is_long && long_smem <= smem_max && !getenv("SIMMR_NO_LONG_KERNEL")) {   /* per-position tables too large for it: byte path */
        pl.lanes = true; pl.threads = DIRECT_THREADS;
        pl.smem = long_smem;
        LongKernel k = long_kernel(quality_mode(m), m.cb);
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl.smem)));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, int(pl.threads), pl.smem));
    } else {
        if (!pl.tiles) { pl.smem = 0; pl.threads = DIRECT_THREADS; }
        FormatKernel k = pl.tiles ? lanes_kernel(replay, quality_mode(m)) : direct_kernel(replay);
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pl.smem)));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, int(pl.threads), pl.smem));
    }
    if (occ < 1) return fail(ctx, SIMMR_E_LIMIT, "format kernel does not fit on an SM");
    pl.grid = occ * ctx->prop.multiProcessorCount;
    if (const char *e = getenv("SIMMR_GRID")) pl.grid = std::max(1, atoi(e));
    if (getenv("SIMMR_DEBUG_SYNC")) fprintf(stderr, "[simmr] tiles=%d threads=%u cap=%u smem=%zu occ=%d grid=%d\n", int(pl.tiles), pl.threads, pl.stage_cap, pl.smem, occ, pl.grid);
    return SIMMR_OK;
}

int ensure_headers(simmr_ctx *ctx, const char *fmt) {
    const std::string f = fmt ? fmt : DEFAULT_HEADER_FORMAT;
    if (ctx->have_headers && f == ctx->header_fmt) return SIMMR_OK;
    std::string err;
    int rc = compile_headers(ctx->host_view, f.c_str(), ctx->plan, err);
    if (rc) return fail(ctx, rc, err);
    if (ctx->plan.hdr_cap > 65535u) return fail(ctx, SIMMR_E_LIMIT, "a read header could exceed 65535 bytes");      /* UnitRec keeps header lengths in 16 bits */
    {   /* device image: block i = 2 * contig + mate of `stride` bytes holds the ContigHdr, its tokens and their literals
         * (literals on 4-byte boundaries, the kernel fetches them by words); token offsets become absolute */
        const HeaderPlan &hp = ctx->plan;
        size_t stride = 32;
        for (const ContigHdr &ch : hp.contigs) {
            size_t need = sizeof(ContigHdr) + size_t(ch.ntok) * sizeof(HdrTok);
            for (uint32_t i = 0; i < ch.ntok; ++i) need += (size_t(hp.toks[ch.tok_off + i].len) + 3) & ~size_t(3);
            stride = std::max(stride, (need + 4 + 31) & ~size_t(31));
        }
        if (stride * hp.contigs.size() > (size_t(2) << 30)) return fail(ctx, SIMMR_E_LIMIT, "header templates need more than 2 GiB");
        std::vector<uint8_t> img(stride * hp.contigs.size() + 64, 0);
        for (size_t i = 0; i < hp.contigs.size(); ++i) {
            uint8_t *blk = img.data() + i * stride;
            ContigHdr ch = hp.contigs[i];
            const uint32_t src_tok = ch.tok_off;
            ch.tok_off = uint32_t((i * stride + sizeof(ContigHdr)) / sizeof(HdrTok));
            memcpy(blk, &ch, sizeof ch);
            size_t lit_at = i * stride + sizeof(ContigHdr) + size_t(ch.ntok) * sizeof(HdrTok);
            for (uint32_t t = 0; t < ch.ntok; ++t) {
                HdrTok k = hp.toks[src_tok + t];
                if (k.type == TOK_LIT) { memcpy(img.data() + lit_at, hp.pool.data() + k.off, k.len); k.off = uint32_t(lit_at); lit_at += (size_t(k.len) + 3) & ~size_t(3); }
                memcpy(blk + sizeof(ContigHdr) + t * sizeof(HdrTok), &k, sizeof k);
            }
        }
        ctx->hdr_stride = uint32_t(stride);
        CU(ctx->d_hdr.upload(img.data(), img.size(), ctx->compute));
    }
    CU(cudaStreamSynchronize(ctx->compute));
    {   /* the same decimal fields in every contig (token lists are canonical: literal, field, ..., literal) let the
         * kernel run the template without per-token dispatch */
        const HeaderPlan &hp = ctx->plan;
        const uint32_t n = hp.contigs.empty() ? 0 : hp.contigs[0].ntok;
        bool same = n >= 1 && (n & 1u) && (n >> 1) <= HDR_MAX_FIELDS;
        for (size_t c = 0; same && c < hp.contigs.size(); ++c) {
            same = hp.contigs[c].ntok == n;
            for (uint32_t i = 0; same && i < n; ++i) {
                const uint8_t t = hp.toks[hp.contigs[c].tok_off + i].type;
                same = (i & 1u) ? (t != TOK_LIT && t == hp.toks[hp.contigs[0].tok_off + i].type) : t == TOK_LIT;
            }
        }
        ctx->hdr_fast = same; ctx->hdr_fields = same ? n >> 1 : 0; ctx->hdr_kinds = 0;
        for (uint32_t i = 0; same && i < (n >> 1); ++i) ctx->hdr_kinds |= uint32_t(hp.toks[hp.contigs[0].tok_off + 2 * i + 1].type) << (4 * i);
    }
    ctx->header_fmt = f; ctx->have_headers = true;
    return SIMMR_OK;
}

void fill_params(simmr_ctx *ctx, RunParams &rp, uint64_t seed, uint32_t id_base, const Plan &pl) {
    const ProfileModel &m = ctx->model;
    memset(&rp, 0, sizeof rp);
    rp.bases = ctx->d_bases.as<uint8_t>(); rp.contig_off = ctx->d_contig_off.as<uint64_t>();
    rp.contig_size = ctx->d_contig_size.as<uint64_t>(); rp.genome_contig = ctx->d_genome_contig.as<uint32_t>();
    rp.impure = ctx->d_impure.as<uint32_t>(); rp.packed = ctx->d_packed.as<uint32_t>() + PACK_PAD_WORDS;
    rp.hdr_fast = ctx->hdr_fast; rp.hdr_fields = ctx->hdr_fields; rp.hdr_kinds = ctx->hdr_kinds;
    rp.hdr = ctx->d_hdr.as<uint8_t>(); rp.hdr_stride = ctx->hdr_stride; rp.toks = ctx->d_hdr.as<HdrTok>(); rp.pool = ctx->d_hdr.as<uint8_t>();
    rp.n_genomes = ctx->host_view.n_genomes; rp.n_contigs = ctx->host_view.n_contigs;
    rp.unit_start = ctx->d_unit_start.as<uint64_t>();
    rp.read_id_base = id_base;
    rp.phred = PhredTab{m.phred_kind, m.const_q, m.cb, m.npos, ctx->d_phred.as<uint32_t>(), ctx->d_gap.as<uint32_t>(), ctx->d_qsub.as<uint32_t>(), m.has_sub ? 1u : 0u, 0u};
    rp.rl = LenTab{m.rl.kind, m.rl.value, m.rl.cb, 0, ctx->d_rl_e.as<uint32_t>(), ctx->d_rl_lo.as<uint32_t>(), ctx->d_rl_hi.as<uint32_t>()};
    rp.is = LenTab{m.is.kind, m.is.value, m.is.cb, 0, ctx->d_is_e.as<uint32_t>(), ctx->d_is_lo.as<uint32_t>(), ctx->d_is_hi.as<uint32_t>()};
    rp.gamma = m.gamma;
    rp.min_size = m.min_size; rp.is_long = m.is_long; rp.flags = m.flags; rp.phred_mult = 1u << m.cb;
    rp.mutates = m.mutates;
    for (uint32_t r = 0; r < 10; ++r) { rp.rk[2 * r] = uint32_t(seed) + r * 0x9E3779B9u; rp.rk[2 * r + 1] = uint32_t(seed >> 32) + r * 0xBB67AE85u; }
    rp.counters = ctx->d_counters.as<unsigned long long>();
    rp.stage_cap = pl.stage_cap; rp.one = 1;
    if (m.err.on) {
        rp.err = ErrTab{1u, m.err.k, uint32_t(m.err.ref.size()), 0u, ctx->d_err_direct.as<uint32_t>(), ctx->d_err_ref.as<uint32_t>(), ctx->d_err_keep.as<uint32_t>(),
                        ctx->d_err_row.as<uint32_t>(), ctx->d_err_cb.as<uint32_t>(), ctx->d_err_entries.as<uint32_t>(), ctx->d_err_alts.as<uint32_t>(), ctx->d_err_direct2.as<uint32_t>(), ctx->d_err_ent4.as<uint32_t>()};
        rp.epacked = ctx->d_epacked.as<uint32_t>(); rp.ebases = ctx->d_ebases.as<uint8_t>();
    }
}

/* ---------------------------------------------------------------------- */
/* the batch pipeline shared by simulate, replay and measure                 */
/* ---------------------------------------------------------------------- */
constexpr int RING = SIMMR_RING_BUFFERS;   /* pinned host buffers per context */
enum : int { EV_START = 0, EV_STOP = 1, EV_F0 /* 2 */ = 2, EV_F1 /* 2 */ = 4, EV_TOTAL /* 2 */ = 6, EV_FMT /* 2 */ = 8, EV_PREP /* 2 */ = 10, EV_DONE /* 2 */ = 12, EV_COPY /* RING */ = 14, EV_COUNT = 14 + SIMMR_RING_BUFFERS };

/* A finished batch on its way from a launch thread to the writer thread. */
struct Ready { int slot; const uint8_t *data; uint64_t bytes, first, units; cudaEvent_t copied; };
/* Launch thread -> writer thread hand-off of one context: the RING pinned buffers are used round-robin, and a sink that
 * retains chunks may give them back in any order, so the producer waits for the very buffer it is about to fill.  A
 * failing sink makes the writer raise `abort`; it keeps releasing buffers so that the producer never blocks, and the
 * producer stops at its next batch. */
struct BatchQueue {
    std::mutex m; std::condition_variable cv;
    std::deque<Ready> q;
    bool busy[RING] = {}; int retained = 0; bool closed = false;   /* retained: chunks a sink kept (SIMMR_SINK_RETAINED) and has not given back */
    std::atomic<bool> abort{false};
    bool acquire(int slot) { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return !busy[slot]; }); busy[slot] = true; return !abort.load(); }
    void push(const Ready &r) { { std::lock_guard<std::mutex> l(m); q.push_back(r); } cv.notify_all(); }
    void close() { { std::lock_guard<std::mutex> l(m); closed = true; } cv.notify_all(); }
    bool pop(Ready &r) { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return !q.empty() || closed; }); if (q.empty()) return false; r = q.front(); q.pop_front(); return true; }
    void release(int slot, bool was_retained = false) { { std::lock_guard<std::mutex> l(m); busy[slot] = false; if (was_retained) --retained; } cv.notify_all(); }
    void retain() { std::lock_guard<std::mutex> l(m); ++retained; }
    void wait_all_released() { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return retained == 0; }); }
};
/* chunks a sink retained (SIMMR_SINK_RETAINED): pinned buffer -> the queue that gets it back */
std::mutex g_retained_mu;
std::map<const void *, std::pair<BatchQueue *, int>> g_retained;      /* ... and the buffer's place in that queue's ring */
/* The writer thread: batches leave in file order — batch j comes from queue j % n — and the sink sees them one at a
 * time.  Returns false if the sink failed. */
bool write_batches(const std::vector<BatchQueue *> &qs, const std::vector<int> &devices, simmr_sink_fn sink, void *user) {
    bool ok = true;
    std::vector<bool> open(qs.size(), true);
    size_t live = qs.size();
    for (size_t j = 0; live; ++j) {
        const size_t g = j % qs.size();
        if (!open[g]) continue;
        Ready r;
        if (!qs[g]->pop(r)) { open[g] = false; --live; continue; }
        cudaSetDevice(devices[g]);
        const bool copied = cudaEventSynchronize(r.copied) == cudaSuccess;
        int verdict = 1;
        if (ok && copied) {
            { std::lock_guard<std::mutex> l(g_retained_mu); g_retained[r.data] = {qs[g], r.slot}; }      /* before the sink can pass it on */
            qs[g]->retain();                                   /* counted before the sink can release it from another thread */
            verdict = sink(user, r.data, r.bytes, r.first, r.units);
            if (verdict != SIMMR_SINK_RETAINED) {
                bool still = false;
                { std::lock_guard<std::mutex> l(g_retained_mu); still = g_retained.erase(r.data) != 0; }
                if (still) { std::lock_guard<std::mutex> l(qs[g]->m); --qs[g]->retained; }
            }
        }
        if (ok && verdict != 0 && verdict != SIMMR_SINK_RETAINED) { ok = false; for (BatchQueue *q : qs) q->abort.store(true); }
        if (verdict != SIMMR_SINK_RETAINED) qs[g]->release(r.slot);
    }
    return ok;
}

struct Runner {
    simmr_ctx *ctx; bool replay; ReplayDev rd; Plan pl; RunParams rp;
    uint64_t ub = 0, ue = 0, batch = 0, out_cap = 0;
    bool errs = false, measure_only = false, guard = false, dbg_sync = false, done_live[2] = {false, false};
    cudaStream_t prep = nullptr;
    uint64_t launches = 0;
    cudaEvent_t *ev;

    void scratch(RunParams &r, int k) const {
        r.sizes = ctx->d_sizes[k].as<uint32_t>(); r.offsets = ctx->d_offsets[k].as<uint64_t>();
        r.desc = ctx->d_desc[k].as<UnitRec>(); r.desc_ext = ctx->d_desc_ext[k].as<UnitExt>(); r.block_sums = ctx->d_block_sums[k].as<uint64_t>();
        r.seg_perm = ctx->d_seg_perm[k].as<uint16_t>();
    }
    /* plan, batch size, device and pinned buffers */
    int init(uint64_t ub_, uint64_t ue_, uint64_t seed, const simmr_run_opts *opts, bool to_host, bool measure, bool dev_sink = false) {
        ub = ub_; ue = ue_; measure_only = measure; ev = ctx->ev;
        int rc = make_plan(ctx, replay, pl);
        if (rc) return rc;
        /* Staging per batch: 16 GiB (of 180) when nothing has to cross PCIe, less when the device is short of memory.  Few
         * large launches beat many small ones twice: every launch fills the tables of 148 CTAs and ends with a partly idle
         * GPU (short reads: 2 -> 16 GiB is +7 %), and a long read is budgeted at its largest possible size (131 KB) although
         * it averages 40 KB — 2 GiB were 16 K reads per launch, 7 per warp, and the last warps' long reads left the GPU half
         * idle (+22 %). */
        errs = ctx->model.err.on && !replay;           /* k-mer error injection between sizing and formatting */
        /* to the host: 8 pinned buffers of 128 MiB.  Error injection walks a read per thread, and a batch is as slow as its
         * longest read: the more reads share a launch the better the GPU is filled (16 GiB = 343 K reads = 1.1 waves of
         * threads ran at 43 % occupancy), so those batches take up to 64 GiB when the device has the room. */
        uint64_t target_bytes = to_host ? (errs ? (1024ull << 20) : (128ull << 20)) : errs ? (65536ull << 20) : (16384ull << 20);   /* to the host with error injection: 1 GiB buffers, 21 K reads per launch */
        if (measure_only) target_bytes = errs ? (4096ull << 20) : (uint64_t(pl.max_unit_bytes) << 24);
        if (!to_host && !measure_only) {
            size_t free_b = 0, total_b = 0;
            CU(cudaMemGetInfo(&free_b, &total_b));
            free_b += ctx->d_out[0].bytes + ctx->d_out[1].bytes + ctx->d_epacked.bytes + ctx->d_ebases.bytes;       /* what this context already holds can be reused */
            /* error injection keeps a batch's edited reads beside the staging buffer: a byte and two bits per FASTQ byte */
            while (target_bytes > (256ull << 20) && target_bytes > free_b / ((dev_sink ? 5 : 3) + (errs ? (ctx->any_impure ? 3 : 1) : 0))) target_bytes /= 2;
        }
        /* Bytes a batch of n units is given.  Short reads: n times the largest unit.  Long reads are 40 KB on average and
         * 131 KB at most, and a batch budgeted at the maximum holds a third of the reads it could — which left the GPU
         * mostly idle in the k-mer error walk (a thread per read: 32 K reads per batch, every batch waiting for its longest
         * read).  They get their typical size and a quarter more, plus sixteen largest units: at least nine standard
         * deviations of a batch's bytes above their mean for gamma lengths, whatever n.  The format kernels write
         * nothing when a batch's bytes exceed its buffer (RunParams::out_cap), and the run fails with SIMMR_E_LIMIT. */
        const uint64_t max_unit = pl.max_unit_bytes, typ_unit = std::min<uint64_t>(max_unit, uint64_t(ctx->plan.hdr_typ) + 2ull * ctx->model.nominal_read_len + 5);
        auto cap_of = [&](uint64_t n) { return ctx->model.is_long ? std::min(n * max_unit, n * typ_unit * 5 / 4 + 16 * max_unit) : n * max_unit; };
        if (opts && opts->batch_units) batch = opts->batch_units;
        else if (ctx->model.is_long && target_bytes > 32 * max_unit) batch = (target_bytes - 16 * max_unit) * 4 / (typ_unit * 5);
        else batch = target_bytes / max_unit;
        batch = std::max<uint64_t>(1, std::min<uint64_t>(batch, std::max<uint64_t>(1, ue - ub)));
        batch = std::min<uint64_t>(batch, 0x7fffffffull);
        while (batch > 1 && cap_of(batch) > (64ull << 30)) batch /= 2;
        out_cap = cap_of(batch) + 64;
        if (const char *e = getenv("SIMMR_DEBUG_CAP")) out_cap = std::min<uint64_t>(out_cap, strtoull(e, nullptr, 10));   /* tests: a batch that does not fit */
        const uint32_t max_blocks = uint32_t((batch + SIZES_THREADS - 1) / SIZES_THREADS);
        const uint64_t max_segs = (batch + SEG_UNITS - 1) / SEG_UNITS;
        for (int k = 0; k < 2; ++k) {
            CU(ctx->d_sizes[k].reserve(batch * 4)); CU(ctx->d_desc[k].reserve(batch * sizeof(UnitRec))); CU(ctx->d_desc_ext[k].reserve(batch * sizeof(UnitExt)));
            CU(ctx->d_offsets[k].reserve((batch + 1) * 8)); CU(ctx->d_block_sums[k].reserve(size_t(max_blocks) * 8));
            if (errs) CU(ctx->d_seg_perm[k].reserve(max_segs * SEG_UNITS * 2));
        }
        if (errs) {      /* scratch for a batch's edited reads: as many bases as the batch has FASTQ bytes, 16 more per read (err_scratch_base) */
            const uint64_t nb = out_cap + 16 * batch + 64;
            CU(ctx->d_epacked.reserve(nb / 4 + 4 * PACK_PAD_WORDS));
            if (ctx->any_impure) CU(ctx->d_ebases.reserve(nb)); else ctx->d_ebases.release();      /* byte-wise scratch: only reads that hold something else than A,C,G,T use it */
        }
        CU(ctx->d_counters.reserve(64)); CU(cudaMemsetAsync(ctx->d_counters.p, 0, 64, ctx->compute));
        guard = ctx->dbg_guard; dbg_sync = ctx->dbg_sync;
        const int nbuf = measure_only ? 0 : (to_host || dev_sink) ? 2 : 1;       /* a device consumer reads batch b while batch b + 1 is being formatted */
        for (int b = 0; b < nbuf; ++b) CU(ctx->d_out[b].reserve(out_cap));
        if (to_host) for (int b = 0; b < RING; ++b) CU(ctx->h_out[b].reserve(out_cap, ctx->numa_node));
        CU(ctx->h_total.reserve(2 * sizeof(uint64_t), ctx->numa_node));
        fill_params(ctx, rp, seed, opts ? opts->read_id_base : 0, pl);
        rp.out_cap = out_cap;
        /* error injection reads and writes one scratch set per context: its batches are prepared on the compute stream, one
         * at a time.  dbg_serial: everything on one stream (kernel timings without co-runners). */
        prep = (errs || measure_only || ctx->dbg_serial) ? ctx->compute : ctx->prep;
        return SIMMR_OK;
    }
    bool ahead() const { return !errs && !measure_only; }   /* batch b + 1 is prepared beside batch b's format kernel */
    /* geometry, sizes and byte offsets of the batch that starts at unit u into scratch set k: the small kernels need no
     * shared memory and share the SMs with the format kernel of the batch before */
    int prepare(uint64_t u, int k) {
        const uint32_t n = uint32_t(std::min<uint64_t>(batch, ue - u));
        RunParams r = rp; scratch(r, k);
        r.batch_begin = u; r.batch_units = n;
        if (done_live[k]) CU(cudaStreamWaitEvent(prep, ev[EV_DONE + k], 0));     /* the format kernel that read this set has finished */
        const uint32_t nblk = (n + SIZES_THREADS - 1) / SIZES_THREADS;
        if (replay) k_unit_sizes<true><<<nblk, SIZES_THREADS, 0, prep>>>(r, rd);
        else k_unit_sizes<false><<<nblk, SIZES_THREADS, 0, prep>>>(r, rd);
        if (dbg_sync) { cudaError_t e_ = cudaStreamSynchronize(prep); if (e_ != cudaSuccess) return fail(ctx, SIMMR_E_CUDA, std::string("k_unit_sizes: ") + cudaGetErrorString(e_)); }
        k_scan_blocks<<<1, 1024, 0, prep>>>(r, nblk);
        if (errs) k_segment_sort<<<(n + SEG_UNITS - 1) / SEG_UNITS, SEG_THREADS, 0, prep>>>(r);     /* offsets + the reads of every segment ordered by size */
        else k_unit_offsets<<<nblk, SIZES_THREADS, 0, prep>>>(r);
        launches += 3;
        if (errs) {      /* edit the reads (size-sorted order), then offsets again from the edited sizes */
            k_kmer_errors<<<(n + 255) / 256, 256, 0, prep>>>(r);
            k_block_sums<<<nblk, SIZES_THREADS, 0, prep>>>(r);
            k_scan_blocks<<<1, 1024, 0, prep>>>(r, nblk);
            k_unit_offsets<<<nblk, SIZES_THREADS, 0, prep>>>(r);
            launches += 4;
        }
        CU(cudaGetLastError());
        if (dbg_sync) { cudaError_t e_ = cudaStreamSynchronize(prep); if (e_ != cudaSuccess) return fail(ctx, SIMMR_E_CUDA, std::string("scan kernels: ") + cudaGetErrorString(e_)); }
        CU(cudaEventRecord(ev[EV_PREP + k], prep));
        return SIMMR_OK;
    }
    /* the format kernel of the batch prepared in set k, into d_out[dslot]; bracketed by EV_F0 / EV_F1 of set k */
    int format(uint64_t u, int k, int dslot) {
        const uint32_t n = uint32_t(std::min<uint64_t>(batch, ue - u));
        scratch(rp, k);
        rp.batch_begin = u; rp.batch_units = n; rp.out = ctx->d_out[dslot].as<uint8_t>(); rp.out_cap = out_cap;
        if (guard) CU(cudaMemsetAsync(ctx->d_out[dslot].p, 0xA5, out_cap, ctx->compute));
        CU(cudaEventRecord(ev[EV_F0 + k], ctx->compute));
        if (pl.tiles) {
            const uint32_t wpc = pl.threads / 32, nctas = ((n + WARP_UNITS - 1) / WARP_UNITS + wpc - 1) / wpc;
            const int grid = int(std::min<uint32_t>(nctas, uint32_t(pl.grid)));
            lanes_kernel(replay, quality_mode(ctx->model))<<<grid, pl.threads, pl.smem, ctx->compute>>>(rp, rd);
        } else if (pl.lanes) {
            CU(cudaMemsetAsync(rp.counters + 4, 0, 8, ctx->compute));       /* the kernel's piece counter */
            long_kernel(quality_mode(ctx->model), ctx->model.cb)<<<pl.grid, pl.threads, pl.smem, ctx->compute>>>(rp);
        } else {
            const int grid = int(std::min<uint32_t>(n, uint32_t(pl.grid)));
            direct_kernel(replay)<<<grid, pl.threads, pl.smem, ctx->compute>>>(rp, rd);
        }
        ++launches;
        CU(cudaGetLastError());
        CU(cudaEventRecord(ev[EV_F1 + k], ctx->compute));
        CU(cudaEventRecord(ev[EV_DONE + k], ctx->compute)); done_live[k] = true;
        if (guard) {   /* SIMMR_DEBUG_GUARD: nothing may be written past the batch's last byte */
            uint64_t total = 0; uint8_t tail[64];
            CU(cudaMemcpyAsync(&total, rp.offsets + n, 8, cudaMemcpyDeviceToHost, ctx->compute));
            CU(cudaStreamSynchronize(ctx->compute));
            if (total + 64 > out_cap) return fail(ctx, SIMMR_E_LIMIT, "batch overflowed its staging buffer");
            CU(cudaMemcpy(tail, rp.out + total, 64, cudaMemcpyDeviceToHost));
            for (int b = 0; b < 64; ++b) if (tail[b] != 0xA5) return fail(ctx, SIMMR_E_CUDA, "guard bytes after the batch output were overwritten");
        }
        return SIMMR_OK;
    }
    int finish(simmr_stats &st, double format_ms) {
        CU(cudaEventRecord(ev[EV_STOP], ctx->compute));
        CU(cudaEventSynchronize(ev[EV_STOP]));
        float ms = 0; CU(cudaEventElapsedTime(&ms, ev[EV_START], ev[EV_STOP]));
        unsigned long long ctr[8];
        CU(cudaMemcpy(ctr, ctx->d_counters.p, sizeof ctr, cudaMemcpyDeviceToHost));
        if (ctr[6]) return fail(ctx, SIMMR_E_LIMIT, "a batch overflowed its staging buffer and was not written: set batch_units lower");   /* raised by the format kernels */
        /* ctr[5]: bases deleted by k-mer error injection — read from the genome, absent from the sequence line */
        st.clamped = ctr[0]; st.algorithmic_bytes = ctr[1] - ctr[5]; st.bases = ctr[2] - ctr[5]; st.records = ctr[3];
        st.bytes = ctr[1] - ctr[2] - ctr[5];                          /* (3L+H+5) - L, less the deleted bases */
        st.device_ms = ms; st.format_ms = format_ms; st.kernel_launches = launches;
        return SIMMR_OK;
    }
};

/* Batches number j0, j0 + stride, ... of the unit range (batch j starts at unit ub + j * batch) — every batch when the
 * context runs alone, every n-th when n contexts share a run.  to_host: finished batches go to the pinned ring and into
 * `q` for the writer thread.  Otherwise (device-resident run, or measure) the bytes stay in HBM staging. */
int produce(Runner &R, uint64_t j0, uint64_t stride, BatchQueue *q, simmr_stats &st, simmr_device_sink_fn dev_sink = nullptr, void *dev_user = nullptr) {
    simmr_ctx *ctx = R.ctx;
    cudaEvent_t *ev = R.ev;
    const uint64_t nbatches = (R.ue - R.ub + R.batch - 1) / R.batch;
    double format_ms = 0.0;
    int rc = SIMMR_OK;
    CU(cudaEventRecord(ev[EV_START], ctx->compute));
    CU(cudaStreamWaitEvent(R.prep, ev[EV_START], 0));
    auto start_of = [&](uint64_t j) { return R.ub + j * R.batch; };
    if (R.ahead() && j0 < nbatches) { rc = R.prepare(start_of(j0), 0); if (rc) return rc; }
    uint64_t seq = 0;                                          /* this context's own batch count */
    /* device-resident runs: batch b's kernel time is read, and its bytes are shown to the device consumer, while batch
     * b + 1 is already running (two staging buffers when there is a consumer) */
    struct Last { bool live = false; int k = 0, dslot = 0; uint64_t first = 0; uint32_t units = 0; } last;
    auto settle = [&]() -> int {
        if (!last.live) return SIMMR_OK;
        last.live = false;
        CU(cudaEventSynchronize(ev[EV_F1 + last.k]));
        float ms = 0; CU(cudaEventElapsedTime(&ms, ev[EV_F0 + last.k], ev[EV_F1 + last.k])); format_ms += ms;

        if (dev_sink) {
            const uint64_t total = static_cast<uint64_t *>(ctx->h_total.p)[last.dslot];
            if (total > R.out_cap) return fail(ctx, SIMMR_E_LIMIT, "batch overflowed its staging buffer: set batch_units lower");
            if (dev_sink(dev_user, ctx->d_out[last.dslot].as<uint8_t>(), total, last.first, last.units)) return fail(ctx, SIMMR_E_SINK, "device sink callback failed");
        }
        return SIMMR_OK;
    };
    for (uint64_t j = j0; j < nbatches; j += stride, ++seq) {
        const uint64_t u = start_of(j);
        const uint32_t n = uint32_t(std::min<uint64_t>(R.batch, R.ue - u));
        const int k = int(seq & 1), dslot = (q || dev_sink) ? int(seq & 1) : 0, slot = int(seq % RING);
        if (!R.ahead()) { rc = R.prepare(u, k); if (rc) return rc; }
        if (R.measure_only) continue;                          /* geometry, sizes (and edits) only: what the units occupy in the FASTQ stream */
        if (q && !q->acquire(slot)) return fail(ctx, SIMMR_E_SINK, "sink callback failed");      /* this batch's pinned buffer is free again */
        if (R.ahead() && j + stride < nbatches) { rc = R.prepare(start_of(j + stride), k ^ 1); if (rc) return rc; }
        CU(cudaStreamWaitEvent(ctx->compute, ev[EV_PREP + k], 0));
        if (q || dev_sink) {      /* the batch's byte count, on its way to the host before the format kernel starts */
            if (q && seq >= 2) CU(cudaStreamWaitEvent(ctx->compute, ev[EV_COPY + int((seq - 2) % RING)], 0));   /* d_out[dslot] has been copied out */
            R.scratch(R.rp, k);
            CU(cudaMemcpyAsync(static_cast<uint64_t *>(ctx->h_total.p) + dslot, R.rp.offsets + n, 8, cudaMemcpyDeviceToHost, ctx->compute));
            CU(cudaEventRecord(ev[EV_TOTAL + dslot], ctx->compute));
        }
        rc = R.format(u, k, dslot); if (rc) return rc;
        if (q) {
            CU(cudaEventRecord(ev[EV_FMT + dslot], ctx->compute));
            CU(cudaEventSynchronize(ev[EV_TOTAL + dslot]));            /* byte count known while the format kernel runs */
            const uint64_t total = static_cast<uint64_t *>(ctx->h_total.p)[dslot];
            if (total > R.out_cap) return fail(ctx, SIMMR_E_LIMIT, "batch overflowed its staging buffer: set batch_units lower");
            CU(cudaStreamWaitEvent(ctx->copy, ev[EV_FMT + dslot], 0));
            CU(cudaMemcpyAsync(ctx->h_out[slot].p, ctx->d_out[dslot].p, total, cudaMemcpyDeviceToHost, ctx->copy));
            CU(cudaEventRecord(ev[EV_COPY + slot], ctx->copy));
            q->push(Ready{slot, static_cast<const uint8_t *>(ctx->h_out[slot].p), total, u, n, ev[EV_COPY + slot]});
            st.d2h_bytes += total + 8;
        } else {
            rc = settle(); if (rc) return rc;                  /* the batch before this one */
            last = Last{true, k, dslot, u, n};
        }
    }
    rc = settle(); if (rc) return rc;
    return R.finish(st, format_ms);
}

/* A run of one context: the caller's thread launches, a writer thread owned by the library hands the batches to the sink. */
int run_units(simmr_ctx *ctx, bool replay, const ReplayDev &rd, uint64_t ub, uint64_t ue, uint64_t seed,
              const simmr_run_opts *opts, simmr_sink_fn sink, void *user, bool measure_only = false,
              simmr_device_sink_fn dev_sink = nullptr) {
    /* debug switches, read once per run */
    ctx->dbg_guard = getenv("SIMMR_DEBUG_GUARD") != nullptr; ctx->dbg_sync = getenv("SIMMR_DEBUG_SYNC") != nullptr; ctx->dbg_serial = getenv("SIMMR_PREP_SERIAL") != nullptr;
    Runner R{ctx, replay, rd};
    int rc = R.init(ub, ue, seed, opts, sink != nullptr, measure_only, dev_sink != nullptr);
    if (rc) return rc;
    simmr_stats st{};
    st.units = ue - ub;
    if (!sink) { rc = produce(R, 0, 1, nullptr, st, dev_sink, user); if (!rc) ctx->stats = st; return rc; }
    BatchQueue q;
    bool sink_ok = true;
    std::thread writer([&] { run_on_node(ctx->numa_node); sink_ok = write_batches({&q}, {ctx->device}, sink, user); });
    rc = produce(R, 0, 1, &q, st);
    q.close();
    writer.join();
    q.wait_all_released();                  /* chunks the sink retained come back before the ring can go away */
    if (rc != SIMMR_OK) { cudaStreamSynchronize(ctx->compute); cudaStreamSynchronize(ctx->copy); cudaStreamSynchronize(ctx->prep); }   /* nothing of the run stays in flight */
    if (!sink_ok) return fail(ctx, SIMMR_E_SINK, "sink callback failed");
    if (!rc) ctx->stats = st;
    return rc;
}

int check_ready(simmr_ctx *ctx) {
    if (!ctx) return fail(nullptr, SIMMR_E_INVALID, "null context");
    if (!ctx->have_genomes) return fail(ctx, SIMMR_E_STATE, "no genomes loaded");
    if (!ctx->have_profile) return fail(ctx, SIMMR_E_STATE, "no error profile set");
    /* simulate.rs:220-225 / main.rs:124: every contig must exceed the minimum size */
    const ProfileModel &m = ctx->model;
    for (uint32_t c = 0; c < ctx->host_view.n_contigs; ++c)
        if (ctx->h_contig_size[c] <= m.min_size)
            return fail(ctx, SIMMR_E_TOO_SMALL, "contig " + std::to_string(c) + " (" + std::to_string(ctx->h_contig_size[c]) +
                        " nt) does not exceed the profile's minimum size " + std::to_string(m.min_size));
    for (uint32_t g = 0; g < ctx->host_view.n_genomes; ++g)
        if (ctx->h_genome_contig[g + 1] == ctx->h_genome_contig[g]) return fail(ctx, SIMMR_E_TOO_SMALL, "genome " + std::to_string(g) + " has no usable contigs");
    return SIMMR_OK;
}

/* the run's unit range: everything, unless opts names one (0,0 counts as a range only with SIMMR_RUN_EXPLICIT_RANGE) */
void unit_range(const simmr_run_opts *opts, uint64_t total, uint64_t &ub, uint64_t &ue) {
    ub = opts ? opts->unit_begin : 0; ue = opts ? opts->unit_end : 0;
    if (ub == 0 && ue == 0 && !(opts && (opts->flags & SIMMR_RUN_EXPLICIT_RANGE))) ue = total;
}
std::vector<uint64_t> unit_prefix(const simmr_ctx *ctx, const uint64_t *reads) {
    const uint32_t n = ctx->host_view.n_genomes;
    std::vector<uint64_t> s(n + 1, 0);
    for (uint32_t g = 0; g < n; ++g) s[g + 1] = s[g] + (ctx->model.is_long ? reads[g] : reads[g] / 2);   /* simulate.rs:179 / :356 */
    return s;
}
int mfail(simmr_multi *m, int code, const std::string &msg) { last_error_slot() = msg; if (m) m->err = msg; return code; }
/* fn(g) on one thread per device; the first failure's code (its message is in the context) */
template <class F> int on_every_device(simmr_multi *m, F fn) {
    std::vector<int> rc(m->ctx.size(), SIMMR_OK);
    std::vector<std::thread> th;
    for (size_t g = 0; g < m->ctx.size(); ++g) th.emplace_back([&, g] { cudaSetDevice(m->devices[g]); run_on_node(m->ctx[g]->numa_node); rc[g] = fn(int(g)); });
    for (auto &t : th) t.join();
    for (size_t g = 0; g < rc.size(); ++g) if (rc[g]) return mfail(m, rc[g], "device " + std::to_string(m->devices[g]) + ": " + m->ctx[g]->err);
    return SIMMR_OK;
}

}  // namespace

extern "C" {

const char *simmr_last_error(const simmr_ctx *ctx) { return ctx ? ctx->err.c_str() : last_error_slot().c_str(); }
void simmr_chunk_release(const uint8_t *data) {
    BatchQueue *q = nullptr;
    int slot = 0;
    { std::lock_guard<std::mutex> l(g_retained_mu); auto it = g_retained.find(data); if (it != g_retained.end()) { q = it->second.first; slot = it->second.second; g_retained.erase(it); } }
    if (q) q->release(slot, true);
}
const char *simmr_version(void) { return "simmr_b200 0.1 (sm_100a)"; }

int simmr_ctx_create(int device, simmr_ctx **out) {
    simmr_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, SIMMR_E_INVALID, "null out pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(nullptr, SIMMR_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, SIMMR_E_INVALID, "device index out of range");
    ctx = new simmr_ctx; ctx->device = device;
    auto bail = [&](const std::string &m) { delete ctx; return fail(nullptr, SIMMR_E_CUDA, m); };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(cudaGetErrorString(e));
    if ((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) return bail(cudaGetErrorString(e));
    if (ctx->prop.major != 10) return bail("device is sm_" + std::to_string(ctx->prop.major * 10 + ctx->prop.minor) + "; this library carries sm_100a code only");
    if ((e = cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking)) != cudaSuccess) return bail(cudaGetErrorString(e));
    if ((e = cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking)) != cudaSuccess) return bail(cudaGetErrorString(e));
    if ((e = cudaStreamCreateWithFlags(&ctx->prep, cudaStreamNonBlocking)) != cudaSuccess) return bail(cudaGetErrorString(e));
    ctx->numa_node = numa_node_of_device(device);
    for (int i = 0; i < EV_COUNT; ++i) {      /* the pipeline's events live as long as the context */
        const bool timed = i < EV_TOTAL;
        if ((e = cudaEventCreateWithFlags(&ctx->ev[i], timed ? cudaEventDefault : cudaEventDisableTiming)) != cudaSuccess) return bail(cudaGetErrorString(e));
    }
    *out = ctx;
    return SIMMR_OK;
}

void simmr_ctx_destroy(simmr_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < EV_COUNT; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->compute) cudaStreamDestroy(ctx->compute);
    if (ctx->copy) cudaStreamDestroy(ctx->copy);
    if (ctx->prep) cudaStreamDestroy(ctx->prep);
    delete ctx;
}

int simmr_device_info(const simmr_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, uint64_t *hbm_bytes) {
    if (!ctx) return fail(nullptr, SIMMR_E_INVALID, "null context");
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    if (hbm_bytes) *hbm_bytes = ctx->prop.totalGlobalMem;
    return SIMMR_OK;
}

int simmr_genomes_load(simmr_ctx *ctx, const simmr_genome_set *s) {
    if (!ctx || !s) return fail(ctx, SIMMR_E_INVALID, "null argument");
    if (!s->n_genomes || !s->n_contigs) return fail(ctx, SIMMR_E_INVALID, "empty genome set");
    CU(cudaSetDevice(ctx->device));
    ctx->h_contig_off.assign(s->contig_off, s->contig_off + s->n_contigs + 1);
    ctx->h_contig_size.assign(s->contig_size, s->contig_size + s->n_contigs);
    ctx->h_genome_contig.assign(s->genome_contig, s->genome_contig + s->n_genomes + 1);
    ctx->h_contig_id_off.assign(s->contig_id_off, s->contig_id_off + s->n_contigs + 1);
    ctx->h_genome_id_off.assign(s->genome_id_off, s->genome_id_off + s->n_genomes + 1);
    ctx->h_contig_ids.assign(s->contig_ids, ctx->h_contig_id_off.back());
    ctx->h_genome_ids.assign(s->genome_ids, ctx->h_genome_id_off.back());
    for (uint32_t c = 0; c < s->n_contigs; ++c)
        if (ctx->h_contig_off[c + 1] < ctx->h_contig_off[c]) return fail(ctx, SIMMR_E_INVALID, "contig offsets must be non-decreasing");
    if (ctx->h_genome_contig.back() != s->n_contigs) return fail(ctx, SIMMR_E_INVALID, "genome_contig does not cover all contigs");
    /* the kernels index with these tables unchecked */
    if (ctx->h_genome_contig.front() != 0) return fail(ctx, SIMMR_E_INVALID, "genome_contig must start at 0");
    for (uint32_t g = 0; g < s->n_genomes; ++g)
        if (ctx->h_genome_contig[g + 1] < ctx->h_genome_contig[g]) return fail(ctx, SIMMR_E_INVALID, "genome_contig must be non-decreasing");
    for (uint32_t c = 0; c < s->n_contigs; ++c)
        if (ctx->h_contig_size[c] > ctx->h_contig_off[c + 1] - ctx->h_contig_off[c]) return fail(ctx, SIMMR_E_INVALID, "a contig's size exceeds its span in bases");
    simmr_genome_set &v = ctx->host_view;
    v = *s; v.bases = nullptr;
    v.contig_off = ctx->h_contig_off.data(); v.contig_size = ctx->h_contig_size.data(); v.genome_contig = ctx->h_genome_contig.data();
    v.contig_ids = ctx->h_contig_ids.data(); v.contig_id_off = ctx->h_contig_id_off.data();
    v.genome_ids = ctx->h_genome_ids.data(); v.genome_id_off = ctx->h_genome_id_off.data();
    const uint64_t nbytes = ctx->h_contig_off.back();
    /* one resident copy of every contig; the reference clones sequences per read (simulate.rs:362-367) */
    CU(ctx->d_bases.reserve(nbytes + 64));
    if (nbytes < (size_t(64) << 20)) CU(cudaMemcpy(ctx->d_bases.p, s->bases, nbytes, cudaMemcpyHostToDevice));
    else {
        /* The caller's buffer is ordinary pageable memory: a plain cudaMemcpy stages it through the driver's own bounce buffer
         * on one thread (13 GB/s; 40 ms of C2's end-to-end step for 0.5 GB of genomes).  Here four threads copy slices into
         * two pinned buffers of the context while the buffer before is on its way over PCIe. */
        constexpr size_t CHUNK = size_t(32) << 20;
        for (int b = 0; b < 2; ++b) CU(ctx->h_in[b].reserve(CHUNK, ctx->numa_node));
        cudaEvent_t done[2] = {ctx->ev[EV_COPY], ctx->ev[EV_COPY + 1]};       /* free outside a run */
        bool used[2] = {false, false};
        int b = 0;
        for (size_t at = 0; at < nbytes; at += CHUNK, b ^= 1) {
            const size_t n = std::min(CHUNK, nbytes - at);
            if (used[b]) CU(cudaEventSynchronize(done[b]));                    /* the copy that read this buffer has finished */
            uint8_t *dst = static_cast<uint8_t *>(ctx->h_in[b].p);
            std::thread th[3];
            const size_t part = (n + 3) / 4;
            for (int t = 0; t < 3; ++t) th[t] = std::thread([=] { const size_t lo = std::min(n, part * (t + 1)), hi = std::min(n, part * (t + 2)); memcpy(dst + lo, s->bases + at + lo, hi - lo); });
            memcpy(dst, s->bases + at, std::min(n, part));
            for (auto &t : th) t.join();
            CU(cudaMemcpyAsync(ctx->d_bases.as<uint8_t>() + at, dst, n, cudaMemcpyHostToDevice, ctx->copy));
            CU(cudaEventRecord(done[b], ctx->copy)); used[b] = true;
        }
        CU(cudaStreamSynchronize(ctx->copy));
    }
    CU(cudaMemset(static_cast<uint8_t *>(ctx->d_bases.p) + nbytes, 'N', 64));
    {   /* 2-bit copy for the fast path + which 32-byte blocks hold only A,C,G,T (anything else takes the byte path) */
        const uint64_t nblocks = (nbytes + 31) / 32, nimpure = (nblocks + 31) / 32 + 1;
        CU(ctx->d_impure.reserve(nimpure * 4));
        CU(ctx->d_packed.reserve((2 * nblocks + 2 * PACK_PAD_WORDS) * 4));     /* readable words on both sides */
        CU(cudaMemsetAsync(ctx->d_impure.p, 0, nimpure * 4, ctx->compute));
        CU(cudaMemsetAsync(ctx->d_packed.p, 0, PACK_PAD_WORDS * 4, ctx->compute));
        CU(cudaMemsetAsync(ctx->d_packed.as<uint32_t>() + PACK_PAD_WORDS + 2 * nblocks, 0, PACK_PAD_WORDS * 4, ctx->compute));
        k_pack<<<unsigned((nblocks + 255) / 256), 256, 0, ctx->compute>>>(ctx->d_bases.as<uint8_t>(), nbytes, ctx->d_packed.as<uint32_t>() + PACK_PAD_WORDS,
                                                                        ctx->d_impure.as<uint32_t>(), nblocks, nimpure);
        CU(cudaGetLastError());
    }
    CU(ctx->d_contig_off.upload(ctx->h_contig_off.data(), ctx->h_contig_off.size(), ctx->compute));
    CU(ctx->d_contig_size.upload(ctx->h_contig_size.data(), ctx->h_contig_size.size(), ctx->compute));
    CU(ctx->d_genome_contig.upload(ctx->h_genome_contig.data(), ctx->h_genome_contig.size(), ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    {   /* does any window take the byte paths at all?  (error injection keeps a byte-wise scratch copy only then) */
        const uint64_t nimpure = ((nbytes + 31) / 32 + 31) / 32 + 1;
        std::vector<uint32_t> map(nimpure);
        CU(cudaMemcpy(map.data(), ctx->d_impure.p, nimpure * 4, cudaMemcpyDeviceToHost));
        ctx->any_impure = std::any_of(map.begin(), map.end(), [](uint32_t w) { return w != 0; });
    }
    ctx->stats.h2d_bytes = nbytes;
    ctx->have_genomes = true; ctx->have_headers = false;
    return SIMMR_OK;
}

int simmr_error_profile_set(simmr_ctx *ctx, const simmr_eprofile_desc *d) {
    if (!ctx || !d) return fail(ctx, SIMMR_E_INVALID, "null argument");
    CU(cudaSetDevice(ctx->device));
    std::string err;
    int rc = build_profile(d, ctx->model, err);
    if (rc) return fail(ctx, rc, err);
    const ProfileModel &m = ctx->model;
    cudaStream_t s = ctx->compute;
    {   /* device layout: quality bytes carry the +33 of util.rs:46-50 already */
        CU(ctx->d_phred.upload(m.phred_entries.data(), m.phred_entries.size(), s));   /* PT_Q32 words hold the printed characters as built */
        CU(ctx->d_qsub.upload(m.qsub_entries.data(), m.qsub_entries.size(), s));
        CU(ctx->d_gap.upload(m.gap_entries.data(), m.gap_entries.size(), s));
        CU(cudaStreamSynchronize(s));
    }
    CU(ctx->d_rl_e.upload(m.rl.entries.data(), m.rl.entries.size(), s)); CU(ctx->d_rl_lo.upload(m.rl.lo.data(), m.rl.lo.size(), s)); CU(ctx->d_rl_hi.upload(m.rl.hi.data(), m.rl.hi.size(), s));
    CU(ctx->d_is_e.upload(m.is.entries.data(), m.is.entries.size(), s)); CU(ctx->d_is_lo.upload(m.is.lo.data(), m.is.lo.size(), s)); CU(ctx->d_is_hi.upload(m.is.hi.data(), m.is.hi.size(), s));
    CU(ctx->d_acc.upload(m.acc_lut, 256, s)); CU(ctx->d_plthr.upload(m.pl_thr, 64, s));
    CU(ctx->d_rodds.upload(m.r_odds.data(), m.r_odds.size(), s)); CU(ctx->d_ralias.upload(m.r_alias.data(), m.r_alias.size(), s));
    CU(ctx->d_roff.upload(m.r_off.data(), m.r_off.size(), s)); CU(ctx->d_rlo.upload(m.r_lo.data(), m.r_lo.size(), s)); CU(ctx->d_rhi.upload(m.r_hi.data(), m.r_hi.size(), s));
    if (m.err.on) {
        CU(ctx->d_err_direct.upload(m.err.direct.data(), m.err.direct.size(), s)); CU(ctx->d_err_ref.upload(m.err.ref.data(), m.err.ref.size(), s));
        CU(ctx->d_err_keep.upload(m.err.keep_thr.data(), m.err.keep_thr.size(), s)); CU(ctx->d_err_row.upload(m.err.row.data(), m.err.row.size(), s));
        CU(ctx->d_err_cb.upload(m.err.cb.data(), m.err.cb.size(), s)); CU(ctx->d_err_entries.upload(m.err.entries.data(), m.err.entries.size(), s));
        CU(ctx->d_err_alts.upload(m.err.alts.data(), m.err.alts.size(), s));
        /* the 2-bit walk's view of the same tables (ErrTab::direct2, ent4) */
        auto packed_alt = [&](uint32_t alt) {                    /* bases as 2-bit codes and their number (N digits are deleted bases) */
            uint32_t bases = 0, n = 0;
            for (uint32_t dg = 0; dg < m.err.k; ++dg) { const uint32_t v = (alt >> (3u * dg)) & 7u; if (v != 4u) { bases |= (v & 3u) << (2u * n); ++n; } }
            return bases | (n << 28);
        };
        if (m.err.entries.size() >= (1u << 28)) return fail(ctx, SIMMR_E_LIMIT, "alt-k-mer table too large");
        std::vector<uint32_t> ent4(m.err.entries.size() * 4, 0), direct2(m.err.direct.size() * 2, 0);
        for (size_t r = 0; r < m.err.row.size(); ++r) {
            const uint32_t ro = m.err.row[r], n = 1u << m.err.cb[r];
            for (uint32_t i = 0; i < n; ++i) {
                const uint32_t e = m.err.entries[ro + i];
                ent4[4 * size_t(ro + i)] = e >> 12; ent4[4 * size_t(ro + i) + 1] = packed_alt(m.err.alts[ro + i]); ent4[4 * size_t(ro + i) + 2] = packed_alt(m.err.alts[ro + (e & 0xfffu)]);
            }
        }
        for (size_t c = 0; c < m.err.direct.size(); ++c) {
            direct2[2 * c] = m.err.direct[c].x;
            if (m.err.direct[c].y) { const uint32_t r = m.err.direct[c].y - 1u; direct2[2 * c + 1] = (m.err.row[r] + 1u) | (m.err.cb[r] << 28); }
        }
        CU(ctx->d_err_direct2.upload(direct2.data(), direct2.size(), s)); CU(ctx->d_err_ent4.upload(ent4.data(), ent4.size(), s));
    }
    CU(cudaStreamSynchronize(s));
    ctx->have_profile = true;
    return SIMMR_OK;
}

uint64_t simmr_total_units(const simmr_ctx *ctx, const uint64_t *reads) {
    if (!ctx || !reads || !ctx->have_genomes || !ctx->have_profile) return 0;
    return unit_prefix(ctx, reads).back();
}

static int simulate_common(simmr_ctx *ctx, const uint64_t *reads, const simmr_run_opts *opts, simmr_sink_fn sink, simmr_device_sink_fn dev_sink, void *user) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!reads) return fail(ctx, SIMMR_E_INVALID, "null reads_per_genome");
    CU(cudaSetDevice(ctx->device));
    rc = ensure_headers(ctx, opts ? opts->header_format : nullptr);
    if (rc) return rc;
    const std::vector<uint64_t> us = unit_prefix(ctx, reads);
    CU(ctx->d_unit_start.upload(us.data(), us.size(), ctx->compute));
    uint64_t ub, ue;
    unit_range(opts, us.back(), ub, ue);
    if (ub > ue || ue > us.back()) return fail(ctx, SIMMR_E_INVALID, "unit range outside the run");
    uint64_t seed;
    if (opts && opts->has_seed) seed = opts->seed;
    else { std::random_device rd; seed = (uint64_t(rd()) << 32) | rd(); }      /* StdRng::from_entropy(), simulate.rs:174 */
    ctx->stats = simmr_stats{};
    if (ub == ue) return SIMMR_OK;
    ReplayDev rd{};
    return run_units(ctx, false, rd, ub, ue, seed, opts, sink, user, false, dev_sink);
}
int simmr_simulate(simmr_ctx *ctx, const uint64_t *reads, const simmr_run_opts *opts, simmr_sink_fn sink, void *user) {
    return simulate_common(ctx, reads, opts, sink, nullptr, user);
}
int simmr_simulate_device(simmr_ctx *ctx, const uint64_t *reads, const simmr_run_opts *opts, simmr_device_sink_fn sink, void *user) {
    return simulate_common(ctx, reads, opts, nullptr, sink, user);
}

int simmr_measure(simmr_ctx *ctx, const uint64_t *reads, const simmr_run_opts *opts, uint64_t *bytes, uint64_t *records) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!reads || !opts || !opts->has_seed) return fail(ctx, SIMMR_E_INVALID, "measuring a shard needs reads_per_genome and a seed");
    CU(cudaSetDevice(ctx->device));
    rc = ensure_headers(ctx, opts->header_format);
    if (rc) return rc;
    const std::vector<uint64_t> us = unit_prefix(ctx, reads);
    CU(ctx->d_unit_start.upload(us.data(), us.size(), ctx->compute));
    uint64_t ub, ue;
    unit_range(opts, us.back(), ub, ue);
    if (ub > ue || ue > us.back()) return fail(ctx, SIMMR_E_INVALID, "unit range outside the run");
    ctx->stats = simmr_stats{};
    if (ub < ue) { ReplayDev rd{}; rc = run_units(ctx, false, rd, ub, ue, opts->seed, opts, nullptr, nullptr, true); if (rc) return rc; }
    if (bytes) *bytes = ctx->stats.bytes;
    if (records) *records = ctx->stats.records;
    return SIMMR_OK;
}

int simmr_replay(simmr_ctx *ctx, const simmr_draws *d, const simmr_run_opts *opts, simmr_sink_fn sink, void *user) {
    int rc = check_ready(ctx);
    if (rc) return rc;
    if (!d) return fail(ctx, SIMMR_E_INVALID, "null draws");
    CU(cudaSetDevice(ctx->device));
    rc = ensure_headers(ctx, opts ? opts->header_format : nullptr);
    if (rc) return rc;
    const ProfileModel &m = ctx->model;
    const uint64_t n = d->n_units;
    if (!d->genome || !d->contig || !d->read_length || !d->start || !d->base_off) return fail(ctx, SIMMR_E_INVALID, "replay needs genome, contig, read_length, start and base_off");
    if (!m.is_long && !d->insert_size) return fail(ctx, SIMMR_E_INVALID, "paired-end replay needs insert_size");
    for (uint64_t u = 0; u < n; ++u) {
        if (d->genome[u] >= ctx->host_view.n_genomes) return fail(ctx, SIMMR_E_INVALID, "replay: genome index out of range");
        const uint32_t g = d->genome[u];
        if (d->contig[u] >= ctx->h_genome_contig[g + 1] - ctx->h_genome_contig[g]) return fail(ctx, SIMMR_E_INVALID, "replay: contig index out of range");
    }
    const uint64_t nslots = d->base_off[n];
    for (uint64_t u = 0; u < n; ++u) {      /* a unit reads rl draw slots per record from base_off[u] on */
        if (d->base_off[u + 1] < d->base_off[u]) return fail(ctx, SIMMR_E_INVALID, "replay: base_off must be non-decreasing");
        if (d->base_off[u + 1] - d->base_off[u] < uint64_t(d->read_length[u]) * (m.is_long ? 1 : 2)) return fail(ctx, SIMMR_E_INVALID, "replay: a unit has fewer draw slots than bases");
    }
    if (m.replay_kind == RP_NORMAL_FLOOR || m.replay_kind == RP_ACCURACY) { if (!d->phred_z) return fail(ctx, SIMMR_E_INVALID, "replay needs phred_z for this profile"); }
    if (m.replay_kind == RP_ALIAS && (!d->phred_cand || !d->phred_u || !d->phred_inbin)) return fail(ctx, SIMMR_E_INVALID, "replay needs phred_cand/phred_u/phred_inbin for custom profiles");
    if (m.mutates && (!d->mut_u || !d->mut_k)) return fail(ctx, SIMMR_E_INVALID, "replay needs mut_u/mut_k for this profile");
    cudaStream_t s = ctx->compute;
    std::vector<uint64_t> zero64(n + 1, 0); std::vector<uint16_t> zero16(n, 0);
    ReplayDev rd{};
    CU(ctx->d_draw[0].upload(d->genome, n, s)); rd.genome = ctx->d_draw[0].as<uint32_t>();
    CU(ctx->d_draw[1].upload(d->contig, n, s)); rd.contig = ctx->d_draw[1].as<uint32_t>();
    CU(ctx->d_draw[2].upload(d->read_length, n, s)); rd.read_length = ctx->d_draw[2].as<uint16_t>();
    CU(ctx->d_draw[3].upload(d->insert_size ? d->insert_size : zero16.data(), n, s)); rd.insert_size = ctx->d_draw[3].as<uint16_t>();
    CU(ctx->d_draw[4].upload(d->start, n, s)); rd.start = ctx->d_draw[4].as<uint64_t>();
    CU(ctx->d_draw[5].upload(d->redraw ? d->redraw : zero64.data(), n, s)); rd.redraw = ctx->d_draw[5].as<uint64_t>();
    CU(ctx->d_draw[6].upload(d->base_off, n + 1, s)); rd.base_off = ctx->d_draw[6].as<uint64_t>();
    if (d->phred_z) { CU(ctx->d_draw[7].upload(d->phred_z, nslots, s)); rd.phred_z = ctx->d_draw[7].as<float>(); }
    if (d->phred_cand) { CU(ctx->d_draw[8].upload(d->phred_cand, nslots, s)); rd.phred_cand = ctx->d_draw[8].as<uint32_t>(); }
    if (d->phred_u) { CU(ctx->d_draw[9].upload(d->phred_u, nslots, s)); rd.phred_u = ctx->d_draw[9].as<double>(); }
    if (d->phred_inbin) { CU(ctx->d_draw[10].upload(d->phred_inbin, nslots, s)); rd.phred_inbin = ctx->d_draw[10].as<uint32_t>(); }
    if (d->mut_u) { CU(ctx->d_draw[11].upload(d->mut_u, nslots, s)); rd.mut_u = ctx->d_draw[11].as<float>(); }
    if (d->mut_k) { CU(ctx->d_draw[12].upload(d->mut_k, nslots, s)); rd.mut_k = ctx->d_draw[12].as<uint8_t>(); }
    rd.acc_lut = ctx->d_acc.as<float>(); rd.pl_thr = ctx->d_plthr.as<float>();
    rd.r_odds = ctx->d_rodds.as<double>(); rd.r_alias = ctx->d_ralias.as<uint32_t>(); rd.r_off = ctx->d_roff.as<uint32_t>();
    rd.r_lo = ctx->d_rlo.as<uint32_t>(); rd.r_hi = ctx->d_rhi.as<uint32_t>();
    rd.r_npos = m.r_off.empty() ? 0 : uint32_t(m.r_off.size() - 1); rd.pl_qmax = m.pl_qmax; rd.kind = m.replay_kind;
    rd.mean_phred = m.mean_phred_f; rd.acc20 = m.acc20;
    /* unit_start is unused in replay but the parameter block wants a valid pointer */
    std::vector<uint64_t> us(ctx->host_view.n_genomes + 1, 0);
    CU(ctx->d_unit_start.upload(us.data(), us.size(), s));
    CU(cudaStreamSynchronize(s));
    uint64_t ub, ue;
    unit_range(opts, n, ub, ue);
    if (ub > ue || ue > n) return fail(ctx, SIMMR_E_INVALID, "unit range outside the draws");
    ctx->stats = simmr_stats{};
    if (ub == ue) return SIMMR_OK;
    return run_units(ctx, true, rd, ub, ue, 0, opts, sink, user);
}

/* ---- several GPUs of one box --------------------------------------------- */
int simmr_multi_create(const int *devices, int n, simmr_multi **out) {
    if (!out || !devices || n < 1) return mfail(nullptr, SIMMR_E_INVALID, "simmr_multi_create needs at least one device");
    for (int a = 0; a < n; ++a) for (int b = 0; b < a; ++b) if (devices[a] == devices[b]) return mfail(nullptr, SIMMR_E_INVALID, "device listed twice");
    simmr_multi *m = new simmr_multi;
    for (int g = 0; g < n; ++g) {
        simmr_ctx *c = nullptr;
        const int rc = simmr_ctx_create(devices[g], &c);
        if (rc) { const std::string why = last_error_slot(); simmr_multi_destroy(m); return mfail(nullptr, rc, why); }
        m->ctx.push_back(c); m->devices.push_back(devices[g]);
    }
    *out = m;
    return SIMMR_OK;
}
void simmr_multi_destroy(simmr_multi *m) {
    if (!m) return;
    for (simmr_ctx *c : m->ctx) simmr_ctx_destroy(c);
    delete m;
}
const char *simmr_multi_last_error(const simmr_multi *m) { return m ? m->err.c_str() : last_error_slot().c_str(); }
int simmr_multi_genomes_load(simmr_multi *m, const simmr_genome_set *set) {
    if (!m) return mfail(nullptr, SIMMR_E_INVALID, "null multi-context");
    return on_every_device(m, [&](int g) { return simmr_genomes_load(m->ctx[g], set); });
}
int simmr_multi_error_profile_set(simmr_multi *m, const simmr_eprofile_desc *d) {
    if (!m) return mfail(nullptr, SIMMR_E_INVALID, "null multi-context");
    return on_every_device(m, [&](int g) { return simmr_error_profile_set(m->ctx[g], d); });
}
uint64_t simmr_multi_total_units(const simmr_multi *m, const uint64_t *reads) { return m && !m->ctx.empty() ? simmr_total_units(m->ctx[0], reads) : 0; }
int simmr_multi_stats_get(const simmr_multi *m, simmr_stats *out) {
    if (!m || !out) return mfail(nullptr, SIMMR_E_INVALID, "null argument");
    *out = m->stats;
    return SIMMR_OK;
}

int simmr_multi_simulate(simmr_multi *m, const uint64_t *reads, const simmr_run_opts *opts, simmr_sink_fn sink, void *user) {
    if (!m) return mfail(nullptr, SIMMR_E_INVALID, "null multi-context");
    if (!reads) return mfail(m, SIMMR_E_INVALID, "null reads_per_genome");
    const size_t G = m->ctx.size();
    simmr_run_opts o = opts ? *opts : simmr_run_opts{};
    if (!o.has_seed) { std::random_device rd; o.seed = (uint64_t(rd()) << 32) | rd(); o.has_seed = 1; }   /* one seed for every device */
    /* per device: headers, unit table */
    uint64_t ub = 0, ue = 0;
    int rc = on_every_device(m, [&](int g) {
        simmr_ctx *ctx = m->ctx[g];
        int r = check_ready(ctx); if (r) return r;
        r = ensure_headers(ctx, o.header_format); if (r) return r;
        const std::vector<uint64_t> us = unit_prefix(ctx, reads);
        CU(ctx->d_unit_start.upload(us.data(), us.size(), ctx->compute));
        if (g == 0) unit_range(&o, us.back(), ub, ue);
        ctx->stats = simmr_stats{};
        return int(SIMMR_OK);
    });
    if (rc) return rc;
    const uint64_t total = simmr_total_units(m->ctx[0], reads);
    if (ub > ue || ue > total) return mfail(m, SIMMR_E_INVALID, "unit range outside the run");
    m->stats = simmr_stats{};
    if (ub == ue) return SIMMR_OK;
    /* every device cuts the range into the same batches: the first context's plan decides their size */
    std::vector<Runner> R;
    for (size_t g = 0; g < G; ++g) R.push_back(Runner{m->ctx[g], false, ReplayDev{}});
    cudaSetDevice(m->devices[0]);
    for (simmr_ctx *c : m->ctx) { c->dbg_guard = getenv("SIMMR_DEBUG_GUARD") != nullptr; c->dbg_sync = getenv("SIMMR_DEBUG_SYNC") != nullptr; c->dbg_serial = getenv("SIMMR_PREP_SERIAL") != nullptr; }
    rc = R[0].init(ub, ue, o.seed, &o, sink != nullptr, false);
    if (rc) return mfail(m, rc, m->ctx[0]->err);
    o.batch_units = R[0].batch;
    std::vector<BatchQueue> q(G);
    std::vector<BatchQueue *> qs;
    for (auto &x : q) qs.push_back(&x);
    bool sink_ok = true;
    std::thread writer;
    if (sink) writer = std::thread([&] { sink_ok = write_batches(qs, m->devices, sink, user); });
    std::vector<simmr_stats> st(G);
    rc = on_every_device(m, [&](int g) {
        int r = g == 0 ? int(SIMMR_OK) : R[g].init(ub, ue, o.seed, &o, sink != nullptr, false);
        if (!r) r = produce(R[g], uint64_t(g), G, sink ? &q[g] : nullptr, st[g]);
        q[g].close();
        if (r) { for (auto &x : q) x.abort.store(true); cudaDeviceSynchronize(); }     /* the other devices stop at their next batch */
        return r;
    });
    if (sink) { writer.join(); for (auto &x : q) x.wait_all_released(); }
    if (!sink_ok) return mfail(m, SIMMR_E_SINK, "sink callback failed");
    if (rc) return rc;
    simmr_stats sum{};
    sum.units = ue - ub;
    for (size_t g = 0; g < G; ++g) {
        sum.records += st[g].records; sum.bases += st[g].bases; sum.bytes += st[g].bytes; sum.clamped += st[g].clamped;
        sum.algorithmic_bytes += st[g].algorithmic_bytes; sum.kernel_launches += st[g].kernel_launches; sum.d2h_bytes += st[g].d2h_bytes;
        sum.device_ms = std::max(sum.device_ms, st[g].device_ms); sum.format_ms = std::max(sum.format_ms, st[g].format_ms);
        m->ctx[g]->stats = st[g];
    }
    m->stats = sum;
    return SIMMR_OK;
}

int simmr_stats_get(const simmr_ctx *ctx, simmr_stats *out) {
    if (!ctx || !out) return fail(nullptr, SIMMR_E_INVALID, "null argument");
    *out = ctx->stats;
    return SIMMR_OK;
}

}  // extern "C"

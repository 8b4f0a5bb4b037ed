/* device_code.cuh — sm_100a kernels of the simmr read-simulation hot path.
 *
 * Path (reference file:line -> kernel):
 *   simulate.rs:165-190, 205-262, 353-375, 464-491  geometry      -> unit_geometry()
 *   fastq.rs:34-57 header interpolation, record size             -> k_unit_sizes, header_tokens(), emit_header()
 *   (no reference equivalent: byte offsets of the FASTQ stream)   -> k_scan_blocks, k_unit_offsets
 *   simulate.rs:261-286 + error_profiles/ *.rs phreds/mutation +
 *   util.rs:15-57 RC / +33 + fastq.rs:61-69 record layout         -> k_format_lanes (short reads), k_format_long,
 *                                                                    k_format_direct (replay of long reads, oversize records)
 *   (load-time) 2-bit copy of the contigs + which 32-byte blocks
 *   hold only A,C,G,T                                            -> k_pack
 *
 * Draw schedule, table formats and the tile layout are specified in DESIGN.md.
 * Compile with -fmad=false: the replay arithmetic (minimal_short.rs:90-100) and
 * the gamma sampler must round exactly like the CPU.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "../../include/simmr_b200.h"
#include "simmr_types.h"

#ifndef SIMMR_KO
#define SIMMR_KO 0          /* development builds knock a piece of work out to see what it costs (profiles/r02_experiment_knockouts.txt); 0 = the product */
#endif

namespace simmr {

/* ---------------------------------------------------------------------- */
/* kernel parameter blocks                                                  */
/* ---------------------------------------------------------------------- */
struct ReplayDev {                 /* device copies of simmr_draws + replay LUTs */
    const uint32_t *genome, *contig;
    const uint16_t *read_length, *insert_size;
    const uint64_t *start, *redraw, *base_off;
    const float *phred_z; const uint32_t *phred_cand; const double *phred_u; const uint32_t *phred_inbin;
    const float *mut_u; const uint8_t *mut_k;
    const float *acc_lut;          /* [256] util.rs:96-98                       */
    const float *pl_thr;           /* [pl_qmax+1] perfect_long thresholds       */
    const double *r_odds; const uint32_t *r_alias, *r_off, *r_lo, *r_hi;   /* rand_distr alias per position */
    uint32_t r_npos, pl_qmax, kind; float mean_phred, acc20;
};

/* long reads are cut into LONG_PIECES pieces of a multiple of 16 bases, each with its own chains: the 32 lanes' work in k_format_long */
constexpr uint32_t LONG_PIECES = 32;
constexpr uint32_t LONG_RING = 8, LONG_AHEAD = 5;   /* k_format_long: words of the 2-bit copy per lane in shared memory, and how far ahead they are fetched */
struct UnitRec; struct UnitExt;
struct RunParams {
    /* genomes (device resident) */
    const uint8_t *bases; const uint64_t *contig_off, *contig_size; const uint32_t *genome_contig;
    const uint32_t *impure;        /* 1 bit per 32-byte block of `bases`: holds a byte other than A,C,G,T */
    const uint32_t *packed;        /* 2 bits per base (A,C,G,T = 0..3), 16 bases per word, same indexing as `bases` */
    /* compiled header templates: block 2 * contig + mate of hdr_stride bytes = ContigHdr, its tokens, their literals (one or
     * two cache lines a record's lane can prefetch from the contig number alone); toks and pool address the same buffer */
    const uint8_t *hdr; uint32_t hdr_stride, hdr_pad_; const HdrTok *toks; const uint8_t *pool;
    uint32_t n_genomes, n_contigs;
    /* unit space */
    const uint64_t *unit_start;    /* [n_genomes+1] first global unit of each genome */
    uint64_t batch_begin; uint32_t batch_units; uint32_t read_id_base;
    /* profile (phred.entries: device layout, quality bytes already +33) */
    PhredTab phred; LenTab rl, is; GammaConst gamma;
    uint32_t min_size, is_long, flags, mutates;
    uint32_t phred_mult;           /* 1 << phred.cb */
    uint32_t rk[20];               /* Philox round keys: seed + r * Weyl constants, r = 0..9 */
    /* per-batch scratch + output */
    uint32_t *sizes; uint64_t *offsets, *block_sums; uint8_t *out;
    uint64_t out_cap;              /* bytes of `out`: a batch whose FASTQ bytes (offsets[batch_units]) exceed it is not written at all */
    struct UnitRec *desc;          /* per-unit geometry written by k_unit_sizes, read by the format kernels (32 bytes) */
    struct UnitExt *desc_ext;      /* what does not fit a UnitRec (clamped windows, positions above 2^32): written and read for flagged units only */
    /* k_kmer_errors takes the reads of every segment of SEG_UNITS units in the order of their sizes (k_segment_sort) */
    uint16_t *seg_perm;
    unsigned long long *counters;  /* [0] clamped [1] algorithmic bytes [2] bases [3] records [4] k_format_long's read counter [5] bases deleted by k-mer errors [6] a batch overflowed */
    /* per-warp stage of k_format_lanes */
    uint32_t stage_cap;            /* bytes of a warp's stage in k_format_lanes */
    /* decimal fields of the header template when every contig has the same ones (fastq.rs:34-57): hdr_fast == 0 means
     * per-contig lists only; kinds 4 bits each, field i in bits 4*i */
    uint32_t hdr_fast, hdr_fields, hdr_kinds;
    uint32_t one;                  /* always 1: a multiplier the compiler cannot fold (mwc1_draw) */
    /* k-mer error injection (custom-long profiles with an alt-k-mer table): the tables, and the edited reads — the
     * format kernel's source instead of the genome — as 2-bit words (reads of A,C,G,T only) or bytes (all others) */
    ErrTab err;
    uint32_t *epacked; uint8_t *ebases;
};

constexpr int SIZES_THREADS = 256;
constexpr int FORMAT_THREADS = 512;   /* most threads of a format CTA (each warp has a private stage) -> register cap of 128 */
constexpr int FORMAT_MIN_CTAS = 1;
constexpr int DIRECT_THREADS = 256;
constexpr uint32_t LONG_EVENTS = 16;      /* k_format_long: substitution events of a piece walked ahead of its lines (more are handled the slow way) */
__host__ __device__ constexpr uint32_t long_extra_smem() {   /* k_format_long: per warp a genome-word ring and two tiles of 8 turns x 32 lines; per lane an event list */
    return uint32_t(DIRECT_THREADS / 32) * (LONG_RING * 128u + 8192u) + LONG_EVENTS * uint32_t(DIRECT_THREADS) * 4u;
}
constexpr uint32_t PACK_PAD_WORDS = 4352; /* readable words before and after `packed`: a lane keeps fetching (and discarding)
                                             for as long as the longest record of its warp lasts, at most 65535 / 16 words */
/* shared-memory layout of a format CTA (byte offsets from the 1024-aligned base): letter tables, two-digit table, gap and
 * quality-given-substitution tables, quality table (per-lane copies or one copy per position), then the warps' stages */
constexpr uint32_t SM_LUT_FWD = 0, SM_LUT_RC = 1024, SM_DEC2 = 2048, SM_GAP = 2304, SM_QSUB = 3328, SM_QTAB = 4352;
constexpr uint32_t STAGE_FRONT = 16;  /* slack before the tile image (a header's leading newline may land there) */
/* how k_format_lanes samples qualities */
enum : int { QM_CONST = 0,            /* one quality for every base                                              */
             QM_LANE = 1,             /* one position, every lane has its own copy of the table (no bank conflicts) */
             QM_POS = 2 };            /* one table per position, a single copy                                   */

/* ---------------------------------------------------------------------- */
/* small PTX helpers                                                        */
/* ---------------------------------------------------------------------- */
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
/* loads from the read-only tables in shared memory: not volatile, the compiler may move them */
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }

__device__ __forceinline__ const ContigHdr &hdr_at(const RunParams &rp, uint32_t contig, uint32_t mate) {
    return *reinterpret_cast<const ContigHdr *>(rp.hdr + size_t(2u * contig + mate) * rp.hdr_stride);
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

/* ---------------------------------------------------------------------- */
/* Philox4x32-10                                                            */
/* ---------------------------------------------------------------------- */
struct Philox4 { uint32_t x, y, z, w; };
/* round keys come from the kernel parameter block (constant bank): no per-call key schedule */
__device__ __forceinline__ Philox4 philox(const uint32_t *rk, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = uint64_t(0xD2511F53u) * c0, p1 = uint64_t(0xCD9E8D57u) * c2;
        c0 = uint32_t(p1 >> 32) ^ c1 ^ rk[2 * r]; c1 = uint32_t(p1); c2 = uint32_t(p0 >> 32) ^ c3 ^ rk[2 * r + 1]; c3 = uint32_t(p0);
    }
    return Philox4{c0, c1, c2, c3};
}
/* draw-schedule addressing: counter = (block, stream, unit) */
__device__ __forceinline__ Philox4 draw(const RunParams &rp, uint64_t unit, uint32_t stream, uint32_t blk) {
    return philox(rp.rk, blk, stream, uint32_t(unit), uint32_t(unit >> 32));
}
__device__ __forceinline__ uint64_t pair64(uint32_t lo, uint32_t hi) { return uint64_t(lo) | (uint64_t(hi) << 32); }

/* Per-base draws: MWC64X generators (D. B. Thomas: x' = lo(A*x + c), c' = hi(A*x + c), A = 4294883355)
 * seeded from Philox words.  Philox block (piece, stream, unit) seeds the two chains of a piece of a read — words 0,1 the quality chain (one step per
 * two bases), words 2,3 the substitution-event chain.  Short reads are one piece (block 0), long reads LONG_PIECES. */
struct Mwc1 { uint32_t x, c; };
/* one draw = (word, level): word = x ^ c picks a table column with its low bits, level = x is compared with the entry;
 * then the chain steps.  The 64-bit addend {c, 0} is made by a widening multiply by `one` (a kernel parameter that is
 * always 1, so the compiler keeps it): both instructions of a step go to the multiply pipe instead of the ALU. */
__device__ __forceinline__ uint32_t mwc1_draw(Mwc1 &m, uint32_t one, uint32_t &level) {
    level = m.x;
    const uint32_t r = m.x ^ m.c;
    asm("{ .reg .b64 cc, t; mul.wide.u32 cc, %1, %2; mad.wide.u32 t, %0, %3, cc; mov.b64 {%0, %1}, t; }" : "+r"(m.x), "+r"(m.c) : "r"(one), "r"(4294883355u));
    return r;
}
__device__ __forceinline__ void mwc1_seed(const RunParams &rp, Mwc1 &qc, Mwc1 &ec, uint32_t piece, uint32_t stream, uint32_t ulo, uint32_t uhi) {
    const Philox4 w = philox(rp.rk, piece, stream, ulo, uhi);
    qc.x = w.x; qc.c = (w.y >> 2) + 1u; ec.x = w.z; ec.c = (w.w >> 2) + 1u;
}
/* PT_Q32 entry of the column the draw's word picked: self iff level < entry */
__device__ __forceinline__ uint32_t q32_pick(uint32_t e, uint32_t level) { return (level < e ? e : e >> 8) & 0xffu; }
/* bases per piece of a read of L bases */
__device__ __forceinline__ uint32_t piece_len(uint32_t L, bool long_read) { return long_read ? 16u * ((L + 511u) >> 9) : L; }

/* ---------------------------------------------------------------------- */
/* samplers                                                                 */
/* ---------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t sample_len(const LenTab &t, uint32_t w_alias, uint32_t w_inbin) {
    if (t.kind == LEN_CONST) return t.value;
    const uint32_t idx = w_alias >> (32 - t.cb);
    const uint32_t frac = (w_alias << t.cb) >> 12;
    const uint32_t e = __ldg(t.entries + idx);
    const uint32_t col = frac < (e >> 12) ? idx : (e & 0xfffu);
    const uint32_t lo = __ldg(t.lo + col), hi = __ldg(t.hi + col);
    return lo + uint32_t((uint64_t(w_inbin) * (uint64_t(hi) - lo + 1)) >> 32);
}

/* ln(x), x>0 normal, from IEEE +,-,*,/,fma only: same bits on CPU and GPU */
__device__ __forceinline__ double det_log(double x) {
    long long bits = __double_as_longlong(x);
    int e = int((bits >> 52) & 0x7ff) - 1022;
    double m = __longlong_as_double((bits & 0x000fffffffffffffll) | 0x3fe0000000000000ll);
    if (m < 0.70710678118654752440) { m = m * 2.0; e -= 1; }
    const double s = (m - 1.0) / (m + 1.0), z = s * s;
    double p = 1.0 / 23.0;
    p = fma(p, z, 1.0 / 21.0); p = fma(p, z, 1.0 / 19.0); p = fma(p, z, 1.0 / 17.0); p = fma(p, z, 1.0 / 15.0);
    p = fma(p, z, 1.0 / 13.0); p = fma(p, z, 1.0 / 11.0); p = fma(p, z, 1.0 / 9.0); p = fma(p, z, 1.0 / 7.0);
    p = fma(p, z, 1.0 / 5.0); p = fma(p, z, 1.0 / 3.0);
    const double t = fma(p * z, s, s);
    return fma(double(e), 0.69314718055994530942, 2.0 * t);
}
__device__ __forceinline__ double u01_open(uint32_t w) { return (double(w) + 0.5) * (1.0 / 4294967296.0); }
__device__ __forceinline__ uint32_t sat_u16(double x) { return !(x == x) || x <= 0.0 ? 0u : (x >= 65535.0 ? 65535u : uint32_t(x)); }

/* one Marsaglia-Tsang attempt (rand_distr gamma.rs GammaLargeShape) fed by one
 * Philox block: polar normal from w.x,w.y, acceptance uniform from w.z */
__device__ __forceinline__ bool gamma_attempt(const GammaConst &g, Philox4 w, uint32_t &rl) {
    const double v1 = 2.0 * u01_open(w.x) - 1.0, v2 = 2.0 * u01_open(w.y) - 1.0;
    const double s = v1 * v1 + v2 * v2;
    if (s >= 1.0 || s == 0.0) return false;
    const double x = v1 * sqrt(-2.0 * det_log(s) / s);
    const double t = 1.0 + g.c * x;
    if (t <= 0.0) return false;
    const double v = t * t * t, uu = u01_open(w.z), x2 = x * x;
    if (uu < 1.0 - 0.0331 * (x2 * x2) || det_log(uu) < 0.5 * x2 + g.d * (1.0 - v + det_log(v))) {
        rl = sat_u16(floor(g.d * v * g.scale));
        return true;
    }
    return false;
}

/* ---------------------------------------------------------------------- */
/* geometry of one unit — simulate.rs:205-262 (PE), :353-375 + :464-491 (long) */
/* ---------------------------------------------------------------------- */
struct UnitGeo {
    uint64_t a[2];        /* absolute offset in `bases` of each window (forward strand) */
    uint64_t sp[2], ep[2];/* start_position / end_position printed in the header        */
    uint32_t L[2];        /* bases in each record                                       */
    uint32_t contig, id, nrec, clamped;
    uint32_t rl, pad_;    /* drawn read length (replay slots of mate 2 start rl after mate 1) */
};

/* A batch of long reads is budgeted below its largest possible size (engine.cu): if its bytes exceed the staging buffer
 * after all, the format kernels write nothing and raise counters[6]; the run then fails with SIMMR_E_LIMIT. */
__device__ __forceinline__ bool batch_overflows(const RunParams &rp) {
    if (rp.offsets[rp.batch_units] <= rp.out_cap) return false;
    if (blockIdx.x == 0 && threadIdx.x == 0) rp.counters[6] = 1ull;
    return true;
}
/* what the format kernels need of a unit, in registers */
struct UnitDesc {
    uint64_t a0, a1;       /* window offsets in `bases`                                     */
    uint64_t p0, p1;       /* PE: fwd_start, rev_end (simulate.rs:233,241-258); long: start, end */
    uint32_t contig;
    uint16_t rl, L0, L1, H0, H1;
    uint16_t flags;        /* 1: window 0 impure, 2: window 1 impure, 4: two records, 8: a position needs more than 32 bits, 16: UnitExt holds a1, p0, p1 */
};
/* ... and in memory: 32 bytes per unit (one sector), written once by k_unit_sizes.  Window 1 starts at a0 - p0 + p1
 * unless a window was clamped at its contig's end; such units, and units of contigs beyond 2^32 bases, also have a UnitExt. */
struct __align__(16) UnitRec {
    uint64_t a0;
    uint32_t p0, p1;
    uint32_t contig;
    uint16_t rl, L0, L1, H0, H1, flags;
};
struct __align__(16) UnitExt { uint64_t a1, p0, p1, pad_; };
static_assert(sizeof(UnitRec) == 32 && sizeof(UnitExt) == 32, "descriptor layout");
/* v0, v1 = the UnitRec of unit i of the batch as stored */
__device__ __forceinline__ UnitDesc unpack_unit(const RunParams &rp, uint4 v0, uint4 v1, uint32_t i) {
    UnitDesc d;
    d.a0 = pair64(v0.x, v0.y); d.p0 = v0.z; d.p1 = v0.w;
    d.contig = v1.x; d.rl = uint16_t(v1.y); d.L0 = uint16_t(v1.y >> 16); d.L1 = uint16_t(v1.z); d.H0 = uint16_t(v1.z >> 16);
    d.H1 = uint16_t(v1.w); d.flags = uint16_t(v1.w >> 16);
    d.a1 = d.a0 - d.p0 + d.p1;
    if (d.flags & 16u) {
        const uint4 *e = reinterpret_cast<const uint4 *>(rp.desc_ext + i);
        const uint4 e0 = __ldg(e), e1 = __ldg(e + 1);
        d.a1 = pair64(e0.x, e0.y); d.p0 = pair64(e0.z, e0.w); d.p1 = pair64(e1.x, e1.y);
    }
    return d;
}
__device__ __forceinline__ UnitDesc load_unit(const RunParams &rp, uint32_t i) {
    const uint4 *src = reinterpret_cast<const uint4 *>(rp.desc + i);
    const uint4 v0 = __ldg(src), v1 = __ldg(src + 1);
    return unpack_unit(rp, v0, v1, i);
}

template <bool REPLAY>
__device__ void unit_geometry(const RunParams &rp, const ReplayDev &rd, uint64_t u, uint32_t g_lo, uint32_t g_hi, UnitGeo &o) {
    o.id = rp.read_id_base + uint32_t(u);                         /* simulate.rs:85-89 */
    o.clamped = 0;
    uint32_t g, c; uint64_t RL, IS = 0, start, redraw = 0;
    Philox4 w0{};
    if (REPLAY) {
        g = rd.genome[u]; c = __ldg(rp.genome_contig + g) + rd.contig[u];
        RL = rd.read_length[u]; start = rd.start[u]; redraw = rd.redraw[u];
        if (!rp.is_long) IS = rd.insert_size[u];
    } else {
        g = g_lo;                                   /* the block's units span genomes g_lo..g_hi */
        while (g < g_hi && __ldg(rp.unit_start + g + 1) <= u) ++g;
        w0 = draw(rp, u, 0, 0);
        const uint32_t c0 = __ldg(rp.genome_contig + g), nseq = __ldg(rp.genome_contig + g + 1) - c0;
        if (!rp.is_long) {
            const Philox4 w1 = draw(rp, u, 0, 1);
            c = c0 + uint32_t(__umul64hi(pair64(w0.x, w0.y), nseq));            /* simulate.rs:182 */
            RL = sample_len(rp.rl, w1.x, w1.y) & 0xffffu;                        /* :211 (`as u16`) */
            IS = sample_len(rp.is, w1.z, w1.w) & 0xffffu;                        /* :212 */
            start = __umul64hi(pair64(w0.z, w0.w), __ldg(rp.contig_size + c) - rp.min_size);   /* :233 */
        } else {
            uint32_t rl = 20000, ncand = 0; bool got = false;
            for (uint32_t a = 0; a < 256 && !got; ++a) {                          /* :356-372 */
                uint32_t r;
                if (!gamma_attempt(rp.gamma, draw(rp, u, 0, 16 + a), r)) continue;
                uint32_t cnt = 0;
                for (uint32_t k = 0; k < nseq; ++k) cnt += __ldg(rp.contig_size + c0 + k) > r;
                if (!cnt) continue;
                rl = r; ncand = cnt; got = true;
            }
            if (!got) for (uint32_t k = 0; k < nseq; ++k) ncand += __ldg(rp.contig_size + c0 + k) > rl;
            uint32_t pick = uint32_t(__umul64hi(pair64(w0.x, w0.y), ncand));     /* :375 */
            c = c0;
            for (uint32_t k = 0, seen = 0; k < nseq; ++k)
                if (__ldg(rp.contig_size + c0 + k) > rl) { if (seen == pick) { c = c0 + k; break; } ++seen; }
            RL = rl; start = 0;
        }
    }
    o.contig = c; o.rl = uint32_t(RL); o.pad_ = 0;
    const uint64_t base = __ldg(rp.contig_off + c), phys = __ldg(rp.contig_off + c + 1) - base;
    const uint64_t n = __ldg(rp.contig_size + c);
    if (!rp.is_long) {
        const uint64_t R = rp.min_size;
        uint64_t fs = start, fe = start + RL, lo, hi;
        if (fs + IS >= n || fs + IS + RL >= n) {                                  /* simulate.rs:241-247 */
            if (!REPLAY) { const Philox4 w2 = draw(rp, u, 0, 2); redraw = fs + __umul64hi(pair64(w2.x, w2.y), (n - R) - fs); }
            lo = redraw; hi = lo + RL;
        } else if (int32_t(uint32_t(fs + IS) - uint32_t(RL)) < 0) {               /* :250-251 */
            lo = 0; hi = RL;
        } else { lo = fs + IS - RL; hi = fs + IS; }                               /* :253-256 */
        o.sp[0] = fs; o.ep[0] = fe;                                               /* :287-292 */
        o.sp[1] = lo + RL; o.ep[1] = lo;                                          /* :293-298 */
        if (fe > phys) { fe = phys; o.clamped = 1; }
        if (fs > fe)   { fs = fe;   o.clamped = 1; }
        if (hi > phys) { hi = phys; o.clamped = 1; }
        if (lo > hi)   { lo = hi;   o.clamped = 1; }
        o.a[0] = base + fs; o.L[0] = uint32_t(fe - fs);
        o.a[1] = base + lo; o.L[1] = uint32_t(hi - lo);
        o.nrec = 2;
    } else {
        uint64_t end;
        if (REPLAY) { end = start + RL; if (end >= n) end = redraw; }             /* :484-489 */
        else if (rp.flags & SIMMR_FLAG_COMPAT_LONG_START) {
            start = RL ? __umul64hi(pair64(w0.z, w0.w), RL) : 0;                  /* :484 */
            end = start + RL;
            if (end >= n) { const Philox4 w1 = draw(rp, u, 0, 1); end = start + __umul64hi(pair64(w1.x, w1.y), n - start); }
        } else { start = __umul64hi(pair64(w0.z, w0.w), n - RL); end = start + RL; }
        if (end > phys) { end = phys; o.clamped = 1; }
        if (start > end) start = end;
        o.sp[0] = start; o.ep[0] = end;                                           /* :513-518 */
        o.a[0] = base + start; o.L[0] = uint32_t(end - start);
        o.a[1] = 0; o.L[1] = 0; o.sp[1] = o.ep[1] = 0;
        o.nrec = 1;
    }
}

/* does [a, a+L) touch a 32-byte block flagged as holding anything but A,C,G,T? */
__device__ __forceinline__ bool window_impure(const RunParams &rp, uint64_t a, uint32_t L) {
    if (!L) return false;
    const uint64_t b0 = a >> 5, b1 = (a + L - 1) >> 5;
    if (b1 - b0 < 32) {                          /* up to 1024 bases: the window's bits sit in two words (the map has a spare word at its end) */
        const uint32_t nbits = uint32_t(b1 - b0) + 1u;
        const uint32_t bits = __funnelshift_r(__ldg(rp.impure + (b0 >> 5)), __ldg(rp.impure + (b0 >> 5) + 1), uint32_t(b0) & 31u);
        return (bits & (0xffffffffu >> (32u - nbits))) != 0u;
    }
    for (uint64_t w = b0 >> 5; w <= (b1 >> 5); ++w) {
        uint32_t bits = __ldg(rp.impure + w);
        if (w == (b0 >> 5)) bits &= 0xffffffffu << (b0 & 31);
        if (w == (b1 >> 5)) bits &= 0xffffffffu >> (31 - (b1 & 31));
        if (bits) return true;
    }
    return false;
}

/* ---------------------------------------------------------------------- */
/* headers — fastq.rs:34-57                                                 */
/* ---------------------------------------------------------------------- */
__constant__ uint32_t POW10[10] = {1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u, 1000000000u};
/* decimal digits of v: floor(log10) guessed from the bit length (1233 / 4096 ~ log10 2), corrected by one compare */
__device__ __forceinline__ uint32_t dec_digits32(uint32_t v) {
    const uint32_t guess = ((32u - __clz(v | 1u)) * 1233u) >> 12;      /* 0..9: digits - 1 or digits */
    return guess + ((v | 1u) >= POW10[guess] ? 1u : 0u);
}
__device__ __forceinline__ uint32_t dec_digits(uint64_t v) {
    if (!(v >> 32)) return dec_digits32(uint32_t(v));
    uint32_t d = 1;
    while (v >= 10) { v /= 10; ++d; }
    return d;
}
__device__ __forceinline__ uint32_t header_len(const ContigHdr &ch, uint32_t id, uint64_t sp, uint64_t ep) {
    return ch.lit_bytes + ch.n_id * dec_digits32(id) + ch.n_sp * dec_digits(sp) + ch.n_ep * dec_digits(ep);
}
__device__ __forceinline__ uint32_t put_dec(uint8_t *dst, uint64_t v) {
    const uint32_t d = dec_digits(v);
    uint32_t i = d;
    while (v >> 32) { dst[--i] = uint8_t('0' + v % 10); v /= 10; }
    uint32_t w = uint32_t(v);
    while (i) { const uint32_t q = __umulhi(w, 0xCCCCCCCDu) >> 3; dst[--i] = uint8_t('0' + (w - q * 10u)); w = q; }
    return d;
}
/* Instantiates the compiled template of a contig (fastq.rs:34-57).  Literals start on
 * 4-byte boundaries of the pool so they can be fetched a word at a time. */
__device__ uint32_t emit_header(uint8_t *dst, const RunParams &rp, const ContigHdr &ch, uint32_t id, uint64_t sp, uint64_t ep) {
    uint32_t n = 0;
    const HdrTok *t = rp.toks + ch.tok_off;
    for (uint32_t i = 0; i < ch.ntok; ++i) {
        const HdrTok k = t[i];
        switch (k.type) {
        case TOK_LIT: {
            const uint32_t *s = reinterpret_cast<const uint32_t *>(rp.pool + k.off);
            for (uint32_t j = 0; j < k.len; j += 4) {
                const uint32_t w = __ldg(s + (j >> 2));
                dst[n + j] = uint8_t(w);
                if (j + 1 < k.len) dst[n + j + 1] = uint8_t(w >> 8);
                if (j + 2 < k.len) dst[n + j + 2] = uint8_t(w >> 16);
                if (j + 3 < k.len) dst[n + j + 3] = uint8_t(w >> 24);
            }
            n += k.len; break; }
        case TOK_READ_ID: n += put_dec(dst + n, id); break;
        case TOK_START: n += put_dec(dst + n, sp); break;
        default: n += put_dec(dst + n, ep); break;                /* TOK_END; rc and pair are part of the literals */
        }
    }
    return n;
}

/* The same header by a whole warp (k_format_long: one record per warp at a time, and lane 0 writing it byte by byte,
 * one dependent load after the other, was 15 % of that kernel).  Lane t sizes token t, a scan places the tokens, then every
 * lane renders bytes lane, lane + 32, ... of the header: a literal's byte from the pool, a field's digit by division.
 * False (nothing written) for what it does not cover — more than 32 tokens, a position beyond 2^32. */
__device__ __forceinline__ bool emit_header_warp(uint8_t *dst, const RunParams &rp, const ContigHdr &ch, uint32_t id, uint64_t sp, uint64_t ep, uint32_t lane) {
    const uint32_t ntok = ch.ntok;
    if (ntok > 32u || ((sp | ep) >> 32)) return false;
    uint32_t type = TOK_LIT, off = 0, len = 0, val = 0;
    if (lane < ntok) {
        const uint2 k = __ldg(reinterpret_cast<const uint2 *>(rp.toks + ch.tok_off) + lane);      /* {off, len | type << 16} */
        type = (k.y >> 16) & 0xffu; off = k.x;
        if (type == TOK_LIT) len = k.y & 0xffffu;
        else { val = type == TOK_READ_ID ? id : type == TOK_START ? uint32_t(sp) : uint32_t(ep); len = dec_digits32(val); }
    }
    uint32_t end = len;
#pragma unroll
    for (uint32_t o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, end, o); if (lane >= o) end += y; }
    const uint32_t start = end - len, H = __shfl_sync(0xffffffffu, end, 31);
    for (uint32_t p0 = 0; p0 < H; p0 += 32u) {
        const uint32_t pos = p0 + lane;
        uint32_t t = 0;                                     /* the token holding byte pos: as many tokens end at or before it */
        for (uint32_t i = 0; i + 1u < ntok; ++i) t += __shfl_sync(0xffffffffu, end, i) <= pos ? 1u : 0u;
        const uint32_t ts = __shfl_sync(0xffffffffu, start, t), tl = __shfl_sync(0xffffffffu, len, t), ty = __shfl_sync(0xffffffffu, type, t);
        const uint32_t to = __shfl_sync(0xffffffffu, off, t), tv = __shfl_sync(0xffffffffu, val, t);
        if (pos < H) {
            const uint32_t idx = pos - ts;
            dst[pos] = ty == TOK_LIT ? __ldg(rp.pool + to + idx) : uint8_t('0' + (tv / POW10[tl - 1u - idx]) % 10u);
        }
    }
    return true;
}

/* ---------------------------------------------------------------------- */
/* per-base work, generic (any byte values, replay, long reads)             */
/* ---------------------------------------------------------------------- */
/* util.rs:15-27 */
__device__ __forceinline__ uint8_t complement(uint8_t b) {
    const uint8_t up = b & 0xDFu;
    if (up == 'A' || up == 'T') return b ^ 0x15u;
    if (up == 'C' || up == 'G') return b ^ 0x04u;
    return b;
}
/* minimal_short.rs:124-131: the three other bases in alphabetical order */
__device__ __forceinline__ uint8_t substitute(uint8_t b, uint32_t k) {
    uint32_t rank;
    switch (b) { case 'A': rank = 0; break; case 'C': rank = 1; break; case 'G': rank = 2; break; case 'T': rank = 3; break; default: return b; }
    const uint32_t pick = k + (k >= rank);
    return uint8_t(0x54474341u >> (8 * pick));                    /* "ACGT" */
}
/* native mode: m in 1..3 picks the substituted base as rank(b) ^ m (A,C,G,T = 0..3) — a relabelling of
 * choose()'s index (minimal_short.rs:125) that keeps it uniform over the three other bases */
__device__ __forceinline__ uint8_t substitute_native(uint8_t b, uint32_t m) {
    uint32_t rank;
    switch (b) { case 'A': rank = 0; break; case 'C': rank = 1; break; case 'G': rank = 2; break; case 'T': rank = 3; break; default: return b; }
    return uint8_t(0x54474341u >> (8 * (rank ^ m)));
}
__device__ __forceinline__ uint32_t sat_u8f(float x) { return !(x == x) || x <= 0.f ? 0u : (x >= 255.f ? 255u : uint32_t(x)); }

/* replay: the reference's arithmetic on one recorded draw slot */
__device__ __forceinline__ void replay_base(const RunParams &rp, const ReplayDev &rd, uint64_t slot, uint32_t i, uint32_t &q, uint32_t &m) {
    switch (rd.kind) {
    case RP_CONST: q = rp.phred.const_q; break;                                   /* perfect_short.rs:43 */
    case RP_NORMAL_FLOOR: q = sat_u8f(floorf(rd.mean_phred + 10.0f * rd.phred_z[slot])); break;   /* minimal_short.rs:90-100 */
    case RP_ACCURACY: {                                                           /* perfect_long.rs:66-77 */
        const float a = fminf(rd.acc20 + 0.05f * rd.phred_z[slot], 0.9999f);
        q = 0;
        for (uint32_t k = 1; k <= rd.pl_qmax; ++k) q += a >= rd.pl_thr[k];
        break; }
    default: {                                                                    /* custom_short.rs:108-129, 339-349 */
        const uint32_t pos = i < rd.r_npos ? i : rd.r_npos - 1;
        const uint32_t o = rd.r_off[pos], nw = rd.r_off[pos + 1] - o;
        const uint32_t cand = rd.phred_cand[slot] % nw;
        const uint32_t bin = rd.phred_u[slot] < rd.r_odds[o + cand] ? cand : rd.r_alias[o + cand];
        const uint32_t lo = rd.r_lo[o + bin], hi = rd.r_hi[o + bin];
        q = (lo + uint32_t((uint64_t(rd.phred_inbin[slot]) * (uint64_t(hi) - lo + 1)) >> 32)) & 0xffu;
        break; }
    }
    m = 0;
    if (rp.mutates && rd.mut_u[slot] > rd.acc_lut[q]) m = 1 + rd.mut_k[slot] % 3;   /* minimal_short.rs:123-125 */
}

/* Native mode with byte stores, one thread per piece of the read (thread thr of nthr takes pieces thr, thr + nthr, ...):
 * the piece's quality chain, then its substitution events.  Used for windows that hold anything but
 * A,C,G,T, for records that do not fit a stage and behind SIMMR_NO_LONG_KERNEL; same bytes as the word paths of
 * k_format_lanes and k_format_long.  Tables are read from global memory (entries hold the characters as printed). */
/* base j of an edited read (k_kmer_errors): 2-bit scratch for reads of A,C,G,T only, byte scratch for the others */
__device__ __forceinline__ uint8_t err_base(const RunParams &rp, bool pure, uint64_t sb, uint32_t j) {
    if (!pure) return rp.ebases[sb + j];
    const uint64_t at = sb + j;
    return uint8_t(0x54474341u >> (8u * ((rp.epacked[at >> 4] >> (2u * (uint32_t(at) & 15u))) & 3u)));
}
template <bool PUNCT>
__device__ void emit_body_pieces(const RunParams &rp, uint8_t *dst, uint64_t unit, uint32_t mate, uint64_t a, uint32_t L, uint32_t thr, uint32_t nthr,
                                 uint32_t Ls = 0xffffffffu, bool pure = false) {
    const bool edited = Ls != 0xffffffffu;            /* long read after k-mer error injection: Ls bases from the scratch copy at a, L qualities */
    if (!edited) Ls = L;
    uint8_t *seq = dst + 1, *qual = dst + Ls + 4;
    if (PUNCT && thr == 0) { dst[0] = '\n'; dst[Ls + 1] = '\n'; dst[Ls + 2] = '+'; dst[Ls + 3] = '\n'; dst[Ls + L + 4] = '\n'; }   /* fastq.rs:61-69 */
    const uint8_t *src = rp.bases + a;
    const bool rc = mate != 0, sampled = rp.phred.kind != PT_CONST;
    const uint32_t mask = rp.phred_mult - 1u, last = rp.phred.npos - 1u, S = piece_len(L, rp.is_long != 0);
    if (edited) for (uint32_t j = thr; j < Ls; j += nthr) seq[j] = err_base(rp, pure, a, j);
    for (uint32_t piece = thr; piece * uint64_t(S) < L; piece += nthr) {
        const uint32_t lo = piece * S, hi = min(L, lo + S);
        Mwc1 qc{}, ec{};
        if (sampled) mwc1_seed(rp, qc, ec, piece, 1 + mate, uint32_t(unit), uint32_t(unit >> 32));
        uint32_t word = 0, lv_odd = 0;                       /* one step of the quality chain serves two bases (quals4) */
        for (uint32_t i = lo; i < hi; ++i) {
            uint32_t qc8 = (rp.phred.const_q + 33u) & 0xffu;
            if (sampled) {
                uint32_t lv = lv_odd, col = (word >> rp.phred.cb) & mask;
                if (!((i - lo) & 1u)) { lv_odd = qc.c; word = mwc1_draw(qc, rp.one, lv); col = word & mask; }
                qc8 = q32_pick(__ldg(rp.phred.entries + (min(i, last) << rp.phred.cb) + col), lv);
            }
            qual[i] = uint8_t(qc8);
            if (edited) continue;
            const uint8_t b = __ldg(src + i);
            if (rc) seq[L - 1 - i] = complement(b); else seq[i] = b;           /* util.rs:31-37 */
        }
        if (!sampled || !rp.phred.has_sub || edited) continue;                  /* profiles with a k-mer table never substitute (custom_long.rs:238-245) */
        for (uint32_t next = lo; next < hi;) {                                  /* three draws of the event chain per step */
            uint32_t l1, l2, l3;
            const uint32_t r1 = mwc1_draw(ec, rp.one, l1), r2 = mwc1_draw(ec, rp.one, l2), r3 = mwc1_draw(ec, rp.one, l3);
            const uint32_t g = q32_pick(__ldg(rp.phred.gap + (r1 & (GAP_COLS - 1u))), l1);
            if (g == GAP_NONE) { next += GAP_NONE; continue; }
            const uint32_t at = next + g;
            if (at >= hi) break;
            qual[at] = uint8_t(q32_pick(__ldg(rp.phred.qsub + (r2 & mask)), l2));
            uint8_t *sb = rc ? seq + (L - 1 - at) : seq + at;                  /* mutated before the reverse complement (simulate.rs:269-286) */
            const uint8_t fwd = rc ? complement(*sb) : *sb, nb = substitute_native(fwd, 1u + __umulhi(r3, 3u));
            *sb = rc ? complement(nb) : nb;
            next = at + 1u;
        }
    }
}

/* Replay: the sequence and quality lines of one record from recorded draws, byte stores (and, when PUNCT, the five
 * fixed bytes around them); `dst` = the byte after the header, in shared or global memory.  Base i of the window
 * takes draw slot slot0 + i; for mate 2 the (mutated) base is complemented and lands at L-1-i, the quality stays at i
 * (simulate.rs:283-285).  Thread thr of nthr takes bases thr, thr + nthr, ... */
template <bool PUNCT>
__device__ __forceinline__ void emit_body_replay(const RunParams &rp, const ReplayDev &rd, uint8_t *dst, uint32_t mate, uint64_t a, uint32_t L,
                                                 uint64_t slot0, uint32_t thr, uint32_t nthr) {
    uint8_t *seq = dst + 1, *qual = dst + L + 4;
    if (PUNCT && thr == 0) { dst[0] = '\n'; dst[L + 1] = '\n'; dst[L + 2] = '+'; dst[L + 3] = '\n'; dst[2 * L + 4] = '\n'; }   /* fastq.rs:61-69 */
    const uint8_t *src = rp.bases + a;
    const bool rc = mate != 0;
    for (uint32_t i = thr; i < L; i += nthr) {
        uint32_t q, m; replay_base(rp, rd, slot0 + i, i, q, m);
        uint8_t b = __ldg(src + i);
        if (m) b = substitute(b, m - 1);
        qual[i] = uint8_t((q + 33) & 0xffu);                                    /* util.rs:46-50 */
        if (rc) seq[L - 1 - i] = complement(b); else seq[i] = b;
    }
}

__device__ __forceinline__ uint32_t lds32m(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds8(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds16m(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint2 lds64m(uint32_t addr) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t addr) { uint32_t v; asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }

/* ---------------------------------------------------------------------- */
/* headers, fast path: one thread per record, aligned word stores           */
/* ---------------------------------------------------------------------- */
/* Byte stream writer into shared memory.  `carry` holds the `fill` (0..3) bytes not yet
 * stored, top-aligned; every append funnel-shifts them in front of the new bytes. */
struct ByteStream { uint32_t addr, carry, fill; };
__device__ __forceinline__ void bs_open(ByteStream &b, uint32_t at) {        /* first byte goes to shared address `at` */
    b.addr = at & ~3u; b.fill = at & 3u;
    b.carry = b.fill ? lds32m(b.addr) << (32u - 8u * b.fill) : 0u;           /* keep the bytes already in that word */
}
__device__ __forceinline__ void bs_word(ByteStream &b, uint32_t v) {         /* 4 bytes */
    sts32(b.addr, __funnelshift_l(b.carry, v, 8u * b.fill)); b.addr += 4u; b.carry = v;
}
__device__ __forceinline__ void bs_put(ByteStream &b, uint32_t v, uint32_t n) {   /* low n (1..4) bytes of v */
    const uint32_t out = __funnelshift_l(b.carry, v, 8u * b.fill), nf = b.fill + n;
    if (nf >= 4u) { sts32(b.addr, out); b.addr += 4u; }
    b.carry = __funnelshift_l(b.carry, v, 32u - 8u * n); b.fill = nf & 3u;
}
__device__ __forceinline__ void bs_close(ByteStream &b) {                     /* the rest of the last word keeps its bytes */
    if (b.fill) { const uint32_t keep = 0xffffffffu << (8u * b.fill); sts32(b.addr, (lds32m(b.addr) & keep) | (b.carry >> (32u - 8u * b.fill))); }
}
/* four decimal digits of c < 10000 as ASCII, most significant in the low byte */
__device__ __forceinline__ uint32_t four_digits(uint32_t dec2_s, uint32_t c) {
    const uint32_t hi = (c * 5243u) >> 19, lo = c - hi * 100u;              /* c / 100 for c < 43699 */
    return lds16(dec2_s + 2u * hi) | (lds16(dec2_s + 2u * lo) << 16);
}
__device__ __forceinline__ void bs_decimal(ByteStream &b, uint32_t dec2_s, uint32_t v) {
    /* digits: floor(log10) guessed from the bit length, corrected by one compare */
    const uint32_t guess = ((32u - __clz(v | 1u)) * 1233u) >> 12;
    const uint32_t d = guess + ((v | 1u) >= lds32(dec2_s + 200u + 4u * guess) ? 1u : 0u);    /* table of 10^k after the digit pairs */
    const uint32_t top = __umulhi(v, 0xABCC7712u) >> 26;                     /* v / 10^8 */
    const uint32_t r8 = v - top * 100000000u, m4 = __umulhi(r8, 0xD1B71759u) >> 13, l4 = r8 - m4 * 10000u;   /* r8 / 10^4 */
    const uint32_t r2 = four_digits(dec2_s, l4);
    if (d <= 4) { bs_put(b, r2 >> (8u * (4u - d)), d); return; }
    const uint32_t r1 = four_digits(dec2_s, m4);
    if (d <= 8) { const uint32_t s = 8u * (8u - d); bs_word(b, __funnelshift_r(r1, r2, s)); if (d > 4) bs_put(b, r2 >> s, d - 4u); return; }
    const uint32_t r0 = lds16(dec2_s + 2u * top) << 16, s = 8u * (12u - d);  /* 12 right-aligned characters r0 r1 r2 */
    bs_word(b, __funnelshift_r(r0, r1, s));
    bs_word(b, __funnelshift_r(r1, r2, s));
    bs_put(b, r2 >> s, d - 8u);
}

/* A record's header template into the stream.  Token lists are canonical (host_model.cpp): literal, field,
 * literal, ..., field, literal with nf <= HDR_MAX_FIELDS decimal fields whose kinds (4 bits each) are the same
 * for every contig; tk = this (contig, mate)'s token list. */
constexpr uint32_t HDR_MAX_FIELDS = 4;
__device__ __forceinline__ void header_tokens(ByteStream &b, const RunParams &rp, uint32_t dec2_s, const uint2 *tk, uint32_t kinds,
                                              uint32_t nf, uint32_t id, uint32_t sp, uint32_t ep) {
    uint2 lit[HDR_MAX_FIELDS + 1];                                               /* {off, len | type << 16}, fetched together */
#pragma unroll
    for (uint32_t i = 0; i <= HDR_MAX_FIELDS; ++i) lit[i] = i <= nf ? __ldg(tk + 2 * i) : make_uint2(0u, 0u);
#pragma unroll
    for (uint32_t i = 0; i <= HDR_MAX_FIELDS; ++i) {
        if (i > nf) break;                                                       /* same for every thread */
        const uint32_t *src = reinterpret_cast<const uint32_t *>(rp.pool + lit[i].x);
        const uint32_t len = lit[i].y & 0xffffu, nw = len >> 2;
#pragma unroll 4
        for (uint32_t j = 0; j < nw; ++j) bs_word(b, __ldg(src + j));
        if (len & 3u) bs_put(b, __ldg(src + nw), len & 3u);
        if (i < nf) {
            const uint32_t kind = (kinds >> (4u * i)) & 15u;
            bs_decimal(b, dec2_s, kind == TOK_READ_ID ? id : kind == TOK_START ? sp : ep);
        }
    }
}

/* ---------------------------------------------------------------------- */
/* kernels                                                                  */
/* ---------------------------------------------------------------------- */
/* load time, one thread per 32-byte block of `bases`: two words of the 2-bit copy and, by
 * ballot, one bit = the block holds a byte that is not A,C,G,T */
__global__ void __launch_bounds__(256) k_pack(const uint8_t *bases, uint64_t nbytes, uint32_t *packed, uint32_t *impure, uint64_t nblocks, uint64_t nimpure) {
    const uint64_t blk = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;      /* grid covers a multiple of 32 blocks */
    bool bad = false;
    if (blk < nblocks) {
        const uint4 *src = reinterpret_cast<const uint4 *>(bases + blk * 32);   /* the tail is padded with 'N' */
        uint32_t out[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint4 v = src[h];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint32_t code = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t c = (w[k] >> (8 * j)) & 0xffu;
                    const uint32_t r = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
                    if (blk * 32 + h * 16 + k * 4 + j < nbytes) bad |= r == 4u;
                    code |= (r & 3u) << (2 * (4 * k + j));
                }
            }
            out[h] = code;
        }
        packed[2 * blk] = out[0]; packed[2 * blk + 1] = out[1];
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && (blk >> 5) < nimpure) impure[blk >> 5] = bits;
}

/* one thread per unit: geometry -> descriptor, FASTQ bytes of the unit, block totals, run statistics */
template <bool REPLAY>
__global__ void __launch_bounds__(SIZES_THREADS) k_unit_sizes(RunParams rp, ReplayDev rd) {
    __shared__ uint64_t sh[32];
    static_assert(SIZES_THREADS == 256, "block totals assume 8 warps");
    const uint32_t i = blockIdx.x * SIZES_THREADS + threadIdx.x;
    uint64_t bytes = 0, alg = 0, bases = 0, clamped = 0, recs = 0;
    /* genome of the block's first and last unit by binary search, once; the threads then walk between the two */
    __shared__ uint32_t g_span[2];
    if (!REPLAY && threadIdx.x < 64) {      /* warp 0: the first unit's genome, warp 1: the last one's — genomes whose first unit is <= u, counted by 32 lanes */
        const uint32_t first = blockIdx.x * SIZES_THREADS, last = min(first + SIZES_THREADS, rp.batch_units) - 1;
        const uint64_t u = rp.batch_begin + (threadIdx.x >> 5 ? last : first);
        uint32_t below = 0;
        for (uint32_t g = threadIdx.x & 31; g < rp.n_genomes; g += 32) below += __ldg(rp.unit_start + g) <= u ? 1u : 0u;
        below = __reduce_add_sync(0xffffffffu, below);
        if ((threadIdx.x & 31) == 0) g_span[threadIdx.x >> 5] = below ? below - 1u : 0u;
    }
    __syncthreads();
    if (i < rp.batch_units) {
        UnitGeo g; unit_geometry<REPLAY>(rp, rd, rp.batch_begin + i, REPLAY ? 0 : g_span[0], REPLAY ? 0 : g_span[1], g);
        uint32_t H[2] = {0, 0};
        uint64_t big = 0;
#pragma unroll
        for (uint32_t r = 0; r < 2; ++r) {                       /* unrolled: the arrays stay in registers */
            if (r >= g.nrec) break;
            H[r] = header_len(hdr_at(rp, g.contig, r), g.id, g.sp[r], g.ep[r]);
            bytes += H[r] + 2ull * g.L[r] + 5;
            alg += H[r] + 3ull * g.L[r] + 5;                     /* BASELINE.md §3 */
            bases += g.L[r];
            big |= g.sp[r] | g.ep[r];
        }
        clamped = g.clamped; recs = g.nrec;
        rp.sizes[i] = uint32_t(bytes);
        const uint64_t p0 = g.sp[0], p1 = g.nrec == 2 ? g.ep[1] : g.ep[0];
        const bool ext = (big >> 32) != 0 || g.clamped != 0;
        const uint32_t flags = (REPLAY || window_impure(rp, g.a[0], g.L[0]) ? 1u : 0u) |
                               (REPLAY || window_impure(rp, g.a[1], g.L[1]) ? 2u : 0u) | (g.nrec == 2 ? 4u : 0u) |
                               ((big >> 32) ? 8u : 0u) | (ext ? 16u : 0u);    /* 8: a position needs more than 32 bits */
        uint4 *dst = reinterpret_cast<uint4 *>(rp.desc + i);
        dst[0] = make_uint4(uint32_t(g.a[0]), uint32_t(g.a[0] >> 32), uint32_t(p0), uint32_t(p1));
        dst[1] = make_uint4(g.contig, (g.rl & 0xffffu) | (g.L[0] << 16), (g.L[1] & 0xffffu) | (H[0] << 16), (H[1] & 0xffffu) | (flags << 16));
        if (ext) {
            uint4 *e = reinterpret_cast<uint4 *>(rp.desc_ext + i);
            e[0] = make_uint4(uint32_t(g.a[1]), uint32_t(g.a[1] >> 32), uint32_t(p0), uint32_t(p0 >> 32));
            e[1] = make_uint4(uint32_t(p1), uint32_t(p1 >> 32), 0u, 0u);
        }
    }
    /* block totals of the five counters: warp sums, one barrier, then five lanes of warp 0 finish (a block's totals fit 32 bits) */
    uint32_t v[5] = {uint32_t(bytes), uint32_t(alg), uint32_t(bases), uint32_t(clamped), uint32_t(recs)};
#pragma unroll
    for (int c = 0; c < 5; ++c) v[c] = __reduce_add_sync(0xffffffffu, v[c]);
    uint32_t *sh32 = reinterpret_cast<uint32_t *>(sh);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 5; ++c) sh32[c * 8 + wid] = v[c];
    }
    __syncthreads();
    if (wid == 0 && lane < 5) {
        uint64_t t = 0;
        for (int w = 0; w < SIZES_THREADS / 32; ++w) t += sh32[lane * 8 + w];
        if (lane == 0) rp.block_sums[blockIdx.x] = t;
        else if (t) atomicAdd(rp.counters + (lane == 1 ? 1 : lane == 2 ? 2 : lane == 3 ? 0 : 3), (unsigned long long)t);
    }
}

/* exclusive scan of the per-block byte totals (single CTA); total -> offsets[batch_units] */
__global__ void __launch_bounds__(1024) k_scan_blocks(RunParams rp, uint32_t nblocks) {
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t carry_sh;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nblocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < nblocks ? rp.block_sums[i] : 0;
        uint64_t x = v;
        for (int o = 1; o < 32; o <<= 1) { const uint64_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint64_t t = warp_tot[lane], s = t;
            for (int o = 1; o < 32; o <<= 1) { const uint64_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            warp_tot[lane] = s - t;
        }
        __syncthreads();
        const uint64_t carry = carry_sh;
        if (i < nblocks) rp.block_sums[i] = carry + warp_tot[wid] + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_sh = carry + warp_tot[wid] + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) rp.offsets[rp.batch_units] = carry_sh;
}

/* per-unit exclusive byte offsets inside the batch */
__global__ void __launch_bounds__(SIZES_THREADS) k_unit_offsets(RunParams rp) {
    __shared__ uint32_t warp_tot[SIZES_THREADS / 32];
    const uint32_t i = blockIdx.x * SIZES_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t v = i < rp.batch_units ? rp.sizes[i] : 0;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < wid; ++w) before += warp_tot[w];
    if (i < rp.batch_units) rp.offsets[i] = rp.block_sums[blockIdx.x] + before + x - v;
}


/* Byte offsets of a segment's SEG_UNITS units (what k_unit_offsets does) and the segment's units in the order of their
 * FASTQ bytes (a counting sort over 256 bins of the segment's range).  k_kmer_errors gives a read to a thread and takes
 * the reads in that order, so that the 32 reads of a warp are about equally long (gamma lengths vary by an order of
 * magnitude).  [A size-ordered variant of the short-read kernel used the same order; it lost more on storing every unit
 * separately than it gained from equal turns, DESIGN.md §6.1.]
 *   seg_perm[seg * SEG_UNITS + r]   unit (index inside the segment) at sorted rank r */
constexpr uint32_t SEG_UNITS = 1024, SEG_THREADS = 256, SEG_PER = SEG_UNITS / SEG_THREADS;
constexpr uint32_t WARP_UNITS = 16;
static_assert(SEG_PER * SIZES_THREADS == SEG_UNITS && SEG_THREADS == SIZES_THREADS, "a segment is SEG_PER blocks of k_unit_sizes");
__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t *warp_tot, uint32_t &total) {   /* SEG_THREADS threads; two barriers */
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    __syncthreads();                                   /* warp_tot may still be read from the previous scan */
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (uint32_t w = 0; w < SEG_THREADS / 32; ++w) { const uint32_t t = warp_tot[w]; all += t; if (w < wid) before += t; }
    total = all;
    return before + x - v;
}
__global__ void __launch_bounds__(SEG_THREADS) k_segment_sort(RunParams rp) {
    __shared__ uint32_t warp_tot[SEG_THREADS / 32], hist[256], red[2][SEG_THREADS / 32];
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const uint32_t seg0 = blockIdx.x * SEG_UNITS, n = min(SEG_UNITS, rp.batch_units - seg0);
    /* 1. sizes -> offsets, block by block of k_unit_sizes (thread t takes unit j * 256 + t of the segment) */
    uint32_t sz[SEG_PER], mn = 0xffffffffu, mx = 0;
#pragma unroll
    for (uint32_t j = 0; j < SEG_PER; ++j) {
        const uint32_t u = j * SEG_THREADS + t;
        sz[j] = u < n ? rp.sizes[seg0 + u] : 0u;
        uint32_t tot;
        const uint32_t before = block_scan_excl(sz[j], warp_tot, tot);
        if (u < n) {
            rp.offsets[seg0 + u] = rp.block_sums[blockIdx.x * SEG_PER + j] + before;
            mn = min(mn, sz[j]); mx = max(mx, sz[j]);
        }
    }
    /* 2. the segment's size range -> 256 bins */
    mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0) { red[0][wid] = mn; red[1][wid] = mx; }
    hist[t] = 0;
    __syncthreads();
#pragma unroll
    for (uint32_t w = 0; w < SEG_THREADS / 32; ++w) { mn = min(mn, red[0][w]); mx = max(mx, red[1][w]); }
    const uint32_t span = mx - mn, sh = span < 256u ? 0u : 24u - __clz(span);     /* (span >> sh) <= 255 */
    uint32_t bin[SEG_PER];
#pragma unroll
    for (uint32_t j = 0; j < SEG_PER; ++j) if (j * SEG_THREADS + t < n) {
        bin[j] = (sz[j] - mn) >> sh;
        atomicAdd(&hist[bin[j]], 1u);
    }
    __syncthreads();
    uint32_t tot;
    const uint32_t mine = hist[t], start = block_scan_excl(mine, warp_tot, tot);
    __syncthreads();
    hist[t] = start;
    __syncthreads();
    /* 3. rank of every unit (any order inside a bin) */
#pragma unroll
    for (uint32_t j = 0; j < SEG_PER; ++j) if (j * SEG_THREADS + t < n) rp.seg_perm[seg0 + atomicAdd(&hist[bin[j]], 1u)] = uint16_t(j * SEG_THREADS + t);
}

/* ---------------------------------------------------------------------- */
/* k-mer error injection — custom_long.rs:187-232 (== custom_short.rs:452-516) */
/* ---------------------------------------------------------------------- */
/* Where read i's edited bases go in the scratch buffers (an index in bases): its FASTQ offset before the edit rounded up
 * to a multiple of 16, plus 16 per read before it — a record holds at least as many bytes as its read has bases, so the
 * regions never overlap, and every region starts on a word of the 2-bit copy. */
__device__ __forceinline__ uint64_t err_scratch_base(uint64_t off0, uint32_t i) { return ((off0 + 15u) & ~uint64_t(15)) + 16ull * i; }
/* one draw of a read's error chain decides window `code` (3-bit digits, row r of the tables): the k-mer itself, or one of the other alternates */
__device__ __forceinline__ uint32_t err_pick_other(const ErrTab &t, uint32_t r, uint32_t word) {
    const uint32_t ro = __ldg(t.row + r), cb = __ldg(t.cb + r);
    const uint32_t idx = cb ? word >> (32u - cb) : 0u, frac = (word << cb) >> 12;
    const uint32_t e = __ldg(t.entries + ro + idx), j = frac < (e >> 12) ? idx : (e & 0xfffu);
    return __ldg(t.alts + ro + j);
}
/* One thread per read, reads taken in the size-sorted order of their segment (k_segment_sort) so that a warp's 32 reads are
 * about equally long.  The reference walks i over the original length, looks at the k bytes at i of the sequence as edited
 * so far and splices the sampled alternate in (N digits of an alternate are deleted bases): everything before i is final,
 * everything from i + k on is still the original's tail, so the state is the window (`win`, `cnt` bases) and how far into
 * the original it reaches (`r`).  One draw of the read's chain (Philox block 0 of stream 3, one step discarded) per
 * window.  Where the reference would slice past the shortened sequence (and panic) the walk stops.
 * Reads of A,C,G,T only: 2-bit codes in, 2-bit words out (epacked), tables by direct index.  Other reads: bytes in,
 * bytes out (ebases), k-mers found by binary search (N is a digit; anything else cannot be encoded, :203-207).
 * Afterwards the unit's descriptor points at the scratch copy (a0), L1 holds the sequence length and sizes[] the
 * record's new bytes. */
__global__ void __launch_bounds__(256) k_kmer_errors(RunParams rp) {
    if (batch_overflows(rp)) return;                   /* the scratch copies are as large as the staging buffer */
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rp.batch_units) return;
    const uint32_t seg0 = t & ~(SEG_UNITS - 1u);
    const uint32_t i = seg0 + __ldg(rp.seg_perm + t);
    UnitDesc d = load_unit(rp, i);
    const ErrTab &et = rp.err;
    const uint32_t L = d.L0, k = et.k;
    const uint64_t unit = rp.batch_begin + i, sb = err_scratch_base(__ldg(rp.offsets + i), i);
    Mwc1 ch;
    { const Philox4 w = philox(rp.rk, 0u, 3u, uint32_t(unit), uint32_t(unit >> 32)); ch.x = w.x; ch.c = (w.y >> 2) + 1u; uint32_t lv; mwc1_draw(ch, rp.one, lv); }
    uint32_t Ls = 0;
    if (!(d.flags & 1u)) {
        /* 2-bit path.  A warp's 32 reads walk in lock step, and with an error in one window of thirteen some lane takes
         * the error branch in nearly every step: the step is kept short — the original's bases wait in a 64-bit buffer
         * (a window is refilled by one masked shift, not base by base), an alternate comes from the tables with its bases
         * already packed and counted (ErrTab::alts2). */
        const uint32_t *src = rp.packed + (d.a0 >> 4);
        const uint32_t off0 = uint32_t(d.a0) & 15u;
        uint64_t sbuf = uint64_t(__ldg(src)) >> (2u * off0);                          /* bases of the original not yet taken, oldest in the low bits */
        uint32_t sleft = 16u - off0;                                                  /* ... how many of them the buffer holds */
        uint32_t nextw = __ldg(++src);                                                /* the word behind the buffer, fetched a refill ahead */
        uint32_t *dst = rp.epacked + (sb >> 4);
        uint32_t ow = 0, ocnt = 0;                                                    /* output word being filled */
        uint32_t win = 0, cnt = 0, r = 0;
        const uint32_t kmask = (1u << (2u * k)) - 1u;
        auto emit = [&](uint32_t b) {
            ow |= b << (2u * ocnt);
            if (++ocnt == 16u) { dst[Ls >> 4] = ow; ow = 0; ocnt = 0; }
            ++Ls;
        };
        for (uint32_t at = 0; at + k <= L; ++at) {                                    /* :196-199 */
            if (sleft <= 16u) { sbuf |= uint64_t(nextw) << (2u * sleft); sleft += 16u; nextw = __ldg(++src); }   /* the padding behind `packed` covers the last fetches */
            const uint32_t need = min(k - cnt, L - r);                                /* k <= 10 < sleft */
            win |= (uint32_t(sbuf) & ((1u << (2u * need)) - 1u)) << (2u * cnt);
            sbuf >>= 2u * need; sleft -= need; cnt += need; r += need;
            if (cnt < k) break;                                                       /* the reference would panic here */
            uint32_t level; const uint32_t word = mwc1_draw(ch, rp.one, level);
            const uint2 ent = __ldg(reinterpret_cast<const uint2 *>(et.direct2) + (win & kmask));
            if (ent.y && level >= ent.x) {                                            /* observed k-mer, and the draw leaves it: :216-228 */
                const uint32_t ro = (ent.y & 0x0fffffffu) - 1u, cb = ent.y >> 28;
                const uint32_t idx = cb ? word >> (32u - cb) : 0u, frac = (word << cb) >> 12;
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(et.ent4) + ro + idx);
                const uint32_t alt = frac < q.x ? q.y : q.z;                          /* the column's alternate or its alias's */
                win = alt & 0x0fffffffu; cnt = alt >> 28;
            }
            if (!cnt) {
                if (r >= L) break;
                win = uint32_t(sbuf) & 3u; cnt = 1; sbuf >>= 2; --sleft; ++r;         /* sleft > 16 - 10 here */
            }
            emit(win & 3u); win >>= 2; --cnt;
        }
        for (; cnt; --cnt) { emit(win & 3u); win >>= 2; }
        while (r < L) {
            if (!sleft) { sbuf = nextw; sleft = 16u; nextw = __ldg(++src); }
            emit(uint32_t(sbuf) & 3u); sbuf >>= 2; --sleft; ++r;
        }
        if (ocnt) dst[Ls >> 4] = ow;
    } else {
        /* byte path */
        const uint8_t *src = rp.bases + d.a0;
        uint8_t *dst = rp.ebases + sb;
        uint8_t win[16]; uint32_t cnt = 0, r = 0;
        for (uint32_t at = 0; at + k <= L; ++at) {
            while (cnt < k && r < L) win[cnt++] = __ldg(src + r++);
            if (cnt < k) break;
            uint32_t level; const uint32_t word = mwc1_draw(ch, rp.one, level);
            uint32_t code = 0; bool ok = true;
            for (uint32_t dg = 0; dg < k; ++dg) {                                     /* encoding.rs:149-179 */
                const uint8_t c = win[dg];
                const uint32_t v = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : c == 'N' ? 4u : 8u;
                ok &= v < 8u; code |= (v & 7u) << (3u * dg);
            }
            if (ok) {
                uint32_t lo = 0, hi = et.n_ref;                                       /* the k-mer's row */
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(et.ref + mid) < code) lo = mid + 1u; else hi = mid; }
                if (lo < et.n_ref && __ldg(et.ref + lo) == code && level >= __ldg(et.keep_thr + lo)) {
                    const uint32_t alt = err_pick_other(et, lo, word);
                    cnt = 0;
                    for (uint32_t dg = 0; dg < k; ++dg) { const uint32_t v = (alt >> (3u * dg)) & 7u; if (v != 4u) win[cnt++] = uint8_t(0x54474341u >> (8u * v)); }
                }
            }
            if (!cnt && r < L) { win[0] = __ldg(src + r++); cnt = 1; }
            if (!cnt) break;
            dst[Ls++] = win[0];
            for (uint32_t j = 1; j < cnt; ++j) win[j - 1] = win[j];
            --cnt;
        }
        for (uint32_t j = 0; j < cnt; ++j) dst[Ls++] = win[j];
        while (r < L) dst[Ls++] = __ldg(src + r++);
    }
    /* the unit now reads from the scratch copy; its record shrinks by the deleted bases */
    uint4 *rec = reinterpret_cast<uint4 *>(rp.desc + i);
    uint4 v0 = rec[0], v1 = rec[1];
    v0.x = uint32_t(sb); v0.y = uint32_t(sb >> 32);
    v1.z = (v1.z & 0xffff0000u) | (Ls & 0xffffu);                                     /* L1 <- sequence length */
    rec[0] = v0; rec[1] = v1;
    rp.sizes[i] -= L - Ls;
    if (L != Ls) atomicAdd(rp.counters + 5, (unsigned long long)(L - Ls));
}
/* per-block byte totals again, after k_kmer_errors changed the sizes */
__global__ void __launch_bounds__(SIZES_THREADS) k_block_sums(RunParams rp) {
    __shared__ uint32_t warp_tot[SIZES_THREADS / 32];
    const uint32_t i = blockIdx.x * SIZES_THREADS + threadIdx.x;
    uint32_t v = i < rp.batch_units ? rp.sizes[i] : 0u;
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { uint64_t s = 0; for (int w = 0; w < SIZES_THREADS / 32; ++w) s += warp_tot[w]; rp.block_sums[blockIdx.x] = s; }
}

/* ---------------------------------------------------------------------- */
/* short reads, word path: one lane per record                              */
/* ---------------------------------------------------------------------- */
/* Quality characters (as printed) of four consecutive bases, the first at window index i0 (the same for every lane of
 * the warp).  One step of the record's quality chain serves two bases: the word x ^ c picks base 2n's column with bits
 * 0..cb-1 and base 2n+1's with bits cb..2cb-1; their levels are x and c.  QM_LANE: the lane's own copy of the single
 * position's table (word col * 32 + lane: every lane stays in its bank); QM_POS: row min(i, npos-1) of the one copy.
 * CBK = cb when it is known at compile time (7: minimal profiles, where the second column's byte offset is then a plain
 * mask of the word; 6: NovaSeq-style blobs), 0 = read from the parameters.  FLAT (QM_POS): none of the four bases is past
 * the table's last row, so their rows follow each other from row i0 on and the address is a column plus a constant —
 * the row arithmetic (a minimum, a shift and an add per look-up) had made the per-position kernel ALU-bound at 79 %. */
template <int QM, int CBK, bool FLAT = false>
__device__ __forceinline__ uint32_t quals4(const RunParams &rp, uint32_t qtab_s, uint32_t mask, Mwc1 &rng, uint32_t i0) {
    uint32_t qw = 0;
    const uint32_t cb = CBK ? uint32_t(CBK) : rp.phred.cb;
#pragma unroll
    for (uint32_t j = 0; j < 4; j += 2) {
        const uint32_t lv0 = rng.x, lv1 = rng.c;
        uint32_t unused;
        const uint32_t r = mwc1_draw(rng, rp.one, unused);
        uint32_t at0, at1;
        if (QM == QM_LANE) {
            at0 = qtab_s + (r & mask) * 128u;
            at1 = CBK == 7 ? qtab_s + (r & 0x3f80u) : qtab_s + ((r >> cb) & mask) * 128u;
        } else if (FLAT) {
            const uint32_t row_s = qtab_s + (i0 << (cb + 2u));                                       /* the same for every lane */
            at0 = row_s + (j << (cb + 2u)) + (r & mask) * 4u;
            at1 = row_s + ((j + 1u) << (cb + 2u)) + ((r >> cb) & mask) * 4u;
        } else {
            const uint32_t last = rp.phred.npos - 1u;                                                /* custom_short.rs:339-345 */
            at0 = qtab_s + ((min(i0 + j, last) << cb) + (r & mask)) * 4u;
            at1 = qtab_s + ((min(i0 + j + 1u, last) << cb) + ((r >> cb) & mask)) * 4u;
        }
        const uint32_t e0 = lds32(at0), e1 = lds32(at1);
        /* byte j of qw <- byte 0 (self) or 1 (alias) of the entry; bytes below j stay */
        const uint32_t keep0 = 0x3210u & ((1u << (4u * j)) - 1u), keep1 = 0x3210u & ((1u << (4u * j + 4u)) - 1u);
        qw = __byte_perm(qw, e0, lv0 < e0 ? (keep0 | (4u << (4u * j))) : (keep0 | (5u << (4u * j))));
        qw = __byte_perm(qw, e1, lv1 < e1 ? (keep1 | (4u << (4u * j + 4u))) : (keep1 | (5u << (4u * j + 4u))));
    }
    return qw;
}

/* Sequence and quality lines of one record by one lane, all 32 lanes of the warp converged (a lane without a record for
 * this path runs with on = false and stores nothing); the warp takes T4 groups of 4 bases of every record, 16 bases per turn.
 * A record body is "\nSEQ\n+\nQUAL\n"; body_s = shared-space address of the '\n' that ends the header.  Both lines are
 * written front to back as 4-byte aligned words: a line's group of 4 characters is funnel-shifted against the previous
 * group onto the word grid of the stage.  The first and last word of a line also cover up to 3 bytes outside it; those
 * belong to this record's header or fixed bytes or to the next record's header, all written in phase 2.
 * Mate 2 (rc) reads its window back to front — 16 bases = one 32-bit window of the 2-bit genome copy, byte-reversed,
 * then the reverse-complement table turns each byte into 4 letters (util.rs:31-37) — while its quality line runs
 * forward like mate 1's: quality i belongs to window base i, never reversed (simulate.rs:283-285).  Substitutions
 * are applied afterwards (substitutions_lane), so the two lines are independent here. */
template <int QM, int CBK>
__device__ __forceinline__ void body_lane(const RunParams &rp, uint32_t qtab_s, uint32_t lut_s, bool on, bool rc, uint32_t body_s,
                                          Mwc1 &rng, uint64_t a, uint32_t L, uint32_t T4, bool pre, uint32_t pre_lo, uint32_t pre_hi) {
    if (!on) { L = 0; a = 0; }
    const uint32_t Ds = body_s + 1u, Dq = body_s + L + 4u;
    const uint32_t os = Ds & 3u, oq = Dq & 3u, sh_s = 8u * os, sh_q = 8u * oq;
    /* word slots of each line still to store: slot k = characters 4k - o .. 4k - o + 3 of the line */
    int32_t ns = on ? int32_t((L + os + 3u) >> 2) : 0, nq = on ? int32_t((L + oq + 3u) >> 2) : 0;
    uint32_t ps = Ds - os, pq = Dq - oq;
    /* 32-bit windows of the 2-bit copy: mate 1 from the window's first base upwards, mate 2 from its last base downwards.
     * The two words under the window slide by one word per turn; the next word is fetched a turn ahead. */
    const int64_t bit = rc ? 2 * int64_t(a + L) - 32 : 2 * int64_t(a);
    const uint32_t *pw = rp.packed + (bit >> 5);
    const uint32_t s2 = uint32_t(bit) & 31u, rev = rc ? 0x0123u : 0x3210u;
    uint32_t w_lo = pre_lo, w_hi = pre_hi;             /* pre: the window's first two words were fetched while the tile before was composed */
    if (!pre) { w_lo = __ldg(pw); w_hi = __ldg(pw + 1); }
    pw += rc ? -1 : 2;
    const int32_t dw = rc ? -1 : 1;
    const uint32_t qconst = ((rp.phred.const_q + 33u) & 0xffu) * 0x01010101u, mask = rp.phred_mult - 1u;
    uint32_t prev_s = 0, prev_q = 0;
    /* one turn = G groups of 4 bases of every record: G = 4 but for the warp's last turn, which is as long as its longest record needs */
    auto turn = [&](auto gc, auto whole, auto flat, uint32_t it) {       /* whole: every lane is inside both lines for all of this turn; flat: quals4 */
        constexpr uint32_t G = decltype(gc)::value;
        constexpr bool WHOLE = decltype(whole)::value, FLAT = decltype(flat)::value;
        const uint32_t X = __byte_perm(__funnelshift_r(w_lo, w_hi, s2), 0u, rev);
        if (G == 4) {   /* mate 1: (lo, hi) <- (hi, new); mate 2: (lo, hi) <- (new, lo).  The padding around `packed` covers the last fetch. */
            const uint32_t nw = __ldg(pw);
            pw += dw;
            const uint32_t keep = rc ? w_lo : w_hi;
            w_lo = rc ? nw : keep; w_hi = rc ? keep : nw;
        }
        uint32_t cs[G], cq[G];                         /* the turn's words of each line, on the stage's word grid */
#pragma unroll
        for (uint32_t h = 0; h < G; ++h) {
            const uint32_t qw = QM == QM_CONST ? qconst : quals4<QM, CBK, FLAT>(rp, qtab_s, mask, rng, 16u * it + 4u * h);
            const uint32_t sw = lds32(lut_s + ((h == 0 ? X << 2 : X >> (8u * h - 2u)) & 0x3fcu));
            cs[h] = __funnelshift_l(prev_s, sw, sh_s); cq[h] = QM == QM_CONST ? qconst : __funnelshift_l(prev_q, qw, sh_q);
            prev_s = sw; prev_q = qw;
        }
        if (WHOLE) {                                   /* no per-word tests (a lane without a record stores nothing at all) */
            if (on) {
#pragma unroll
                for (uint32_t h = 0; h < G; ++h) { sts32(ps + 4u * h, cs[h]); sts32(pq + 4u * h, cq[h]); }
            }
        } else {
#pragma unroll
            for (uint32_t h = 0; h < G; ++h) {
                if (ns > int32_t(h)) sts32(ps + 4u * h, cs[h]);
                if (nq > int32_t(h)) sts32(pq + 4u * h, cq[h]);
            }
        }
        ps += 4u * G; pq += 4u * G; ns -= int32_t(G); nq -= int32_t(G);
    };
    const uint32_t full = (T4 - 1u) >> 2;              /* T4 >= 1 */
    /* turns during which every lane that has a record stores whole groups in both lines: the shortest record's */
    const uint32_t whole = min(full, __reduce_min_sync(0xffffffffu, on ? uint32_t(min(ns, nq)) >> 2 : 0xffffffffu));
    uint32_t it = 0;
    if (QM == QM_POS) {                                /* the turns whose sixteen positions all have a row of their own */
        const uint32_t nflat = min(whole, rp.phred.npos >> 4);
        for (; it < nflat; ++it) turn(std::integral_constant<uint32_t, 4>{}, std::true_type{}, std::true_type{}, it);
    }
    for (; it < whole; ++it) turn(std::integral_constant<uint32_t, 4>{}, std::true_type{}, std::false_type{}, it);
    for (; it < full; ++it) turn(std::integral_constant<uint32_t, 4>{}, std::false_type{}, std::false_type{}, it);
    switch (T4 - 4u * full) {
    case 1: turn(std::integral_constant<uint32_t, 1>{}, std::false_type{}, std::false_type{}, full); break;
    case 2: turn(std::integral_constant<uint32_t, 2>{}, std::false_type{}, std::false_type{}, full); break;
    case 3: turn(std::integral_constant<uint32_t, 3>{}, std::false_type{}, std::false_type{}, full); break;
    default: turn(std::integral_constant<uint32_t, 4>{}, std::false_type{}, std::false_type{}, full); break;
    }
}

/* The substituted bases of one record (minimal_short.rs:104-140 through the factorisation of DESIGN.md §4): geometric
 * gaps along the window; each event replaces the base (rank ^ m, the same for both strands because complementing is
 * rank ^ 3) and its quality character.  Byte stores into the lines body_lane wrote. */
__device__ __forceinline__ void substitutions_lane(const RunParams &rp, uint32_t gap_s, uint32_t qsub_s, bool on, bool rc, uint32_t body_s,
                                                   Mwc1 &ev, uint32_t L) {
    const uint32_t Ds = body_s + 1u, Dq = body_s + L + 4u, mask = rp.phred_mult - 1u;
    uint32_t next = on ? 0u : L;
    /* a step's three draws and two table entries are fetched one step ahead of their use */
    uint32_t l1, l2, l3;
    uint32_t r1 = mwc1_draw(ev, rp.one, l1), r2 = mwc1_draw(ev, rp.one, l2), r3 = mwc1_draw(ev, rp.one, l3);
    uint32_t eg = lds32(gap_s + 4u * (r1 & (GAP_COLS - 1u))), eq = lds32(qsub_s + 4u * (r2 & mask));
    while (next < L) {
        const uint32_t g = q32_pick(eg, l1), qn = q32_pick(eq, l2), m = 1u + __umulhi(r3, 3u);
        r1 = mwc1_draw(ev, rp.one, l1); r2 = mwc1_draw(ev, rp.one, l2); r3 = mwc1_draw(ev, rp.one, l3);
        eg = lds32(gap_s + 4u * (r1 & (GAP_COLS - 1u))); eq = lds32(qsub_s + 4u * (r2 & mask));
        if (g == GAP_NONE) { next += GAP_NONE; continue; }
        const uint32_t at = next + g;
        if (at >= L) break;
        sts8(Dq + at, qn);
        const uint32_t sa = Ds + (rc ? L - 1u - at : at);
        uint32_t b; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(sa) : "memory");
        uint32_t rank = (b >> 1) & 3u; rank ^= rank >> 1;                       /* A,C,G,T -> 0..3 */
        sts8(sa, (0x54474341u >> (8u * (rank ^ m))) & 0xffu);
        next = at + 1u;
    }
}

/* Persistent warps, one lane per record.  A warp owns WARP_UNITS consecutive units (32 records) at a time: it composes
 * their FASTQ bytes in its private stage in shared memory and sends them off with one bulk async copy (TMA); only the
 * tables are shared by the CTA, so there is no CTA barrier in the loop.  A group of units whose bytes exceed the stage is
 * cut into sub-tiles.  Lane 2k + m holds mate m of unit k.
 *   phase 1  sequence and quality lines (body_lane, then substitutions_lane); byte stores for windows holding anything
 *            but A,C,G,T and for replay
 *   phase 2  headers and the fixed newline / '+' bytes, by the same lane
 * QM = how qualities are sampled (QM_*). */
/* header fields of record r of a unit (simulate.rs:287-298, :513-518) */
__device__ __forceinline__ void record_positions(const UnitDesc &d, uint32_t r, uint64_t &sp, uint64_t &ep) {
    if (!(d.flags & 4u)) { sp = d.p0; ep = d.p1; }
    else if (r == 0) { sp = d.p0; ep = d.p0 + d.rl; }
    else { sp = d.p1 + d.rl; ep = d.p1; }
}
__host__ __device__ constexpr uint32_t warp_region_bytes(uint32_t stage_cap) { return (STAGE_FRONT + stage_cap + 48u + 15u) & ~15u; }
/* bytes of the quality table in shared memory */
__host__ __device__ constexpr uint32_t qtab_bytes(int qm, uint32_t cb, uint32_t npos) {
    return qm == QM_LANE ? (128u << cb) : qm == QM_POS ? ((npos << cb) * 4u + 127u) & ~127u : 0u;
}

template <bool REPLAY, int QM>
__global__ void __launch_bounds__(FORMAT_THREADS, FORMAT_MIN_CTAS) k_format_lanes(RunParams rp, ReplayDev rd) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (batch_overflows(rp)) return;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t base_s = uint32_t(__cvta_generic_to_shared(smem));
    const uint32_t qbytes = qtab_bytes(QM, rp.phred.cb, rp.phred.npos);
    uint8_t *region = smem + SM_QTAB + qbytes + warp * warp_region_bytes(rp.stage_cap);
    uint8_t *stage = region + STAGE_FRONT;
    const uint32_t dec2_s = base_s + SM_DEC2, stage_s = uint32_t(__cvta_generic_to_shared(stage));
    const uint32_t qtab_s = base_s + SM_QTAB + (QM == QM_LANE ? 4u * lane : 0u);
    {   /* tables */
        uint32_t *qt = reinterpret_cast<uint32_t *>(smem + SM_QTAB);
        if (QM == QM_LANE) for (uint32_t i = threadIdx.x; i < (32u << rp.phred.cb); i += blockDim.x) qt[i] = rp.phred.entries[i >> 5];
        if (QM == QM_POS) for (uint32_t i = threadIdx.x; i < (rp.phred.npos << rp.phred.cb); i += blockDim.x) qt[i] = rp.phred.entries[i];
        if (QM != QM_CONST && rp.phred.has_sub) {
            for (uint32_t i = threadIdx.x; i < GAP_COLS; i += blockDim.x) reinterpret_cast<uint32_t *>(smem + SM_GAP)[i] = rp.phred.gap[i];
            for (uint32_t i = threadIdx.x; i < rp.phred_mult; i += blockDim.x) reinterpret_cast<uint32_t *>(smem + SM_QSUB)[i] = rp.phred.qsub[i];
        }
        for (uint32_t b = threadIdx.x; b < 256; b += blockDim.x) {           /* 4 bases per byte -> ASCII, both strands */
            uint32_t f = 0, r = 0;
            for (uint32_t j = 0; j < 4; ++j) {
                const uint32_t c = (b >> (2 * j)) & 3u;
                f |= ((0x54474341u >> (8 * c)) & 0xffu) << (8 * j);          /* "ACGT"[c] at byte j */
                r |= ((0x41434754u >> (8 * c)) & 0xffu) << (8 * (3 - j));    /* "TGCA"[c] at byte 3-j (util.rs:15-37) */
            }
            reinterpret_cast<uint32_t *>(smem + SM_LUT_FWD)[b] = f;
            reinterpret_cast<uint32_t *>(smem + SM_LUT_RC)[b] = r;
            if (b < 100) reinterpret_cast<uint16_t *>(smem + SM_DEC2)[b] = uint16_t(('0' + b / 10) | (('0' + b % 10) << 8));
            if (b < 10) { uint32_t p = 1; for (uint32_t k = 0; k < b; ++k) p *= 10u; reinterpret_cast<uint32_t *>(smem + SM_DEC2 + 200)[b] = p; }
        }
    }
    __syncthreads();

    const uint32_t mate = lane & 1u, slot = lane >> 1;
    const uint32_t ntiles = (rp.batch_units + WARP_UNITS - 1) / WARP_UNITS;
    /* What a tile needs from global memory before it can start — this lane's unit descriptor, the tile's first byte in the
     * batch, the first two words of the 2-bit copy under this lane's window — is fetched into registers while the tile
     * before it is composed (the descriptors after its lines, the words after its headers): a warp has nothing else to
     * do during those round trips, and 16 warps per SM do not hide them. */
    struct TileIn { uint4 r0, r1; uint64_t base; uint32_t w_lo, w_hi; };
    auto fetch_in = [&](uint32_t tile, TileIn &t) {
        t.r0 = t.r1 = make_uint4(0u, 0u, 0u, 0u); t.base = 0; t.w_lo = t.w_hi = 0;
        if (tile < ntiles) {
            const uint32_t t0 = tile * WARP_UNITS, tn = min(WARP_UNITS, rp.batch_units - t0);
            if (slot < tn) { const uint4 *src = reinterpret_cast<const uint4 *>(rp.desc + t0 + slot); t.r0 = __ldg(src); t.r1 = __ldg(src + 1); }
            t.base = __ldg(rp.offsets + t0);
        }
    };
    auto fetch_words = [&](TileIn &t) {      /* same address arithmetic as body_lane; units with a UnitExt (flag 16) fetch theirs there */
        const uint32_t L = mate ? (t.r1.z & 0xffffu) : (t.r1.y >> 16);
        const uint64_t a = pair64(t.r0.x, t.r0.y) + (mate ? uint64_t(t.r0.w) - t.r0.z : 0u);
        const int64_t bit = mate ? 2 * int64_t(a + L) - 32 : 2 * int64_t(a);
        if (!((t.r1.w >> 16) & 16u)) { const uint32_t *pw = rp.packed + (bit >> 5); t.w_lo = __ldg(pw); t.w_hi = __ldg(pw + 1); }
    };
    TileIn nx{}; bool nx_ok = false;
    const uint32_t tile_step = gridDim.x * nwarps;
    for (uint32_t tile = blockIdx.x * nwarps + warp; tile < ntiles; tile += tile_step) {
        const uint32_t t0 = tile * WARP_UNITS, tn = min(WARP_UNITS, rp.batch_units - t0);
        TileIn cur = nx;
        const bool words_ahead = nx_ok;
        if (!nx_ok) fetch_in(tile, cur);
        nx_ok = false;
        const UnitDesc d_tile = slot < tn ? unpack_unit(rp, cur.r0, cur.r1, t0 + slot) : UnitDesc{};
        const uint32_t tile_next = tile + tile_step;
        uint64_t tile_base = cur.base;                           /* first byte of the sub-tile in the batch */
        for (uint32_t s0 = 0; s0 < tn;) {
            /* sub-tile = longest run of units from s0 that fits the stage; a record's place in it by a scan of the records' bytes */
            const uint32_t u0 = t0 + s0, left = tn - s0;
            UnitDesc d = d_tile;
            if (s0 != 0) { d = UnitDesc{}; if (slot < left) d = load_unit(rp, u0 + slot); }
            const bool present = slot < left && (mate == 0 || (d.flags & 4u));
            const uint32_t L = mate ? d.L1 : d.L0, H = mate ? d.H1 : d.H0;
            const uint32_t rbytes = present ? H + 2u * L + 5u : 0u;
            uint32_t rend = rbytes;                                                       /* bytes of the records up to and including this lane's */
#pragma unroll
            for (uint32_t o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, rend, o); if (lane >= o) rend += y; }
            const uint32_t pad = uint32_t(tile_base & 15);
            const uint32_t nu = __popc(__ballot_sync(0xffffffffu, mate == 1u && slot < left && rend + pad <= rp.stage_cap));
            const uint32_t total = __shfl_sync(0xffffffffu, rend, 2u * nu - 1u);
            const bool last_sub = s0 + nu >= tn;
            /* this lane's record */
            const bool mine = slot < nu;
            const bool have = mine && present;
            const uint64_t a = mate ? d.a1 : d.a0;
            const uint32_t rec = pad + rend - rbytes;                                 /* stage offset of the record */
            const bool odd = have && ((d.flags & 8u) || H < 2u || L < 4u);            /* rules out the word paths for the sub-tile */
            const bool words_ok = !__any_sync(0xffffffffu, odd) && !REPLAY;
            const bool hdr_fast = words_ok && rp.hdr_fast != 0;
            const bool fast = have && words_ok && !(d.flags & (1u << mate));
            const uint64_t unit = rp.batch_begin + u0 + slot;
            /* the previous bulk copy must have finished reading the stage before it is rewritten */
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();

            /* ---- phase 1: sequence and quality lines ---- */
            {
                const uint32_t body_s = stage_s + rec + H;
                const uint32_t slots = ((L + 3u + 3u) >> 2);                          /* word slots of a line, at most */
                const uint32_t T4 = __reduce_max_sync(0xffffffffu, fast ? slots : 0u);
                Mwc1 qc{}, ec{};
                if (T4) {
                    if (QM != QM_CONST) mwc1_seed(rp, qc, ec, 0u, 1u + mate, uint32_t(unit), uint32_t(unit >> 32));
                    const uint32_t lut_s = base_s + (mate ? SM_LUT_RC : SM_LUT_FWD);
                    const bool pre = words_ahead && s0 == 0 && !(d.flags & 16u);
                    if (QM == QM_LANE && rp.phred.cb == 7u) body_lane<QM, 7>(rp, qtab_s, lut_s, fast, mate != 0, body_s, qc, a, L, T4, pre, cur.w_lo, cur.w_hi);
                    else if (QM == QM_POS && rp.phred.cb == 6u) body_lane<QM, 6>(rp, qtab_s, lut_s, fast, mate != 0, body_s, qc, a, L, T4, pre, cur.w_lo, cur.w_hi);
                    else body_lane<QM, 0>(rp, qtab_s, lut_s, fast, mate != 0, body_s, qc, a, L, T4, pre, cur.w_lo, cur.w_hi);
                }
                if (last_sub && tile_next < ntiles) fetch_in(tile_next, nx);           /* on their way during the substitutions and the headers */
                if (SIMMR_KO != 3 && T4 && QM != QM_CONST && rp.phred.has_sub) substitutions_lane(rp, base_s + SM_GAP, base_s + SM_QSUB, fast, mate != 0, body_s, ec, L);
                if (have && !fast) {
                    if (REPLAY) emit_body_replay<false>(rp, rd, stage + rec + H, mate, a, L, rd.base_off[unit] + (mate ? d.rl : 0), 0, 1);
                    else emit_body_pieces<false>(rp, stage + rec + H, unit, mate, a, L, 0, 1);
                }
            }
            __syncwarp();

            /* ---- phase 2: header and the fixed bytes ---- */
            if (have && SIMMR_KO != 1) {
                uint64_t sp, ep; record_positions(d, mate, sp, ep);
                const uint32_t id = rp.read_id_base + uint32_t(unit);
                const uint32_t rec_s = stage_s + rec;
                const ContigHdr ch = hdr_at(rp, d.contig, mate);
                if (hdr_fast) {      /* newline that ends the previous record, the tokens, newline */
                    ByteStream b; bs_open(b, rec_s - 1u); bs_put(b, '\n', 1);
                    header_tokens(b, rp, dec2_s, reinterpret_cast<const uint2 *>(rp.toks + ch.tok_off), rp.hdr_kinds, rp.hdr_fields,
                                  id, uint32_t(sp), uint32_t(ep));
                    bs_put(b, '\n', 1); bs_close(b);
                } else {
                    emit_header(stage + rec, rp, ch, id, sp, ep);
                    sts8(rec_s + H, '\n');
                    if (rec > pad) sts8(rec_s - 1u, '\n');                    /* ends the previous record */
                }
                const uint32_t mid = rec_s + H + L + 1u;                      /* fastq.rs:61-69 */
                sts8(mid, '\n'); sts8(mid + 1u, '+'); sts8(mid + 2u, '\n');
            }
            if (lane == 31 && total) sts8(stage_s + pad + total - 1u, '\n');
            if (last_sub && tile_next < ntiles) { fetch_words(nx); nx_ok = true; }   /* the descriptors have arrived by now */
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            /* stage[pad, pad+total) -> out[tile_base, +total): 16-byte aligned middle by TMA, edges by byte stores */
            const uint32_t head = min(total, (16 - pad) & 15), mid = (total - head) & ~15u, tail = total - head - mid;
            if (lane == 0 && mid) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(rp.out + tile_base + head), "r"(stage_s + pad + head), "r"(mid) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (lane >= 1 && lane < 1 + head) rp.out[tile_base + (lane - 1)] = stage[pad + (lane - 1)];
            if (lane >= 16 && lane < 16 + tail) rp.out[tile_base + head + mid + (lane - 16)] = stage[pad + head + mid + (lane - 16)];
            s0 += nu; tile_base += total;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}


/* One CTA per unit, bytes written straight to global memory with byte stores: replay of long reads, records that do not
 * fit a stage, and native long reads behind SIMMR_NO_LONG_KERNEL.  Not a fast path. */
template <bool REPLAY>
__global__ void __launch_bounds__(DIRECT_THREADS) k_format_direct(RunParams rp, ReplayDev rd) {
    if (batch_overflows(rp)) return;
    for (uint32_t i = blockIdx.x; i < rp.batch_units; i += gridDim.x) {
        const uint64_t u = rp.batch_begin + i;
        const UnitDesc d = load_unit(rp, i);
        uint8_t *dst = rp.out + rp.offsets[i];
        const uint64_t slot0 = REPLAY ? rd.base_off[u] : 0;
        const uint32_t nrec = (d.flags & 4u) ? 2 : 1;
        for (uint32_t r = 0; r < nrec; ++r) {
            const uint32_t H = r ? d.H1 : d.H0, L = r ? d.L1 : d.L0;
            if (threadIdx.x == 0) {
                uint64_t sp, ep; record_positions(d, r, sp, ep);
                emit_header(dst, rp, hdr_at(rp, d.contig, r), rp.read_id_base + uint32_t(u), sp, ep);
            }
            if (REPLAY) emit_body_replay<true>(rp, rd, dst + H, r, r ? d.a1 : d.a0, L, slot0 + (r ? d.rl : 0), threadIdx.x, DIRECT_THREADS);
            else if (rp.err.on) emit_body_pieces<true>(rp, dst + H, u, r, d.a0, L, threadIdx.x, DIRECT_THREADS, d.L1, !(d.flags & 1u));   /* edited long read */
            else emit_body_pieces<true>(rp, dst + H, u, r, r ? d.a1 : d.a0, L, threadIdx.x, DIRECT_THREADS);   /* a thread per piece */
            dst += H + 2 * L + 5;
        }
    }
}

/* ---------------------------------------------------------------------- */
/* long reads, native mode                                                  */
/* ---------------------------------------------------------------------- */
/* One line (sequence or quality) of one run, written straight to global memory.  Word k of the
 * run = funnel(W[k-1], W[k]) where W[k] is the run's k-th 4-byte group and word 0 is the aligned
 * word holding the run's first byte.  Four groups arrive at a time; the words are regrouped (by
 * q = the run's phase inside a 16-byte line) so that whole 16-byte lines leave as one vector
 * store.  Only whole words inside the run's byte range are stored here; the run's first and last
 * bytes that share a word with something else go by byte stores in rl_close, so runs (and
 * records) never touch each other's bytes. */
struct RunLine {
    uint8_t *word0;        /* address of aligned word 0 */
    uint32_t o, n;         /* the run's first byte is byte o of word 0; n bytes */
    uint32_t kend;         /* words [1, kend) are whole and stored by rl_quad; word 0 and word kend by rl_close */
    uint32_t sh, q;        /* 8 * o; (word0 / 4) % 4 */
    uint32_t c0, c1, c2;   /* words k-3, k-2, k-1 */
    uint32_t g[5];         /* W[k-1], W[k..k+3] of the last quad */
};
__device__ __forceinline__ void rl_open(RunLine &r, uint8_t *at, uint32_t n) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(at);
    r.o = uint32_t(a) & 3u; r.word0 = at - r.o; r.n = n; r.sh = 8u * r.o; r.q = (uint32_t(a) >> 2) & 3u;
    r.kend = (n + r.o) >> 2; r.c0 = r.c1 = r.c2 = 0;
    r.g[0] = r.g[1] = r.g[2] = r.g[3] = r.g[4] = 0;
}
/* groups k..k+3 of the run (k a multiple of 4) */
__device__ __forceinline__ void rl_quad(RunLine &r, uint32_t k, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    const uint32_t a0 = __funnelshift_l(r.g[4], w0, r.sh), a1 = __funnelshift_l(w0, w1, r.sh), a2 = __funnelshift_l(w1, w2, r.sh), a3 = __funnelshift_l(w2, w3, r.sh);
    r.g[0] = r.g[4]; r.g[1] = w0; r.g[2] = w1; r.g[3] = w2; r.g[4] = w3;
    /* the 16-byte line that ends inside this quad holds words k-q .. k-q+3 */
    const bool s1 = r.q & 1u, s2 = r.q & 2u;
    const uint32_t v0 = s1 ? r.c1 : r.c2, v1 = s1 ? r.c2 : a0, v2 = s1 ? a0 : a1, v3 = s1 ? a1 : a2, v4 = s1 ? a2 : a3;   /* window shifted by q & 1 ... */
    const uint32_t u0 = s1 ? r.c0 : r.c1;
    const uint32_t o0 = s2 ? u0 : v1, o1 = s2 ? v0 : v2, o2 = s2 ? v1 : v3, o3 = s2 ? v2 : v4;                           /* ... and by q & 2 */
    const int32_t cb = int32_t(k) - int32_t(r.q);
    uint32_t *dst = reinterpret_cast<uint32_t *>(r.word0) + cb;
    if (cb >= 1 && uint32_t(cb) + 3u < r.kend) *reinterpret_cast<uint4 *>(dst) = make_uint4(o0, o1, o2, o3);
    else {
        if (cb >= 1 && uint32_t(cb) < r.kend) dst[0] = o0;
        if (cb + 1 >= 1 && uint32_t(cb + 1) < r.kend) dst[1] = o1;
        if (cb + 2 >= 1 && uint32_t(cb + 2) < r.kend) dst[2] = o2;
        if (cb + 3 >= 1 && uint32_t(cb + 3) < r.kend) dst[3] = o3;
    }
    r.c0 = a1; r.c1 = a2; r.c2 = a3;
}
/* Whole 16-byte lines of eight consecutive turns are collected in shared memory and written out by the warp together, a
 * lane's eight lines (128 contiguous bytes) by eight neighbouring lanes: 32 lanes on 32 different lines per store
 * instruction become four full 128-byte rows.  Tile layout: line of turn slot s, lane l at s * 512 + ((l + s) & 31) * 16
 * (both the lane-wise writes and the row-wise reads are free of bank conflicts). */
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory"); return v; }
/* rl_quad with the whole-line store going to the tile (slot = turn & 7; `valid` collects the slots that hold a line) */
__device__ __forceinline__ void rl_quad_staged(RunLine &r, uint32_t k, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3,
                                               uint32_t tile_s, uint32_t lane, uint32_t slot, uint32_t &valid) {
    const uint32_t a0 = __funnelshift_l(r.g[4], w0, r.sh), a1 = __funnelshift_l(w0, w1, r.sh), a2 = __funnelshift_l(w1, w2, r.sh), a3 = __funnelshift_l(w2, w3, r.sh);
    r.g[0] = r.g[4]; r.g[1] = w0; r.g[2] = w1; r.g[3] = w2; r.g[4] = w3;
    const bool s1 = r.q & 1u, s2 = r.q & 2u;
    const uint32_t v0 = s1 ? r.c1 : r.c2, v1 = s1 ? r.c2 : a0, v2 = s1 ? a0 : a1, v3 = s1 ? a1 : a2, v4 = s1 ? a2 : a3;
    const uint32_t u0 = s1 ? r.c0 : r.c1;
    const uint32_t o0 = s2 ? u0 : v1, o1 = s2 ? v0 : v2, o2 = s2 ? v1 : v3, o3 = s2 ? v2 : v4;
    const int32_t cb = int32_t(k) - int32_t(r.q);
    uint32_t *dst = reinterpret_cast<uint32_t *>(r.word0) + cb;
    if (cb >= 1 && uint32_t(cb) + 3u < r.kend) { sts128(tile_s + slot * 512u + ((lane + slot) & 31u) * 16u, make_uint4(o0, o1, o2, o3)); valid |= 1u << slot; }
    else {
        if (cb >= 1 && uint32_t(cb) < r.kend) dst[0] = o0;
        if (cb + 1 >= 1 && uint32_t(cb + 1) < r.kend) dst[1] = o1;
        if (cb + 2 >= 1 && uint32_t(cb + 2) < r.kend) dst[2] = o2;
        if (cb + 3 >= 1 && uint32_t(cb + 3) < r.kend) dst[3] = o3;
    }
    r.c0 = a1; r.c1 = a2; r.c2 = a3;
}
/* Writes the tile out: `base` = this lane's global address of slot 0's line, `valid` = its slots that hold a line.  All 32 lanes. */
__device__ __forceinline__ void rl_flush(uint32_t tile_s, uint8_t *base, uint32_t valid, uint32_t lane) {
    __syncwarp();
    const uint32_t line = lane & 7u;
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
        const uint32_t row = 4u * i + (lane >> 3);
        const uint64_t rb = __shfl_sync(0xffffffffu, uint64_t(reinterpret_cast<uintptr_t>(base)), row);
        const uint32_t rv = __shfl_sync(0xffffffffu, valid, row);
        if ((rv >> line) & 1u)
            *reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(uintptr_t(rb)) + 16u * line) = lds128(tile_s + line * 512u + ((row + line) & 31u) * 16u);
    }
    __syncwarp();
}
/* k = first group of the last quad, ng = groups in the run */
__device__ __forceinline__ void rl_close(const RunLine &r, uint32_t k, uint32_t ng) {
    uint32_t *w = reinterpret_cast<uint32_t *>(r.word0);
    /* the last q words of the last quad have not left yet */
    if (r.q >= 3 && k + 1 >= 1 && k + 1 < r.kend) w[k + 1] = r.c0;
    if (r.q >= 2 && k + 2 < r.kend) w[k + 2] = r.c1;
    if (r.q >= 1 && k + 3 < r.kend) w[k + 3] = r.c2;
    /* word 0: whole only if the run starts on a word boundary and covers it.  W[0] is g[1] of the first quad; it is
     * kept by the caller in `first` */
}
__device__ __forceinline__ void rl_edges(const RunLine &r, uint32_t k, uint32_t ng, uint32_t first) {
    const uint32_t a0 = __funnelshift_l(0u, first, r.sh);
    if (r.o == 0 && r.kend >= 1) *reinterpret_cast<uint32_t *>(r.word0) = a0;
    else for (uint32_t b = r.o; b < min(4u, r.o + r.n); ++b) r.word0[b] = uint8_t(a0 >> (8u * b));
    /* word kend: the run's last bytes when they do not fill it; W[ng-1] and W[ng-2] are among the last quad's groups */
    const uint32_t rest = (r.n + r.o) & 3u;
    if (rest && r.kend >= 1) {
        const uint32_t i = ng - 1u - k;                                   /* 0..3: position of W[ng-1] in the last quad */
        const uint32_t last = i == 0 ? r.g[1] : i == 1 ? r.g[2] : i == 2 ? r.g[3] : r.g[4];
        const uint32_t before = i == 0 ? r.g[0] : i == 1 ? r.g[1] : i == 2 ? r.g[2] : r.g[3];
        const uint32_t ak = r.kend == ng ? __funnelshift_l(last, 0u, r.sh) : __funnelshift_l(before, last, r.sh);
        for (uint32_t b = 0; b < rest; ++b) r.word0[4u * r.kend + b] = uint8_t(ak >> (8u * b));
    }
}

/* Long reads (simulate.rs:323-406, 464-523), native mode.  A record is far larger than a stage and read lengths vary by
 * an order of magnitude (gamma, CV 0.75), so a warp takes one read at a time from a shared counter and its 32 lanes take
 * the read's LONG_PIECES pieces (16 * ceil(L / 512) bases each): all lanes loop the same number of times
 * whatever the read's length.  A lane seeds its piece's chains, writes that part of the sequence and quality lines
 * straight to global memory in 16-byte lines (RunLine), then applies the piece's substitutions to the bytes it wrote.
 * Lane 0 also writes the header and the fixed bytes.  Reads whose window holds anything but A,C,G,T take the byte
 * path, a lane per piece.  QM and CBK as in k_format_lanes / quals4. */
template <int QM, int CBK>
__global__ void __launch_bounds__(DIRECT_THREADS, 2) k_format_long(RunParams rp) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (batch_overflows(rp)) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t base_s = uint32_t(__cvta_generic_to_shared(smem));
    const uint32_t qtab_s = base_s + SM_QTAB + (QM == QM_LANE ? 4u * lane : 0u), lut_s = base_s + SM_LUT_FWD;
    {
        uint32_t *qt = reinterpret_cast<uint32_t *>(smem + SM_QTAB);
        if (QM == QM_LANE) for (uint32_t i = threadIdx.x; i < (32u << rp.phred.cb); i += blockDim.x) qt[i] = rp.phred.entries[i >> 5];
        if (QM == QM_POS) for (uint32_t i = threadIdx.x; i < (rp.phred.npos << rp.phred.cb); i += blockDim.x) qt[i] = rp.phred.entries[i];
        if (QM != QM_CONST && rp.phred.has_sub) {
            for (uint32_t i = threadIdx.x; i < GAP_COLS; i += blockDim.x) reinterpret_cast<uint32_t *>(smem + SM_GAP)[i] = rp.phred.gap[i];
            for (uint32_t i = threadIdx.x; i < rp.phred_mult; i += blockDim.x) reinterpret_cast<uint32_t *>(smem + SM_QSUB)[i] = rp.phred.qsub[i];
        }
        for (uint32_t b = threadIdx.x; b < 256; b += blockDim.x) {
            uint32_t f = 0;
            for (uint32_t j = 0; j < 4; ++j) f |= ((0x54474341u >> (8 * ((b >> (2 * j)) & 3u))) & 0xffu) << (8 * j);
            reinterpret_cast<uint32_t *>(smem + SM_LUT_FWD)[b] = f;
        }
    }
    __syncthreads();

    const uint32_t qconst = ((rp.phred.const_q + 33u) & 0xffu) * 0x01010101u, mask = rp.phred_mult - 1u;
    for (;;) {
        unsigned long long next = 0;
        if (lane == 0) next = atomicAdd(rp.counters + 4, 1ull);
        const uint32_t i = uint32_t(__shfl_sync(0xffffffffu, next, 0));
        if (i >= rp.batch_units) break;
        const UnitDesc d = load_unit(rp, i);
        const uint32_t L = d.L0, H = d.H0;
        const uint32_t Ls = rp.err.on ? uint32_t(d.L1) : L;            /* bases of the sequence line: fewer than qualities after k-mer errors deleted some (simulate.rs:497-503) */
        const uint64_t unit = rp.batch_begin + i;
        uint8_t *rec = rp.out + rp.offsets[i];
        const bool bytes_path = (d.flags & 9u) || L < 4u;
        {                                                 /* header and the five fixed bytes (fastq.rs:34-69) */
            const ContigHdr &ch = hdr_at(rp, d.contig, 0);
            const bool done = emit_header_warp(rec, rp, ch, rp.read_id_base + uint32_t(unit), d.p0, d.p1, lane);
            if (lane == 0) {
                if (!done) emit_header(rec, rp, ch, rp.read_id_base + uint32_t(unit), d.p0, d.p1);
                rec[H] = '\n'; rec[H + Ls + 1] = '\n'; rec[H + Ls + 2] = '+'; rec[H + Ls + 3] = '\n'; rec[H + Ls + L + 4] = '\n';
            }
        }
        if (bytes_path) {
            if (rp.err.on) emit_body_pieces<false>(rp, rec + H, unit, 0, d.a0, L, lane, 32, Ls, !(d.flags & 1u));
            else emit_body_pieces<false>(rp, rec + H, unit, 0, d.a0, L, lane, 32);
            __syncwarp(); continue;
        }
        const uint32_t S = piece_len(L, true), lo = lane * S;
        const uint32_t nb = lo < L ? min(S, L - lo) : 0u, ng = (nb + 3u) >> 2;        /* qualities and their 4-base groups in this lane's piece */
        const uint32_t nbs = lo < Ls ? min(S, Ls - lo) : 0u, ngs = (nbs + 3u) >> 2;   /* ... and bases */
        uint8_t *seq = rec + H + 1 + lo, *qual = rec + H + Ls + 4 + lo;
        RunLine ls, lq;
        rl_open(ls, seq, nbs); rl_open(lq, qual, nb);
        Mwc1 qc{}, ec{};
        if (QM != QM_CONST && nb) mwc1_seed(rp, qc, ec, lane, 1, uint32_t(unit), uint32_t(unit >> 32));
        /* The piece's first LONG_EVENTS substitution events are walked now (position | new quality << 16 | m << 24, a list
         * per lane in shared memory): the turn that reaches a position XORs m into the base's 2-bit code before the
         * letters are looked up, and the new qualities are plain byte stores after the lines are out — no base is read
         * back from memory.  What a longer chain holds beyond the list is applied the slow way at the end. */
        const uint32_t ev_s = base_s + SM_QTAB + qtab_bytes(QM, rp.phred.cb, rp.phred.npos) + uint32_t(DIRECT_THREADS / 32) * (LONG_RING * 128u + 8192u) + 4u * threadIdx.x;
        uint32_t nev = 0, ev_at0 = 0;
        if (QM != QM_CONST && rp.phred.has_sub) {
            const uint32_t gap_s = base_s + SM_GAP, qsub_s = base_s + SM_QSUB;
            while (ev_at0 < nb && nev < LONG_EVENTS) {
                uint32_t l1, l2, l3;
                const uint32_t r1 = mwc1_draw(ec, rp.one, l1), r2 = mwc1_draw(ec, rp.one, l2), r3 = mwc1_draw(ec, rp.one, l3);
                const uint32_t g = q32_pick(lds32(gap_s + 4u * (r1 & (GAP_COLS - 1u))), l1);
                if (g == GAP_NONE) { ev_at0 += GAP_NONE; continue; }
                const uint32_t at = ev_at0 + g;
                if (at >= nb) { ev_at0 = nb; break; }
                sts32(ev_s + nev * (DIRECT_THREADS * 4u), at | (q32_pick(lds32(qsub_s + 4u * (r2 & mask)), l2) << 16) | ((1u + __umulhi(r3, 3u)) << 24));
                ++nev; ev_at0 = at + 1u;
            }
        }
        uint32_t ev_i = 0, ev_cur = nev ? lds32m(ev_s) : 0xffffffffu;
        const uint64_t a = d.a0 + lo;
        const uint32_t *pw = (rp.err.on ? rp.epacked : rp.packed) + (a >> 4);      /* edited reads come from the scratch copy */
        const uint32_t s2 = 2u * (uint32_t(a) & 15u);
        /* Words of the 2-bit copy reach the lane through a ring in shared memory filled by asynchronous copies LONG_AHEAD
         * turns ahead of their use (with 32 lanes on 32 different lines some lane misses L1 nearly every turn, and a
         * register loaded that far ahead cannot be kept out of the way of the compiler's moves).  Word k of lane l sits at
         * ring word (k % LONG_RING) * 32 + l: copies and reads stay in the lane's bank.  One copy group per word. */
        const uint32_t ring_s = base_s + SM_QTAB + qtab_bytes(QM, rp.phred.cb, rp.phred.npos) + (threadIdx.x >> 5) * (LONG_RING * 128u) + 4u * lane;
#pragma unroll
        for (uint32_t k = 0; k <= LONG_AHEAD; ++k) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ring_s + (k % LONG_RING) * 128u), "l"(pw + k) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(LONG_AHEAD) : "memory");           /* word 0 has landed */
        uint32_t w_lo = lds32m(ring_s);
        const uint32_t quads = (ng + 3u) >> 2, quads_s = (ngs + 3u) >> 2, T = __reduce_max_sync(0xffffffffu, quads);
        uint32_t first_s = 0, first_q = 0, vs = 0, vq = 0;
        const uint32_t tile_s = base_s + SM_QTAB + qtab_bytes(QM, rp.phred.cb, rp.phred.npos) + uint32_t(DIRECT_THREADS / 32) * (LONG_RING * 128u) + (threadIdx.x >> 5) * 8192u;
        for (uint32_t qd = 0; qd < T; ++qd) {             /* sixteen bases = one word of the 2-bit copy */
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ring_s + ((qd + LONG_AHEAD + 1u) % LONG_RING) * 128u), "l"(pw + qd + LONG_AHEAD + 1u) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(LONG_AHEAD) : "memory");       /* words up to qd + 1 have landed */
            const uint32_t w_hi = lds32m(ring_s + ((qd + 1u) % LONG_RING) * 128u);
            if (qd < quads) {
                uint32_t X = __funnelshift_r(w_lo, w_hi, s2);
                while (((ev_cur & 0xffffu) >> 4) == qd) {         /* substitutions among this turn's 16 bases: rank ^ m */
                    X ^= (ev_cur >> 24) << (2u * (ev_cur & 15u));
                    ++ev_i; ev_cur = ev_i < nev ? lds32m(ev_s + ev_i * (DIRECT_THREADS * 4u)) : 0xffffffffu;
                }
                uint32_t sw[4], qw[4];
#pragma unroll
                for (uint32_t h = 0; h < 4; ++h) {
                    qw[h] = QM == QM_CONST ? qconst : quals4<QM, CBK>(rp, qtab_s, mask, qc, lo + 16u * qd + 4u * h);
                    sw[h] = lds32(lut_s + ((h == 0 ? X << 2 : X >> (8u * h - 2u)) & 0x3fcu));
                }
                if (qd == 0) { first_s = sw[0]; first_q = qw[0]; }
                if (qd < quads_s) rl_quad_staged(ls, 4u * qd, sw[0], sw[1], sw[2], sw[3], tile_s, lane, qd & 7u, vs);
                rl_quad_staged(lq, 4u * qd, qw[0], qw[1], qw[2], qw[3], tile_s + 4096u, lane, qd & 7u, vq);
            }
            w_lo = w_hi;
            if ((qd & 7u) == 7u || qd + 1u == T) {        /* eight turns collected (or the read is done): the warp writes the lines out */
                const int32_t first = int32_t(16u * (qd & ~7u));
                rl_flush(tile_s, ls.word0 + first - 4 * int32_t(ls.q), vs, lane);
                rl_flush(tile_s + 4096u, lq.word0 + first - 4 * int32_t(lq.q), vq, lane);
                vs = vq = 0;
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");      /* nothing of this read's copies may land in the next read's ring */
        if (ngs) { const uint32_t kl = 4u * (quads_s - 1u); rl_close(ls, kl, ngs); rl_edges(ls, kl, ngs, first_s); }
        if (ng) { const uint32_t kl = 4u * (quads - 1u); rl_close(lq, kl, ng); rl_edges(lq, kl, ng, first_q); }
        if (QM != QM_CONST && rp.phred.has_sub) {
            for (uint32_t k = 0; k < nev; ++k) {          /* the listed events' qualities */
                const uint32_t e = lds32m(ev_s + k * (DIRECT_THREADS * 4u));
                qual[e & 0xffffu] = uint8_t(e >> 16);
            }
            const uint32_t gap_s = base_s + SM_GAP, qsub_s = base_s + SM_QSUB;
            for (uint32_t at0 = ev_at0; at0 < nb;) {      /* events past the list: the base is read back from the line */
                uint32_t l1, l2, l3;
                const uint32_t r1 = mwc1_draw(ec, rp.one, l1), r2 = mwc1_draw(ec, rp.one, l2), r3 = mwc1_draw(ec, rp.one, l3);
                const uint32_t g = q32_pick(lds32(gap_s + 4u * (r1 & (GAP_COLS - 1u))), l1);
                if (g == GAP_NONE) { at0 += GAP_NONE; continue; }
                const uint32_t at = at0 + g;
                if (at >= nb) break;
                qual[at] = uint8_t(q32_pick(lds32(qsub_s + 4u * (r2 & mask)), l2));
                const uint32_t b = seq[at];
                uint32_t rank = (b >> 1) & 3u; rank ^= rank >> 1;                   /* A,C,G,T -> 0..3 */
                seq[at] = uint8_t(0x54474341u >> (8u * (rank ^ (1u + __umulhi(r3, 3u)))));
                at0 = at + 1u;
            }
        }
    }
}

}  // namespace simmr

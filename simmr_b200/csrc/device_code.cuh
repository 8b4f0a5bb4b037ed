/* device_code.cuh — sm_100a kernels of the simmr read-simulation hot path.
 *
 * Path (reference file:line -> kernel):
 *   simulate.rs:165-190, 205-262, 353-375, 464-491  geometry      -> unit_geometry()
 *   fastq.rs:34-57 header interpolation, record size             -> k_unit_sizes, emit_header()
 *   (no reference equivalent: byte offsets of the FASTQ stream)   -> k_scan_blocks, k_unit_offsets
 *   simulate.rs:261-286 + error_profiles/ *.rs phreds/mutation +
 *   util.rs:15-57 RC / +33 + fastq.rs:61-69 record layout         -> k_format_tiles, k_format_direct
 *   (load-time) which 32-byte blocks hold only A,C,G,T           -> k_purity
 *
 * Draw schedule, table formats and the tile layout are specified in DESIGN.md.
 * Compile with -fmad=false: the replay arithmetic (minimal_short.rs:90-100) and
 * the gamma sampler must round exactly like the CPU.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/simmr_b200.h"
#include "simmr_types.h"

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

struct UnitDesc;
struct RunParams {
    /* genomes (device resident) */
    const uint8_t *bases; const uint64_t *contig_off, *contig_size; const uint32_t *genome_contig;
    const uint32_t *impure;        /* 1 bit per 32-byte block of `bases`: holds a byte other than A,C,G,T */
    const ContigHdr *chdr; const HdrTok *toks; const uint8_t *pool;
    uint32_t n_genomes, n_contigs;
    /* unit space */
    const uint64_t *unit_start;    /* [n_genomes+1] first global unit of each genome */
    uint64_t batch_begin; uint32_t batch_units; uint32_t read_id_base;
    /* profile (phred.entries: device layout, quality bytes already +33) */
    PhredTab phred; LenTab rl, is; GammaConst gamma;
    uint32_t min_size, is_long, flags, mutates;
    uint32_t rk[20];               /* Philox round keys: seed + r * Weyl constants, r = 0..9 */
    /* per-batch scratch + output */
    uint32_t *sizes; uint64_t *offsets, *block_sums; uint8_t *out;
    struct UnitDesc *desc;         /* per-unit geometry written by k_unit_sizes, read by the format kernels */
    unsigned long long *counters;  /* [0] clamped [1] algorithmic bytes [2] bases [3] records */
    /* tile geometry of k_format_tiles */
    uint32_t upt, stage_cap;
};

constexpr int SIZES_THREADS = 256;
constexpr int FORMAT_THREADS = 128;   /* threads of a tile CTA                          */
constexpr int FORMAT_MIN_CTAS = 8;    /* occupancy target -> register cap of 64          */
constexpr int DIRECT_THREADS = 256;
constexpr int SUB = 8;             /* lanes that share one record in the tile kernel */

/* ---------------------------------------------------------------------- */
/* small PTX helpers                                                        */
/* ---------------------------------------------------------------------- */
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t addr) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v; }

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

/* what the format kernels need of a unit: 48 bytes written once by k_unit_sizes */
struct __align__(16) UnitDesc {
    uint64_t a0, a1;       /* window offsets in `bases`                                     */
    uint64_t p0, p1;       /* PE: fwd_start, rev_end (simulate.rs:233,241-258); long: start, end */
    uint32_t contig;
    uint16_t rl, L0, L1, H0, H1;
    uint16_t flags;        /* 1: window 0 impure, 2: window 1 impure, 4: two records        */
};
static_assert(sizeof(UnitDesc) == 48, "UnitDesc layout");

__device__ __forceinline__ uint32_t genome_of(const RunParams &rp, uint64_t u) {
    uint32_t lo = 0, hi = rp.n_genomes;            /* last g with unit_start[g] <= u */
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(rp.unit_start + mid) <= u) lo = mid; else hi = mid; }
    return lo;
}

template <bool REPLAY>
__device__ void unit_geometry(const RunParams &rp, const ReplayDev &rd, uint64_t u, UnitGeo &o) {
    o.id = rp.read_id_base + uint32_t(u);                         /* simulate.rs:85-89 */
    o.clamped = 0;
    uint32_t g, c; uint64_t RL, IS = 0, start, redraw = 0;
    Philox4 w0{};
    if (REPLAY) {
        g = rd.genome[u]; c = __ldg(rp.genome_contig + g) + rd.contig[u];
        RL = rd.read_length[u]; start = rd.start[u]; redraw = rd.redraw[u];
        if (!rp.is_long) IS = rd.insert_size[u];
    } else {
        g = genome_of(rp, u);
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
__device__ __forceinline__ uint32_t dec_digits32(uint32_t v) {
    return 1u + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u) + (v >= 100000u) + (v >= 1000000u) +
           (v >= 10000000u) + (v >= 100000000u) + (v >= 1000000000u);
}
__device__ __forceinline__ uint32_t dec_digits(uint64_t v) {
    if (!(v >> 32)) return dec_digits32(uint32_t(v));
    uint32_t d = 1;
    while (v >= 10) { v /= 10; ++d; }
    return d;
}
__device__ __forceinline__ uint32_t header_len(const ContigHdr &ch, uint32_t id, uint64_t sp, uint64_t ep) {
    return ch.lit_bytes + ch.n_id * dec_digits32(id) + ch.n_sp * dec_digits(sp) + ch.n_ep * dec_digits(ep) + ch.n_rc + ch.n_pair;
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
__device__ uint32_t emit_header(uint8_t *dst, const RunParams &rp, const ContigHdr &ch, uint32_t id, uint64_t sp, uint64_t ep, bool rc, int mate) {
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
        case TOK_END: n += put_dec(dst + n, ep); break;
        case TOK_RC: dst[n++] = rc ? 't' : 'f'; break;
        default: dst[n++] = mate ? '2' : '1'; break;              /* fastq.rs:57, :106 */
        }
    }
    return n;
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

/* (quality byte as printed, mutation code) of base i from its Philox word */
__device__ __forceinline__ void native_base(const RunParams &rp, const uint32_t *tab, uint32_t r, uint32_t i, uint32_t &qc, uint32_t &m) {
    const uint32_t cb = rp.phred.cb, idx = r >> (32 - cb), x = r << cb;
    const uint32_t pos = i < rp.phred.npos ? i : rp.phred.npos - 1;               /* custom_short.rs:339-345 */
    const uint32_t col = (pos << cb) + idx;
    if (rp.phred.kind == PT_T64) {
        const uint32_t thr = tab[2 * col], pay = tab[2 * col + 1];
        const uint32_t s = x < thr ? pay : pay >> 16;
        qc = s & 0xffu; m = (s >> 8) & 0xffu;
    } else {
        const uint32_t e = tab[col];
        qc = ((x >> 16) < (e >> 16) ? e : e >> 8) & 0xffu; m = 0;
    }
}

/* Writes one record body (everything after the header) with byte stores; `dst`
 * may be shared or global memory.  Base i of the window gets quality slot i; for
 * mate 2 the (mutated) base is complemented and lands at L-1-i, the quality stays
 * at i (simulate.rs:283-285). */
template <bool REPLAY>
__device__ __forceinline__ void emit_body_generic(const RunParams &rp, const ReplayDev &rd, const uint32_t *tab, uint8_t *dst,
                                                  uint64_t unit, uint32_t mate, uint64_t a, uint32_t L, uint64_t slot0, uint32_t thr, uint32_t nthr) {
    uint8_t *seq = dst + 1, *qual = dst + L + 4;
    if (thr == 0) { dst[0] = '\n'; dst[L + 1] = '\n'; dst[L + 2] = '+'; dst[L + 3] = '\n'; dst[2 * L + 4] = '\n'; }   /* fastq.rs:61-69 */
    const uint8_t *src = rp.bases + a;
    const bool rc = mate != 0;
    for (uint32_t g = thr; g * 4 < L; g += nthr) {
        Philox4 w{};
        if (!REPLAY && rp.phred.kind != PT_CONST) w = draw(rp, unit, 1 + mate, g);
        const uint32_t wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t i = 4 * g + j;
            if (i >= L) break;
            uint32_t qc, m;
            if (REPLAY) { uint32_t q; replay_base(rp, rd, slot0 + i, i, q, m); qc = (q + 33) & 0xffu; }   /* util.rs:46-50 */
            else if (rp.phred.kind == PT_CONST) { qc = (rp.phred.const_q + 33) & 0xffu; m = 0; }
            else native_base(rp, tab, wv[j], i, qc, m);
            uint8_t b = __ldg(src + i);
            if (m) b = substitute(b, m - 1);
            qual[i] = uint8_t(qc);
            if (rc) seq[L - 1 - i] = complement(b); else seq[i] = b;
        }
    }
}

/* ---------------------------------------------------------------------- */
/* per-base work, fast path: windows of pure A/C/G/T, SWAR on 4 bases/word  */
/* ---------------------------------------------------------------------- */
/* One output line written through aligned 32-bit shared-memory stores.  Words
 * arrive per 4-base group; the line starts at an arbitrary byte address, so two
 * consecutive group words are funnel-shifted into each aligned word. */
struct LineW {
    uint32_t D;         /* shared-space address of the first byte of the line        */
    int32_t L;          /* bytes in the line                                         */
    int32_t off0;       /* start of composed word 0 relative to D (-6..0)            */
    uint32_t sh;        /* 8 * (virtual start & 3)                                   */
    uint32_t first, prev;
};
__device__ __forceinline__ void lw_init(LineW &w, uint32_t D, uint32_t L, uint32_t lead) {
    const uint32_t V = D - lead;           /* `lead` garbage bytes precede the line (reversed lines) */
    w.D = D; w.L = int32_t(L); w.off0 = -int32_t(lead + (V & 3u)); w.sh = 8u * (V & 3u); w.first = w.prev = 0;
}
/* composed word k (may be far out of range on either side: nothing is written then) */
__device__ __forceinline__ void lw_emit(const LineW &w, int32_t k, uint32_t c) {
    const int32_t p = w.off0 + 4 * k;
    if (p >= 0 && p + 4 <= w.L) sts32(w.D + uint32_t(p), c);
    else if (p + 4 > 0 && p < w.L)
        for (int32_t b = max(p, 0); b < min(p + 4, w.L); ++b) sts8(w.D + uint32_t(b), c >> (8 * (b - p)));
}
/* ascending: group g's word holds line bytes [4g, 4g+4) */
__device__ __forceinline__ void lw_step_asc(LineW &w, uint32_t t, uint32_t g, uint32_t x) {
    if (t == 0) w.first = x; else lw_emit(w, int32_t(g), __funnelshift_l(w.prev, x, w.sh));
    w.prev = x;
}
__device__ __forceinline__ void lw_finish_asc(LineW &w, uint32_t g0, uint32_t sl) {
    uint32_t left = __shfl_up_sync(0xffffffffu, w.prev, 1, SUB);
    if (sl == 0) left = 0;
    lw_emit(w, int32_t(g0), __funnelshift_l(left, w.first, w.sh));
}
/* descending: group g's word holds line bytes [L-4g-4, L-4g), i.e. composed index NG-1-g */
__device__ __forceinline__ void lw_step_desc(LineW &w, uint32_t t, uint32_t g, uint32_t NG, uint32_t x) {
    if (t == 0) w.first = x; else lw_emit(w, int32_t(NG) - int32_t(g), __funnelshift_l(x, w.prev, w.sh));
    w.prev = x;
}
__device__ __forceinline__ void lw_finish_desc(LineW &w, uint32_t g0, uint32_t NG, uint32_t sl) {
    uint32_t left = __shfl_up_sync(0xffffffffu, w.prev, 1, SUB);
    if (sl == 0) left = 0;
    lw_emit(w, int32_t(NG) - int32_t(g0), __funnelshift_l(w.first, left, w.sh));
}

/* complement of four A/C/G/T bytes: A<->T is ^0x15, C<->G is ^0x04; bit 1 tells which */
__device__ __forceinline__ uint32_t comp4(uint32_t b) {
    const uint32_t t = (b >> 1) & 0x01010101u;
    return b ^ 0x15151515u ^ (t * 0x11u);
}
/* substitutions on a word of pure A/C/G/T, branch-free; mw byte j = 0 or choose()+1.
 * With rank(A,C,G,T) = 0..3 the replacement has rank (m-1) + [m > rank]: the m-th of the
 * other three bases in alphabetical order (minimal_short.rs:124-131). */
__device__ __forceinline__ uint32_t mutate4(uint32_t bw, uint32_t mw) {
    const uint32_t rank = ((bw >> 1) ^ (bw >> 2)) & 0x03030303u;
    const uint32_t gt = ((mw + (rank ^ 0x03030303u)) >> 2) & 0x01010101u;     /* m > rank, per byte */
    const uint32_t pick = (mw | 0x04040404u) + gt - 0x01010101u;               /* low 2 bits; garbage where m == 0 */
    const uint32_t sel = (pick & 0x0003u) | ((pick >> 4) & 0x0030u) | ((pick >> 8) & 0x0300u) | ((pick >> 12) & 0x3000u);
    const uint32_t nb = __byte_perm(0x54474341u, 0, sel);                      /* "ACGT"[pick] per byte */
    const uint32_t hit = (((mw + 0x7f7f7f7fu) & 0x80808080u) >> 7) * 0xffu;    /* 0xff where m != 0 */
    return (bw & ~hit) | (nb & hit);
}

/* quality characters and mutation codes of bases 4g..4g+3 */
template <int TK>
__device__ __forceinline__ void sample4(const RunParams &rp, uint32_t tab_s, Philox4 w, uint32_t g, uint32_t &qw, uint32_t &mw) {
    const uint32_t cb = rp.phred.cb;
    const uint32_t r[4] = {w.x, w.y, w.z, w.w};
    if (TK == PT_T64) {
        uint32_t p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint2 e = lds64(tab_s + ((r[j] >> (32 - cb)) << 3));
            p[j] = (r[j] << cb) < e.x ? e.y : e.y >> 16;
        }
        const uint32_t p01 = __byte_perm(p[0], p[1], 0x5410), p23 = __byte_perm(p[2], p[3], 0x5410);
        qw = __byte_perm(p01, p23, 0x6420); mw = __byte_perm(p01, p23, 0x7531);
    } else {
        uint32_t v[4];
        const uint32_t last = rp.phred.npos - 1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t pos = min(4 * g + j, last);
            const uint32_t e = lds32(tab_s + ((((pos << cb) + (r[j] >> (32 - cb)))) << 2));
            v[j] = ((r[j] << cb) >> 16) < (e >> 16) ? e : e >> 8;
        }
        qw = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410); mw = 0;
    }
}

/* Record bodies by a whole warp: each sub-group of SUB lanes owns one record (`on` =
 * this sub-group has one), lane `sl` handles the 4-base groups [sl*T, (sl+1)*T) with T
 * common to the warp so that all 32 lanes stay converged.  body_s = shared-space
 * address of the '\n' that ends the header. */
template <int TK>
__device__ __forceinline__ void emit_body_fast(const RunParams &rp, uint32_t tab_s, bool on, uint32_t body_s, uint64_t unit, uint32_t mate,
                                               uint64_t a, uint32_t L, uint32_t sl) {
    if (!on) { L = 0; body_s = 0; a = 0; }
    if (on && sl == 0) { sts8(body_s, '\n'); sts8(body_s + L + 1, '\n'); sts8(body_s + L + 2, '+'); sts8(body_s + L + 3, '\n'); sts8(body_s + 2 * L + 4, '\n'); }
    const uint32_t NG = (L + 3) >> 2;
    const uint32_t T = __reduce_max_sync(0xffffffffu, (NG + SUB) / SUB);   /* SUB*T >= NG+1 composed words */
    const bool rc = mate != 0;
    LineW ls, lq;
    lw_init(ls, body_s + 1, L, rc ? ((4u - (L & 3u)) & 3u) : 0u);
    lw_init(lq, body_s + L + 4, L, 0);
    const uint32_t al = uint32_t(a & 3u);                               /* rp.bases is 256-byte aligned */
    const uint32_t *src = reinterpret_cast<const uint32_t *>(rp.bases + (a - al));
    const uint32_t g0 = sl * T;
    const uint32_t qconst = ((rp.phred.const_q + 33u) & 0xffu) * 0x01010101u;
    for (uint32_t t = 0; t < T; ++t) {
        const uint32_t g = g0 + t;
        uint32_t sw = 0, qw = 0;
        if (g < NG) {
            const uint32_t w0 = __ldg(src + g), w1 = __ldg(src + g + 1);
            sw = __funnelshift_r(w0, w1, 8u * al);
            if (TK == PT_CONST) qw = qconst;
            else {
                uint32_t mw;
                sample4<TK>(rp, tab_s, draw(rp, unit, 1 + mate, g), g, qw, mw);
                if (TK == PT_T64) sw = mutate4(sw, mw);
            }
            if (rc) sw = __byte_perm(comp4(sw), 0, 0x0123);
        }
        if (rc) lw_step_desc(ls, t, g, NG, sw); else lw_step_asc(ls, t, g, sw);
        lw_step_asc(lq, t, g, qw);
    }
    if (rc) lw_finish_desc(ls, g0, NG, sl); else lw_finish_asc(ls, g0, sl);
    lw_finish_asc(lq, g0, sl);
}

/* ---------------------------------------------------------------------- */
/* kernels                                                                  */
/* ---------------------------------------------------------------------- */
/* load time: bit per 32-byte block = block holds a byte that is not A,C,G,T */
__global__ void k_purity(const uint8_t *bases, uint64_t nbytes, uint32_t *impure, uint64_t nwords) {
    const uint64_t w = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (w >= nwords) return;
    uint32_t bits = 0;
    for (uint32_t b = 0; b < 32; ++b) {
        const uint64_t off = (w * 32 + b) * 32;
        bool bad = false;
        for (uint32_t k = 0; k < 32 && off + k < nbytes; k += 4) {
            const uint32_t v = *reinterpret_cast<const uint32_t *>(bases + off + k);   /* tail is padded with 'N' */
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t c = (v >> (8 * j)) & 0xffu;
                if (off + k + j < nbytes) bad |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
            }
        }
        bits |= uint32_t(bad) << b;
    }
    impure[w] = bits;
}

__device__ __forceinline__ uint64_t block_sum_u64(uint64_t v, uint64_t *sh) {
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    uint64_t r = 0;
    if (wid == 0) {
        r = lane < (blockDim.x >> 5) ? sh[lane] : 0;
        for (int o = 16; o; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;   /* valid in thread 0 */
}

/* one thread per unit: geometry -> descriptor, FASTQ bytes of the unit, block totals, run statistics */
template <bool REPLAY>
__global__ void __launch_bounds__(SIZES_THREADS) k_unit_sizes(RunParams rp, ReplayDev rd) {
    __shared__ uint64_t sh[32];
    const uint32_t i = blockIdx.x * SIZES_THREADS + threadIdx.x;
    uint64_t bytes = 0, alg = 0, bases = 0, clamped = 0, recs = 0;
    if (i < rp.batch_units) {
        UnitGeo g; unit_geometry<REPLAY>(rp, rd, rp.batch_begin + i, g);
        const ContigHdr ch = rp.chdr[g.contig];
        uint32_t H[2] = {0, 0};
        for (uint32_t r = 0; r < g.nrec; ++r) {
            H[r] = header_len(ch, g.id, g.sp[r], g.ep[r]);
            bytes += H[r] + 2ull * g.L[r] + 5;
            alg += H[r] + 3ull * g.L[r] + 5;                     /* BASELINE.md §3 */
            bases += g.L[r];
        }
        clamped = g.clamped; recs = g.nrec;
        rp.sizes[i] = uint32_t(bytes);
        UnitDesc d;
        d.a0 = g.a[0]; d.a1 = g.a[1];
        d.p0 = g.sp[0]; d.p1 = g.nrec == 2 ? g.ep[1] : g.ep[0];
        d.contig = g.contig; d.rl = uint16_t(g.rl); d.L0 = uint16_t(g.L[0]); d.L1 = uint16_t(g.L[1]);
        d.H0 = uint16_t(H[0]); d.H1 = uint16_t(H[1]);
        d.flags = uint16_t((REPLAY || window_impure(rp, g.a[0], g.L[0]) ? 1u : 0u) |
                           (REPLAY || window_impure(rp, g.a[1], g.L[1]) ? 2u : 0u) | (g.nrec == 2 ? 4u : 0u));
        rp.desc[i] = d;
    }
    const uint64_t tb = block_sum_u64(bytes, sh), ta = block_sum_u64(alg, sh), tn = block_sum_u64(bases, sh),
                   tc = block_sum_u64(clamped, sh), tr = block_sum_u64(recs, sh);
    if (threadIdx.x == 0) {
        rp.block_sums[blockIdx.x] = tb;
        if (tc) atomicAdd(rp.counters + 0, (unsigned long long)tc);
        atomicAdd(rp.counters + 1, (unsigned long long)ta);
        atomicAdd(rp.counters + 2, (unsigned long long)tn);
        atomicAdd(rp.counters + 3, (unsigned long long)tr);
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

struct TileUnit {      /* one unit of the current tile, in shared memory */
    UnitDesc d;
    uint32_t off[2];   /* stage offsets of its records */
    uint32_t pad_[2];
};
/* header fields of record r of a unit (simulate.rs:287-298, :513-518) */
__device__ __forceinline__ void record_positions(const UnitDesc &d, uint32_t r, uint64_t &sp, uint64_t &ep) {
    if (!(d.flags & 4u)) { sp = d.p0; ep = d.p1; }
    else if (r == 0) { sp = d.p0; ep = d.p0 + d.rl; }
    else { sp = d.p1 + d.rl; ep = d.p1; }
}

/* Persistent CTAs; each tile = up to `upt` consecutive units whose FASTQ bytes are
 * composed in shared memory and leave through one bulk async copy (TMA).  A tile whose
 * bytes exceed the stage is cut into sub-tiles.  TK = table kind the fast path is
 * specialised for (REPLAY always takes the generic per-base path). */
template <bool REPLAY, int TK>
__global__ void __launch_bounds__(FORMAT_THREADS, FORMAT_MIN_CTAS) k_format_tiles(RunParams rp, ReplayDev rd) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t next_unit;
    const uint32_t tab_words = rp.phred.kind == PT_CONST ? 0 : (rp.phred.npos << rp.phred.cb) * (rp.phred.kind == PT_T64 ? 2u : 1u);
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
    TileUnit *tu = reinterpret_cast<TileUnit *>(smem + ((tab_words * 4 + 15) & ~15u));
    uint8_t *stage = reinterpret_cast<uint8_t *>(tu + rp.upt);
    for (uint32_t i = threadIdx.x; i < tab_words; i += FORMAT_THREADS) tab[i] = rp.phred.entries[i];
    const uint32_t tab_s = uint32_t(__cvta_generic_to_shared(tab)), stage_s = uint32_t(__cvta_generic_to_shared(stage));

    const uint32_t ntiles = (rp.batch_units + rp.upt - 1) / rp.upt;
    const uint32_t lane = threadIdx.x & 31, sl = threadIdx.x & (SUB - 1);
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t t0 = tile * rp.upt, tn = min(rp.upt, rp.batch_units - t0);
        for (uint32_t s0 = 0; s0 < tn;) {
            /* sub-tile = longest run of units from s0 that fits the stage */
            const uint32_t u0 = t0 + s0;
            const uint64_t tile_base = rp.offsets[u0];
            const uint32_t pad = uint32_t(tile_base & 15);
            const bool fits = threadIdx.x < tn - s0 && rp.offsets[u0 + threadIdx.x + 1] - tile_base + pad <= rp.stage_cap;
            /* the previous bulk copy must have finished reading the stage before it is rewritten */
            if (threadIdx.x == 0) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); next_unit = 0; }
            const uint32_t nu = __syncthreads_count(fits);
            const uint32_t total = uint32_t(rp.offsets[u0 + nu] - tile_base);
            if (threadIdx.x < nu) {
                TileUnit &t = tu[threadIdx.x];
                t.d = rp.desc[u0 + threadIdx.x];
                t.off[0] = pad + uint32_t(rp.offsets[u0 + threadIdx.x] - tile_base);
                t.off[1] = t.off[0] + t.d.H0 + 2u * t.d.L0 + 5u;
            }
            __syncthreads();
            const uint32_t nrec = 2 * nu;
            /* headers: one thread per record */
            for (uint32_t r = threadIdx.x; r < nrec && !(rp.flags & 0x200u); r += FORMAT_THREADS) {
                const TileUnit &t = tu[r >> 1]; const uint32_t mate = r & 1;
                if (mate == 0 || (t.d.flags & 4u)) {
                    uint64_t sp, ep; record_positions(t.d, mate, sp, ep);
                    emit_header(stage + t.off[mate], rp, rp.chdr[t.d.contig], rp.read_id_base + uint32_t(rp.batch_begin + u0 + (r >> 1)),
                                sp, ep, mate != 0, mate);
                }
            }
            /* bodies: each warp pulls 32/SUB units at a time from a shared counter; its sub-groups
             * of SUB lanes take one unit each and run converged, mate 1 of all of them, then mate 2 */
            for (;;) {
                uint32_t ul = 0;
                if (lane == 0) ul = atomicAdd(&next_unit, 32u / SUB);
                ul = __shfl_sync(0xffffffffu, ul, 0);
                if (ul >= nu || (rp.flags & 0x400u)) break;
                ul += lane / SUB;
                const TileUnit &t = tu[min(ul, nu - 1)];
                const uint64_t unit = rp.batch_begin + u0 + ul;
#pragma unroll 1
                for (uint32_t mate = 0; mate < 2; ++mate) {
                    const bool have = ul < nu && (mate == 0 || (t.d.flags & 4u));
                    const bool fast = have && !REPLAY && !(t.d.flags & (1u << mate));
                    const uint32_t L = mate ? t.d.L1 : t.d.L0, body = t.off[mate] + (mate ? t.d.H1 : t.d.H0);
                    const uint64_t a = mate ? t.d.a1 : t.d.a0;
                    emit_body_fast<TK>(rp, tab_s, fast, stage_s + body, unit, mate, a, L, sl);
                    if (have && !fast)
                        emit_body_generic<REPLAY>(rp, rd, tab, stage + body, unit, mate, a, L,
                                                  (REPLAY ? rd.base_off[unit] : 0) + (mate ? t.d.rl : 0), sl, SUB);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            /* stage[pad, pad+total) -> out[tile_base, +total): 16-byte aligned middle by TMA, edges by byte stores */
            const uint32_t head = min(total, (16 - pad) & 15), mid = (total - head) & ~15u, tail = total - head - mid;
            if (rp.flags & 0x100u) { for (uint32_t b = threadIdx.x; b < mid; b += FORMAT_THREADS) rp.out[tile_base + head + b] = stage[pad + head + b]; }
            else if (threadIdx.x == 0 && mid) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(rp.out + tile_base + head), "r"(stage_s + pad + head), "r"(mid) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (threadIdx.x >= 32 && threadIdx.x < 32 + head) rp.out[tile_base + (threadIdx.x - 32)] = stage[pad + (threadIdx.x - 32)];
            if (threadIdx.x >= 64 && threadIdx.x < 64 + tail) rp.out[tile_base + head + mid + (threadIdx.x - 64)] = stage[pad + head + mid + (threadIdx.x - 64)];
            s0 += nu;
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

/* One CTA per unit, bytes written straight to global memory.  Used for long
 * reads (records far larger than a tile) and as the fallback when a unit does
 * not fit the shared-memory stage. */
template <bool REPLAY>
__global__ void __launch_bounds__(DIRECT_THREADS) k_format_direct(RunParams rp, ReplayDev rd) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tab_words = rp.phred.kind == PT_CONST ? 0 : (rp.phred.npos << rp.phred.cb) * (rp.phred.kind == PT_T64 ? 2u : 1u);
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
    for (uint32_t i = threadIdx.x; i < tab_words; i += DIRECT_THREADS) tab[i] = rp.phred.entries[i];
    __syncthreads();
    for (uint32_t i = blockIdx.x; i < rp.batch_units; i += gridDim.x) {
        const uint64_t u = rp.batch_begin + i;
        const UnitDesc d = rp.desc[i];
        uint8_t *dst = rp.out + rp.offsets[i];
        const uint64_t slot0 = REPLAY ? rd.base_off[u] : 0;
        const uint32_t nrec = (d.flags & 4u) ? 2 : 1;
        for (uint32_t r = 0; r < nrec; ++r) {
            const uint32_t H = r ? d.H1 : d.H0, L = r ? d.L1 : d.L0;
            if (threadIdx.x == 0) {
                uint64_t sp, ep; record_positions(d, r, sp, ep);
                emit_header(dst, rp, rp.chdr[d.contig], rp.read_id_base + uint32_t(u), sp, ep, r != 0, int(r));
            }
            emit_body_generic<REPLAY>(rp, rd, tab, dst + H, u, r, r ? d.a1 : d.a0, L, slot0 + (r ? d.rl : 0), threadIdx.x, DIRECT_THREADS);
            dst += H + 2 * L + 5;
        }
    }
}

}  // namespace simmr

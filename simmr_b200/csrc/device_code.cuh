/* device_code.cuh — sm_100a kernels of the simmr read-simulation hot path.
 *
 * Path (reference file:line -> kernel):
 *   simulate.rs:165-190, 205-262, 353-375, 464-491  geometry      -> unit_geometry()
 *   fastq.rs:34-57 header interpolation, record size             -> k_unit_sizes, emit_header()
 *   (no reference equivalent: byte offsets of the FASTQ stream)   -> k_scan_blocks, k_unit_offsets
 *   simulate.rs:261-286 + error_profiles/ *.rs phreds/mutation +
 *   util.rs:15-57 RC / +33 + fastq.rs:61-69 record layout         -> k_format_tiles, k_format_direct
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

struct RunParams {
    /* genomes (device resident) */
    const uint8_t *bases; const uint64_t *contig_off, *contig_size; const uint32_t *genome_contig;
    const ContigHdr *chdr; const HdrTok *toks; const uint8_t *pool;
    uint32_t n_genomes, n_contigs;
    /* unit space */
    const uint64_t *unit_start;    /* [n_genomes+1] first global unit of each genome */
    uint64_t batch_begin; uint32_t batch_units; uint32_t read_id_base;
    /* profile */
    PhredTab phred; LenTab rl, is; GammaConst gamma;
    uint32_t min_size, is_long, flags, mutates;
    uint32_t key0, key1;           /* Philox key = seed                          */
    /* per-batch scratch + output */
    uint32_t *sizes; uint64_t *offsets, *block_sums; uint8_t *out;
    unsigned long long *counters;  /* [0] clamped [1] algorithmic bytes [2] bases [3] records */
    /* tile geometry of k_format_tiles */
    uint32_t upt, stage_cap;
};

constexpr int SIZES_THREADS = 256;
constexpr int FORMAT_THREADS = 256;

/* ---------------------------------------------------------------------- */
/* Philox4x32-10                                                            */
/* ---------------------------------------------------------------------- */
struct Philox4 { uint32_t x, y, z, w; };
__device__ __forceinline__ Philox4 philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}
/* draw-schedule addressing: counter = (block, stream, unit) */
__device__ __forceinline__ Philox4 draw(const RunParams &rp, uint64_t unit, uint32_t stream, uint32_t blk) {
    return philox(rp.key0, rp.key1, blk, stream, uint32_t(unit), uint32_t(unit >> 32));
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

/* ---------------------------------------------------------------------- */
/* headers — fastq.rs:34-57                                                 */
/* ---------------------------------------------------------------------- */
__device__ __forceinline__ uint32_t dec_digits(uint64_t v) {
    uint32_t d = 1;
    while (v >= 10) { v /= 10; ++d; }
    return d;
}
__device__ __forceinline__ uint32_t header_len(const ContigHdr &ch, uint32_t id, uint64_t sp, uint64_t ep) {
    return ch.lit_bytes + ch.n_id * dec_digits(id) + ch.n_sp * dec_digits(sp) + ch.n_ep * dec_digits(ep) + ch.n_rc + ch.n_pair;
}
__device__ __forceinline__ uint32_t put_dec(uint8_t *dst, uint64_t v) {
    const uint32_t d = dec_digits(v);
    for (uint32_t i = d; i-- > 0;) { dst[i] = uint8_t('0' + v % 10); v /= 10; }
    return d;
}
__device__ uint32_t emit_header(uint8_t *dst, const RunParams &rp, const ContigHdr &ch, uint32_t id, uint64_t sp, uint64_t ep, bool rc, int mate) {
    uint32_t n = 0;
    const HdrTok *t = rp.toks + ch.tok_off;
    for (uint32_t i = 0; i < ch.ntok; ++i) {
        const HdrTok k = t[i];
        switch (k.type) {
        case TOK_LIT: { const uint8_t *s = rp.pool + k.off; for (uint32_t j = 0; j < k.len; ++j) dst[n + j] = __ldg(s + j); n += k.len; break; }
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
/* per-base work                                                            */
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

/* Writes one record body (everything after the header) for bases handled by
 * the calling group of `nthr` threads; `dst` may be shared or global memory.
 * Base i of the window gets quality slot i; for mate 2 the (mutated) base is
 * complemented and lands at L-1-i, the quality stays at i (simulate.rs:283-285). */
template <bool REPLAY>
__device__ __forceinline__ void emit_body(const RunParams &rp, const ReplayDev &rd, const uint32_t *tab, uint8_t *dst,
                                          uint64_t unit, uint32_t mate, uint64_t a, uint32_t L, uint64_t slot0, uint32_t thr, uint32_t nthr) {
    uint8_t *seq = dst + 1, *qual = dst + L + 4;
    if (thr == 0) { dst[0] = '\n'; dst[L + 1] = '\n'; dst[L + 2] = '+'; dst[L + 3] = '\n'; dst[2 * L + 4] = '\n'; }   /* fastq.rs:61-69 */
    const uint8_t *src = rp.bases + a;
    const bool rc = mate != 0;
    const uint32_t cb = rp.phred.cb, mshift = rp.phred.mshift, mmask = (1u << mshift) - 1u, npos = rp.phred.npos;
    for (uint32_t g = thr; g * 4 < L; g += nthr) {
        Philox4 w{};
        if (!REPLAY && !rp.phred.constant) w = draw(rp, unit, 1 + mate, g);
        const uint32_t wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t i = 4 * g + j;
            if (i >= L) break;
            uint32_t q, m;
            if (REPLAY) replay_base(rp, rd, slot0 + i, i, q, m);
            else if (rp.phred.constant) { q = rp.phred.const_q; m = 0; }
            else {
                const uint32_t r = wv[j], pos = i < npos ? i : npos - 1;
                const uint32_t idx = r >> (32 - cb), frac = (r << cb) >> 10;
                const uint32_t e = tab[(pos << cb) + idx];
                const uint32_t o = frac < (e >> 10) ? idx : (e & 0x3ffu);
                q = rp.phred.qlo + (o >> mshift); m = o & mmask;
            }
            uint8_t b = __ldg(src + i);
            if (m) b = substitute(b, m - 1);
            qual[i] = uint8_t(q + 33);                                            /* util.rs:46-50 */
            if (rc) seq[L - 1 - i] = complement(b); else seq[i] = b;
        }
    }
}

/* ---------------------------------------------------------------------- */
/* kernels                                                                  */
/* ---------------------------------------------------------------------- */
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

/* one thread per unit: FASTQ bytes of the unit, block totals, run statistics */
template <bool REPLAY>
__global__ void __launch_bounds__(SIZES_THREADS) k_unit_sizes(RunParams rp, ReplayDev rd) {
    __shared__ uint64_t sh[32];
    const uint32_t i = blockIdx.x * SIZES_THREADS + threadIdx.x;
    uint64_t bytes = 0, alg = 0, bases = 0, clamped = 0, recs = 0;
    if (i < rp.batch_units) {
        UnitGeo g; unit_geometry<REPLAY>(rp, rd, rp.batch_begin + i, g);
        const ContigHdr ch = rp.chdr[g.contig];
        for (uint32_t r = 0; r < g.nrec; ++r) {
            const uint32_t H = header_len(ch, g.id, g.sp[r], g.ep[r]);
            bytes += H + 2ull * g.L[r] + 5;
            alg += H + 3ull * g.L[r] + 5;                        /* BASELINE.md §3 */
            bases += g.L[r];
        }
        clamped = g.clamped; recs = g.nrec;
        rp.sizes[i] = uint32_t(bytes);
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

struct TileUnit {      /* geometry of one unit, kept in shared memory for the tile */
    UnitGeo g;
    uint32_t off[2], H[2];
    uint64_t slot0;
};

/* Persistent CTAs; each tile = `upt` consecutive units whose FASTQ bytes are
 * composed in shared memory and leave through one bulk async copy (TMA). */
template <bool REPLAY>
__global__ void __launch_bounds__(FORMAT_THREADS) k_format_tiles(RunParams rp, ReplayDev rd) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t tab_words = rp.phred.constant ? 0 : (rp.phred.npos << rp.phred.cb);
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
    TileUnit *tu = reinterpret_cast<TileUnit *>(smem + ((tab_words * 4 + 15) & ~15u));
    uint8_t *stage = reinterpret_cast<uint8_t *>(tu + rp.upt);
    stage += (16 - (reinterpret_cast<uintptr_t>(stage) & 15)) & 15;
    for (uint32_t i = threadIdx.x; i < tab_words; i += FORMAT_THREADS) tab[i] = rp.phred.entries[i];

    const uint32_t ntiles = (rp.batch_units + rp.upt - 1) / rp.upt;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t u0 = tile * rp.upt, nu = min(rp.upt, rp.batch_units - u0);
        /* the previous tile's bulk copy must have finished reading the stage */
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
        const uint64_t tile_base = rp.offsets[u0], tile_end = rp.offsets[u0 + nu];
        const uint32_t pad = uint32_t(tile_base & 15), total = uint32_t(tile_end - tile_base);
        if (threadIdx.x < nu) {
            TileUnit &t = tu[threadIdx.x];
            const uint64_t u = rp.batch_begin + u0 + threadIdx.x;
            unit_geometry<REPLAY>(rp, rd, u, t.g);
            const ContigHdr ch = rp.chdr[t.g.contig];
            uint32_t off = pad + uint32_t(rp.offsets[u0 + threadIdx.x] - tile_base);
            for (uint32_t r = 0; r < t.g.nrec; ++r) {
                t.H[r] = header_len(ch, t.g.id, t.g.sp[r], t.g.ep[r]);
                t.off[r] = off; off += t.H[r] + 2 * t.g.L[r] + 5;
            }
            t.slot0 = REPLAY ? rd.base_off[u] : 0;
        }
        __syncthreads();
        /* headers: one thread per record */
        for (uint32_t r = threadIdx.x; r < 2 * nu; r += FORMAT_THREADS) {
            const TileUnit &t = tu[r >> 1]; const uint32_t mate = r & 1;
            if (mate < t.g.nrec)
                emit_header(stage + t.off[mate], rp, rp.chdr[t.g.contig], t.g.id, t.g.sp[mate], t.g.ep[mate], mate != 0, mate);
        }
        /* bodies: one warp per record */
        for (uint32_t r = warp; r < 2 * nu; r += FORMAT_THREADS / 32) {
            const TileUnit &t = tu[r >> 1]; const uint32_t mate = r & 1;
            if (mate < t.g.nrec)
                emit_body<REPLAY>(rp, rd, tab, stage + t.off[mate] + t.H[mate], rp.batch_begin + u0 + (r >> 1), mate,
                                  t.g.a[mate], t.g.L[mate], t.slot0 + (mate ? t.g.rl : 0), lane, 32);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        /* stage[pad, pad+total) -> out[tile_base, tile_end): 16-byte aligned middle by TMA, edges by byte stores */
        const uint32_t head = min(total, (16 - pad) & 15), mid = (total - head) & ~15u, tail = total - head - mid;
        if (threadIdx.x == 0 && mid) {
            const uint32_t saddr = uint32_t(__cvta_generic_to_shared(stage + pad + head));
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(rp.out + tile_base + head), "r"(saddr), "r"(mid) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (threadIdx.x >= 32 && threadIdx.x < 32 + head) rp.out[tile_base + (threadIdx.x - 32)] = stage[pad + (threadIdx.x - 32)];
        if (threadIdx.x >= 64 && threadIdx.x < 64 + tail) rp.out[tile_base + head + mid + (threadIdx.x - 64)] = stage[pad + head + mid + (threadIdx.x - 64)];
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

/* One CTA per unit, bytes written straight to global memory.  Used for long
 * reads (records far larger than a tile) and as the fallback when a unit does
 * not fit the shared-memory stage. */
template <bool REPLAY>
__global__ void __launch_bounds__(FORMAT_THREADS) k_format_direct(RunParams rp, ReplayDev rd) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tab_words = rp.phred.constant ? 0 : (rp.phred.npos << rp.phred.cb);
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem);
    __shared__ UnitGeo g;
    __shared__ uint32_t H[2];
    for (uint32_t i = threadIdx.x; i < tab_words; i += FORMAT_THREADS) tab[i] = rp.phred.entries[i];
    for (uint32_t i = blockIdx.x; i < rp.batch_units; i += gridDim.x) {
        __syncthreads();
        const uint64_t u = rp.batch_begin + i;
        if (threadIdx.x == 0) {
            unit_geometry<REPLAY>(rp, rd, u, g);
            const ContigHdr ch = rp.chdr[g.contig];
            for (uint32_t r = 0; r < g.nrec; ++r) H[r] = header_len(ch, g.id, g.sp[r], g.ep[r]);
        }
        __syncthreads();
        uint8_t *dst = rp.out + rp.offsets[i];
        const uint64_t slot0 = REPLAY ? rd.base_off[u] : 0;
        for (uint32_t r = 0; r < g.nrec; ++r) {
            if (threadIdx.x == 0) emit_header(dst, rp, rp.chdr[g.contig], g.id, g.sp[r], g.ep[r], r != 0, int(r));
            emit_body<REPLAY>(rp, rd, tab, dst + H[r], u, r, g.a[r], g.L[r], slot0 + (r ? g.rl : 0), threadIdx.x, FORMAT_THREADS);
            dst += H[r] + 2 * g.L[r] + 5;
        }
    }
}

}  // namespace simmr

/*
 * simmr_oracle.cpp — CPU restatement of simmr's read-simulation path.
 *
 * TEST INFRASTRUCTURE ONLY (see simmr_oracle.h).  Plain C++17, no CUDA, no
 * dependency on anything under simmr_b200/.  Every function cites the
 * reference file:line it follows (paths relative to /root/reference/).
 *
 * PARITY: pinned to the reference's own known-answer vectors (SURVEY.md §8c,
 * tests/test_oracle_golden.py); the rand / rand_distr / needletail / bincode
 * semantics are restated from the crates' published algorithms and are
 * "parity unpinned" because the Rust reference cannot be run in this image.
 *
 * Build with -ffp-contract=off: the native-mode tables and the long-read
 * gamma sampler must give the same IEEE-754 results as the CUDA side.
 */
#include "simmr_oracle.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <thread>
#include <atomic>
#include <chrono>

namespace {

thread_local std::string g_err;
int fail(const std::string &m) { g_err = m; return -1; }

/* ------------------------------------------------------------------ */
/* Rust cast semantics                                                  */
/* ------------------------------------------------------------------ */
/* `x as u8` / `as u16` / `as usize` from a float: truncate toward zero,
 * saturate, NaN -> 0 (Rust reference, "Numeric cast" semantics). */
template <typename F> inline uint8_t sat_u8(F x) {
    if (!(x == x)) return 0;
    if (x <= (F)0) return 0;
    if (x >= (F)255) return 255;
    return (uint8_t)x;
}
template <typename F> inline uint16_t sat_u16(F x) {
    if (!(x == x)) return 0;
    if (x <= (F)0) return 0;
    if (x >= (F)65535) return 65535;
    return (uint16_t)x;
}
inline uint64_t sat_usize(double x) {
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 18446744073709551615.0) return UINT64_MAX;
    return (uint64_t)x;
}

/* ------------------------------------------------------------------ */
/* primitives — simmr/src/util.rs                                       */
/* ------------------------------------------------------------------ */
/* util.rs:15-27 (and shared/src/util.rs:15-27) */
inline uint8_t complement(uint8_t n) {
    switch (n) {
    case 'A': return 'T'; case 'a': return 't';
    case 'T': return 'A'; case 't': return 'a';
    case 'C': return 'G'; case 'c': return 'g';
    case 'G': return 'C'; case 'g': return 'c';
    default: return n;
    }
}
/* util.rs:32-37 */
std::vector<uint8_t> reverse_complement(const uint8_t *s, size_t n) {
    std::vector<uint8_t> r(n);
    for (size_t i = 0; i < n; i++) r[i] = complement(s[n - 1 - i]);
    return r;
}
/* util.rs:46-50 — u8 add; release builds wrap */
inline uint8_t encode_quality_score(uint8_t s) { return (uint8_t)(s + 33); }
/* util.rs:69-71 */
inline float phred_to_probability(uint8_t s) { return powf(10.0f, -((float)s / 10.0f)); }
/* util.rs:83-85 */
inline uint8_t probability_to_phred(float p) { return sat_u8(-10.0f * log10f(p)); }
/* util.rs:96-98 */
inline float phred_to_accuracy(uint8_t s) { return 1.0f - phred_to_probability(s); }
/* util.rs:109-111 — f32::round = half away from zero = roundf */
inline uint8_t accuracy_to_phred(float acc) { return sat_u8(roundf(-10.0f * log10f(1.0f - acc))); }

/* shared/src/util.rs:49-51 — from_utf8(..).unwrap_or("") */
bool utf8_valid(const uint8_t *s, size_t n) {
    size_t i = 0;
    while (i < n) {
        uint8_t c = s[i];
        if (c < 0x80) { i++; continue; }
        int len; uint32_t cp;
        if (c >= 0xC2 && c <= 0xDF) { len = 2; cp = c & 0x1F; }
        else if (c >= 0xE0 && c <= 0xEF) { len = 3; cp = c & 0x0F; }
        else if (c >= 0xF0 && c <= 0xF4) { len = 4; cp = c & 0x07; }
        else return false;
        if (i + len > n) return false;
        for (int k = 1; k < len; k++) {
            if ((s[i + k] & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (s[i + k] & 0x3F);
        }
        if (len == 3 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return false;
        if (len == 4 && (cp < 0x10000 || cp > 0x10FFFF)) return false;
        i += len;
    }
    return true;
}
std::string bytes_to_string(const uint8_t *s, size_t n) {
    return utf8_valid(s, n) ? std::string((const char *)s, n) : std::string();
}

/* custom_short.rs:41-49, custom_long.rs (same) */
double gaussian(double x, const double *xs, size_t n, double bw) {
    double sum = 0.0;
    for (size_t i = 0; i < n; i++) {
        double f = (x - xs[i]) / bw;
        sum += std::exp(-0.5 * (f * f));
    }
    return sum / (std::sqrt(2.0 * M_PI) * (double)n * bw);
}

/* shared/src/encoding.rs:149-179 */
int kmer_encode(const uint8_t *kmer, uint32_t k, uint32_t esize, uint32_t *out) {
    uint32_t enc = 0;
    for (uint32_t i = 0; i < k; i++) {
        uint32_t v;
        switch (kmer[i]) {
        case 'A': continue;
        case 'C': v = 1; break;
        case 'G': v = 2; break;
        case 'T': v = 3; break;
        case 'N': if (esize < 3) return -1; v = 4; break;
        default: return -1;
        }
        uint32_t mask = 0xffffffffu;
        for (uint32_t b = 0; b < esize; b++) mask &= ~(1u << (i * esize + b));
        enc = (enc & mask) | (v << (i * esize));
    }
    *out = enc;
    return 0;
}
/* shared/src/encoding.rs:186-210; returns decoded length or -1 */
int kmer_decode(uint32_t code, uint32_t k, uint32_t esize, bool skip_n, uint8_t *out) {
    int n = 0;
    for (uint32_t i = 0; i < k; i++) {
        uint32_t v = ((1u << esize) - 1u) & (code >> (i * esize));
        switch (v) {
        case 0: out[n++] = 'A'; break;
        case 1: out[n++] = 'C'; break;
        case 2: out[n++] = 'G'; break;
        case 3: out[n++] = 'T'; break;
        case 4: if (!skip_n) out[n++] = 'N'; break;
        default: return -1;
        }
    }
    return n;
}

/* Rust `{}` for f64: shortest round-trip digits, never scientific. */
std::string format_f64(double v) {
    if (v != v) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[512];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11; Random123)          */
/* ------------------------------------------------------------------ */
struct Philox {
    uint32_t k0, k1;
    explicit Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    void block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) const {
        uint32_t a = k0, b = k1;
        for (int r = 0; r < 10; r++) {
            uint64_t p0 = (uint64_t)0xD2511F53u * c0;
            uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
            uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a;
            uint32_t n1 = (uint32_t)p1;
            uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ b;
            uint32_t n3 = (uint32_t)p0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
    /* draw-schedule addressing: ctr = (blk, stream, unit) */
    void at(uint64_t unit, uint32_t stream, uint32_t blk, uint32_t out[4]) const {
        block(blk, stream, (uint32_t)unit, (uint32_t)(unit >> 32), out);
    }
};
/* Per-base draws of native mode.  A read is cut into pieces (short reads: one piece; long reads: 32 pieces
 * of 16 * ceil(L / 512) bases); Philox block (piece, stream, unit) seeds the piece's two MWC64X chains (D. B. Thomas:
 * x' = lo(A*x + c), c' = hi(A*x + c), A = 4294883355) — words 0,1 the quality chain (one step per two bases), words 2,3
 * the substitution-event chain. */
/* word = x ^ c (its low bits pick a table column), level = x (compared with the entry); the quality chain serves a second
 * base from the same step: column from the next cb bits of the word, level2 = c */
struct Draw { uint32_t word, level, level2; };
struct Chain {
    uint32_t x, c;
    Chain(uint32_t w_even, uint32_t w_odd) : x(w_even), c((w_odd >> 2) + 1) {}
    Draw next() {
        const Draw d{x ^ c, x, c};
        const uint64_t t = (uint64_t)x * 4294883355u + c;
        x = (uint32_t)t; c = (uint32_t)(t >> 32);
        return d;
    }
};
inline uint64_t u64_of(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }
inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* ------------------------------------------------------------------ */
/* deterministic ln(x), x>0 normal: only IEEE +,-,*,/,fma */
/* ------------------------------------------------------------------ */
double det_log(double x) {
    uint64_t bits; memcpy(&bits, &x, 8);
    int e = (int)((bits >> 52) & 0x7ff) - 1022;
    bits = (bits & 0x000fffffffffffffull) | 0x3fe0000000000000ull;  /* m in [0.5,1) */
    double m; memcpy(&m, &bits, 8);
    if (m < 0.70710678118654752440) { m = m * 2.0; e -= 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 23.0;
    p = std::fma(p, z, 1.0 / 21.0);
    p = std::fma(p, z, 1.0 / 19.0);
    p = std::fma(p, z, 1.0 / 17.0);
    p = std::fma(p, z, 1.0 / 15.0);
    p = std::fma(p, z, 1.0 / 13.0);
    p = std::fma(p, z, 1.0 / 11.0);
    p = std::fma(p, z, 1.0 / 9.0);
    p = std::fma(p, z, 1.0 / 7.0);
    p = std::fma(p, z, 1.0 / 5.0);
    p = std::fma(p, z, 1.0 / 3.0);
    double t = std::fma(p * z, s, s);           /* atanh(s) */
    return std::fma((double)e, 0.69314718055994530942, 2.0 * t);
}

/* ------------------------------------------------------------------ */
/* custom-profile blob — shared/src/encoding.rs:82-117 (bincode 1.3.3)  */
/* ------------------------------------------------------------------ */
struct Bins {
    uint64_t num_bins = 0, bin_width = 0;
    std::vector<double> density;
    std::vector<std::pair<uint32_t, uint32_t>> ranges;
};
struct Blob {
    uint64_t bin_size = 0;
    std::vector<Bins> quality;
    uint8_t bit_encoding = 0;
    uint64_t kmer_size = 0;
    std::vector<std::pair<uint32_t, std::vector<std::pair<uint32_t, float>>>> probs;
    double is_mean = 0, is_std = 0;
    bool has_is_bins = false;
    Bins is_bins;
    double rl_mean = 0, rl_std = 0;
    Bins rl_bins;
    bool is_long = false;
};
struct Cur {
    const uint8_t *p; uint64_t n, i = 0; bool ok = true;
    template <typename T> T get() {
        T v{};
        if (i + sizeof(T) > n) { ok = false; return v; }
        memcpy(&v, p + i, sizeof(T)); i += sizeof(T); return v;
    }
};
bool read_bins(Cur &c, Bins &b) {
    b.num_bins = c.get<uint64_t>(); b.bin_width = c.get<uint64_t>();
    uint64_t n = c.get<uint64_t>();
    if (!c.ok || n > (c.n - c.i) / 8) return false;
    b.density.resize(n);
    for (auto &d : b.density) d = c.get<double>();
    uint64_t m = c.get<uint64_t>();
    if (!c.ok || m > (c.n - c.i) / 8) return false;
    b.ranges.resize(m);
    for (auto &r : b.ranges) { r.first = c.get<uint32_t>(); r.second = c.get<uint32_t>(); }
    return c.ok;
}
bool parse_blob(const uint8_t *p, uint64_t n, Blob &b) {
    Cur c{p, n};
    b.bin_size = c.get<uint64_t>();
    uint64_t np = c.get<uint64_t>();
    if (!c.ok || np > n) return false;
    b.quality.resize(np);
    for (auto &q : b.quality) if (!read_bins(c, q)) return false;
    b.bit_encoding = c.get<uint8_t>();
    b.kmer_size = c.get<uint64_t>();
    uint64_t nk = c.get<uint64_t>();
    if (!c.ok || nk > n) return false;
    b.probs.resize(nk);
    for (auto &kp : b.probs) {
        kp.first = c.get<uint32_t>();
        uint64_t na = c.get<uint64_t>();
        if (!c.ok || na > n) return false;
        kp.second.resize(na);
        for (auto &a : kp.second) { a.first = c.get<uint32_t>(); a.second = c.get<float>(); }
    }
    b.is_mean = c.get<double>(); b.is_std = c.get<double>();
    uint8_t tag = c.get<uint8_t>();
    b.has_is_bins = tag != 0;
    if (b.has_is_bins && !read_bins(c, b.is_bins)) return false;
    b.rl_mean = c.get<double>(); b.rl_std = c.get<double>();
    if (!read_bins(c, b.rl_bins)) return false;
    b.is_long = c.get<uint8_t>() != 0;
    return c.ok;
}
/* bin range for density index j; j beyond the ranges continues unit bins
 * (the reference would panic: simmrd emits 71 densities / 70 ranges,
 * SURVEY.md §A.4) */
inline std::pair<uint32_t, uint32_t> bin_range(const Bins &b, size_t j) {
    if (j < b.ranges.size()) return b.ranges[j];
    if (b.ranges.empty()) return {(uint32_t)j, (uint32_t)j};
    uint32_t v = b.ranges.back().second + 1 + (uint32_t)(j - b.ranges.size());
    return {v, v};
}

/* ------------------------------------------------------------------ */
/* rand_distr 0.4.3 WeightedAliasIndex<f64> (weighted_alias.rs)          */
/* ------------------------------------------------------------------ */
double pairwise_sum(const double *v, size_t n) {
    if (n <= 32) { double s = 0.0; for (size_t i = 0; i < n; i++) s += v[i]; return s; }
    size_t mid = n / 2;
    return pairwise_sum(v, mid) + pairwise_sum(v + mid, n - mid);
}
struct RustAlias {
    std::vector<double> odds; std::vector<uint32_t> alias; double sum = 0;
    bool build(const std::vector<double> &w) {
        size_t n = w.size();
        if (n == 0) return false;
        sum = pairwise_sum(w.data(), n);
        if (!(sum > 0.0)) return false;
        odds = w;
        for (auto &o : odds) o *= (double)n;
        alias.assign(n, 0);
        const uint32_t NIL = UINT32_MAX;
        uint32_t smalls = NIL, bigs = NIL;
        auto push = [&](uint32_t &head, uint32_t i) { alias[i] = head; head = i; };
        auto pop = [&](uint32_t &head) { uint32_t p = head; head = alias[p]; return p; };
        for (size_t i = 0; i < n; i++) { if (odds[i] < sum) push(smalls, (uint32_t)i); else push(bigs, (uint32_t)i); }
        while (smalls != NIL && bigs != NIL) {
            uint32_t s = pop(smalls), b = pop(bigs);
            alias[s] = b;
            odds[b] = odds[b] - sum + odds[s];
            if (odds[b] < sum) push(smalls, b); else push(bigs, b);
        }
        while (smalls != NIL) odds[pop(smalls)] = sum;
        while (bigs != NIL) odds[pop(bigs)] = sum;
        return true;
    }
    /* weighted_alias.rs sample(): candidate, then `u < no_alias_odds[cand]` */
    uint32_t pick(uint32_t cand, double u) const { return u < odds[cand] ? cand : alias[cand]; }
};

/* ------------------------------------------------------------------ */
/* native-mode quantised alias tables                    */
/* ------------------------------------------------------------------ */
/* Vose alias over N=2^cb columns: keep-probability and alias per column. */
bool vose(const std::vector<double> &w, uint32_t cb, std::vector<double> &prob, std::vector<uint32_t> &alias) {
    const size_t N = (size_t)1 << cb;
    if (w.size() > N) return false;
    double S = 0.0;
    for (double x : w) S += x;
    if (!(S > 0.0)) return false;
    std::vector<double> p(N, 0.0);
    for (size_t i = 0; i < w.size(); i++) p[i] = w[i] * (double)N / S;
    std::vector<uint32_t> small, large;
    alias.resize(N); prob.assign(N, 1.0);
    for (size_t i = 0; i < N; i++) { alias[i] = (uint32_t)i; if (p[i] < 1.0) small.push_back((uint32_t)i); else large.push_back((uint32_t)i); }
    while (!small.empty() && !large.empty()) {
        uint32_t s = small.back(); small.pop_back();
        uint32_t l = large.back(); large.pop_back();
        prob[s] = p[s]; alias[s] = l;
        p[l] = (p[l] + p[s]) - 1.0;
        if (p[l] < 1.0) small.push_back(l); else large.push_back(l);
    }
    for (size_t i = 0; i < N; i++) if (prob[i] >= 1.0) alias[i] = (uint32_t)i;
    return true;
}
/* threshold = floor(prob * 2^bits), saturated to bits; prob>=1 -> all ones */
inline uint32_t quant_thr(double prob, uint32_t bits) {
    const uint32_t top = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
    if (prob >= 1.0) return top;
    double t = std::floor(prob * std::ldexp(1.0, (int)bits));
    return t <= 0.0 ? 0u : (t >= (double)top ? top : (uint32_t)t);
}
/* length tables: entry = thr20<<12 | alias12 */
bool build_len_alias(const std::vector<double> &w, uint32_t cb, std::vector<uint32_t> &out) {
    std::vector<double> prob; std::vector<uint32_t> alias;
    if (!vose(w, cb, prob, alias)) return false;
    out.resize(prob.size());
    for (size_t i = 0; i < prob.size(); i++) out[i] = (quant_thr(prob[i], 20) << 12) | alias[i];
    return true;
}
inline uint32_t ceil_log2(size_t n) { uint32_t b = 0; while (((size_t)1 << b) < n) b++; return b; }
inline double normal_cdf(double x, double mu, double sd) { return 0.5 * std::erfc(-(x - mu) / (sd * 1.4142135623730951)); }

/* pmf of `floor(Normal(mu,sd)) as uN` (saturating) on the window [vlo,vhi];
 * tails beyond 8 sd are lumped on the window edges. */
void normal_floor_pmf(double mu, double sd, uint32_t vmax, uint32_t &vlo, std::vector<double> &pmf) {
    double lo = std::floor(mu - 8.0 * sd), hi = std::ceil(mu + 8.0 * sd);
    vlo = lo <= 0.0 ? 0u : (lo >= (double)vmax ? vmax : (uint32_t)lo);
    uint32_t vhi = hi <= 0.0 ? 0u : (hi >= (double)vmax ? vmax : (uint32_t)hi);
    pmf.assign(vhi - vlo + 1, 0.0);
    for (uint32_t v = vlo; v <= vhi; v++) {
        double a = (v == vlo) ? 0.0 : normal_cdf((double)v, mu, sd);
        double b = (v == vhi) ? 1.0 : normal_cdf((double)v + 1.0, mu, sd);
        pmf[v - vlo] = b - a;
    }
}
/* P(mutate | q): #{k in [0,2^24): k*2^-24 > accuracy(q)} / 2^24
 * (minimal_short.rs:123 with rand's 24-bit gen::<f32>()) */
inline double mutation_prob(uint8_t q) {
    double a = (double)phred_to_accuracy(q) * 16777216.0;
    double T = std::floor(a);
    if (T < 0.0) return 1.0;
    if (T >= 16777215.0) return 0.0;
    return (16777215.0 - T) / 16777216.0;
}

struct PhredTable {      /* per-position alias tables */
    uint32_t kind = 0;   /* 0: constant quality; 3: thr16<<16 | alias<<8 | self per column (kinds 1 and 2 are retired) */
    uint8_t const_q = 0;
    uint32_t cb = 0, npos = 0;
    std::vector<uint32_t> entries;   /* one word per column, npos rows */
    /* kind 3: sampled with column = low cb bits of the draw's word and "self iff the draw's level < entry";
     * a mutating profile adds gap[256] (bases until the next substitution, 255 = none within 255) and
     * qsub[1 << cb] (quality given a substitution) */
    bool has_sub = false;
    std::vector<uint32_t> gap, qsub;
};
struct LenTable {
    uint32_t kind = 0;   /* 0 const, 1 alias over bins, 2 Marsaglia-Tsang gamma */
    uint32_t value = 0;  /* const */
    uint32_t cb = 0;
    std::vector<uint32_t> entries;   /* thr20<<12 | alias12 */
    std::vector<uint32_t> lo, hi;    /* per column */
};

/* k-mer error injection (custom_long.rs:187-232 == custom_short.rs:452-516), native tables: per observed
 * reference k-mer the chance that the sampled alternate is the k-mer itself, as a 32-bit threshold on a draw's level, and
 * an alias row (entry = thr20<<12 | alias12, as the length tables) over the other alternates.  Rows are padded to
 * 1 << cb columns in both `entries` and `alts`, so one offset addresses both. */
struct ErrTable {
    bool on = false;
    uint32_t k = 0;
    std::vector<uint32_t> ref;        /* distinct reference k-mers (3-bit codes), ascending */
    std::vector<uint32_t> keep_thr, row, cb;
    std::vector<uint32_t> entries, alts;
    /* for the tests: position in the blob's alternate list of (a) the first alternate equal to the k-mer, (b) every other alternate in row order */
    std::vector<uint32_t> blob_entry, id_choice; std::vector<std::vector<uint32_t>> other_choice;
    int find(uint32_t code) const { auto it = std::lower_bound(ref.begin(), ref.end(), code); return it != ref.end() && *it == code ? int(it - ref.begin()) : -1; }
};
bool build_err_table(const Blob &b, ErrTable &t) {
    t = ErrTable();
    if (b.probs.empty()) return true;
    if (b.bit_encoding != 3 || b.kmer_size < 1 || b.kmer_size > 10) return false;      /* simmrd/src/main.rs:382, cli.rs:146-156 */
    t.k = (uint32_t)b.kmer_size; t.on = true;
    const uint32_t mask = t.k == 10 ? 0x3fffffffu : ((1u << (3 * t.k)) - 1u);
    std::map<uint32_t, uint32_t> last;                       /* HashMap collect(): the last entry of a k-mer wins */
    for (uint32_t i = 0; i < b.probs.size(); i++) last[b.probs[i].first] = i;
    for (auto &kv : last) {
        const auto &alts = b.probs[kv.second].second;
        double w_id = 0.0, w_all = 0.0; std::vector<double> w_other; std::vector<uint32_t> c_other, ch_other; uint32_t id_at = UINT32_MAX;
        for (uint32_t j = 0; j < alts.size(); j++) {
            const uint32_t code = alts[j].first & mask;
            for (uint32_t d = 0; d < t.k; d++) if (((code >> (3 * d)) & 7u) > 4u) return false;     /* decode_kmer would fail (encoding.rs:205) */
            const double w = (double)alts[j].second;
            if (!(w >= 0.0)) return false;
            w_all += w;
            if (code == (kv.first & mask)) { w_id += w; if (id_at == UINT32_MAX) id_at = j; }
            else if (w > 0.0) { w_other.push_back(w); c_other.push_back(code); ch_other.push_back(j); }
        }
        if (!(w_all > 0.0)) continue;                        /* WeightedAliasIndex::new would fail: treated as "not observed" */
        if (w_other.size() > 4096) return false;
        t.ref.push_back(kv.first); t.blob_entry.push_back(kv.second); t.id_choice.push_back(id_at); t.other_choice.push_back(ch_other);
        t.row.push_back((uint32_t)t.entries.size());
        if (w_other.empty()) {                               /* always the k-mer itself */
            t.keep_thr.push_back(0xffffffffu); t.cb.push_back(0);
            t.entries.push_back(0xfffff000u); t.alts.push_back(kv.first & mask);
            continue;
        }
        t.keep_thr.push_back(quant_thr(w_id / w_all, 32));
        const uint32_t cb = ceil_log2(w_other.size());
        t.cb.push_back(cb);
        std::vector<uint32_t> e;
        if (!build_len_alias(w_other, cb, e)) return false;
        t.entries.insert(t.entries.end(), e.begin(), e.end());
        for (uint32_t j = 0; j < (1u << cb); j++) t.alts.push_back(j < c_other.size() ? c_other[j] : (kv.first & mask));
    }
    if (t.ref.empty()) t.on = false;
    return true;
}

struct Profile {
    int kind = 0;
    uint16_t read_length = 150, insert_size = 150;
    uint8_t mean_phred = 30;
    uint32_t flags = 0;
    Blob blob; bool has_blob = false;
    bool is_long = false, mutates = false;
    uint16_t min_size = 0;
    PhredTable phred;
    LenTable rl, is;
    ErrTable err;                    /* custom-long: k-mer error injection */
    /* replay helpers */
    std::vector<RustAlias> q_alias;
};

/* kind 3 word: a column whose 16-bit threshold is zero names its alias twice, so that the payload bits of the word
 * cannot make the sampler keep a column that has no mass */
inline uint32_t q32_entry(double prob, uint32_t self_v, uint32_t alias_v) {
    const uint32_t thr = quant_thr(prob, 16);
    if (thr == 0) self_v = alias_v;
    return (thr << 16) | ((alias_v & 0xffu) << 8) | (self_v & 0xffu);
}
bool q32_table(const std::vector<double> &w, uint32_t vlo, uint32_t cb, std::vector<uint32_t> &out) {
    std::vector<double> prob; std::vector<uint32_t> alias;
    if (!vose(w, cb, prob, alias)) return false;
    out.resize(prob.size());
    for (size_t i = 0; i < prob.size(); i++) out[i] = q32_entry(prob[i], vlo + (uint32_t)i, vlo + alias[i]);
    return true;
}
/* Mutating profiles.  minimal_short.rs:83-140 (and its twins minimal_long.rs:78-140, perfect_long.rs:60-119) draws q, then substitutes with probability pm(q),
 * independently per base.  The same joint law, factored the other way: a base is substituted with the marginal
 * probability s = sum_q P(q) pm(q); a substituted base's quality follows P(q | substituted), any other base's
 * P(q | kept).  The substituted bases of a record are then found by geometric gaps. */
bool phred_table_from_pmf(const std::vector<double> &pq, uint32_t qlo, PhredTable &t) {
    const size_t n = pq.size();
    std::vector<double> kept(n), subst(n);
    double s = 0.0;
    for (size_t i = 0; i < n; i++) {
        const double pm = mutation_prob((uint8_t)(qlo + i));
        kept[i] = pq[i] * (1.0 - pm); subst[i] = pq[i] * pm;
        s += subst[i];
    }
    t.kind = 3; t.npos = 1; t.cb = std::max(1u, ceil_log2(n));
    if (t.cb > 8) return false;
    if (!q32_table(kept, qlo + 33u, t.cb, t.entries)) return false;    /* kind 3 words hold the printed characters (util.rs:46-50) */
    t.has_sub = s > 0.0;
    if (t.has_sub) {
        if (!q32_table(subst, qlo + 33u, t.cb, t.qsub)) return false;
        std::vector<double> gap(256);
        double none = 1.0;
        for (uint32_t k = 0; k < 255; k++) { gap[k] = s * none; none = none * (1.0 - s); }
        gap[255] = none;
        if (!q32_table(gap, 0, 8, t.gap)) return false;
    }
    return true;
}
/* per-position tables from the blob: alias over densities, uniform in-bin,
 * `as u8` truncation (custom_short.rs:108-129, 328-352); no mutation */
bool phred_table_from_blob(const Blob &b, PhredTable &t) {
    size_t np = b.quality.size();
    if (np == 0) return false;
    std::vector<std::vector<double>> pm(np, std::vector<double>(256, 0.0));
    uint32_t qmin = 255, qmax = 0;
    for (size_t p = 0; p < np; p++) {
        const Bins &bn = b.quality[p];
        if (bn.density.empty()) return false;     /* reference: NoItem -> unwrap panic */
        double W = 0.0;
        for (double d : bn.density) W += d;
        if (!(W > 0.0)) return false;
        for (size_t j = 0; j < bn.density.size(); j++) {
            auto r = bin_range(bn, j);
            uint64_t lo = r.first, hi = std::max(r.first, r.second), cnt = hi - lo + 1;
            if (!(bn.density[j] > 0.0)) continue;
            /* each value of the bin is equally likely; `as u8` folds it mod 256 */
            uint64_t full = cnt / 256, rem = cnt % 256;
            for (uint64_t v = 0; v < 256; v++) {
                uint64_t hits = full + (((v + 256 - (lo & 255)) & 255) < rem ? 1 : 0);
                if (!hits) continue;
                pm[p][v] += bn.density[j] / W * ((double)hits / (double)cnt);
                qmin = std::min<uint32_t>(qmin, (uint32_t)v); qmax = std::max<uint32_t>(qmax, (uint32_t)v);
            }
        }
    }
    if (qmin > qmax) return false;
    /* columns of a row = its qualities with mass above 2^-24 of the row (rarer ones would get a zero threshold), ascending;
     * table width = the richest row */
    std::vector<double> cut(np);
    size_t widest = 1;
    for (size_t p = 0; p < np; p++) {
        double row = 0.0;
        for (uint32_t v = qmin; v <= qmax; v++) row += pm[p][v];
        cut[p] = std::ldexp(row, -24);
        size_t n = 0;
        for (uint32_t v = qmin; v <= qmax; v++) if (pm[p][v] > cut[p]) n++;
        widest = std::max(widest, n);
    }
    t.kind = 3; t.npos = (uint32_t)np;
    t.cb = std::max(1u, ceil_log2(widest));
    t.entries.clear();
    for (size_t p = 0; p < np; p++) {
        std::vector<double> w, prob; std::vector<uint32_t> val, alias;
        for (uint32_t v = qmin; v <= qmax; v++) if (pm[p][v] > cut[p]) { w.push_back(pm[p][v]); val.push_back(v); }
        if (!vose(w, t.cb, prob, alias)) return false;
        val.resize(prob.size(), val.back());               /* empty columns always take their alias */
        for (size_t i = 0; i < prob.size(); i++)
            t.entries.push_back(q32_entry(prob[i], val[i] + 33u, val[alias[i]] + 33u));
    }
    return true;
}
bool len_table_from_pmf(const std::vector<double> &pmf, uint32_t vlo, LenTable &t) {
    t.kind = 1; t.cb = std::max(1u, ceil_log2(pmf.size()));
    if (t.cb > 12) return false;
    size_t N = (size_t)1 << t.cb;
    t.lo.assign(N, vlo); t.hi.assign(N, vlo);
    for (size_t i = 0; i < pmf.size(); i++) t.lo[i] = t.hi[i] = vlo + (uint32_t)i;
    return build_len_alias(pmf, t.cb, t.entries);
}
bool len_table_from_bins(const Bins &b, LenTable &t) {
    if (b.density.empty()) return false;
    t.kind = 1; t.cb = std::max(1u, ceil_log2(b.density.size()));
    if (t.cb > 12) return false;
    size_t N = (size_t)1 << t.cb;
    t.lo.assign(N, 0); t.hi.assign(N, 0);
    for (size_t j = 0; j < b.density.size(); j++) { auto r = bin_range(b, j); t.lo[j] = r.first; t.hi[j] = std::max(r.first, r.second); }
    return build_len_alias(b.density, t.cb, t.entries);
}

/* perfect_long.rs:60-78 as a pmf over q: a = N(acc(20),0.05) (f32), min 0.9999,
 * q = round(-10 log10(1-a)) */
void perfect_long_pmf(std::vector<double> &pq) {
    const double mu = (double)phred_to_accuracy(20), sd = (double)0.05f;
    uint8_t qcap = accuracy_to_phred(0.9999f);
    pq.assign((size_t)qcap + 1, 0.0);
    auto a_of = [](int k) { return 1.0 - std::pow(10.0, -((double)k - 0.5) / 10.0); };   /* smallest a with q>=k */
    for (int k = 0; k <= qcap; k++) {
        double lo = (k == 0) ? 0.0 : normal_cdf(a_of(k), mu, sd);
        double hi = (k == qcap) ? 1.0 : normal_cdf(a_of(k + 1), mu, sd);
        pq[k] = hi - lo;
    }
}

bool make_profile(const orc_profile *in, Profile &p) {
    p.kind = in->kind; p.read_length = in->read_length; p.insert_size = in->insert_size;
    p.mean_phred = in->mean_phred; p.flags = in->flags;
    if (in->kind == ORC_CUSTOM_SHORT || in->kind == ORC_CUSTOM_LONG) {
        if (!in->blob || !parse_blob(in->blob, in->blob_len, p.blob)) { fail("cannot parse custom profile blob"); return false; }
        p.has_blob = true;
    }
    switch (in->kind) {
    case ORC_PERFECT_SHORT:   /* perfect_short.rs:22-59 */
        p.is_long = false; p.mutates = false;
        p.min_size = (uint16_t)(2 * p.read_length + p.insert_size);
        p.phred.kind = 0; p.phred.const_q = 60;
        p.rl.kind = 0; p.rl.value = p.read_length; p.is.kind = 0; p.is.value = p.insert_size;
        break;
    case ORC_MINIMAL_SHORT: { /* minimal_short.rs:33-145; sd 15 / 75 from cli.rs:239-240 */
        p.is_long = false; p.mutates = true;
        p.min_size = (uint16_t)(2 * p.read_length + p.insert_size);
        uint32_t vlo; std::vector<double> pmf;
        normal_floor_pmf((double)(float)p.mean_phred, 10.0, 255, vlo, pmf);
        if (!phred_table_from_pmf(pmf, vlo, p.phred)) { fail("phred table"); return false; }
        normal_floor_pmf((double)p.read_length, 15.0, 65535, vlo, pmf);
        if (!len_table_from_pmf(pmf, vlo, p.rl)) { fail("rl table"); return false; }
        normal_floor_pmf((double)p.insert_size, 75.0, 65535, vlo, pmf);
        if (!len_table_from_pmf(pmf, vlo, p.is)) { fail("is table"); return false; }
        break; }
    case ORC_PERFECT_LONG: {  /* perfect_long.rs:36-135 */
        p.is_long = true; p.mutates = true; p.min_size = 20000;
        std::vector<double> pq; perfect_long_pmf(pq);
        if (!phred_table_from_pmf(pq, 0, p.phred)) { fail("phred table"); return false; }
        p.rl.kind = 2; p.is.kind = 0; p.is.value = 0;
        break; }
    case ORC_MINIMAL_LONG: {  /* minimal_long.rs:44-158 */
        p.is_long = true; p.mutates = true; p.min_size = 20000;
        uint32_t vlo; std::vector<double> pmf;
        normal_floor_pmf((double)(float)p.mean_phred, 10.0, 255, vlo, pmf);
        if (!phred_table_from_pmf(pmf, vlo, p.phred)) { fail("phred table"); return false; }
        p.rl.kind = 2; p.is.kind = 0; p.is.value = 0;
        break; }
    case ORC_CUSTOM_SHORT: {  /* custom_short.rs:155-229, 237-270, 522-542 */
        p.is_long = p.blob.is_long; p.mutates = false;
        p.min_size = sat_u16(2.0 * p.blob.rl_mean + p.blob.is_mean);
        if (!phred_table_from_blob(p.blob, p.phred)) { fail("custom phred table"); return false; }
        if (!len_table_from_bins(p.blob.rl_bins, p.rl)) { fail("custom rl table"); return false; }
        if (p.blob.has_is_bins) { if (!len_table_from_bins(p.blob.is_bins, p.is)) { fail("custom is table"); return false; } }
        else { p.is.kind = 0; p.is.value = 0; }
        break; }
    case ORC_CUSTOM_LONG: {   /* phreds as custom-short, gamma lengths */
        p.is_long = true; p.mutates = false; p.min_size = 20000;
        if (!phred_table_from_blob(p.blob, p.phred)) { fail("custom phred table"); return false; }
        if (!build_err_table(p.blob, p.err)) { fail("bad k-mer table in the custom profile"); return false; }
        p.rl.kind = 2; p.is.kind = 0; p.is.value = 0;
        break; }
    default: fail("unknown profile kind"); return false;
    }
    if (p.has_blob) {
        p.q_alias.resize(p.blob.quality.size());
        for (size_t i = 0; i < p.q_alias.size(); i++)
            if (!p.q_alias[i].build(p.blob.quality[i].density)) { fail("empty quality bins"); return false; }
    }
    return true;
}

/* ------------------------------------------------------------------ */
/* sampling from the native tables                                      */
/* ------------------------------------------------------------------ */
inline uint32_t sample_len(const LenTable &t, uint32_t w_alias, uint32_t w_inbin) {
    if (t.kind == 0) return t.value;
    uint32_t idx = t.cb ? (w_alias >> (32 - t.cb)) : 0;
    uint32_t thr_draw = (uint32_t)(((uint64_t)w_alias << t.cb) & 0xffffffffu) >> 12;
    uint32_t e = t.entries[idx];
    uint32_t col = (thr_draw < (e >> 12)) ? idx : (e & 0xfffu);
    uint32_t lo = t.lo[col], hi = t.hi[col];
    uint64_t span = (uint64_t)hi - lo + 1;
    return lo + (uint32_t)(((uint64_t)w_inbin * span) >> 32);
}
/* kind 3 entry against a draw: column = low cb bits of the draw's word, self iff its level < entry */
inline uint32_t sample_q32(const uint32_t *row, uint32_t cb, Draw d) {
    const uint32_t e = row[d.word & ((1u << cb) - 1u)];
    return (d.level < e ? e : e >> 8) & 0xffu;
}
/* minimal_short.rs:124-131 substitution table */
inline uint8_t substitute(uint8_t nt, uint32_t k) {
    switch (nt) {
    case 'A': return "CGT"[k];
    case 'C': return "AGT"[k];
    case 'T': return "ACG"[k];
    case 'G': return "ACT"[k];
    default: return nt;
    }
}
/* native mode: the table's mutation code m in 1..3 picks the substituted base as
 * rank(nt) ^ m with A,C,G,T = 0..3 — choose()'s index (minimal_short.rs:125) relabelled, still
 * uniform over the three other bases; replay applies substitute() to the recorded index */
inline uint8_t substitute_native(uint8_t nt, uint32_t m) {
    uint32_t rank;
    switch (nt) { case 'A': rank = 0; break; case 'C': rank = 1; break; case 'G': rank = 2; break; case 'T': rank = 3; break; default: return nt; }
    return (uint8_t)"ACGT"[rank ^ (m & 3u)];
}
inline double u01_open(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }

/* Marsaglia-Tsang gamma (rand_distr gamma.rs GammaLargeShape), shape and scale
 * from perfect_long.rs:48-54 in f32, evaluated in f64 with det_log; one Philox
 * block per attempt.  Returns false if the attempt rejects. */
bool gamma_attempt(const uint32_t w[4], uint16_t &rl) {
    const float meanf = 20000.0f, sdf = 15000.0f;
    const float shapef = powf(meanf / sdf, 2.0f), scalef = powf(sdf, 2.0f) / meanf;
    const double d = (double)shapef - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
    double v1 = 2.0 * u01_open(w[0]) - 1.0, v2 = 2.0 * u01_open(w[1]) - 1.0;
    double s = v1 * v1 + v2 * v2;
    if (s >= 1.0 || s == 0.0) return false;
    double x = v1 * std::sqrt(-2.0 * det_log(s) / s);
    double t = 1.0 + c * x;
    if (t <= 0.0) return false;
    double v = t * t * t, uu = u01_open(w[2]), x2 = x * x;
    if (uu < 1.0 - 0.0331 * (x2 * x2) || det_log(uu) < 0.5 * x2 + d * (1.0 - v + det_log(v))) {
        rl = sat_u16(std::floor(d * v * (double)scalef));
        return true;
    }
    return false;
}

/* ------------------------------------------------------------------ */
/* genomes                                                              */
/* ------------------------------------------------------------------ */
struct GenomeView {
    const orc_genomes *g;
    uint32_t nseq(uint32_t gi) const { return g->genome_contig[gi + 1] - g->genome_contig[gi]; }
    uint32_t contig(uint32_t gi, uint32_t k) const { return g->genome_contig[gi] + k; }
    uint64_t size(uint32_t c) const { return g->contig_size[c]; }
    const uint8_t *seq(uint32_t c) const { return g->bases + g->contig_off[c]; }
    uint64_t phys(uint32_t c) const { return g->contig_off[c + 1] - g->contig_off[c]; }
    std::string cid(uint32_t c) const { return bytes_to_string((const uint8_t *)g->contig_ids + g->contig_id_off[c], g->contig_id_off[c + 1] - g->contig_id_off[c]); }
    std::string gid(uint32_t gi) const { return std::string(g->genome_ids + g->genome_id_off[gi], g->genome_id_off[gi + 1] - g->genome_id_off[gi]); }
};

/* ------------------------------------------------------------------ */
/* FASTQ record — simmr/src/fastq.rs:34-69                              */
/* ------------------------------------------------------------------ */
const char *DEFAULT_HEADER =
    "@{:read_id:}|{:genome_id:}/{:pair:} metadata:sid={:sequence_id:}|sp={:start_position:}|ep={:end_position:}|rc={:reverse_complement:}";
void replace_all(std::string &s, const std::string &from, const std::string &to) {
    /* Rust str::replace: non-overlapping matches, left to right */
    std::string r; r.reserve(s.size() + to.size());
    size_t i = 0;
    for (;;) {
        size_t j = s.find(from, i);
        if (j == std::string::npos) { r.append(s, i, std::string::npos); break; }
        r.append(s, i, j - i); r.append(to); i = j + from.size();
    }
    s.swap(r);
}
void emit_record(std::vector<uint8_t> &out, const std::string &fmt, const std::string &gid, uint32_t read_id,
                 const std::string &sid, uint64_t sp, uint64_t ep, bool rc, int pair,
                 const std::vector<uint8_t> &seq, const std::vector<uint8_t> &qual) {
    std::string h = fmt;
    replace_all(h, "{:genome_id:}", gid);
    replace_all(h, "{:read_id:}", std::to_string(read_id));
    replace_all(h, "{:sequence_id:}", sid);
    replace_all(h, "{:start_position:}", std::to_string(sp));
    replace_all(h, "{:end_position:}", std::to_string(ep));
    replace_all(h, "{:reverse_complement:}", rc ? "t" : "f");
    replace_all(h, "{:pair:}", pair == 1 ? "1" : "2");
    out.insert(out.end(), h.begin(), h.end()); out.push_back('\n');
    out.insert(out.end(), seq.begin(), seq.end()); out.push_back('\n');
    out.push_back('+'); out.push_back('\n');
    for (uint8_t q : qual) out.push_back(encode_quality_score(q));   /* util.rs:55-57 */
    out.push_back('\n');
}

/* ------------------------------------------------------------------ */
/* PE geometry — simulate.rs:211-262                                    */
/* ------------------------------------------------------------------ */
struct PeGeom { uint64_t fs, fe, lo, hi; bool clamped; };
template <typename Redraw>
PeGeom pe_geometry(uint64_t n, uint64_t phys, uint64_t RL, uint64_t IS, uint64_t R, uint64_t fwd_start, Redraw redraw) {
    PeGeom g{}; g.fs = fwd_start; g.fe = fwd_start + RL;
    if (fwd_start + IS >= n || fwd_start + IS + RL >= n) {            /* :241-247 */
        g.lo = redraw(fwd_start, n - R); g.hi = g.lo + RL;
    } else if ((int32_t)((uint32_t)(fwd_start + IS) - (uint32_t)RL) < 0) {   /* :250-251, `as i32`, release-mode wrap */
        g.lo = 0; g.hi = RL;
    } else {                                                          /* :253-256 */
        g.lo = fwd_start + IS - RL; g.hi = fwd_start + IS;
    }
    /* The reference would panic on an out-of-range slice (RL drawn larger than
     * the nominal minimum size); the rebuild clamps and counts. */
    g.clamped = false;
    if (g.fe > phys) { g.fe = phys; g.clamped = true; }
    if (g.fs > g.fe) { g.fs = g.fe; g.clamped = true; }
    if (g.hi > phys) { g.hi = phys; g.clamped = true; }
    if (g.lo > g.hi) { g.lo = g.hi; g.clamped = true; }
    return g;
}

struct UnitMap {    /* global unit index -> genome (simulate.rs:135, :179) */
    std::vector<uint64_t> start;   /* [n_genomes+1] */
    bool is_long;
    UnitMap(const uint64_t *reads, uint32_t n, bool lng) : start(n + 1, 0), is_long(lng) {
        for (uint32_t i = 0; i < n; i++) start[i + 1] = start[i] + (lng ? reads[i] : reads[i] / 2);
    }
    uint64_t total() const { return start.back(); }
    uint32_t genome_of(uint64_t u) const { return (uint32_t)(std::upper_bound(start.begin(), start.end(), u) - start.begin() - 1); }
};

/* ------------------------------------------------------------------ */
/* native driver                                                        */
/* ------------------------------------------------------------------ */
struct NativeGeom { uint32_t genome, contig; uint32_t RL, IS; uint64_t start, lo_or_end; PeGeom pe; uint64_t ls, le; };

bool native_geometry(const GenomeView &gv, const Profile &p, const UnitMap &um, const Philox &ph, uint64_t u, NativeGeom &ng) {
    uint32_t gi = um.genome_of(u);
    ng.genome = gi;
    uint32_t w0[4], w1[4];
    ph.at(u, 0, 0, w0);
    if (!p.is_long) {
        ph.at(u, 0, 1, w1);
        uint32_t k = (uint32_t)mulhi64(u64_of(w0[0], w0[1]), gv.nseq(gi));               /* simulate.rs:182 */
        uint32_t c = gv.contig(gi, k);
        ng.contig = c;
        ng.RL = (uint16_t)sample_len(p.rl, w1[0], w1[1]);                                /* :211, `as u16` */
        ng.IS = (uint16_t)sample_len(p.is, w1[2], w1[3]);                                /* :212 */
        uint64_t n = gv.size(c), R = p.min_size;
        if (n <= R) { fail("contig smaller than minimum_genome_size (filter first, main.rs:124)"); return false; }
        uint64_t fs = mulhi64(u64_of(w0[2], w0[3]), n - R);                              /* :233 */
        ng.start = fs;
        ng.pe = pe_geometry(n, gv.phys(c), ng.RL, ng.IS, R, fs, [&](uint64_t a, uint64_t b) {
            uint32_t w2[4]; ph.at(u, 0, 2, w2);
            return a + mulhi64(u64_of(w2[0], w2[1]), b - a);                             /* :245 */
        });
        ng.lo_or_end = ng.pe.lo;
        return true;
    }
    /* long: simulate.rs:356-375, 484-491 */
    uint32_t nseq = gv.nseq(gi), c0 = gv.contig(gi, 0);
    uint16_t rl = 20000; uint32_t ncand = 0;
    bool got = false;
    for (uint32_t a = 0; a < 256 && !got; a++) {
        uint32_t w[4]; ph.at(u, 0, 16 + a, w);
        uint16_t r;
        if (!gamma_attempt(w, r)) continue;
        uint32_t cnt = 0;
        for (uint32_t k = 0; k < nseq; k++) if (gv.size(c0 + k) > r) cnt++;               /* :362-367 */
        if (cnt == 0) continue;                                                          /* :370-372 */
        rl = r; ncand = cnt; got = true;
    }
    if (!got) { ncand = 0; for (uint32_t k = 0; k < nseq; k++) if (gv.size(c0 + k) > rl) ncand++; }
    if (ncand == 0) { fail("no contig longer than the read length"); return false; }
    uint32_t pick = (uint32_t)mulhi64(u64_of(w0[0], w0[1]), ncand), c = c0;              /* :375 */
    for (uint32_t k = 0, seen = 0; k < nseq; k++) if (gv.size(c0 + k) > rl) { if (seen == pick) { c = c0 + k; break; } seen++; }
    ng.contig = c; ng.RL = rl; ng.IS = 0;
    uint64_t n = gv.size(c), start, end;
    if (p.flags & ORC_FLAG_COMPAT_LONG_START) {
        start = rl ? mulhi64(u64_of(w0[2], w0[3]), rl) : 0;                              /* :484 */
        end = start + rl;
        if (end >= n) { ph.at(u, 0, 1, w1); end = start + mulhi64(u64_of(w1[0], w1[1]), n - start); }   /* :488-489 */
    } else {
        start = mulhi64(u64_of(w0[2], w0[3]), n - rl);                                   /* SURVEY §A.4 */
        end = start + rl;
    }
    uint64_t phys = gv.phys(c);
    if (end > phys) end = phys;
    if (start > end) start = end;
    ng.start = start; ng.lo_or_end = end; ng.ls = start; ng.le = end;
    return true;
}

/* pieces of a read and their length: short reads are one piece, long reads 32 pieces of a multiple of 16 bases */
inline uint64_t piece_len(uint64_t L, bool long_read) { return long_read ? 16 * ((L + 511) / 512) : L; }

/* one mate / long read in native mode: per piece, qualities from the piece's quality chain, then the
 * substituted bases by geometric gaps along the piece (forward-strand order, simulate.rs:265-270 / 497-503) */
void native_read(const Profile &p, const Philox &ph, uint64_t u, uint32_t stream, const uint8_t *src, uint64_t L,
                 std::vector<uint8_t> &seq, std::vector<uint8_t> &qual, orc_stats &st) {
    seq.assign(src, src + L); qual.resize(L);
    const PhredTable &t = p.phred;
    if (t.kind == 0) { std::fill(qual.begin(), qual.end(), t.const_q); return; }
    const uint64_t S = piece_len(L, p.is_long);
    for (uint64_t lo = 0, piece = 0; lo < L; lo += S, piece++) {
        const uint64_t hi = std::min(L, lo + S);
        uint32_t w[4]; ph.at(u, stream, (uint32_t)piece, w);
        Chain qc(w[0], w[1]), ec(w[2], w[3]);
        Draw d{};
        for (uint64_t i = lo; i < hi; i++) {               /* one step of the quality chain per two bases */
            const uint64_t pos = std::min<uint64_t>(i, t.npos - 1);             /* custom_short.rs:339-345 */
            if (!((i - lo) & 1)) d = qc.next(); else d = Draw{d.word >> t.cb, d.level2, 0};
            qual[i] = (uint8_t)(sample_q32(t.entries.data() + (pos << t.cb), t.cb, d) - 33u);   /* the table holds characters */
        }
        if (!t.has_sub) continue;
        for (uint64_t next = lo; next < hi;) {             /* every step takes three draws of the event chain */
            const Draw r1 = ec.next(), r2 = ec.next(), r3 = ec.next();
            const uint32_t g = sample_q32(t.gap.data(), 8, r1);
            if (g == 255) { next += 255; continue; }
            const uint64_t at = next + g;
            if (at >= hi) break;
            qual[at] = (uint8_t)(sample_q32(t.qsub.data(), t.cb, r2) - 33u);
            const uint8_t nb = substitute_native(seq[at], 1u + mulhi32(r3.word, 3u));
            if (nb != seq[at]) st.mutated++;
            seq[at] = nb;
            next = at + 1;
        }
    }
}

/* simulate_errors in native mode (custom_long.rs:187-232).  The reference walks i over the ORIGINAL length, looks at the
 * k bytes at i of the sequence as edited so far, and splices the sampled alternate in (an alternate's N digits are deleted
 * bases).  Everything before i is final and everything from i + k on is still the original's tail, so the state is the k
 * bytes under the window (`win`) and how far into the original they reach (`r`).  One draw of the read's error chain
 * (Philox block 0 of stream 3) per window, used or not.  Where the reference would index past the shortened sequence
 * (and panic) the walk stops.  `choices`, when given, receives what orc_simulate_errors' `choice` array must hold to
 * reproduce the run (the alternate's position in the blob's list, one per window that had alternates). */
void native_errors(const Profile &p, const Philox &ph, uint64_t u, const std::vector<uint8_t> &orig, std::vector<uint8_t> &out,
                   std::vector<uint32_t> *choices = nullptr) {
    const ErrTable &t = p.err;
    const size_t L = orig.size(), k = t.k;
    out.clear();
    uint32_t w[4]; ph.at(u, 3, 0, w);
    Chain ch(w[0], w[1]);
    ch.next();                                                                 /* a fresh chain's carry has 30 bits: one step fills it */
    std::vector<uint8_t> win; size_t r = 0;
    for (size_t i = 0; i + k <= L; i++) {                                      /* :196-199 */
        while (win.size() < k && r < L) win.push_back(orig[r++]);
        if (win.size() < k) break;                                             /* the reference would panic here */
        const Draw d = ch.next();
        uint32_t code;
        if (kmer_encode(win.data(), (uint32_t)k, 3, &code) == 0) {             /* :203-207 */
            const int row = t.find(code);
            if (row >= 0) {                                                    /* :210-213 */
                uint32_t alt = code, choice = t.id_choice[row];
                if (d.level >= t.keep_thr[row]) {
                    const uint32_t cb = t.cb[row], wd = d.word;
                    const uint32_t idx = cb ? wd >> (32 - cb) : 0u, frac = (uint32_t)(((uint64_t)wd << cb) & 0xffffffffu) >> 12;
                    const uint32_t e = t.entries[t.row[row] + idx], j = frac < (e >> 12) ? idx : (e & 0xfffu);
                    alt = t.alts[t.row[row] + j];
                    if (j < t.other_choice[row].size()) choice = t.other_choice[row][j];
                }
                if (choices) choices->push_back(choice);
                if (alt != code) {                                             /* :221-228 decode with the N digits skipped, splice */
                    uint8_t dec[16]; const int n = kmer_decode(alt, (uint32_t)k, 3, true, dec);
                    win.assign(dec, dec + n);
                }
            }
        }
        if (win.empty() && r < L) win.push_back(orig[r++]);
        if (win.empty()) break;
        out.push_back(win.front()); win.erase(win.begin());
    }
    out.insert(out.end(), win.begin(), win.end());
    out.insert(out.end(), orig.begin() + r, orig.end());
}

int64_t finish(std::vector<uint8_t> &buf, uint8_t *out, uint64_t cap) {
    if (out) { if (buf.size() > cap) return -(int64_t)buf.size(); memcpy(out, buf.data(), buf.size()); }
    return (int64_t)buf.size();
}

/* ------------------------------------------------------------------ */
/* FASTA — genome.rs:89-162 with needletail 0.4.1 semantics */
/* ------------------------------------------------------------------ */
struct Seq { std::string id; std::vector<uint8_t> seq; uint64_t size; };
struct Genome { std::vector<Seq> seqs; uint64_t size = 0; };
/* needletail sequence.rs normalize(seq, false) */
inline int normalize_byte(uint8_t c) {
    switch (c) {
    case 'A': case 'C': case 'G': case 'T': case 'N': case '-': return c;
    case 'a': return 'A'; case 'c': return 'C'; case 'g': return 'G'; case 't': return 'T'; case 'n': return 'N';
    case 'u': case 'U': return 'T';
    case '.': case '~': return '-';
    case ' ': case '\t': case '\r': case '\n': return -1;
    default: return 'N';
    }
}
bool load_fasta(const std::string &path, bool contiguous, Genome &g, std::string &err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "cannot open " + path; return false; }
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<Seq> seqs;
    size_t i = 0, n = data.size();
    while (i < n && (data[i] == '\n' || data[i] == '\r')) i++;
    if (i < n && data[i] != '>') { err = "not a FASTA file (expected '>')"; return false; }
    while (i < n) {
        size_t e = data.find('\n', i);
        if (e == std::string::npos) e = n;
        size_t he = e; if (he > i && data[he - 1] == '\r') he--;
        Seq s; s.id = data.substr(i + 1, he - i - 1);                 /* record.id(): whole header line */
        i = std::min(n, e + 1);
        while (i < n && !(data[i] == '>' && (i == 0 || data[i - 1] == '\n'))) {
            int c = normalize_byte((uint8_t)data[i]);
            if (c >= 0) s.seq.push_back((uint8_t)c);
            i++;
        }
        s.size = s.seq.size();
        seqs.push_back(std::move(s));
    }
    uint64_t total = 0; for (auto &s : seqs) total += s.seq.size();
    g.size = total;
    if (contiguous) {                                                 /* genome.rs:117-136 */
        Seq w; w.id = "whole genome";
        for (auto &s : seqs) { w.seq.insert(w.seq.end(), s.seq.begin(), s.seq.end()); w.seq.push_back('N'); }
        w.size = total;                                               /* size excludes the joining Ns */
        g.seqs.clear(); g.seqs.push_back(std::move(w));
    } else g.seqs = std::move(seqs);
    return true;
}

/* ------------------------------------------------------------------ */
/* independent generator for the reference-arithmetic samplers          */
/* ------------------------------------------------------------------ */
struct Xo {   /* xoshiro256** seeded by splitmix64 */
    uint64_t s[4];
    explicit Xo(uint64_t seed) { for (auto &x : s) { seed += 0x9e3779b97f4a7c15ull; uint64_t z = seed; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; x = z ^ (z >> 31); } }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() { uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return r; }
    double u01() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
    float f32_24() { return (float)(uint32_t)(next() >> 40) * (1.0f / 16777216.0f); }     /* rand gen::<f32>() */
    double normal() { double a = u01(), b = u01(); return std::sqrt(-2.0 * std::log(a)) * std::cos(2.0 * M_PI * b); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

#include "ref_stream.inc"

}  // namespace

/* ==================================================================== */
/* C ABI                                                                 */
/* ==================================================================== */
extern "C" {

const char *orc_last_error(void) { return g_err.c_str(); }

uint8_t orc_complement(uint8_t n) { return complement(n); }
void orc_reverse_complement(const uint8_t *in, uint64_t n, uint8_t *out) { auto r = reverse_complement(in, n); if (n) memcpy(out, r.data(), n); }
void orc_encode_quality_scores(const uint8_t *in, uint64_t n, uint8_t *out) { for (uint64_t i = 0; i < n; i++) out[i] = encode_quality_score(in[i]); }
float orc_phred_to_probability(uint8_t q) { return phred_to_probability(q); }
float orc_phred_to_accuracy(uint8_t q) { return phred_to_accuracy(q); }
uint8_t orc_probability_to_phred(float p) { return probability_to_phred(p); }
uint8_t orc_accuracy_to_phred(float a) { return accuracy_to_phred(a); }
int orc_bytes_to_string(const uint8_t *in, uint64_t n, char *out) { std::string s = bytes_to_string(in, n); memcpy(out, s.data(), s.size()); return (int)s.size(); }
double orc_gaussian(double x, const double *xs, uint64_t n, double bw) { return gaussian(x, xs, n, bw); }
int orc_kmer_encode(const uint8_t *kmer, uint32_t k, uint32_t esize, uint32_t *out) { return kmer_encode(kmer, k, esize, out); }
int orc_kmer_decode(uint32_t code, uint32_t k, uint32_t esize, int skip_n, uint8_t *out) { return kmer_decode(code, k, esize, skip_n != 0, out); }
void orc_format_f64(double v, char *out, int cap) { std::string s = format_f64(v); snprintf(out, cap, "%s", s.c_str()); }
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    Philox p(u64_of(key[0], key[1])); p.block(ctr[0], ctr[1], ctr[2], ctr[3], out);
}

/* custom_long.rs:187-232 (== custom_short.rs:452-516).  `choice[j]` is the alt
 * index the alias sampler returned for the j-th window that had alternatives. */
int64_t orc_simulate_errors(const uint8_t *seq, uint64_t n, uint32_t k, const uint32_t *ref_kmers,
                            const uint32_t *alt_off, const uint32_t *alt_kmers, uint32_t n_ref,
                            const uint32_t *choice, uint64_t n_choice, uint8_t *out, uint64_t out_cap) {
    std::vector<uint8_t> ns(seq, seq + n);
    std::map<uint32_t, uint32_t> idx;
    for (uint32_t i = 0; i < n_ref; i++) idx[ref_kmers[i]] = i;
    uint64_t used = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (i + k > n) break;                                        /* :197-199 */
        if (i + k > ns.size()) break;                               /* the reference would panic (unguarded slice): the walk stops, as in native mode */
        uint32_t enc;
        if (kmer_encode(ns.data() + i, k, 3, &enc) != 0) continue;   /* :203-207 */
        auto it = idx.find(enc);
        if (it == idx.end()) continue;                               /* :210-213 */
        uint32_t a0 = alt_off[it->second], a1 = alt_off[it->second + 1];
        if (used >= n_choice) return fail("not enough choices");
        uint32_t ch = choice[used++];
        if (a0 + ch >= a1) return fail("choice out of range");
        uint8_t alt[16];
        int an = kmer_decode(alt_kmers[a0 + ch], k, 3, true, alt);   /* :221-226 */
        if (an < 0) return fail("bad alt k-mer");
        ns.erase(ns.begin() + i, ns.begin() + i + k);                /* splice :228 */
        ns.insert(ns.begin() + i, alt, alt + an);
    }
    if (ns.size() > out_cap) return -(int64_t)ns.size();
    memcpy(out, ns.data(), ns.size());
    return (int64_t)ns.size();
}

/* abundance_profiles/{uniform.rs:18-35,46-95; exact.rs:17-35; custom.rs:20-96} */
int orc_abundance(int kind, int size_adjusted, uint64_t total_reads, uint32_t n, const double *abund_in,
                  const uint64_t *genome_sizes, uint64_t *out_reads, double *out_abund) {
    std::vector<std::pair<uint64_t, double>> ra(n);
    if (kind == ORC_AB_UNIFORM) {
        uint64_t per = sat_usize(std::ceil((double)total_reads / (double)n));   /* uniform.rs:28 */
        double ab = 100.0 / (double)n;                                          /* :29 */
        for (auto &x : ra) x = {per, ab};
    } else if (kind == ORC_AB_EXACT) {
        double ab = 100.0 / (double)n;                                          /* exact.rs:18 */
        for (auto &x : ra) x = {total_reads, ab};
        size_adjusted = 0;                                                      /* exact.rs:13-15 */
    } else if (kind == ORC_AB_CUSTOM) {
        if (!abund_in) return fail("custom abundance profile needs abundances");
        double tot = 0.0;
        for (uint32_t i = 0; i < n; i++) tot += abund_in[i];                    /* custom.rs:26 */
        if (tot < 0.99 || tot > 1.01) {
            for (uint32_t i = 0; i < n; i++)
                ra[i] = {sat_usize(std::ceil((double)total_reads * (abund_in[i] / tot))), abund_in[i] / tot};   /* :30-38 */
        } else {
            for (uint32_t i = 0; i < n; i++)
                ra[i] = {sat_usize(std::ceil((double)total_reads * abund_in[i])), abund_in[i]};                 /* :40-43 */
        }
    } else return fail("unknown abundance kind");
    if (size_adjusted) {                                                        /* uniform.rs:60-94 == custom.rs:61-95 */
        if (!genome_sizes) return fail("size adjustment needs genome sizes");
        double total = 0.0, adjusts = 0.0;
        for (uint32_t i = 0; i < n; i++) total += (double)ra[i].first;
        for (uint32_t i = 0; i < n; i++) adjusts += (double)genome_sizes[i] * ra[i].second;
        for (uint32_t i = 0; i < n; i++)
            ra[i].first = sat_usize(std::ceil(total * ((ra[i].second * (double)genome_sizes[i]) / adjusts)));
    }
    for (uint32_t i = 0; i < n; i++) { out_reads[i] = ra[i].first; out_abund[i] = ra[i].second; }
    return 0;
}

int orc_profile_min_genome_size(const orc_profile *pi, uint16_t *out) { Profile p; if (!make_profile(pi, p)) return -1; *out = p.min_size; return 0; }
int orc_profile_is_long(const orc_profile *pi, int *out) { Profile p; if (!make_profile(pi, p)) return -1; *out = p.is_long; return 0; }
int orc_profile_phred_table(const orc_profile *pi, uint32_t *kind, uint32_t *cb, uint32_t *npos, uint32_t *const_q, uint32_t *entries, uint64_t cap) {
    Profile p; if (!make_profile(pi, p)) return -1;
    *kind = p.phred.kind; *cb = p.phred.cb; *npos = p.phred.npos; *const_q = p.phred.const_q;
    std::vector<uint32_t> all = p.phred.entries;      /* kind 3 of a mutating profile: followed by gap[256] and qsub[1 << cb] */
    if (p.phred.has_sub) { all.insert(all.end(), p.phred.gap.begin(), p.phred.gap.end()); all.insert(all.end(), p.phred.qsub.begin(), p.phred.qsub.end()); }
    if (all.size() > cap) return fail("table buffer too small");
    if (!all.empty()) memcpy(entries, all.data(), all.size() * 4);
    return (int)all.size();
}
int orc_profile_err_table(const orc_profile *pi, uint32_t *k, uint32_t *n_ref, uint32_t *words, uint64_t cap) {
    Profile p; if (!make_profile(pi, p)) return -1;
    const ErrTable &e = p.err;
    *k = e.on ? e.k : 0; *n_ref = (uint32_t)e.ref.size();
    std::vector<uint32_t> all;
    for (const std::vector<uint32_t> *v : {&e.ref, &e.keep_thr, &e.row, &e.cb, &e.entries, &e.alts}) all.insert(all.end(), v->begin(), v->end());
    if (all.size() > cap) return fail("buffer too small");
    if (!all.empty()) memcpy(words, all.data(), all.size() * 4);
    return (int)all.size();
}
int orc_profile_len_table(const orc_profile *pi, int which, uint32_t *kind, uint32_t *cb, uint32_t *entries, uint32_t *lo, uint32_t *hi, uint64_t cap) {
    Profile p; if (!make_profile(pi, p)) return -1;
    const LenTable &t = which ? p.is : p.rl;
    *kind = t.kind; *cb = t.cb;
    if (t.kind == 0) { lo[0] = hi[0] = t.value; return 0; }
    if (t.kind == 2) return 0;
    if (t.entries.size() > cap) return fail("table buffer too small");
    memcpy(entries, t.entries.data(), t.entries.size() * 4);
    memcpy(lo, t.lo.data(), t.lo.size() * 4); memcpy(hi, t.hi.data(), t.hi.size() * 4);
    return (int)t.entries.size();
}

/* ---- FASTA / files ---- */
struct orc_fasta { Genome g; };
int orc_fasta_load(const char *path, int contiguous, orc_fasta **out, char *err, int errcap) {
    auto *f = new orc_fasta; std::string e;
    if (!load_fasta(path, contiguous != 0, f->g, e)) { if (err) snprintf(err, errcap, "%s", e.c_str()); delete f; return -1; }
    *out = f; return 0;
}
void orc_fasta_free(orc_fasta *f) { delete f; }
uint32_t orc_fasta_num_seqs(const orc_fasta *f) { return (uint32_t)f->g.seqs.size(); }
uint64_t orc_fasta_size(const orc_fasta *f) { return f->g.size; }
uint64_t orc_fasta_seq_size(const orc_fasta *f, uint32_t i) { return f->g.seqs[i].size; }
uint64_t orc_fasta_seq_len(const orc_fasta *f, uint32_t i) { return f->g.seqs[i].seq.size(); }
const uint8_t *orc_fasta_seq(const orc_fasta *f, uint32_t i) { return f->g.seqs[i].seq.data(); }
const char *orc_fasta_seq_id(const orc_fasta *f, uint32_t i, uint32_t *len) { *len = (uint32_t)f->g.seqs[i].id.size(); return f->g.seqs[i].id.data(); }

/* files.rs:19-86.  The shipped is_simple_variant (files.rs:32-45) misdetects
 * every list as "complex"; the documented behaviour (readme.md:99-100) is
 * restated: >=2 tab-separated fields on line 1 => TSV with a header row. */
int64_t orc_parse_genome_file(const char *path, char *out, uint64_t cap) {
    std::ifstream f(path);
    if (!f) return fail("Genome file does not exist");
    std::vector<std::string> lines; std::string ln;
    while (std::getline(f, ln)) { if (!ln.empty() && ln.back() == '\r') ln.pop_back(); lines.push_back(ln); }
    auto split = [](const std::string &s) { std::vector<std::string> v; size_t i = 0; for (;;) { size_t j = s.find('\t', i); if (j == std::string::npos) { v.push_back(s.substr(i)); break; } v.push_back(s.substr(i, j - i)); i = j + 1; } return v; };
    std::string res; int64_t nrec = 0;
    if (lines.empty()) return 0;
    auto first = split(lines[0]);
    if (first.size() < 2) {   /* simple list: positional, filepath only */
        for (auto &l : lines) { if (l.empty()) continue; auto c = split(l); res += c[0] + "\t\tnan\n"; nrec++; }
    } else {
        int cp = -1, cu = -1, ca = -1;
        for (size_t i = 0; i < first.size(); i++) {
            const std::string &h = first[i];
            if (h == "filepath" || h == "path") cp = (int)i;                        /* files.rs:21 */
            else if (h == "uuid" || h == "id" || h == "genome_id") cu = (int)i;     /* :23 */
            else if (h == "abundance") ca = (int)i;
        }
        if (cp < 0) return fail("genome file header has no path/filepath column");
        for (size_t r = 1; r < lines.size(); r++) {
            if (lines[r].empty()) continue;
            auto c = split(lines[r]);
            std::string pth = cp < (int)c.size() ? c[cp] : "", uu = (cu >= 0 && cu < (int)c.size()) ? c[cu] : "";
            std::string ab = (ca >= 0 && ca < (int)c.size() && !c[ca].empty()) ? c[ca] : "nan";
            res += pth + "\t" + uu + "\t" + ab + "\n"; nrec++;
        }
    }
    if (res.size() + 1 > cap) return fail("buffer too small");
    memcpy(out, res.c_str(), res.size() + 1);
    return nrec;
}
/* files.rs:100-134 */
int64_t orc_metadata_tsv(uint32_t n, const char *const *uuid, const char *const *path, const uint64_t *num_reads,
                         const double *abundance, char *out, uint64_t cap) {
    std::string s = "genome_id\tfilepath\tnum_reads\tabundance\n";
    for (uint32_t i = 0; i < n; i++) s += std::string(uuid[i]) + "\t" + path[i] + "\t" + std::to_string(num_reads[i]) + "\t" + format_f64(abundance[i]) + "\n";
    if (s.size() + 1 > cap) return fail("buffer too small");
    memcpy(out, s.c_str(), s.size() + 1);
    return (int64_t)s.size();
}

int orc_blob_summary(const uint8_t *blob, uint64_t len, uint64_t *bin_size, uint64_t *n_pos, uint64_t *kmer_size,
                     uint64_t *n_kmers, double *is_mean, double *rl_mean, int *has_is_bins, int *is_long) {
    Blob b; if (!parse_blob(blob, len, b)) return fail("cannot parse blob");
    *bin_size = b.bin_size; *n_pos = b.quality.size(); *kmer_size = b.kmer_size; *n_kmers = b.probs.size();
    *is_mean = b.is_mean; *rl_mean = b.rl_mean; *has_is_bins = b.has_is_bins; *is_long = b.is_long;
    return 0;
}

/* ---- drivers ---- */
uint64_t orc_total_units(const orc_profile *pi, const uint64_t *reads, uint32_t n) {
    Profile p; if (!make_profile(pi, p)) return 0;
    return UnitMap(reads, n, p.is_long).total();
}

int orc_native_geometry(const orc_genomes *g, const orc_profile *pi, const uint64_t *reads, const orc_opts *o,
                        uint32_t *genome, uint32_t *contig, uint32_t *rl, uint32_t *is_, uint64_t *start, uint64_t *lo_or_end) {
    Profile p; if (!make_profile(pi, p)) return -1;
    GenomeView gv{g}; UnitMap um(reads, g->n_genomes, p.is_long); Philox ph(o->seed);
    uint64_t b = o->unit_begin, e = o->unit_end; if (b == 0 && e == 0) e = um.total();
    for (uint64_t u = b; u < e; u++) {
        NativeGeom ng{}; if (!native_geometry(gv, p, um, ph, u, ng)) return -1;
        size_t k = u - b;
        genome[k] = ng.genome; contig[k] = ng.contig; rl[k] = ng.RL; is_[k] = ng.IS; start[k] = ng.start; lo_or_end[k] = ng.lo_or_end;
    }
    return 0;
}

/* units [b, e) of a native-mode run appended to `buf` (the body of orc_simulate; also timed by orc_bench_pe) */
static bool simulate_range(const GenomeView &gv, const Profile &p, const UnitMap &um, const Philox &ph, const std::string &fmt,
                           uint32_t id_base, uint64_t b, uint64_t e, std::vector<uint8_t> &buf, orc_stats &st) {
    std::vector<uint8_t> s1, q1, s2, q2;
    for (uint64_t u = b; u < e; u++) {
        NativeGeom ng{}; if (!native_geometry(gv, p, um, ph, u, ng)) return false;
        uint32_t id = id_base + (uint32_t)u;                                          /* simulate.rs:85-89, one id per unit */
        std::string gid = gv.gid(ng.genome), sid = gv.cid(ng.contig);
        const uint8_t *cs = gv.seq(ng.contig);
        if (!p.is_long) {
            const PeGeom &pg = ng.pe;
            if (pg.clamped) st.clamped++;
            native_read(p, ph, u, 1, cs + pg.fs, pg.fe - pg.fs, s1, q1, st);     /* :261,265,269 */
            native_read(p, ph, u, 2, cs + pg.lo, pg.hi - pg.lo, s2, q2, st);     /* :262,266,270 */
            auto rc = reverse_complement(s2.data(), s2.size());                       /* :283; quality NOT reversed :285 */
            emit_record(buf, fmt, gid, id, sid, pg.fs, ng.start + ng.RL, false, 1, s1, q1);   /* :287-292: end_pos = fwd_end */
            emit_record(buf, fmt, gid, id, sid, pg.lo + (uint64_t)ng.RL, pg.lo, true, 2, rc, q2);   /* :293-298: start=rev_start, end=rev_end */
            st.records += 2; st.bases += s1.size() + s2.size();
        } else {
            native_read(p, ph, u, 1, cs + ng.ls, ng.le - ng.ls, s1, q1, st);           /* :494-503 */
            if (p.err.on) { native_errors(p, ph, u, s1, s2); s1.swap(s2); }            /* :500, before the (identity) point mutations :503 */
            emit_record(buf, fmt, gid, id, sid, ng.ls, ng.le, false, 1, s1, q1);       /* fastq.rs:57: {:pair:} -> 1 */
            st.records += 1; st.bases += s1.size();
        }
        st.units++;
    }
    return true;
}

int64_t orc_simulate(const orc_genomes *g, const orc_profile *pi, const uint64_t *reads, const orc_opts *o,
                     uint8_t *out, uint64_t cap, orc_stats *stp) {
    Profile p; if (!make_profile(pi, p)) return -1;
    GenomeView gv{g}; UnitMap um(reads, g->n_genomes, p.is_long); Philox ph(o->seed);
    std::string fmt = o->header_format ? o->header_format : DEFAULT_HEADER;
    uint64_t b = o->unit_begin, e = o->unit_end; if (b == 0 && e == 0) e = um.total();
    if (e > um.total()) return fail("unit range past the end");
    orc_stats st{}; std::vector<uint8_t> buf;
    if (!simulate_range(gv, p, um, ph, fmt, o->read_id_base, b, e, buf, st)) return -1;
    st.bytes = buf.size();
    if (stp) *stp = st;
    return finish(buf, out, cap);
}

/* CPU baseline of bench.py: units [unit_begin, unit_begin + n_units) of a paired-end run split evenly over `threads`
 * std::threads, each writing FASTQ into its own buffer (allocated and touched before the clock starts).
 *   mode 0  native: this library's restatement of the GPU's draw schedule (tables, MWC chains, sparse substitutions)
 *   mode 1  reference stream: simmr's own plumbing — ChaCha12 StdRng re-seeded per pair, a ziggurat normal and an f32
 *           powf per base (ref_stream.inc; minimal-short only).  Every slice seeds its own per-genome generator, as the
 *           reference does once per genome.
 * Returns 0 and the wall time between releasing the threads and the last one finishing. */
int orc_bench_pe(const orc_genomes *g, const orc_profile *pi, const uint64_t *reads, uint64_t seed, uint64_t unit_begin,
                 uint64_t n_units, int mode, int threads, double *seconds, uint64_t *bases, uint64_t *records, uint64_t *bytes) {
    Profile p; if (!make_profile(pi, p)) return -1;
    if (p.is_long) return fail("orc_bench_pe: paired-end profiles only");
    if (mode == 1 && pi->kind != ORC_MINIMAL_SHORT) return fail("orc_bench_pe: the reference stream is restated for minimal-short");
    GenomeView gv{g}; UnitMap um(reads, g->n_genomes, false); Philox ph(seed);
    if (unit_begin + n_units > um.total()) return fail("unit range past the end");
    if (threads < 1) threads = 1;
    const std::string fmt = DEFAULT_HEADER;
    const RefMinimalShort rprof{pi->read_length, pi->insert_size, pi->mean_phred};
    std::vector<std::vector<uint8_t>> bufs(threads);
    std::vector<orc_stats> sts(threads);
    std::vector<int> ok(threads, 1);
    const uint64_t per = (n_units + threads - 1) / threads;
    const uint64_t per_unit = 2 * (256 + 2 * (uint64_t)(pi->read_length + 8 * 15) + 5);
    for (int t = 0; t < threads; t++) { bufs[t].resize(per * per_unit); bufs[t].clear(); }     /* pages mapped before the clock */
    std::atomic<int> go{0};
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back([&, t] {
        while (!go.load(std::memory_order_acquire)) {}
        const uint64_t a = unit_begin + std::min(n_units, per * t), b = unit_begin + std::min(n_units, per * (t + 1));
        orc_stats st{};
        if (mode == 0) ok[t] = simulate_range(gv, p, um, ph, fmt, 0, a, b, bufs[t], st);
        else {
            for (uint64_t u = a; u < b && ok[t];) {
                const uint32_t gi = um.genome_of(u);
                const uint64_t e = std::min(b, um.start[gi + 1]);
                ChaCha12Rng rng_g = ChaCha12Rng::seed_from_u64(seed + 0x9E3779B97F4A7C15ull * u);   /* simulate.rs:172-175, per slice */
                ok[t] = ref_stream_pairs(gv, rprof, gi, e - u, (uint32_t)u, rng_g, fmt, bufs[t], st);
                u = e;
            }
        }
        st.bytes = bufs[t].size(); sts[t] = st;
    });
    const auto t0 = std::chrono::steady_clock::now();
    go.store(1, std::memory_order_release);
    for (auto &x : th) x.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t nb = 0, nr = 0, ny = 0;
    for (int t = 0; t < threads; t++) { if (!ok[t]) return -1; nb += sts[t].bases; nr += sts[t].records; ny += sts[t].bytes; }
    if (seconds) *seconds = dt;
    if (bases) *bases = nb;
    if (records) *records = nr;
    if (bytes) *bytes = ny;
    return 0;
}
/* native-mode k-mer error injection of one read (unit `unit` of a run seeded `seed`) on an arbitrary sequence, with the
 * choices orc_simulate_errors needs to reproduce it */
int64_t orc_native_errors(const orc_profile *pi, uint64_t seed, uint64_t unit, const uint8_t *seq, uint64_t n,
                          uint8_t *out, uint64_t cap, uint32_t *choices, uint64_t choice_cap, uint64_t *n_choices) {
    Profile p; if (!make_profile(pi, p)) return -1;
    if (!p.err.on) return fail("profile has no k-mer table");
    Philox ph(seed);
    std::vector<uint8_t> o; std::vector<uint32_t> ch;
    const std::vector<uint8_t> src(seq, seq + n);
    if (!out) {      /* law of the picks: `cap` consecutive units on the same sequence, the first pick of each into choices */
        for (uint64_t i = 0; i < cap && i < choice_cap; i++) { ch.clear(); native_errors(p, ph, unit + i, src, o, &ch); choices[i] = ch.empty() ? UINT32_MAX : ch[0]; }
        return 0;
    }
    native_errors(p, ph, unit, src, o, &ch);
    if (o.size() > cap || ch.size() > choice_cap) return fail("orc_native_errors: buffer too small");
    if (!o.empty()) memcpy(out, o.data(), o.size());
    if (!ch.empty()) memcpy(choices, ch.data(), ch.size() * 4);
    if (n_choices) *n_choices = ch.size();
    return (int64_t)o.size();
}
/* one ChaCha12 block (rand_chacha layout: 64-bit counter in words 12-13, stream id 0) — known-answer hook for the tests */
void orc_chacha12_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    ChaCha12Rng r; memcpy(r.key, key, sizeof r.key); r.counter = counter; r.refill();
    memcpy(out, r.buf, 64);
}
/* n standard-normal deviates of the restated ziggurat from StdRng::seed_from_u64(seed) */
void orc_ref_ziggurat(uint64_t seed, uint64_t n, double *out) {
    ChaCha12Rng r = ChaCha12Rng::seed_from_u64(seed);
    for (uint64_t i = 0; i < n; i++) out[i] = zig().sample(r);
}
/* first `n` bytes of a reference-stream run (pairs [0, n_pairs) of genome 0 seeded with `seed`): lets the tests check the
 * stream's output as FASTQ and its distributions */
int64_t orc_ref_stream(const orc_genomes *g, const orc_profile *pi, uint64_t n_pairs, uint64_t seed, uint8_t *out, uint64_t cap, orc_stats *stp) {
    if (pi->kind != ORC_MINIMAL_SHORT) return fail("orc_ref_stream: minimal-short only");
    GenomeView gv{g};
    const RefMinimalShort rprof{pi->read_length, pi->insert_size, pi->mean_phred};
    ChaCha12Rng rng_g = ChaCha12Rng::seed_from_u64(seed);
    std::vector<uint8_t> buf; orc_stats st{};
    if (!ref_stream_pairs(gv, rprof, 0, n_pairs, 0, rng_g, DEFAULT_HEADER, buf, st)) return -1;
    st.bytes = buf.size();
    if (stp) *stp = st;
    return finish(buf, out, cap);
}

/* replay: the reference's arithmetic on supplied draws */
int64_t orc_replay(const orc_genomes *g, const orc_profile *pi, const orc_draws *d, const orc_opts *o,
                   uint8_t *out, uint64_t cap, orc_stats *stp) {
    Profile p; if (!make_profile(pi, p)) return -1;
    GenomeView gv{g};
    std::string fmt = o->header_format ? o->header_format : DEFAULT_HEADER;
    orc_stats st{}; std::vector<uint8_t> buf;
    auto phred_at = [&](uint64_t slot, uint64_t i) -> uint8_t {
        switch (p.kind) {
        case ORC_PERFECT_SHORT: return 60;                                             /* perfect_short.rs:43 */
        case ORC_MINIMAL_SHORT: case ORC_MINIMAL_LONG: {                               /* minimal_short.rs:90-100 */
            float v = (float)p.mean_phred + 10.0f * d->phred_z[slot];
            return sat_u8(floorf(v)); }
        case ORC_PERFECT_LONG: {                                                       /* perfect_long.rs:66-77 */
            float a = phred_to_accuracy(20) + 0.05f * d->phred_z[slot];
            a = fminf(a, 0.9999f);
            return accuracy_to_phred(a); }
        default: {                                                                     /* custom_short.rs:108-129, 339-349 */
            size_t pos = std::min<uint64_t>(i, p.q_alias.size() - 1);
            const RustAlias &al = p.q_alias[pos];
            uint32_t cand = d->phred_cand[slot] % (uint32_t)al.odds.size();
            uint32_t bin = al.pick(cand, d->phred_u[slot]);
            auto r = bin_range(p.blob.quality[pos], bin);
            uint64_t span = (uint64_t)std::max(r.first, r.second) - r.first + 1;
            uint32_t v = r.first + (uint32_t)(((uint64_t)d->phred_inbin[slot] * span) >> 32);
            return (uint8_t)v; }                                                       /* `as u8` truncates */
        }
    };
    auto one_read = [&](const uint8_t *src, uint64_t L, uint64_t slot0, std::vector<uint8_t> &seq, std::vector<uint8_t> &qual) {
        seq.resize(L); qual.resize(L);
        for (uint64_t i = 0; i < L; i++) {
            uint8_t q = phred_at(slot0 + i, i); qual[i] = q;
            uint8_t b = src[i];
            if (p.mutates && d->mut_u[slot0 + i] > phred_to_accuracy(q)) {             /* minimal_short.rs:123 */
                uint8_t nb = substitute(b, d->mut_k[slot0 + i] % 3); if (nb != b) st.mutated++; b = nb;
            }
            seq[i] = b;
        }
    };
    uint64_t b = o->unit_begin, e = o->unit_end; if (b == 0 && e == 0) e = d->n_units;
    std::vector<uint8_t> s1, q1, s2, q2;
    for (uint64_t u = b; u < e; u++) {
        uint32_t gi = d->genome[u];
        if (gi >= g->n_genomes || d->contig[u] >= gv.nseq(gi)) return fail("replay: genome/contig out of range");
        uint32_t c = gv.contig(gi, d->contig[u]);
        uint32_t id = o->read_id_base + (uint32_t)u;
        std::string gid = gv.gid(gi), sid = gv.cid(c);
        const uint8_t *cs = gv.seq(c);
        uint64_t n = gv.size(c), RL = d->read_length[u], slot = d->base_off[u];
        if (!p.is_long) {
            uint64_t IS = d->insert_size[u], R = p.min_size;
            PeGeom pg = pe_geometry(n, gv.phys(c), RL, IS, R, d->start[u], [&](uint64_t, uint64_t) { return d->redraw[u]; });
            if (pg.clamped) st.clamped++;
            one_read(cs + pg.fs, pg.fe - pg.fs, slot, s1, q1);
            one_read(cs + pg.lo, pg.hi - pg.lo, slot + RL, s2, q2);
            auto rc = reverse_complement(s2.data(), s2.size());
            emit_record(buf, fmt, gid, id, sid, pg.fs, d->start[u] + RL, false, 1, s1, q1);
            emit_record(buf, fmt, gid, id, sid, pg.lo + RL, pg.lo, true, 2, rc, q2);
            st.records += 2; st.bases += s1.size() + s2.size();
        } else {
            uint64_t start = d->start[u], end = start + RL;                            /* simulate.rs:484-485 */
            if (end >= n) end = d->redraw[u];                                          /* :488-489 */
            uint64_t phys = gv.phys(c);
            if (end > phys) { end = phys; st.clamped++; }
            if (start > end) start = end;
            one_read(cs + start, end - start, slot, s1, q1);
            emit_record(buf, fmt, gid, id, sid, start, end, false, 1, s1, q1);
            st.records += 1; st.bases += s1.size();
        }
        st.units++;
    }
    st.bytes = buf.size();
    if (stp) *stp = st;
    return finish(buf, out, cap);
}

/* ---- reference-arithmetic samplers (independent generator) ---- */
void orc_ref_minimal_phred(uint8_t mean_phred, uint64_t n, uint64_t seed, uint8_t *out) {
    Xo r(seed);
    for (uint64_t i = 0; i < n; i++) { float z = (float)r.normal(); out[i] = sat_u8(floorf((float)mean_phred + 10.0f * z)); }   /* minimal_short.rs:90-100 */
}
void orc_ref_perfect_long_phred(uint64_t n, uint64_t seed, uint8_t *out) {
    Xo r(seed);
    for (uint64_t i = 0; i < n; i++) { float a = phred_to_accuracy(20) + 0.05f * (float)r.normal(); out[i] = accuracy_to_phred(fminf(a, 0.9999f)); }
}
void orc_ref_normal_len(double mean, double sd, uint64_t n, uint64_t seed, uint16_t *out) {
    Xo r(seed);
    for (uint64_t i = 0; i < n; i++) out[i] = sat_u16(std::floor(mean + sd * r.normal()));    /* minimal_short.rs:40-41 */
}
void orc_ref_gamma_len(uint64_t n, uint64_t seed, uint16_t *out) {   /* perfect_long.rs:40-55, rand_distr GammaLargeShape<f32> */
    Xo r(seed);
    const float mean = 20000.0f, sd = 15000.0f, shape = powf(mean / sd, 2.0f), scale = powf(sd, 2.0f) / mean;
    const float d = shape - 1.0f / 3.0f, c = 1.0f / sqrtf(9.0f * d);
    for (uint64_t i = 0; i < n; i++) {
        for (;;) {
            float x = (float)r.normal(), vc = 1.0f + c * x;
            if (vc <= 0.0f) continue;
            float v = vc * vc * vc, u = (float)r.u01(), x2 = x * x;
            if (u < 1.0f - 0.0331f * x2 * x2 || logf(u) < 0.5f * x2 + d * (1.0f - v + logf(v))) { out[i] = sat_u16(floorf(d * v * scale)); break; }
        }
    }
}
void orc_ref_mutate(const uint8_t *seq, const uint8_t *q, uint64_t n, uint64_t seed, uint8_t *out) {   /* minimal_short.rs:104-140 */
    Xo r(seed);
    for (uint64_t i = 0; i < n; i++) {
        uint8_t b = seq[i];
        if (r.f32_24() > phred_to_accuracy(q[i])) b = substitute(b, r.below(3));
        out[i] = b;
    }
}
int orc_ref_custom_phred(const uint8_t *blob, uint64_t blob_len, uint32_t pos, uint64_t n, uint64_t seed, uint8_t *out) {
    Blob b; if (!parse_blob(blob, blob_len, b)) return fail("blob");
    if (b.quality.empty()) return fail("no positions");
    if (pos >= b.quality.size()) pos = (uint32_t)b.quality.size() - 1;
    RustAlias al; if (!al.build(b.quality[pos].density)) return fail("alias");
    Xo r(seed);
    for (uint64_t i = 0; i < n; i++) {
        uint32_t bin = al.pick(r.below((uint32_t)al.odds.size()), r.u01() * al.sum);
        auto rg = bin_range(b.quality[pos], bin);
        uint32_t span = std::max(rg.first, rg.second) - rg.first + 1;
        out[i] = (uint8_t)(rg.first + r.below(span));
    }
    return 0;
}
int orc_ref_custom_len(const uint8_t *blob, uint64_t blob_len, int which, uint64_t n, uint64_t seed, uint16_t *out) {
    Blob b; if (!parse_blob(blob, blob_len, b)) return fail("blob");
    const Bins &bn = which ? b.is_bins : b.rl_bins;
    if (which && !b.has_is_bins) { for (uint64_t i = 0; i < n; i++) out[i] = 0; return 0; }   /* custom_short.rs:255-257 */
    RustAlias al; if (!al.build(bn.density)) return fail("alias");
    Xo r(seed);
    for (uint64_t i = 0; i < n; i++) {
        uint32_t bin = al.pick(r.below((uint32_t)al.odds.size()), r.u01() * al.sum);
        auto rg = bin_range(bn, bin);
        uint32_t span = std::max(rg.first, rg.second) - rg.first + 1;
        out[i] = (uint16_t)(rg.first + r.below(span));
    }
    return 0;
}

}  // extern "C"

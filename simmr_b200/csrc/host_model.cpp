/* host_model.cpp — see host_model.hpp.  Compile with -ffp-contract=off: the
 * sampling tables are specified as exact IEEE-754 double computations
 *  so that independently written builders agree bit for bit. */
#include "host_model.hpp"

#include <zlib.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>

namespace simmr {

std::string &last_error_slot() { thread_local std::string slot; return slot; }

const char *const DEFAULT_HEADER_FORMAT =
    "@{:read_id:}|{:genome_id:}/{:pair:} metadata:sid={:sequence_id:}|sp={:start_position:}|ep={:end_position:}|rc={:reverse_complement:}";

/* ====================================================================== */
/* Philox4x32-10 (Salmon et al., SC'11) — host copy for the parity tests    */
/* ====================================================================== */
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        const uint64_t lo = 0xD2511F53ull * c[0], hi = 0xCD9E8D57ull * c[2];
        const uint32_t t0 = uint32_t(hi >> 32) ^ c[1] ^ k0, t2 = uint32_t(lo >> 32) ^ c[3] ^ k1;
        c[1] = uint32_t(hi); c[3] = uint32_t(lo); c[0] = t0; c[2] = t2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof c);
}

/* ====================================================================== */
/* blob (bincode 1.3.3 default: LE, u64 lengths, 1-byte Option tag / bool)   */
/* ====================================================================== */
namespace {
class Reader {
public:
    Reader(const uint8_t *p, uint64_t n) : p_(p), n_(n) {}
    template <class T> bool take(T &v) {
        if (n_ - at_ < sizeof(T)) return false;
        memcpy(&v, p_ + at_, sizeof(T)); at_ += sizeof(T); return true;
    }
    uint64_t left() const { return n_ - at_; }
private:
    const uint8_t *p_; uint64_t n_, at_ = 0;
};
bool take_bins(Reader &r, BinSet &b) {
    uint64_t n = 0, m = 0;
    if (!r.take(b.num_bins) || !r.take(b.bin_width) || !r.take(n) || n > r.left() / 8) return false;
    b.density.resize(n);
    for (uint64_t i = 0; i < n; ++i) if (!r.take(b.density[i])) return false;
    if (!r.take(m) || m > r.left() / 8) return false;
    b.lo.resize(m); b.hi.resize(m);
    for (uint64_t i = 0; i < m; ++i) if (!r.take(b.lo[i]) || !r.take(b.hi[i])) return false;
    return true;
}
}  // namespace

/* Range of density column j.  simmrd writes 71 densities but 70 ranges
 * (simmrd/src/probability.rs:119-174), which makes the reference index out
 * of bounds (custom_short.rs:122-123); columns past the ranges continue as
 * unit-width bins. */
void BinSet::range_of(size_t j, uint32_t &l, uint32_t &h) const {
    if (j < lo.size()) { l = lo[j]; h = hi[j] < lo[j] ? lo[j] : hi[j]; return; }
    l = h = lo.empty() ? uint32_t(j) : hi.back() + 1 + uint32_t(j - lo.size());
}

bool decode_blob(const uint8_t *p, uint64_t n, ProfileBlob &o) {
    Reader r(p, n);
    uint64_t npos = 0, nk = 0;
    if (!r.take(o.bin_size) || !r.take(npos) || npos > n) return false;
    o.quality.resize(npos);
    for (auto &q : o.quality) if (!take_bins(r, q)) return false;
    if (!r.take(o.bit_encoding) || !r.take(o.kmer_size) || !r.take(nk) || nk > n) return false;
    o.alt_off.assign(1, 0);
    for (uint64_t i = 0; i < nk; ++i) {
        uint32_t ref; uint64_t na;
        if (!r.take(ref) || !r.take(na) || na > r.left() / 8) return false;
        o.ref_kmer.push_back(ref);
        for (uint64_t a = 0; a < na; ++a) {
            uint32_t alt; float pr;
            if (!r.take(alt) || !r.take(pr)) return false;
            o.alt_kmer.push_back(alt); o.alt_prob.push_back(pr);
        }
        o.alt_off.push_back(uint32_t(o.alt_kmer.size()));
    }
    uint8_t tag = 0, lng = 0;
    if (!r.take(o.insert_mean) || !r.take(o.insert_std) || !r.take(tag)) return false;
    o.has_insert_bins = tag != 0;
    if (o.has_insert_bins && !take_bins(r, o.insert_bins)) return false;
    if (!r.take(o.length_mean) || !r.take(o.length_std) || !take_bins(r, o.length_bins) || !r.take(lng)) return false;
    o.is_long = lng != 0;
    return true;
}

/* ====================================================================== */
/* sampling tables                                                          */
/* ====================================================================== */
namespace {

/* util.rs:69-71, 96-98, 109-111 in f32 exactly as the reference */
inline float accuracy_of(uint8_t q) { return 1.0f - powf(10.0f, -(float(q) / 10.0f)); }
inline uint8_t cast_u8(float x) { return !(x == x) || x <= 0.f ? 0 : (x >= 255.f ? 255 : uint8_t(x)); }
inline uint8_t phred_of_accuracy(float a) { return cast_u8(roundf(-10.0f * log10f(1.0f - a))); }
inline uint16_t cast_u16(double x) { return !(x == x) || x <= 0.0 ? 0 : (x >= 65535.0 ? 65535 : uint16_t(x)); }

inline uint32_t bits_for(size_t n) { uint32_t b = 0; while ((size_t(1) << b) < n) ++b; return b ? b : 1; }

/* Vose alias method over 2^cb columns: keep[i] is the probability of
 * staying in column i, other[i] the column taken otherwise. */
bool alias_columns(const std::vector<double> &weight, uint32_t cb, std::vector<double> &keep, std::vector<uint32_t> &other) {
    const size_t cols = size_t(1) << cb;
    if (weight.size() > cols) return false;
    double total = 0.0;
    for (double w : weight) total += w;
    if (!(total > 0.0)) return false;
    std::vector<double> height(cols, 0.0);
    keep.assign(cols, 1.0); other.resize(cols);
    for (size_t i = 0; i < weight.size(); ++i) height[i] = weight[i] * double(cols) / total;
    std::vector<uint32_t> under, over;
    for (size_t i = 0; i < cols; ++i) {
        other[i] = uint32_t(i);
        (height[i] < 1.0 ? under : over).push_back(uint32_t(i));
    }
    while (!under.empty() && !over.empty()) {
        const uint32_t u = under.back(), v = over.back();
        under.pop_back(); over.pop_back();
        keep[u] = height[u]; other[u] = v;
        height[v] = (height[v] + height[u]) - 1.0;
        (height[v] < 1.0 ? under : over).push_back(v);
    }
    for (size_t i = 0; i < cols; ++i) if (keep[i] >= 1.0) other[i] = uint32_t(i);
    return true;
}
/* floor(p * 2^bits) saturated to `bits` bits; certain columns get all ones */
inline uint32_t threshold(double p, uint32_t bits) {
    const uint32_t ones = bits >= 32 ? 0xffffffffu : (1u << bits) - 1u;
    if (p >= 1.0) return ones;
    const double f = std::floor(std::ldexp(p, int(bits)));
    return f <= 0.0 ? 0u : (f >= double(ones) ? ones : uint32_t(f));
}
bool len_words(const std::vector<double> &weight, uint32_t cb, std::vector<uint32_t> &words) {
    std::vector<double> keep; std::vector<uint32_t> other;
    if (!alias_columns(weight, cb, keep, other)) return false;
    words.resize(keep.size());
    for (size_t i = 0; i < keep.size(); ++i) words[i] = (threshold(keep[i], 20) << 12) | other[i];
    return true;
}

inline double phi(double x, double mu, double sd) { return 0.5 * std::erfc(-(x - mu) / (sd * 1.4142135623730951)); }

/* distribution of `Normal(mu,sd).floor() as uN` on [mu-8sd, mu+8sd], tails lumped */
void floor_normal(double mu, double sd, uint32_t vmax, uint32_t &first, std::vector<double> &mass) {
    const double a = std::floor(mu - 8.0 * sd), b = std::ceil(mu + 8.0 * sd);
    first = a <= 0.0 ? 0u : (a >= double(vmax) ? vmax : uint32_t(a));
    const uint32_t last = b <= 0.0 ? 0u : (b >= double(vmax) ? vmax : uint32_t(b));
    mass.assign(last - first + 1, 0.0);
    for (uint32_t v = first; v <= last; ++v) {
        const double below = v == first ? 0.0 : phi(double(v), mu, sd);
        const double upto = v == last ? 1.0 : phi(double(v) + 1.0, mu, sd);
        mass[v - first] = upto - below;
    }
}

/* probability that a 24-bit uniform k/2^24 exceeds accuracy(q) (minimal_short.rs:123) */
double flip_probability(uint8_t q) {
    const double t = std::floor(double(accuracy_of(q)) * 16777216.0);
    if (t < 0.0) return 1.0;
    if (t >= 16777215.0) return 0.0;
    return (16777215.0 - t) / 16777216.0;
}

/* PT_Q32 word of a column: thr16<<16 | alias<<8 | self.  The sampler keeps `self` iff word < entry, so a column whose
 * threshold rounds to zero names its alias twice (the payload bits must not open a window for `self`). */
inline uint32_t q32_word(double keep, uint32_t self_val, uint32_t alias_val) {
    const uint32_t thr = threshold(keep, 16);
    if (thr == 0) self_val = alias_val;
    return (thr << 16) | ((alias_val & 0xffu) << 8) | (self_val & 0xffu);
}
/* `first` = value of column 0; quality tables pass first = lowest quality + 33: PT_Q32 words hold the characters as
 * printed (util.rs:46-50), because the whole word is what a draw is compared with */
bool q32_words(const std::vector<double> &w, uint32_t first, uint32_t cb, std::vector<uint32_t> &out) {
    std::vector<double> keep; std::vector<uint32_t> other;
    if (!alias_columns(w, cb, keep, other)) return false;
    out.resize(keep.size());
    for (size_t i = 0; i < keep.size(); ++i) out[i] = q32_word(keep[i], first + uint32_t(i), first + other[i]);
    return true;
}
/* Mutating profiles (minimal_short.rs:83-140, minimal_long.rs:78-140, perfect_long.rs:60-119), factored the other way round: a base is substituted with
 * the marginal probability s = sum_q P(q) flip(q); its quality then follows P(q | substituted), every other base's
 * P(q | kept).  Bases are independent, so the substituted ones are found by geometric gaps. */
bool phred_from_mass(const std::vector<double> &mass, uint32_t first, ProfileModel &m) {
    const size_t n = mass.size();
    std::vector<double> kept(n), subst(n);
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double f = flip_probability(uint8_t(first + i));
        kept[i] = mass[i] * (1.0 - f); subst[i] = mass[i] * f;
        s += subst[i];
    }
    m.phred_kind = PT_Q32; m.npos = 1; m.cb = bits_for(n);
    if (m.cb > 8) return false;
    if (!q32_words(kept, first + 33u, m.cb, m.phred_entries)) return false;
    m.has_sub = s > 0.0;
    if (m.has_sub) {
        if (!q32_words(subst, first + 33u, m.cb, m.qsub_entries)) return false;
        std::vector<double> gap(GAP_COLS);
        double none = 1.0;                                   /* P(no substitution among the first k bases) */
        for (uint32_t k = 0; k < GAP_NONE; ++k) { gap[k] = s * none; none = none * (1.0 - s); }
        gap[GAP_NONE] = none;
        if (!q32_words(gap, 0, 8, m.gap_entries)) return false;
    }
    return true;
}

bool phred_from_blob(const ProfileBlob &b, ProfileModel &m) {
    const size_t np = b.quality.size();
    if (!np) return false;
    std::vector<std::vector<double>> mass(np, std::vector<double>(256, 0.0));
    uint32_t qmin = 255, qmax = 0;
    for (size_t p = 0; p < np; ++p) {
        const BinSet &bs = b.quality[p];
        if (bs.density.empty()) return false;
        double total = 0.0;
        for (double d : bs.density) total += d;
        if (!(total > 0.0)) return false;
        for (size_t j = 0; j < bs.density.size(); ++j) {
            if (!(bs.density[j] > 0.0)) continue;
            uint32_t l, h; bs.range_of(j, l, h);
            const uint64_t span = uint64_t(h) - l + 1, whole = span / 256, part = span % 256;
            for (uint32_t v = 0; v < 256; ++v) {       /* `as u8` folds values mod 256 */
                const uint64_t hits = whole + ((((v + 256u) - (l & 255u)) & 255u) < part ? 1 : 0);
                if (!hits) continue;
                mass[p][v] += bs.density[j] / total * (double(hits) / double(span));
                qmin = std::min(qmin, v); qmax = std::max(qmax, v);
            }
        }
    }
    if (qmin > qmax) return false;
    /* A row's columns are its qualities that can occur at all — mass above 2^-24 of the row, anything rarer would get a
     * zero 16-bit threshold anyway — in increasing order (the words hold the values themselves), so the table is as wide
     * as the richest position, not as the range of all positions. */
    std::vector<double> cut(np);
    size_t widest = 1;
    for (size_t p = 0; p < np; ++p) {
        double row = 0.0;
        for (uint32_t v = qmin; v <= qmax; ++v) row += mass[p][v];
        cut[p] = std::ldexp(row, -24);
        size_t n = 0;
        for (uint32_t v = qmin; v <= qmax; ++v) n += mass[p][v] > cut[p];
        widest = std::max(widest, n);
    }
    m.phred_kind = PT_Q32; m.npos = uint32_t(np); m.cb = bits_for(widest);
    m.phred_entries.clear();
    for (size_t p = 0; p < np; ++p) {
        std::vector<double> w, keep; std::vector<uint32_t> value, other;
        for (uint32_t v = qmin; v <= qmax; ++v) if (mass[p][v] > cut[p]) { w.push_back(mass[p][v]); value.push_back(v); }
        if (!alias_columns(w, m.cb, keep, other)) return false;
        value.resize(keep.size(), value.back());               /* columns without an outcome always take their alias */
        for (size_t i = 0; i < keep.size(); ++i)
            m.phred_entries.push_back(q32_word(keep[i], value[i] + 33u, value[other[i]] + 33u));
    }
    return true;
}

bool len_from_mass(const std::vector<double> &mass, uint32_t first, LenModel &t) {
    t.kind = LEN_ALIAS; t.cb = bits_for(mass.size());
    if (t.cb > 12) return false;
    const size_t cols = size_t(1) << t.cb;
    t.lo.assign(cols, first); t.hi.assign(cols, first);
    for (size_t i = 0; i < mass.size(); ++i) t.lo[i] = t.hi[i] = first + uint32_t(i);
    t.vmax = first + uint32_t(mass.size()) - 1;
    return len_words(mass, t.cb, t.entries);
}
bool len_from_bins(const BinSet &b, LenModel &t) {
    if (b.density.empty()) return false;
    t.kind = LEN_ALIAS; t.cb = bits_for(b.density.size());
    if (t.cb > 12) return false;
    const size_t cols = size_t(1) << t.cb;
    t.lo.assign(cols, 0); t.hi.assign(cols, 0); t.vmax = 0;
    for (size_t j = 0; j < b.density.size(); ++j) { b.range_of(j, t.lo[j], t.hi[j]); t.vmax = std::max(t.vmax, t.hi[j]); }
    return len_words(b.density, t.cb, t.entries);
}
void len_const(LenModel &t, uint32_t v) { t = LenModel(); t.kind = LEN_CONST; t.value = v; t.vmax = v; }

/* rand_distr 0.4.3 weighted_alias.rs WeightedAliasIndex<f64>::new, for replay */
double tree_sum(const double *v, size_t n) {
    if (n <= 32) { double s = 0.0; for (size_t i = 0; i < n; ++i) s += v[i]; return s; }
    return tree_sum(v, n / 2) + tree_sum(v + n / 2, n - n / 2);
}
bool rust_alias(const std::vector<double> &w, std::vector<double> &odds, std::vector<uint32_t> &alias, double &sum) {
    const uint32_t n = uint32_t(w.size()), nil = UINT32_MAX;
    if (!n) return false;
    sum = tree_sum(w.data(), n);
    if (!(sum > 0.0)) return false;
    odds = w;
    for (double &o : odds) o *= double(n);
    alias.assign(n, 0);
    uint32_t small_head = nil, big_head = nil;   /* intrusive stacks threaded through `alias` */
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t &head = odds[i] < sum ? small_head : big_head;
        alias[i] = head; head = i;
    }
    while (small_head != nil && big_head != nil) {
        const uint32_t s = small_head; small_head = alias[s];
        const uint32_t b = big_head;   big_head = alias[b];
        alias[s] = b;
        odds[b] = odds[b] - sum + odds[s];
        uint32_t &head = odds[b] < sum ? small_head : big_head;
        alias[b] = head; head = b;
    }
    for (; small_head != nil; ) { const uint32_t s = small_head; small_head = alias[s]; odds[s] = sum; }
    for (; big_head != nil; )   { const uint32_t b = big_head;   big_head = alias[b];   odds[b] = sum; }
    return true;
}
/* The alt-k-mer table of a blob as sampling tables (ErrModel).  A k-mer listed twice keeps its last entry (the
 * reference collects the list into a HashMap, custom_short.rs:462-467); an alternate list without weight would make the
 * reference's WeightedAliasIndex::new fail and is treated as "k-mer not observed"; alternates that cannot be decoded
 * (a digit above 4, encoding.rs:205) make the profile invalid. */
bool err_model(const ProfileBlob &b, ErrModel &e) {
    e = ErrModel();
    const size_t n = b.ref_kmer.size();
    if (!n) return true;
    if (b.bit_encoding != 3 || b.kmer_size < 1 || b.kmer_size > 10) return false;
    e.k = uint32_t(b.kmer_size);
    const uint32_t digits = e.k == 10 ? 0x3fffffffu : (1u << (3 * e.k)) - 1u;
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return b.ref_kmer[x] < b.ref_kmer[y]; });
    for (size_t at = 0; at < n;) {
        size_t last = at;
        while (last + 1 < n && b.ref_kmer[order[last + 1]] == b.ref_kmer[order[at]]) ++last;
        const uint32_t src = order[last], kmer = b.ref_kmer[src];      /* the stable sort keeps the list's order among equals */
        at = last + 1;
        double self = 0.0, all = 0.0;
        std::vector<double> weight; std::vector<uint32_t> code;
        for (uint32_t j = b.alt_off[src]; j < b.alt_off[src + 1]; ++j) {
            const uint32_t alt = b.alt_kmer[j] & digits;
            for (uint32_t d = 0; d < e.k; ++d) if (((alt >> (3 * d)) & 7u) > 4u) return false;
            const double w = double(b.alt_prob[j]);
            if (!(w >= 0.0)) return false;
            all += w;
            if (alt == (kmer & digits)) self += w;
            else if (w > 0.0) { weight.push_back(w); code.push_back(alt); }
        }
        if (!(all > 0.0)) continue;
        if (weight.size() > 4096) return false;
        e.ref.push_back(kmer); e.row.push_back(uint32_t(e.entries.size()));
        if (weight.empty()) {
            e.keep_thr.push_back(0xffffffffu); e.cb.push_back(0);
            e.entries.push_back(0xfffff000u); e.alts.push_back(kmer & digits);
            continue;
        }
        e.keep_thr.push_back(threshold(self / all, 32));
        uint32_t cb = 0;
        while ((size_t(1) << cb) < weight.size()) ++cb;
        e.cb.push_back(cb);
        std::vector<uint32_t> words;
        if (!len_words(weight, cb, words)) return false;
        e.entries.insert(e.entries.end(), words.begin(), words.end());
        for (uint32_t j = 0; j < (1u << cb); ++j) e.alts.push_back(j < code.size() ? code[j] : (kmer & digits));
    }
    e.on = !e.ref.empty();
    if (e.on) {
        e.direct.assign(size_t(1) << (2 * e.k), uint2_host{0u, 0u});
        for (uint32_t r = 0; r < e.ref.size(); ++r) {
            uint32_t two = 0; bool acgt = true;
            for (uint32_t d = 0; d < e.k; ++d) { const uint32_t v = (e.ref[r] >> (3 * d)) & 7u; if (v > 3u) acgt = false; two |= (v & 3u) << (2 * d); }
            if (acgt && !(e.ref[r] & ~digits)) e.direct[two] = uint2_host{e.keep_thr[r], r + 1u};
        }
    }
    return true;
}

}  // namespace

int build_profile(const simmr_eprofile_desc *d, ProfileModel &m, std::string &err) {
    m = ProfileModel();
    m.kind = d->kind; m.flags = d->flags;
    for (int q = 0; q < 256; ++q) m.acc_lut[q] = accuracy_of(uint8_t(q));
    m.acc20 = accuracy_of(20);
    m.mean_phred_f = float(d->mean_phred);
    {   /* Marsaglia-Tsang constants: Gamma::new(shape, scale) of perfect_long.rs:48-54 */
        const float mean = 20000.0f, sd = 15000.0f, shape = powf(mean / sd, 2.0f), scale = powf(sd, 2.0f) / mean;
        m.gamma.d = double(shape) - 1.0 / 3.0;
        m.gamma.c = 1.0 / std::sqrt(9.0 * m.gamma.d);
        m.gamma.scale = double(scale);
    }
    ProfileBlob blob;
    const bool custom = d->kind == SIMMR_EP_CUSTOM_SHORT || d->kind == SIMMR_EP_CUSTOM_LONG;
    if (custom && (!d->blob || !decode_blob(d->blob, d->blob_len, blob))) { err = "cannot parse the custom error profile"; return SIMMR_E_PARSE; }
    std::vector<double> mass; uint32_t first = 0;
    auto long_common = [&]() { m.is_long = true; m.min_size = 20000; m.rl = LenModel(); m.rl.kind = LEN_GAMMA; m.rl.vmax = 65535; len_const(m.is, 0); };
    switch (d->kind) {
    case SIMMR_EP_PERFECT_SHORT:    /* perfect_short.rs:22-63 */
        m.min_size = uint16_t(2 * d->read_length + d->insert_size);
        m.phred_kind = PT_CONST; m.const_q = 60; m.replay_kind = RP_CONST;
        len_const(m.rl, d->read_length); len_const(m.is, d->insert_size);
        break;
    case SIMMR_EP_MINIMAL_SHORT:    /* minimal_short.rs:33-149; sigmas from cli.rs:239-240 */
        m.min_size = uint16_t(2 * d->read_length + d->insert_size);
        m.mutates = true; m.replay_kind = RP_NORMAL_FLOOR;
        floor_normal(double(m.mean_phred_f), 10.0, 255, first, mass);
        if (!phred_from_mass(mass, first, m)) { err = "phred table"; return SIMMR_E_INVALID; }
        floor_normal(double(d->read_length), 15.0, 65535, first, mass);
        if (!len_from_mass(mass, first, m.rl)) { err = "read-length table"; return SIMMR_E_INVALID; }
        floor_normal(double(d->insert_size), 75.0, 65535, first, mass);
        if (!len_from_mass(mass, first, m.is)) { err = "insert-size table"; return SIMMR_E_INVALID; }
        break;
    case SIMMR_EP_PERFECT_LONG: {   /* perfect_long.rs:36-135 */
        long_common(); m.mutates = true; m.replay_kind = RP_ACCURACY;
        const double mu = double(m.acc20), sd = double(0.05f);
        const uint8_t cap = phred_of_accuracy(0.9999f);
        auto edge = [](int k) { return 1.0 - std::pow(10.0, -(double(k) - 0.5) / 10.0); };
        mass.assign(size_t(cap) + 1, 0.0);
        for (int k = 0; k <= cap; ++k)
            mass[k] = (k == cap ? 1.0 : phi(edge(k + 1), mu, sd)) - (k == 0 ? 0.0 : phi(edge(k), mu, sd));
        if (!phred_from_mass(mass, 0, m)) { err = "phred table"; return SIMMR_E_INVALID; }
        /* replay: smallest f32 accuracy whose phred is >= k, by bisection on the bit pattern */
        m.pl_qmax = cap;
        for (uint32_t k = 1; k <= cap; ++k) {
            uint32_t lo = 0, hi; const float top = 0.9999f; memcpy(&hi, &top, 4);
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo) / 2; float a; memcpy(&a, &mid, 4);
                if (phred_of_accuracy(a) >= k) hi = mid; else lo = mid + 1;
            }
            memcpy(&m.pl_thr[k], &lo, 4);
        }
        m.pl_thr[0] = -INFINITY;
        break; }
    case SIMMR_EP_MINIMAL_LONG:     /* minimal_long.rs:44-158 */
        long_common(); m.mutates = true; m.replay_kind = RP_NORMAL_FLOOR;
        floor_normal(double(m.mean_phred_f), 10.0, 255, first, mass);
        if (!phred_from_mass(mass, first, m)) { err = "phred table"; return SIMMR_E_INVALID; }
        break;
    case SIMMR_EP_CUSTOM_SHORT:     /* custom_short.rs:155-229, 237-270, 522-542 */
        m.is_long = blob.is_long; m.replay_kind = RP_ALIAS;
        m.min_size = cast_u16(2.0 * blob.length_mean + blob.insert_mean);
        if (!phred_from_blob(blob, m)) { err = "custom profile has an empty quality distribution"; return SIMMR_E_PARSE; }
        if (!len_from_bins(blob.length_bins, m.rl)) { err = "custom profile has no read-length bins"; return SIMMR_E_PARSE; }
        if (blob.has_insert_bins) { if (!len_from_bins(blob.insert_bins, m.is)) { err = "custom profile has empty insert-size bins"; return SIMMR_E_PARSE; } }
        else len_const(m.is, 0);
        break;
    case SIMMR_EP_CUSTOM_LONG:      /* SURVEY.md §A.5 */
        long_common(); m.replay_kind = RP_ALIAS;
        if (!phred_from_blob(blob, m)) { err = "custom profile has an empty quality distribution"; return SIMMR_E_PARSE; }
        if (!err_model(blob, m.err)) { err = "custom profile has a malformed alt-k-mer table"; return SIMMR_E_PARSE; }
        break;
    default: err = "unknown error profile kind"; return SIMMR_E_INVALID;
    }
    m.max_read_len = m.is_long ? 65535u : std::min<uint32_t>(m.rl.vmax, 65535u);   /* `as u16` */
    m.nominal_read_len = m.is_long ? 20000u : (custom ? uint32_t(std::min(65535.0, std::max(1.0, blob.length_mean))) : d->read_length);
    m.nominal_read_len = std::min(m.nominal_read_len, m.max_read_len);
    if (custom) {
        m.r_off.assign(1, 0);
        for (const BinSet &bs : blob.quality) {
            std::vector<double> odds; std::vector<uint32_t> al; double sum;
            if (!rust_alias(bs.density, odds, al, sum)) { err = "custom profile has an empty quality distribution"; return SIMMR_E_PARSE; }
            for (size_t j = 0; j < odds.size(); ++j) { uint32_t l, h; bs.range_of(j, l, h); m.r_lo.push_back(l); m.r_hi.push_back(h); }
            m.r_odds.insert(m.r_odds.end(), odds.begin(), odds.end());
            m.r_alias.insert(m.r_alias.end(), al.begin(), al.end());
            m.r_sum.push_back(sum);
            m.r_off.push_back(uint32_t(m.r_odds.size()));
        }
    }
    return SIMMR_OK;
}

/* ====================================================================== */
/* abundance — abundance_profiles/{uniform,exact,custom}.rs                 */
/* ====================================================================== */
namespace {
inline uint64_t cast_usize(double x) { return !(x == x) || x <= 0.0 ? 0 : (x >= 18446744073709551615.0 ? UINT64_MAX : uint64_t(x)); }
}
int abundance_counts(const simmr_aprofile_desc *d, uint64_t total_reads, uint32_t n, const uint64_t *sizes,
                     uint64_t *out_reads, double *out_abund, std::string &err) {
    bool adjust = d->size_adjusted != 0;
    const double total = double(total_reads), count = double(n);
    switch (d->kind) {
    case SIMMR_AP_UNIFORM:                      /* uniform.rs:28-34 */
        for (uint32_t i = 0; i < n; ++i) { out_reads[i] = cast_usize(std::ceil(total / count)); out_abund[i] = 100.0 / count; }
        break;
    case SIMMR_AP_EXACT:                        /* exact.rs:17-24; never size aware (exact.rs:13-15) */
        for (uint32_t i = 0; i < n; ++i) { out_reads[i] = total_reads; out_abund[i] = 100.0 / count; }
        adjust = false;
        break;
    case SIMMR_AP_CUSTOM: {                     /* custom.rs:20-45 */
        if (!d->abundances) { err = "custom abundance profile without abundances"; return SIMMR_E_INVALID; }
        double s = 0.0;
        for (uint32_t i = 0; i < n; ++i) s += d->abundances[i];
        const bool renorm = s < 0.99 || s > 1.01;
        for (uint32_t i = 0; i < n; ++i) {
            const double a = renorm ? d->abundances[i] / s : d->abundances[i];
            out_reads[i] = cast_usize(std::ceil(total * a)); out_abund[i] = a;
        }
        break; }
    default: err = "unknown abundance profile kind"; return SIMMR_E_INVALID;
    }
    if (adjust) {                               /* uniform.rs:60-94, custom.rs:61-95 */
        if (!sizes) { err = "size adjustment needs genome sizes"; return SIMMR_E_INVALID; }
        double reads = 0.0, weight = 0.0;
        for (uint32_t i = 0; i < n; ++i) reads += double(out_reads[i]);
        for (uint32_t i = 0; i < n; ++i) weight += double(sizes[i]) * out_abund[i];
        for (uint32_t i = 0; i < n; ++i) out_reads[i] = cast_usize(std::ceil(reads * ((out_abund[i] * double(sizes[i])) / weight)));
    }
    return SIMMR_OK;
}

/* ====================================================================== */
/* genome file + metadata — files.rs                                        */
/* ====================================================================== */
bool utf8_ok(const uint8_t *s, size_t n) {
    for (size_t i = 0; i < n;) {
        const uint8_t c = s[i];
        if (c < 0x80) { ++i; continue; }
        const int need = (c >= 0xC2 && c <= 0xDF) ? 1 : (c >= 0xE0 && c <= 0xEF) ? 2 : (c >= 0xF0 && c <= 0xF4) ? 3 : -1;
        if (need < 0 || n - i <= size_t(need)) return false;
        uint32_t cp = need == 1 ? (c & 0x1Fu) : need == 2 ? (c & 0x0Fu) : (c & 0x07u);
        for (int k = 1; k <= need; ++k) { if ((s[i + k] & 0xC0) != 0x80) return false; cp = (cp << 6) | (s[i + k] & 0x3Fu); }
        if (need == 2 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return false;
        if (need == 3 && (cp < 0x10000 || cp > 0x10FFFF)) return false;
        i += size_t(need) + 1;
    }
    return true;
}

std::string rust_f64(double v) {   /* Rust `{}`: shortest round-trip, positional */
    if (v != v) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char b[400];
    auto r = std::to_chars(b, b + sizeof b, v, std::chars_format::fixed);
    return std::string(b, r.ptr);
}

namespace {
std::vector<std::string> tab_split(const std::string &s) {
    std::vector<std::string> f; size_t at = 0;
    for (;;) { const size_t t = s.find('\t', at); f.push_back(s.substr(at, t == std::string::npos ? t : t - at)); if (t == std::string::npos) break; at = t + 1; }
    return f;
}
}
/* files.rs:56-86.  is_simple_variant (files.rs:32-45) as shipped classifies
 * every non-trivial first line as the TSV form; the documented rule
 * (readme.md:99-100) is implemented: two or more tab-separated fields on the
 * first line mean a TSV with a header row, otherwise a plain list of paths. */
int64_t parse_genome_file(const char *path, std::string &out, std::string &err) {
    std::ifstream f(path);
    if (!f) { err = "Genome file does not exist"; return SIMMR_E_IO; }
    std::vector<std::string> rows; std::string ln;
    while (std::getline(f, ln)) { if (!ln.empty() && ln.back() == '\r') ln.pop_back(); if (!ln.empty()) rows.push_back(ln); }
    out.clear();
    if (rows.empty()) return 0;
    const auto head = tab_split(rows[0]);
    int64_t n = 0;
    if (head.size() < 2) {
        for (const auto &r : rows) { out += tab_split(r)[0] + "\t\tnan\n"; ++n; }
        return n;
    }
    int cpath = -1, cid = -1, cab = -1;
    for (size_t i = 0; i < head.size(); ++i) {
        if (head[i] == "filepath" || head[i] == "path") cpath = int(i);                       /* files.rs:21 */
        else if (head[i] == "uuid" || head[i] == "id" || head[i] == "genome_id") cid = int(i); /* files.rs:23 */
        else if (head[i] == "abundance") cab = int(i);
    }
    if (cpath < 0) { err = "genome file: no `path`/`filepath` column"; return SIMMR_E_PARSE; }
    for (size_t r = 1; r < rows.size(); ++r) {
        const auto c = tab_split(rows[r]);
        auto cell = [&](int k) { return k >= 0 && size_t(k) < c.size() ? c[size_t(k)] : std::string(); };
        const std::string ab = cell(cab);
        out += cell(cpath) + "\t" + cell(cid) + "\t" + (ab.empty() ? "nan" : ab) + "\n"; ++n;
    }
    return n;
}
/* files.rs:100-134 */
std::string metadata_tsv(uint32_t n, const char *const *uuid, const char *const *path, const uint64_t *reads, const double *abund) {
    std::string s = "genome_id\tfilepath\tnum_reads\tabundance\n";
    for (uint32_t i = 0; i < n; ++i) { s += uuid[i]; s += '\t'; s += path[i]; s += '\t'; s += std::to_string(reads[i]); s += '\t'; s += rust_f64(abund[i]); s += '\n'; }
    return s;
}

/* ====================================================================== */
/* header templates — fastq.rs:34-57                                        */
/* ====================================================================== */
namespace {
std::string subst(const std::string &s, const char *what, const std::string &with) {   /* str::replace */
    std::string r; const size_t wl = strlen(what);
    for (size_t at = 0;;) {
        const size_t hit = s.find(what, at);
        if (hit == std::string::npos) { r.append(s, at, std::string::npos); return r; }
        r.append(s, at, hit - at); r += with; at = hit + wl;
    }
}
}
/* The reference applies seven String::replace calls in a fixed order per record.
 * genome_id and sequence_id are expanded here per contig, honouring that order
 * ({:read_id:} is resolved before {:sequence_id:}, so a placeholder spelled
 * inside a contig id only survives if it is one of the later four).  The last
 * two replacements, {:reverse_complement:} and {:pair:}, take one of two fixed
 * values per mate (fastq.rs:54-57, :103-106), so each contig gets two token
 * lists — entry 2c for mate 1 / long reads ("f", "1"), entry 2c+1 for mate 2
 * ("t", "2") — with those characters folded into the literals.  What is left
 * per record are the three decimal fields. */
int compile_headers(const simmr_genome_set &g, const char *fmt, HeaderPlan &plan, std::string &err) {
    static const struct { const char *text; uint8_t tok; } later[] = {
        {"{:start_position:}", TOK_START}, {"{:end_position:}", TOK_END}, {"{:reverse_complement:}", TOK_RC}, {"{:pair:}", TOK_PAIR}};
    struct Piece { uint8_t kind; std::string text; };
    plan = HeaderPlan();
    const std::string base = fmt ? fmt : DEFAULT_HEADER_FORMAT;
    for (uint32_t gi = 0; gi < g.n_genomes; ++gi) {
        const std::string gid(g.genome_ids + g.genome_id_off[gi], g.genome_id_off[gi + 1] - g.genome_id_off[gi]);
        const std::string with_gid = subst(base, "{:genome_id:}", gid);
        for (uint32_t c = g.genome_contig[gi]; c < g.genome_contig[gi + 1]; ++c) {
            const uint8_t *idp = reinterpret_cast<const uint8_t *>(g.contig_ids) + g.contig_id_off[c];
            const size_t idn = g.contig_id_off[c + 1] - g.contig_id_off[c];
            const std::string sid = utf8_ok(idp, idn) ? std::string(reinterpret_cast<const char *>(idp), idn) : std::string();   /* shared/src/util.rs:49-51 */
            std::vector<Piece> pieces;
            size_t at = 0;
            for (bool more = true; more;) {
                const size_t hit = with_gid.find("{:read_id:}", at);
                std::string seg = with_gid.substr(at, hit == std::string::npos ? hit : hit - at);
                seg = subst(seg, "{:sequence_id:}", sid);
                for (size_t p = 0; p < seg.size();) {      /* left-to-right scan for the four later placeholders */
                    size_t best = std::string::npos; int which = -1;
                    for (int k = 0; k < 4; ++k) { const size_t h = seg.find(later[k].text, p); if (h < best) { best = h; which = k; } }
                    const std::string piece = seg.substr(p, best == std::string::npos ? best : best - p);
                    if (!piece.empty()) pieces.push_back(Piece{TOK_LIT, piece});
                    if (which < 0) break;
                    pieces.push_back(Piece{later[which].tok, std::string()});
                    p = best + strlen(later[which].text);
                }
                if (hit == std::string::npos) more = false;
                else { pieces.push_back(Piece{TOK_READ_ID, std::string()}); at = hit + 11; }
            }
            for (int mate = 0; mate < 2; ++mate) {
                ContigHdr ch{}; ch.tok_off = uint32_t(plan.toks.size());
                size_t lit_total = 0;
                std::string run;                           /* literal text gathered up to the next decimal field */
                /* canonical form: literal, field, literal, ..., field, literal — a literal (possibly empty) on every
                 * even place, so the kernel needs no per-token dispatch */
                auto flush = [&]() {
                    if (run.size() > 65532) return false;
                    plan.pool.resize((plan.pool.size() + 3) & ~size_t(3), '\0');   /* the kernel fetches literals by words */
                    plan.toks.push_back(HdrTok{uint32_t(plan.pool.size()), uint16_t(run.size()), TOK_LIT, 0});
                    plan.pool.append(run);
                    lit_total += run.size(); run.clear();
                    return true;
                };
                for (const Piece &pc : pieces) {
                    if (pc.kind == TOK_LIT) run += pc.text;
                    else if (pc.kind == TOK_RC) run += mate ? 't' : 'f';
                    else if (pc.kind == TOK_PAIR) run += mate ? '2' : '1';
                    else {
                        if (!flush()) { err = "read header too long"; return SIMMR_E_LIMIT; }
                        plan.toks.push_back(HdrTok{0, 0, pc.kind, 0});
                        uint8_t *ctr = pc.kind == TOK_READ_ID ? &ch.n_id : pc.kind == TOK_START ? &ch.n_sp : &ch.n_ep;
                        if (*ctr == 255) { err = "header format repeats a placeholder more than 255 times"; return SIMMR_E_LIMIT; }
                        ++*ctr;
                    }
                }
                if (!flush() || lit_total > 60000 || plan.toks.size() - ch.tok_off > 60000) { err = "read header too long"; return SIMMR_E_LIMIT; }
                ch.ntok = uint16_t(plan.toks.size() - ch.tok_off); ch.lit_bytes = uint16_t(lit_total);
                plan.contigs.push_back(ch);
                const uint32_t cap = uint32_t(lit_total) + 10u * ch.n_id + 20u * ch.n_sp + 20u * ch.n_ep;
                plan.hdr_cap = std::max(plan.hdr_cap, cap);
                uint32_t pos_digits = 1;
                for (uint64_t v = g.contig_size[c]; v >= 10; v /= 10) ++pos_digits;
                plan.hdr_typ = std::max(plan.hdr_typ, uint32_t(lit_total) + 10u * ch.n_id + pos_digits * (ch.n_sp + ch.n_ep));
            }
        }
    }
    return SIMMR_OK;
}

}  // namespace simmr

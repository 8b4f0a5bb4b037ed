"""Shared test helpers: synthetic genomes, synthetic simmrd-style profile blobs
(shared/src/encoding.rs:82-117 bincode layout) and replay draws."""
import math
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import simmr_b200 as S                      # noqa: E402
from oracle import pyoracle as po           # noqa: E402

ORACLE_KIND = {S._native.EP_PERFECT_SHORT: po.PERFECT_SHORT, S._native.EP_MINIMAL_SHORT: po.MINIMAL_SHORT,
               S._native.EP_PERFECT_LONG: po.PERFECT_LONG, S._native.EP_MINIMAL_LONG: po.MINIMAL_LONG,
               S._native.EP_CUSTOM_SHORT: po.CUSTOM_SHORT, S._native.EP_CUSTOM_LONG: po.CUSTOM_LONG}


def oracle_profile(prof):
    """The oracle-side descriptor of a simmr_b200 error profile."""
    blob = None if prof._blob is None else prof._blob.tobytes()
    return po.Profile(ORACLE_KIND[prof.kind], prof.read_length, prof.insert_size, prof.mean_phred_score,
                      prof.flags, blob)


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def synth_genomes(n_genomes, contigs_range=(1, 4), len_range=(2000, 6000), seed=1, with_n=False, long_ids=False):
    """i.i.d. ACGT genomes; ids follow BASELINE.md §4 (g0042_c003, %016x splitmix uuids)."""
    rng = np.random.default_rng(seed)
    genomes = []
    for g in range(n_genomes):
        nc = int(rng.integers(contigs_range[0], contigs_range[1] + 1))
        contigs = []
        for c in range(nc):
            ln = int(rng.integers(len_range[0], len_range[1] + 1))
            seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, ln)]
            if with_n and ln > 100:
                at = int(rng.integers(0, ln - 50))
                seq = seq.copy()
                seq[at:at + 37] = ord("N")
            cid = f"g{g:04d}_c{c:03d}"
            if long_ids:
                cid += " Escherichia coli str. K-12 substr. MG1655, complete genome " + "x" * int(rng.integers(0, 40))
            contigs.append((cid.encode(), seq))
        genomes.append((f"{splitmix64(g):016x}", contigs))
    return S.GenomeSet.from_contigs(genomes)


# ----------------------------------------------------------------------------
# bincode writers
# ----------------------------------------------------------------------------
def _bins(num_bins, width, density, ranges):
    b = struct.pack("<QQQ", num_bins, width, len(density))
    b += struct.pack(f"<{len(density)}d", *density)
    b += struct.pack("<Q", len(ranges))
    for lo, hi in ranges:
        b += struct.pack("<II", lo, hi)
    return b


def make_blob(quality, rl_bins, is_bins=None, kmers=(), kmer_size=7, bin_size=1, is_mean=300.0, is_std=60.0,
              rl_mean=150.0, rl_std=3.0, is_long=False):
    """quality: list of (density list, ranges list) per position; *_bins: (density, ranges)."""
    b = struct.pack("<QQ", bin_size, len(quality))
    for dens, rng_ in quality:
        b += _bins(len(rng_), 1, dens, rng_)
    b += struct.pack("<BQQ", 3, kmer_size, len(kmers))
    for ref, alts in kmers:
        b += struct.pack("<IQ", ref, len(alts))
        for alt, pr in alts:
            b += struct.pack("<If", alt, pr)
    b += struct.pack("<dd", is_mean, is_std)
    if is_bins is None:
        b += struct.pack("<B", 0)
    else:
        b += struct.pack("<B", 1) + _bins(len(is_bins[1]), 10, is_bins[0], is_bins[1])
    b += struct.pack("<dd", rl_mean, rl_std)
    b += _bins(len(rl_bins[1]), 2, rl_bins[0], rl_bins[1])
    b += struct.pack("<B", 1 if is_long else 0)
    return b


def novaseq_blob(npos=151):
    """Synthetic NovaSeq-like profile as simmrd would write it: 71 KDE
    densities and 70 (i,i) ranges per position, mean Q36 decaying to Q30."""
    quality = []
    for p in range(npos):
        mu = 36.0 - 6.0 * p / max(1, npos - 1)
        dens = [math.exp(-0.5 * ((q - mu) / 3.5) ** 2) + (0.02 if q in (2, 11, 25, 37) else 0.0) for q in range(71)]
        quality.append((dens, [(i, i) for i in range(70)]))
    rl = ([0.02, 0.03, 0.05, 0.1, 0.8], [(141, 142), (143, 144), (145, 146), (147, 148), (149, 151)])
    isz = ([math.exp(-0.5 * ((150 + 10 * i + 5 - 350) / 80.0) ** 2) for i in range(45)],
           [(150 + 10 * i, 150 + 10 * i + 9) for i in range(45)])
    return make_blob(quality, rl, isz, is_mean=350.0, is_std=80.0, rl_mean=149.0, rl_std=2.5)


def ont_blob(npos=64):
    quality = []
    for p in range(npos):
        mu = 12.0 + 2.0 * math.sin(p / 9.0)
        dens = [math.exp(-0.5 * ((q - mu) / 4.0) ** 2) for q in range(71)]
        quality.append((dens, [(i, i) for i in range(70)]))
    rl = ([1.0], [(20000, 20000)])
    return make_blob(quality, rl, None, rl_mean=19800.0, rl_std=15000.0, is_long=True)


def kmer_code(s):
    """3-bit code of shared/src/encoding.rs:149-179: base i at bits [3i, 3i+3), A C G T N = 0..4."""
    return sum(b"ACGTN".index(c) << (3 * i) for i, c in enumerate(s))


def ont_kmers(k=7, n_alts=20, seed=8, keep=0.9):
    """An alt-k-mer table as simmrd writes it (simmrd/src/main.rs:315-328): every A,C,G,T k-mer with itself (weight `keep`)
    and n_alts - 1 other k-mers: single substitutions and single / double deletions (N digits, alignment.rs:333-343)."""
    rng = np.random.default_rng(seed)
    out = []
    for v in range(4 ** k):
        digits = [(v >> (2 * i)) & 3 for i in range(k)]
        ref = sum(d << (3 * i) for i, d in enumerate(digits))
        alts, seen = [(ref, keep)], {ref}
        w = rng.random(n_alts - 1)
        w = (1.0 - keep) * w / w.sum()
        for j in range(n_alts - 1):
            for _ in range(8):
                d = list(digits)
                kind = rng.integers(0, 3)
                at = int(rng.integers(0, k))
                if kind == 0:
                    d[at] = (d[at] + 1 + int(rng.integers(0, 3))) & 3
                elif kind == 1:
                    d[at] = 4
                else:
                    d[at] = 4
                    d[int(rng.integers(0, k))] = 4
                c = sum(x << (3 * i) for i, x in enumerate(d))
                if c not in seen:
                    break
            seen.add(c)
            alts.append((c, float(np.float32(w[j]))))
        out.append((ref, alts))
    return out


_ONT_CACHE = {}


def ont_blob_with_kmers(k=5, n_alts=20, npos=64, seed=8):
    """ont_blob plus an alt-k-mer table (~n_alts alternates per k-mer) -> (blob, kmers)."""
    key = (k, n_alts, npos, seed)
    if key not in _ONT_CACHE:
        kmers = ont_kmers(k, n_alts, seed)
        quality = []
        for p in range(npos):
            mu = 12.0 + 2.0 * math.sin(p / 9.0)
            quality.append(([math.exp(-0.5 * ((q - mu) / 4.0) ** 2) for q in range(71)], [(i, i) for i in range(70)]))
        blob = make_blob(quality, ([1.0], [(20000, 20000)]), None, kmers=kmers, kmer_size=k, rl_mean=19800.0, rl_std=15000.0, is_long=True)
        _ONT_CACHE[key] = (blob, kmers)
    return _ONT_CACHE[key]


# ----------------------------------------------------------------------------
# replay draws
# ----------------------------------------------------------------------------
def make_pe_draws(gs, prof, n_units, seed=3, force_cases=True):
    """Random PE draws in the §A.6 schema, exercising the three mate-2 cases."""
    rng = np.random.default_rng(seed)
    R = prof.minimum_genome_size()
    genome = rng.integers(0, gs.n_genomes, n_units).astype(np.uint32)
    ncont = (gs.genome_contig[1:] - gs.genome_contig[:-1]).astype(np.int64)
    contig = (rng.random(n_units) * ncont[genome]).astype(np.uint32)
    cglob = gs.genome_contig[genome].astype(np.int64) + contig
    n = gs.contig_size[cglob].astype(np.int64)
    kind = prof.kind
    if kind == S._native.EP_PERFECT_SHORT:
        rl = np.full(n_units, prof.read_length, np.int64)
        isz = np.full(n_units, prof.insert_size, np.int64)
    else:
        rl = np.clip(np.floor(rng.normal(prof.read_length, 15, n_units)), 1, min(R, 65535)).astype(np.int64)
        isz = np.clip(np.floor(rng.normal(prof.insert_size, 75, n_units)), 0, 65535).astype(np.int64)
    start = (rng.random(n_units) * (n - R)).astype(np.int64)
    if force_cases and n_units >= 4:
        start[0] = 0
        isz[0] = 0            # (fs+IS) - RL < 0  -> (0, RL)
        start[1] = n[1] - R - 1
        isz[1] = min(65535, R + 500)   # out of bounds -> redraw
        start[2] = 0
        isz[2] = rl[2]        # boundary: lo = 0
    redraw = start + (rng.random(n_units) * (n - R - start)).astype(np.int64)
    base_off = np.zeros(n_units + 1, np.uint64)
    base_off[1:] = np.cumsum(2 * rl)
    ns = int(base_off[-1])
    d = dict(genome=genome, contig=contig, read_length=rl.astype(np.uint16), insert_size=isz.astype(np.uint16),
             start=start.astype(np.uint64), redraw=redraw.astype(np.uint64), base_off=base_off)
    d.update(per_base_draws(rng, ns))
    return d


def make_long_draws(gs, prof, n_units, seed=5):
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, gs.n_genomes, n_units).astype(np.uint32)
    ncont = (gs.genome_contig[1:] - gs.genome_contig[:-1]).astype(np.int64)
    contig = (rng.random(n_units) * ncont[genome]).astype(np.uint32)
    cglob = gs.genome_contig[genome].astype(np.int64) + contig
    n = gs.contig_size[cglob].astype(np.int64)
    rl = np.minimum(np.floor(rng.gamma(1.7777779, 11250, n_units)), n - 1).astype(np.int64)
    rl = np.clip(rl, 0, 65535)
    start = (rng.random(n_units) * np.maximum(rl, 1)).astype(np.int64)       # simulate.rs:484 [0, RL)
    end = start + rl
    over = end >= n
    redraw = np.where(over, start + (rng.random(n_units) * (n - start)).astype(np.int64), 0)
    length = np.where(over, redraw - start, rl)
    base_off = np.zeros(n_units + 1, np.uint64)
    base_off[1:] = np.cumsum(np.maximum(length, rl))     # room for RL slots per read
    d = dict(genome=genome, contig=contig, read_length=rl.astype(np.uint16), start=start.astype(np.uint64),
             redraw=redraw.astype(np.uint64), base_off=base_off)
    d.update(per_base_draws(rng, int(base_off[-1])))
    return d


def per_base_draws(rng, ns):
    z = rng.standard_normal(ns).astype(np.float32)
    if ns > 8:
        z[:8] = [-10.0, 10.0, -3.0, 3.0, 22.5, 0.0, -0.05, 7.0]       # saturation / boundary deviates
    return dict(phred_z=z, phred_cand=rng.integers(0, 2**32, ns, dtype=np.uint64).astype(np.uint32),
                phred_u=rng.random(ns), phred_inbin=rng.integers(0, 2**32, ns, dtype=np.uint64).astype(np.uint32),
                mut_u=(rng.integers(0, 1 << 24, ns).astype(np.float32) / np.float32(1 << 24)),
                mut_k=rng.integers(0, 3, ns).astype(np.uint8))


def parse_fastq(data):
    """-> list of (header, seq, qual) byte strings; asserts the 4-line layout."""
    lines = data.split(b"\n")
    assert lines[-1] == b""
    lines = lines[:-1]
    assert len(lines) % 4 == 0
    out = []
    for i in range(0, len(lines), 4):
        assert lines[i + 2] == b"+"
        out.append((lines[i], lines[i + 1], lines[i + 3]))
    return out


def q32_distribution(entries, cb, size=256, chars=True):
    """Law of the short-read sampler over one row of kind-3 words for a uniform draw: the low cb bits of
    the draw's word pick a column, the column's own value is kept iff the draw's 32-bit level < entry, else its alias is
    taken.  Quality tables hold the characters as printed (chars=True: the law is returned over q = character - 33)."""
    cols = 1 << cb
    out = np.zeros(size)
    for c, e in enumerate(int(v) for v in entries):
        keep = e / 2.0 ** 32
        out[e & 0xFF] += keep / cols
        out[(e >> 8) & 0xFF] += (1.0 - keep) / cols
    return np.roll(out, -33) if chars else out


def short_read_law(table):
    """(P(q) over all bases, substitution rate per base) of a kind-3 table: qualities of kept bases from the first
    (1 << cb) words, then — mutating profiles — the gap table (256 words; 255 = no substitution within 255 bases) and
    the quality-given-substitution table."""
    cb, cols = table["cb"], 1 << table["cb"]
    assert table["kind"] == 3 and table["npos"] == 1
    kept = q32_distribution(table["entries"][:cols], cb)
    if len(table["entries"]) == cols:
        return kept, 0.0
    gap = q32_distribution(table["entries"][cols:cols + 256], 8, chars=False)
    sub = q32_distribution(table["entries"][cols + 256:], cb)
    # distance between substituted bases: gap + 1, where 255 means "255 bases kept, draw again"
    if gap[255] >= 1.0:                      # the quantised gap row never ends a gap: no substitutions at all
        return kept, 0.0
    mean_d = (float((gap[:255] * (np.arange(255) + 1)).sum()) + 255.0 * gap[255]) / (1.0 - gap[255])
    rate = 1.0 / mean_d
    return (1.0 - rate) * kept + rate * sub, rate



from simmr_b200.synth import synthetic_genomes as config_genomes, lognormal_abundances    # noqa: E402,F401

"""Native (Philox) mode against the reference's arithmetic, statistically.

The GPU output is byte-identical to the oracle's native mode (tests/test_gpu_parity.py), so the
distributions are checked here on the CPU: draws of the native schedule (oracle) against draws made
with the reference's own formulas and an independent generator (orc_ref_*: f32 `mean + sd*z`, `floor`,
saturating casts, rand_distr's Marsaglia-Tsang gamma, WeightedAliasIndex + Uniform::new_inclusive).
Stated tolerances: two-sample KS with p > 1e-4 at the sample sizes below; chi-square within 6 sigma;
per-genome read counts exact."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

import helpers as H
import simmr_b200 as S
from oracle import pyoracle as po

KS_P = 1e-4


def _genomes(n_g=6, contigs=(2, 5), lens=(30000, 120000), seed=5):
    gs = H.synth_genomes(n_g, contigs, lens, seed=seed)
    return gs, po.Genomes.from_set(gs)


def _u16(fn, n, *args):
    out = np.zeros(n, np.uint16)
    fn(*args, C.c_uint64(n), C.c_uint64(99), out.ctypes.data_as(po.u16p))
    return out


def test_minimal_short_lengths_inserts_and_positions():
    """minimal_short.rs:33-81 (RL = floor N(150,15), IS = floor N(150,75) as u16), simulate.rs:182 (contig uniform
    over the genome's contigs, not length-weighted), :233 (start uniform in [0, n-R))."""
    gs, og = _genomes()
    prof = po.Profile(po.MINIMAL_SHORT, 150, 150, 30)
    reads = [60000] * gs.n_genomes
    geo = po.native_geometry(og, prof, reads, 7)
    n = len(geo["rl"])
    L = po.lib()
    ref_rl = _u16(lambda *a: L.orc_ref_normal_len(C.c_double(150.0), C.c_double(15.0), *a), n)
    ref_is = _u16(lambda *a: L.orc_ref_normal_len(C.c_double(150.0), C.c_double(75.0), *a), n)
    assert stats.ks_2samp(geo["rl"], ref_rl).pvalue > KS_P
    assert stats.ks_2samp(geo["is_"], ref_is).pvalue > KS_P
    assert abs(geo["rl"].mean() - 149.5) < 0.15 and abs(geo["rl"].std() - 15.0) < 0.15
    # contig choice: uniform over the contigs of the unit's genome
    R = prof.min_genome_size()
    for g in range(gs.n_genomes):
        sel = geo["genome"] == g
        c0, c1 = int(gs.genome_contig[g]), int(gs.genome_contig[g + 1])
        counts = np.bincount(geo["contig"][sel] - c0, minlength=c1 - c0)
        assert stats.chisquare(counts).pvalue > KS_P
        # start / (n - R) is uniform on [0, 1)
        span = (gs.contig_size[geo["contig"][sel]] - R).astype(np.float64)
        assert stats.kstest(geo["start"][sel] / span, "uniform").pvalue > KS_P
    # units per genome = reads // 2, in genome order (simulate.rs:135, :179)
    assert np.bincount(geo["genome"]).tolist() == [r // 2 for r in reads]
    assert np.all(np.diff(geo["genome"].astype(np.int64)) >= 0)


def test_long_read_lengths_follow_the_reference_gamma():
    """perfect_long.rs:40-55: Gamma(shape (20000/15000)^2, scale 15000^2/20000) in f32, floor, saturating u16."""
    gs, og = _genomes(3, (1, 2), (120000, 200000), seed=6)
    prof = po.Profile(po.MINIMAL_LONG, mean_phred=30)
    reads = [20000, 20000, 20000]
    geo = po.native_geometry(og, prof, reads, 11)
    n = len(geo["rl"])
    ref = _u16(po.lib().orc_ref_gamma_len, n)
    assert stats.ks_2samp(geo["rl"], ref).pvalue > KS_P
    sat = float((geo["rl"] == 65535).mean())
    assert abs(sat - 0.0142) < 0.003                                   # P(gamma > 65535) = 1.42 %
    span = (gs.contig_size[geo["contig"]] - geo["rl"]).astype(np.float64)
    assert stats.kstest(geo["start"] / span, "uniform").pvalue > KS_P   # start in [0, n - RL)
    assert np.all(geo["lo_or_end"] - geo["start"] == geo["rl"])


def test_perfect_long_phreds():
    """perfect_long.rs:60-78: accuracy ~ N(acc(20), 0.05) capped at 0.9999 -> round(-10 log10(1-a)) in [0, 40]."""
    prof = po.Profile(po.PERFECT_LONG)
    pq, _ = H.short_read_law(prof.phred_table())               # P(q) over all bases: kept and substituted ones
    n = 2_000_000
    q = np.zeros(n, np.uint8)
    po.lib().orc_ref_perfect_long_phred(C.c_uint64(n), C.c_uint64(3), q.ctypes.data_as(po.u8p))
    emp = np.bincount(q, minlength=256) / n
    assert emp[41:].sum() == 0 and pq[41:].sum() == 0
    sel = pq * n >= 20
    chi2 = float((((emp - pq) ** 2)[sel] / pq[sel]).sum() * n)
    assert chi2 < sel.sum() + 6 * np.sqrt(2 * sel.sum()), chi2


@pytest.mark.parametrize("which,blob", [("novaseq", H.novaseq_blob()), ("ont", H.ont_blob())])
def test_custom_profile_tables_follow_the_blob(which, blob):
    """custom_short.rs:75-151, 237-270, 328-352: per-position quality (alias over the 71 densities, uniform inside the
    bin, `as u8`), read-length and insert-size bins — exact table distribution vs draws made the reference's way."""
    kind = po.CUSTOM_SHORT if which == "novaseq" else po.CUSTOM_LONG
    prof = po.Profile(kind, blob=blob)
    t = prof.phred_table()
    assert t["kind"] == 3
    cols, npos = 1 << t["cb"], t["npos"]
    n = 400_000
    for pos in (0, npos // 2, npos - 1):
        e = t["entries"][pos * cols:(pos + 1) * cols]
        pq = H.q32_distribution(e, t["cb"])
        q = np.zeros(n, np.uint8)
        assert po.lib().orc_ref_custom_phred(prof._blob.ctypes.data_as(po.u8p), C.c_uint64(len(blob)), pos, C.c_uint64(n), C.c_uint64(5),
                                             q.ctypes.data_as(po.u8p)) == 0
        emp = np.bincount(q, minlength=256) / n
        sel = pq * n >= 20
        chi2 = float((((emp - pq) ** 2)[sel] / pq[sel]).sum() * n)
        assert chi2 < sel.sum() + 6 * np.sqrt(2 * sel.sum()), (pos, chi2)
        assert abs(emp[~sel].sum() - pq[~sel].sum()) < 2e-3
    if which == "novaseq":
        gs, og = _genomes(4, (1, 3), (20000, 60000), seed=8)
        geo = po.native_geometry(og, prof, [50000] * 4, 21)
        for field, w in (("rl", 0), ("is_", 1)):
            ref = np.zeros(len(geo[field]), np.uint16)
            assert po.lib().orc_ref_custom_len(prof._blob.ctypes.data_as(po.u8p), C.c_uint64(len(blob)), w, C.c_uint64(len(ref)),
                                               C.c_uint64(17), ref.ctypes.data_as(po.u16p)) == 0
            assert stats.ks_2samp(geo[field], ref).pvalue > KS_P, field


def test_abundance_counts_are_exact():
    """North-star item 2: per-genome read counts bit-identical to the reference's rounding — also what the native run
    produces (records per genome in the FASTQ)."""
    gs, og = _genomes(5, (1, 2), (4000, 9000), seed=9)
    ab = np.array([0.4, 0.3, 0.15, 0.1, 0.05])
    reads, _ = S.CustomAbundanceProfile(ab).determine_abundances(10001, 5)
    assert reads.tolist() == [int(np.ceil(10001 * a)) for a in ab]                  # custom.rs:42
    data = po.simulate(og, po.Profile(po.PERFECT_SHORT, 100, 100), reads, 3)
    per_genome = np.zeros(5, np.int64)
    for hdr, _, _ in H.parse_fastq(data):
        per_genome[[gs.uuid(g) for g in range(5)].index(hdr.split(b"|")[1].split(b"/")[0].decode())] += 1
    assert per_genome.tolist() == [2 * (int(r) // 2) for r in reads]                # simulate.rs:179

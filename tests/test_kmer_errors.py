"""k-mer error injection (simulate_errors, custom_long.rs:187-232 == custom_short.rs:452-516): the oracle's native-mode
walk against the restated reference loop with explicit alias picks, the reference's own known answer, and the law of the
picks against the blob's weights."""
import numpy as np

import helpers as H
import simmr_b200 as S
from oracle import pyoracle as po


def _profile(kmers, k):
    quality = [([1.0] * 71, [(i, i) for i in range(70)])]
    blob = H.make_blob(quality, ([1.0], [(20000, 20000)]), None, kmers=kmers, kmer_size=k, rl_mean=19800.0, rl_std=15000.0, is_long=True)
    return po.Profile(po.CUSTOM_LONG, blob=blob)


def test_reference_known_answer_through_the_native_walk():
    # custom_long.rs:299-343: k = 3, ACC -> CAT, ATC -> ATC, TCG -> TGT: "ACCCG" -> "CATGT" (every k-mer has one alternate)
    kmers = [(H.kmer_code(b"ACC"), [(H.kmer_code(b"CAT"), 1.0)]), (H.kmer_code(b"ATC"), [(H.kmer_code(b"ATC"), 1.0)]),
             (H.kmer_code(b"TCG"), [(H.kmer_code(b"TGT"), 1.0)])]
    prof = _profile(kmers, 3)
    for seed in (0, 1, 99):
        out, ch = po.native_errors(prof, seed, 5, b"ACCCG")
        assert out == b"CATGT" and ch.tolist() == [0, 0, 0]
    assert po.simulate_errors(b"ACCCG", 3, kmers, [0, 0, 0]) == b"CATGT"


def test_native_walk_equals_the_reference_loop_with_its_picks():
    blob, kmers = H.ont_blob_with_kmers(k=5, n_alts=20)
    prof = po.Profile(po.CUSTOM_LONG, blob=blob)
    rng = np.random.default_rng(2)
    for unit, n in ((0, 5000), (7, 4), (8, 5), (9, 6), (123456789012, 20000), (3, 0)):
        seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].tobytes()
        if n == 20000:
            seq = seq[:700] + b"NNNN" + seq[704:9000] + b"-" + seq[9001:]             # N is a k-mer digit, '-' cannot be encoded
        out, ch = po.native_errors(prof, 42, unit, seq)
        assert po.simulate_errors(seq, 5, kmers, ch) == out
        if n >= 5000:
            assert len(out) < len(seq) and out != seq[:len(out)]                   # deletions happened
            assert len(ch) >= (n - 4) * 0.85                                        # one pick per window that could be encoded and found, until the shrinking sequence ends the walk
    # different units, different picks; same unit, same picks
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 3000)].tobytes()
    a = po.native_errors(prof, 42, 1, seq)[0]
    assert a != po.native_errors(prof, 42, 2, seq)[0] and a == po.native_errors(prof, 42, 1, seq)[0]


def test_picks_follow_the_blob_weights():
    blob, kmers = H.ont_blob_with_kmers(k=5, n_alts=20)
    prof = po.Profile(po.CUSTOM_LONG, blob=blob)
    ref, alts = kmers[[r for r, _ in kmers].index(H.kmer_code(b"ACGTA"))]
    w = np.array([x for _, x in alts], np.float64)
    n = 400_000
    picks = po.native_error_first_picks(prof, 7, 1000, n, b"ACGTA")      # the first window of reads that start with the k-mer
    counts = np.bincount(picks, minlength=len(alts)).astype(float)
    exp = w / w.sum() * n
    chi = float(((counts - exp) ** 2 / exp).sum())
    assert chi < len(alts) + 6 * np.sqrt(2 * len(alts)), chi


def test_long_read_run_applies_errors_after_the_qualities():
    """simulate.rs:497-503: qualities for the window's length, then errors: the sequence line may be shorter than the
    quality line."""
    blob, kmers = H.ont_blob_with_kmers(k=5, n_alts=20)
    gs = H.synth_genomes(2, (1, 2), (21000, 40000), seed=5)
    og = po.Genomes.from_set(gs)
    prof = po.Profile(po.CUSTOM_LONG, blob=blob)
    data = po.simulate(og, prof, [6, 5], 11)
    recs = H.parse_fastq(data)
    assert len(recs) == 11
    plain = po.Profile(po.CUSTOM_LONG, blob=H.ont_blob())
    base = H.parse_fastq(po.simulate(og, plain, [6, 5], 11))
    for u, ((h, s, q), (h0, s0, q0)) in enumerate(zip(recs, base)):
        assert h == h0 and q == q0 and len(s) <= len(q)                            # same geometry and qualities, edited bases
        assert po.native_errors(prof, 11, u, s0)[0] == s


def test_host_tables_match_the_oracle():
    """The library's host model and the oracle build the alt-k-mer sampling tables independently; they must agree word for
    word (the device samples from the library's, the parity tests compare with the oracle's)."""
    for k, n_alts in ((5, 20), (3, 4), (7, 6)):
        blob, kmers = H.ont_blob_with_kmers(k=k, n_alts=n_alts)
        a, b = S.CustomLongErrorProfile(blob).err_table(), po.Profile(po.CUSTOM_LONG, blob=blob).err_table()
        assert a["k"] == b["k"] == k and a["n_ref"] == b["n_ref"] == 4 ** k
        assert np.array_equal(a["words"], b["words"])
    # a k-mer listed twice keeps its last list; an all-zero list is dropped; a blob without a table has none
    ref = H.kmer_code(b"ACG")
    kmers = [(ref, [(ref, 1.0)]), (H.kmer_code(b"CCC"), [(H.kmer_code(b"CCC"), 0.0)]), (ref, [(H.kmer_code(b"ANG"), 0.25), (ref, 0.75)])]
    blob = H.make_blob([([1.0] * 71, [(i, i) for i in range(70)])], ([1.0], [(20000, 20000)]), None, kmers=kmers, kmer_size=3, is_long=True)
    a, b = S.CustomLongErrorProfile(blob).err_table(), po.Profile(po.CUSTOM_LONG, blob=blob).err_table()
    assert a["n_ref"] == 1 and np.array_equal(a["words"], b["words"]) and a["words"][1] == 0xC0000000       # keep threshold 0.75
    assert S.CustomLongErrorProfile(H.ont_blob()).err_table()["k"] == 0

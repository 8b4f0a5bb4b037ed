"""The reference's own stream plumbing as restated in oracle/ref_stream.inc (ChaCha12 StdRng re-seeded per pair,
ziggurat normal, f32 powf per base) — the algorithm bench.py times as the CPU reference arm.  Checked here: the
ChaCha12 core against the published ECRYPT known answer, the ziggurat's law, and that the stream's FASTQ has the
reference's marginals (simulate.rs:205-302, minimal_short.rs:33-140), which are also the native tables' law."""
import ctypes as C

import numpy as np

import helpers as H
import simmr_b200 as S
from oracle import pyoracle as po


def test_chacha12_known_answer():
    # ECRYPT / eSTREAM ChaCha12, 256-bit all-zero key, zero IV, block 0
    want = bytes.fromhex("9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
                         "0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be")
    key = (C.c_uint32 * 8)()
    out = (C.c_uint32 * 16)()
    po.lib().orc_chacha12_block(key, C.c_uint64(0), out)
    assert bytes(out) == want
    po.lib().orc_chacha12_block(key, C.c_uint64(1), out)
    assert bytes(out) != want


def test_ziggurat_is_standard_normal():
    from scipy import stats
    n = 400_000
    z = np.empty(n)
    po.lib().orc_ref_ziggurat(C.c_uint64(11), C.c_uint64(n), z.ctypes.data_as(C.POINTER(C.c_double)))
    assert stats.kstest(z, "norm").pvalue > 1e-3
    assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.std() - 1) < 0.005 and np.abs(z).max() > 4.0      # the tail branch is reached


def test_reference_stream_marginals_match_the_native_tables():
    gs = H.synth_genomes(2, (1, 3), (40000, 90000), seed=3)
    og = po.Genomes.from_set(gs)
    prof = S.MinimalShortErrorProfile(150, 150, 30)
    data, st = po.ref_stream(og, H.oracle_profile(prof), 60_000, 9)
    recs = H.parse_fastq(data)
    assert len(recs) == 120_000 == st.records
    L = np.array([len(s) for _, s, _ in recs[::2]])
    assert abs(L.mean() - 149.5) < 5 * 15 / np.sqrt(len(L)) and abs(L.std() - 15) < 0.3                # floor N(150, 15)
    assert all(len(s) == len(q) for _, s, q in recs)
    q = np.frombuffer(b"".join(q for _, _, q in recs), np.uint8) - 33
    law, rate = H.short_read_law(prof.phred_table())                                                # the GPU's tables
    hist = np.bincount(q, minlength=256)[:256].astype(float)
    keep = law * len(q) >= 20
    chi = float((((hist - law * len(q)) ** 2)[keep] / (law * len(q))[keep]).sum())
    assert chi < keep.sum() + 6 * np.sqrt(2 * keep.sum()), chi
    # substitutions: mate 1 against its source window
    seq = gs.bases.tobytes()
    coff = {gs.contig_id(c): int(gs.contig_off[c]) for c in range(gs.n_contigs)}
    mm = tot = 0
    for h, s, _ in recs[::2]:
        hd = h.decode()
        o = coff[hd.split("sid=")[1].split("|")[0].encode()] + int(hd.split("|sp=")[1].split("|")[0])
        a, b = np.frombuffer(seq[o:o + len(s)], np.uint8), np.frombuffer(s, np.uint8)
        mm += int((a != b).sum())
        tot += len(s)
    assert abs(mm / tot - rate) < 5 * np.sqrt(rate / tot)                                            # 1.340 % at mean Q30


def test_bench_entry_counts_and_threads():
    gs = H.synth_genomes(3, (1, 2), (30000, 50000), seed=4)
    og, prof = po.Genomes.from_set(gs), H.oracle_profile(S.MinimalShortErrorProfile())
    reads = np.array([4000, 6000, 2000], np.uint64)
    want, st = po.simulate(og, prof, reads, 42, unit_begin=100, unit_end=3100, want_stats=True)
    for threads in (1, 3):
        r = po.bench_pe(og, prof, reads, 42, 100, 3000, 0, threads)
        assert (r["bases"], r["records"], r["bytes"]) == (st.bases, st.records, len(want)) and r["seconds"] > 0
        r = po.bench_pe(og, prof, reads, 42, 100, 3000, 1, threads)
        assert r["records"] == 6000 and abs(r["bases"] / r["records"] - 149.5) < 2

"""The CPU oracle against every known-answer vector the reference's own tests hold for
the read-simulation path.  Vectors: tests/golden/reference_kats.json,
written by tests/golden/make_golden.py from the reference's test sources (each entry
cites file:line).  CPU only."""
import ctypes as C
import json
import os
import struct

import numpy as np
import pytest

import helpers as H
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(GOLD, "reference_kats.json")))


def f32(x):
    return struct.unpack("<f", struct.pack("<f", x))[0]


def test_complement_and_reverse_complement():
    L = po.lib()
    for src, want in KATS["complement"]["cases"]:          # util_tests.rs:8-50, shared/src/util.rs:107-138
        assert bytes(L.orc_complement(b) for b in src.encode()) == want.encode()
        out = np.zeros(len(src), np.uint8)
        a = np.frombuffer(src.encode(), np.uint8).copy()
        L.orc_reverse_complement(a.ctypes.data_as(po.u8p), C.c_uint64(len(a)), out.ctypes.data_as(po.u8p))
        assert out.tobytes() == want.encode()[::-1]         # util.rs:31-37


def test_encode_quality_scores():
    L = po.lib()
    k = KATS["encode_quality_score"]
    for q, ch in k["cases"]:                                # util_tests.rs:69-74
        a, out = np.array([q], np.uint8), np.zeros(1, np.uint8)
        L.orc_encode_quality_scores(a.ctypes.data_as(po.u8p), C.c_uint64(1), out.ctypes.data_as(po.u8p))
        assert out.tobytes() == ch.encode()
    a = np.array(k["vector"][0], np.uint8)                   # util_tests.rs:77-81
    out = np.zeros(len(a), np.uint8)
    L.orc_encode_quality_scores(a.ctypes.data_as(po.u8p), C.c_uint64(len(a)), out.ctypes.data_as(po.u8p))
    assert out.tobytes() == k["vector"][1].encode()


def test_phred_maths_is_exact_in_f32():
    """util_tests.rs:84-109 asserts `==` on f32 values; so does this."""
    L = po.lib()
    k = KATS["phred_f32"]
    for p, q in k["probability_to_phred"]:
        assert L.orc_probability_to_phred(C.c_float(p)) == q
    for q, p in k["phred_to_probability"]:
        assert L.orc_phred_to_probability(q) == f32(p)
    for q, a in k["phred_to_accuracy"]:
        assert L.orc_phred_to_accuracy(q) == f32(a)
    for a, q in k["accuracy_to_phred"]:
        assert L.orc_accuracy_to_phred(C.c_float(a)) == q
    assert L.orc_accuracy_to_phred(C.c_float(0.9999)) == 40   # perfect_long.rs:70-76 cap


def test_genome_from_fasta_fixture():
    k = KATS["genome_from_fasta"]                            # genome_tests.rs:8-20
    g = po.load_fasta(os.path.join(GOLD, k["file"]))
    assert g["num_seqs"] == k["num_seqs"] and g["size"] == k["size"]
    assert [s["id"].decode() for s in g["seqs"]] == k["ids"]
    assert [s["size"] for s in g["seqs"]] == k["sizes"]
    c = po.load_fasta(os.path.join(GOLD, k["file"]), contiguous=True)   # genome.rs:117-136
    assert c["num_seqs"] == 1 and c["seqs"][0]["id"] == b"whole genome"
    assert c["seqs"][0]["size"] == 320 and len(c["seqs"][0]["seq"]) == 322 and c["seqs"][0]["seq"][160:161] == b"N"


def test_perfect_short_profile_and_slices():
    """error_profile_tests.rs:8-21 and the slice / reverse-complement literals of
    simulate_tests.rs:36-45, 82-91 on the reference's own fixture, produced by the oracle's
    replay driver (draws chosen so the windows are the ones the literals name)."""
    k = KATS["slices"]
    g = po.load_fasta(os.path.join(GOLD, k["file"]))
    assert g["num_seqs"] == 1 and g["seqs"][0]["size"] == 7920
    seq = g["seqs"][0]["seq"]
    import simmr_b200 as S
    gs = S.GenomeSet.from_contigs([("u", [(g["seqs"][0]["id"], seq)])])
    og = po.Genomes.from_set(gs)
    prof = po.Profile(po.PERFECT_SHORT, 20, 20)
    assert prof.min_genome_size() == 60 and not prof.is_long()
    for (fs, fe, fwd), (lo, hi, rev) in zip(k["forward"], k["reverse_complement"]):
        assert seq[fs:fe].decode() == fwd
        # mate-2 window [lo, hi) = [fs+IS-RL, fs+IS) (simulate.rs:253-256): IS = hi - fs
        d = po.Draws(genome=[0], contig=[0], read_length=[20], insert_size=[hi - fs], start=[fs], redraw=[0],
                     base_off=[0, 40])
        recs = H.parse_fastq(po.replay(og, prof, d))
        assert recs[0][1].decode() == fwd and recs[1][1].decode() == rev
        assert recs[0][2] == recs[1][2] == bytes([k["quality"] + 33]) * 20
        assert recs[0][0].endswith(f"|sp={fs}|ep={fe}|rc=f".encode()) and recs[1][0].endswith(f"|sp={hi}|ep={lo}|rc=t".encode())
    p = KATS["perfect_short"]
    prof = po.Profile(po.PERFECT_SHORT, p["read_length"], p["insert_size"])
    t = prof.phred_table()
    assert t["kind"] == 0 and t["const_q"] == p["quality"]
    assert prof.len_table(0)["kind"] == 0 and prof.min_genome_size() == 2 * p["read_length"] + p["insert_size"]
    d = po.Draws(genome=[0], contig=[0], read_length=[150], insert_size=[150], start=[100], redraw=[0], base_off=[0, 300])
    recs = H.parse_fastq(po.replay(og, prof, d))
    assert all(len(q) == p["quality_len"] and q == bytes([p["quality"] + 33]) * 150 for _, _, q in recs)


def test_uniform_abundance():
    k = KATS["uniform_abundance"]                            # abundance_profile_tests.rs:8-30
    reads, ab = po.abundance(po.AB_UNIFORM, k["total_reads"], k["genomes"])
    assert reads.tolist() == k["reads"] and ab.tolist() == k["abundance"]


def test_gaussian_kat():
    k = KATS["gaussian"]                                     # custom_short.rs:552-561
    xs = np.array(k["xs"], np.float64)
    assert po.lib().orc_gaussian(k["x"], xs.ctypes.data_as(po.f64p), C.c_uint64(len(xs)), k["bandwidth"]) == k["value"]


def _enc(s, esize):
    out = C.c_uint32()
    a = np.frombuffer(s.encode(), np.uint8).copy()
    assert po.lib().orc_kmer_encode(a.ctypes.data_as(po.u8p), len(a), esize, C.byref(out)) == 0
    return out.value


def test_kmer_codecs():
    L = po.lib()
    for s, code in KATS["two_bit_codec"]["cases"]:           # encoding.rs:288-314
        assert _enc(s, 2) == code
        out = np.zeros(len(s), np.uint8)
        assert L.orc_kmer_decode(code, len(s), 2, 0, out.ctypes.data_as(po.u8p)) == len(s)
        assert out.tobytes().decode() == s
    assert _enc("ACC", 3) == 0x48 and _enc("CAT", 3) == 0xC1
    out = np.zeros(4, np.uint8)
    n = L.orc_kmer_decode(_enc("ANGT", 3), 4, 3, 1, out.ctypes.data_as(po.u8p))     # N = deleted base (custom_long.rs:221)
    assert out[:n].tobytes() == b"AGT"


def test_kmer_splice_kat():
    k = KATS["kmer_splice"]                                  # custom_long.rs:299-343
    refs = sorted(k["map"])
    ref_codes = np.array([_enc(r, 3) for r in refs], np.uint32)
    order = np.argsort(ref_codes)
    ref_codes = ref_codes[order]
    alts, off = [], [0]
    for i in order:
        alts += [_enc(a, 3) for a in k["map"][refs[i]]]
        off.append(len(alts))
    alts, off = np.array(alts, np.uint32), np.array(off, np.uint32)
    src = np.frombuffer(k["input"].encode(), np.uint8).copy()
    choice = np.zeros(16, np.uint32)
    out = np.zeros(64, np.uint8)
    n = po.lib().orc_simulate_errors(src.ctypes.data_as(po.u8p), C.c_uint64(len(src)), 3, ref_codes.ctypes.data_as(po.u32p),
                                     off.ctypes.data_as(po.u32p), alts.ctypes.data_as(po.u32p), len(ref_codes),
                                     choice.ctypes.data_as(po.u32p), C.c_uint64(len(choice)), out.ctypes.data_as(po.u8p),
                                     C.c_uint64(len(out)))
    assert n == len(k["output"]) and out[:n].tobytes().decode() == k["output"]


def test_bytes_to_string_and_default_header():
    buf = C.create_string_buffer(16)
    for s, want in KATS["bytes_to_string"]["cases"]:         # shared/src/util.rs:140-143
        a = np.frombuffer(s.encode(), np.uint8).copy()
        n = po.lib().orc_bytes_to_string(a.ctypes.data_as(po.u8p), C.c_uint64(len(a)), buf)
        assert buf.raw[:n].decode() == want
    import simmr_b200 as S
    assert S.DEFAULT_HEADER_FORMAT == KATS["default_header_format"]["value"]        # cli.rs:196


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10 (Salmon et al., SC'11) — the generator the
    native mode is specified on (not a reference vector: the reference uses ChaCha12)."""
    assert po.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert po.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert po.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_rust_float_formatting():
    """files.rs:128 prints abundances with `{}` (shortest round-trip, never exponent)."""
    assert po.format_f64(1.0) == "1" and po.format_f64(0.1) == "0.1" and po.format_f64(100.0 / 3) == "33.333333333333336"
    assert po.format_f64(1e-7) == "0.0000001" and po.format_f64(20.0) == "20"


@pytest.mark.parametrize("mean_q", [30, 17, 90])       # 90: more than 128 outcomes, 256-column tables (cb == 8)
def test_native_tables_match_reference_arithmetic(mean_q):
    """Native mode samples quantised alias tables.  Their exact distribution must
    match draws made with the reference's own arithmetic (minimal_short.rs:83-140: f32
    mean + 10*z, floor, saturating u8; gen::<f32>() > accuracy) within sampling error."""
    prof = po.Profile(po.MINIMAL_SHORT, 150, 150, mean_q)
    pq, rate = H.short_read_law(prof.phred_table())
    n = 2_000_000
    q = np.zeros(n, np.uint8)
    po.lib().orc_ref_minimal_phred(mean_q, C.c_uint64(n), C.c_uint64(7), q.ctypes.data_as(po.u8p))
    emp = np.bincount(q, minlength=256) / n
    sel = pq * n >= 20
    chi2 = float((((emp - pq) ** 2)[sel] / pq[sel]).sum() * n)
    assert chi2 < sel.sum() + 6 * np.sqrt(2 * sel.sum()), chi2          # chi-square, ~6 sigma
    assert abs(emp[~sel].sum() - pq[~sel].sum()) < 1e-3
    seq = np.frombuffer(b"ACGT", np.uint8)[np.arange(n) % 4].copy()
    out = np.zeros(n, np.uint8)
    po.lib().orc_ref_mutate(seq.ctypes.data_as(po.u8p), q.ctypes.data_as(po.u8p), C.c_uint64(n), C.c_uint64(9),
                            out.ctypes.data_as(po.u8p))
    rate_ref = float((out != seq).mean())
    assert abs(rate_ref - rate) <= 6 * np.sqrt(rate / n) + 2.0 / n, (rate_ref, rate)
    if rate * n < 1000:                                                  # mean 90: a substitution in 10^7 bases, nothing to compare
        return
    if mean_q == 30:
        assert abs(rate - 0.01340) < 2e-4                                # 1.340 %
    # quality of the substituted bases (the reference substitutes after drawing q: P(q | substituted))
    cols = 1 << prof.phred_table()["cb"]
    sub = H.q32_distribution(prof.phred_table()["entries"][cols + 256:], prof.phred_table()["cb"])
    qs = q[out != seq]
    emp_s = np.bincount(qs, minlength=256) / len(qs)
    sel = sub * len(qs) >= 20
    chi2 = float((((emp_s - sub) ** 2)[sel] / sub[sel]).sum() * len(qs))
    assert chi2 < sel.sum() + 6 * np.sqrt(2 * sel.sum()), chi2


def test_native_stream_statistics():
    """The short-read draw schedule (a Philox-seeded MWC64X quality chain and an event chain per record,
    DESIGN.md §4) as it comes out of a native run: Phred histogram against the tables' exact distribution
    (chi-square), substitution rate, no correlation between neighbouring qualities or between the two
    mates, equal use of the three alternative bases, substitutions spread evenly along the read."""
    import simmr_b200 as S
    rng = np.random.default_rng(8)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 400000)]
    gs = S.GenomeSet.from_contigs([("u", [(b"c", seq)])])
    og = po.Genomes.from_set(gs)
    prof = po.Profile(po.MINIMAL_SHORT, 150, 150, 30)
    pq, pm = H.short_read_law(prof.phred_table())
    recs = H.parse_fastq(po.simulate(og, prof, [30000], 1234))
    q = np.frombuffer(b"".join(r[2] for r in recs), np.uint8).astype(np.int64) - 33
    n = len(q)
    assert n > 4_000_000
    emp = np.bincount(q, minlength=256) / n
    sel = pq * n >= 20
    chi2 = float((((emp - pq) ** 2)[sel] / pq[sel]).sum() * n)
    assert chi2 < sel.sum() + 6 * np.sqrt(2 * sel.sum()), chi2
    # substitutions: compare every mate with its source window (mate 2 is reverse-complemented)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    src = seq.tobytes()
    subs = total = 0
    alt = np.zeros(4, np.int64)
    where = np.zeros(3, np.int64)                            # substitutions in the thirds of the window (forward-strand order)
    q_at_sub = []
    for hdr, s, ql in recs[:20000]:
        f = dict(kv.split(b"=") for kv in hdr.split(b"metadata:")[1].split(b"|"))
        sp, ep, rc = int(f[b"sp"]), int(f[b"ep"]), f[b"rc"] == b"t"
        ref = src[ep:sp].translate(comp)[::-1] if rc else src[sp:ep]
        a, b = np.frombuffer(s, np.uint8), np.frombuffer(ref, np.uint8)
        diff = a != b
        subs += int(diff.sum()); total += len(a)
        at = np.nonzero(diff)[0]
        where += np.bincount(((len(a) - 1 - at if rc else at) * 3) // len(a), minlength=3)
        q_at_sub.append(np.frombuffer(ql, np.uint8)[len(a) - 1 - at if rc else at].astype(np.int64) - 33)   # quality i belongs to window base i
        # rank(new) xor rank(old) in 1..3 must be uniform
        rank = np.zeros(256, np.int64); rank[list(b"ACGT")] = range(4)
        ra, rb = rank[a[diff]], rank[b[diff]]
        if rc:
            ra, rb = 3 - ra, 3 - rb
        alt += np.bincount(ra ^ rb, minlength=4)
    assert abs(subs / total - pm) < 6 * np.sqrt(pm / total), (subs / total, pm)
    assert alt[0] == 0 and np.all(np.abs(alt[1:] - subs / 3) < 6 * np.sqrt(subs / 3)), alt
    assert np.all(np.abs(where - subs / 3) < 6 * np.sqrt(subs / 3)), where
    # substituted bases carry qualities drawn from P(q | substituted): far below the read's mean
    t = prof.phred_table(); cols = 1 << t["cb"]
    sub_law = H.q32_distribution(t["entries"][cols + 256:], t["cb"])
    qs = np.concatenate(q_at_sub)
    assert abs(qs.mean() - float((sub_law * np.arange(256)).sum())) < 6 * np.sqrt(float((sub_law * np.arange(256) ** 2).sum()) / len(qs))
    qf = q.astype(np.float64) - q.mean()
    assert abs(float((qf[:-1] * qf[1:]).mean() / qf.var())) < 5 / np.sqrt(n)         # lag-1: next step of the chain
    assert abs(float((qf[:-4] * qf[4:]).mean() / qf.var())) < 5 / np.sqrt(n)         # lag-4
    # the two bases served by one chain step (columns from adjacent bit fields of the word, levels x and c) are
    # independent: joint histogram of (q[2k], q[2k+1]) in quartile bins against the product of its margins
    L0 = np.array([len(r[2]) for r in recs])
    starts = np.concatenate([[0], np.cumsum(L0)[:-1]])
    even = np.concatenate([np.arange(s0, s0 + (l0 & ~1), 2) for s0, l0 in zip(starts[:4000], L0[:4000])])
    edges = np.quantile(q, [0.25, 0.5, 0.75])
    a, b = np.digitize(q[even], edges), np.digitize(q[even + 1], edges)
    joint = np.bincount(4 * a + b, minlength=16).reshape(4, 4).astype(np.float64)
    expect = joint.sum(1)[:, None] * joint.sum(0)[None, :] / joint.sum()
    assert float(((joint - expect) ** 2 / expect).sum()) < 9 + 6 * np.sqrt(18), joint       # chi-square, 9 degrees of freedom
    m1 = np.frombuffer(b"".join(r[2][:100] for r in recs[0::2]), np.uint8).astype(np.float64)
    m2 = np.frombuffer(b"".join(r[2][:100] for r in recs[1::2]), np.uint8).astype(np.float64)
    assert abs(float(np.corrcoef(m1, m2)[0, 1])) < 5 / np.sqrt(len(m1))


def test_native_stream_statistics_long_reads():
    """Long reads are cut into 32 pieces with their own chains: qualities against the tables' law,
    substitution rate, and no seam between pieces (substitutions as frequent in the 64 bases around piece boundaries
    as elsewhere)."""
    import simmr_b200 as S
    rng = np.random.default_rng(12)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 300000)]
    gs = S.GenomeSet.from_contigs([("u", [(b"c", seq)])])
    og = po.Genomes.from_set(gs)
    prof = po.Profile(po.MINIMAL_LONG, mean_phred=20)
    pq, pm = H.short_read_law(prof.phred_table())
    recs = H.parse_fastq(po.simulate(og, prof, [300], 77))
    q = np.frombuffer(b"".join(r[2] for r in recs), np.uint8).astype(np.int64) - 33
    n = len(q)
    assert n > 3_000_000
    emp = np.bincount(q, minlength=256) / n
    sel = pq * n >= 20
    chi2 = float((((emp - pq) ** 2)[sel] / pq[sel]).sum() * n)
    assert chi2 < sel.sum() + 6 * np.sqrt(2 * sel.sum()), chi2
    src = seq.tobytes()
    subs = total = near = near_total = 0
    for hdr, s, _ in recs:
        f = dict(kv.split(b"=") for kv in hdr.split(b"metadata:")[1].split(b"|"))
        sp, ep = int(f[b"sp"]), int(f[b"ep"])
        a, b = np.frombuffer(s, np.uint8), np.frombuffer(src[sp:ep], np.uint8)
        diff = a != b
        subs += int(diff.sum()); total += len(a)
        piece = 16 * ((len(a) + 511) // 512)
        if piece >= 128:
            pos = np.arange(len(a)) % piece
            edge = (pos < 32) | (pos >= piece - 32)
            near += int(diff[edge].sum()); near_total += int(edge.sum())
    assert abs(subs / total - pm) < 6 * np.sqrt(pm / total), (subs / total, pm)
    assert abs(near / near_total - pm) < 6 * np.sqrt(pm / near_total), (near / near_total, pm)

"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU
oracle on the same seeded inputs.  Bar: byte-identical FASTQ."""
import numpy as np
import pytest

import helpers as H
import simmr_b200 as S
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def first_diff(a, b):
    n = min(len(a), len(b))
    x = np.frombuffer(a[:n], np.uint8) != np.frombuffer(b[:n], np.uint8)
    i = int(np.argmax(x)) if x.any() else n
    lo = max(0, i - 200)
    return f"len {len(a)} vs {len(b)}; first difference at byte {i}\nGPU   : {a[lo:i + 80]!r}\nORACLE: {b[lo:i + 80]!r}"


def check_native(engine, gs, prof, reads, seed, **kw):
    engine.load_genomes(gs)
    engine.set_error_profile(prof)
    got = engine.simulate_bytes(reads, seed, **kw)
    okw = {k: v for k, v in kw.items() if k in ("header_format", "unit_begin", "unit_end", "read_id_base")}
    want, st = po.simulate(po.Genomes.from_set(gs), H.oracle_profile(prof), reads, seed, want_stats=True, **okw)
    assert got == want, first_diff(got, want)
    s = engine.stats()
    assert s["records"] == st.records and s["bases"] == st.bases and s["bytes"] == st.bytes == len(got)
    assert s["clamped"] == st.clamped
    return got


def test_c1_perfect_short_fixture_genome(engine):
    """BASELINE config C1 shape on the reference's own 7 920 bp fixture is not possible here
    (/root/reference is absent on the GPU box), so the same geometry runs on a synthetic contig."""
    gs = H.synth_genomes(1, (1, 1), (50000, 50000), seed=11)
    data = check_native(engine, gs, S.PerfectShortErrorProfile(150, 150), [5000], 42)
    recs = H.parse_fastq(data)
    assert len(recs) == 5000
    assert all(q == b"]" * 150 for _, _, q in recs)           # Q60 + 33
    # with RL == IS mate 2 is the reverse complement of mate 1
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for i in range(0, 200, 2):
        assert recs[i + 1][1] == recs[i][1].translate(comp)[::-1]


@pytest.mark.parametrize("seed", [0, 42, 2**63 + 12345])
def test_minimal_short_multi_genome(engine, seed):
    gs = H.synth_genomes(7, (1, 5), (1500, 9000), seed=2, with_n=True)
    check_native(engine, gs, S.MinimalShortErrorProfile(150, 150, 30), [2001, 0, 1, 640, 2, 3000, 999], seed)


def test_minimal_short_other_params(engine):
    gs = H.synth_genomes(3, (2, 3), (4000, 9000), seed=5)
    check_native(engine, gs, S.MinimalShortErrorProfile(100, 400, 17), [3000, 2000, 1000], 7)
    check_native(engine, gs, S.MinimalShortErrorProfile(250, 50, 40), [500, 500, 500], 8)


def test_batching_and_sharding_do_not_change_bytes(engine):
    gs = H.synth_genomes(4, (1, 3), (3000, 8000), seed=9)
    prof = S.MinimalShortErrorProfile()
    reads = [1500, 700, 0, 2200]
    whole = check_native(engine, gs, prof, reads, 99)
    small = engine.simulate_bytes(reads, 99, batch_units=257)
    assert small == whole
    total = engine.total_units(reads)
    parts = [engine.simulate_bytes(reads, 99, unit_begin=a, unit_end=b) for a, b in
             [(0, 1), (1, total // 3), (total // 3, total // 3), (total // 3, total - 1), (total - 1, total)] if b > a]
    assert b"".join(parts) == whole


def test_header_format_and_ids(engine):
    gs = H.synth_genomes(2, (2, 2), (3000, 4000), seed=4, long_ids=True)
    prof = S.PerfectShortErrorProfile(75, 200)
    for fmt in ["@{:read_id:}/{:pair:}", "@r{:read_id:}-{:read_id:} {:pair:}{:pair:} g={:genome_id:} rc={:reverse_complement:} "
                "{:end_position:}..{:start_position:} {:sequence_id:} literal{:nothing:}", "plain"]:
        check_native(engine, gs, prof, [300, 200], 5, header_format=fmt, read_id_base=4294967000)


def test_placeholder_inside_ids(engine):
    """fastq.rs:34-57 replaces sequentially: a placeholder spelled inside an id is only
    expanded if its replace() comes later in the chain."""
    rng = np.random.default_rng(0)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 3000)]
    gs = S.GenomeSet.from_contigs([("G{:read_id:}|{:sequence_id:}", [(b"c{:read_id:}{:pair:}{:start_position:}{:genome_id:}", seq)]),
                                   ("u2", [(b"bad\xff\xfeutf8", seq), (b"ok\xc3\xa9", seq)])])
    check_native(engine, gs, S.PerfectShortErrorProfile(50, 120), [40, 60], 3)


def test_replay_perfect_and_minimal_short(engine):
    gs = H.synth_genomes(3, (1, 3), (3000, 7000), seed=21)
    og = po.Genomes.from_set(gs)
    for prof in (S.PerfectShortErrorProfile(150, 150), S.MinimalShortErrorProfile(150, 150, 30), S.MinimalShortErrorProfile(90, 260, 8)):
        d = H.make_pe_draws(gs, prof, 400, seed=13)
        engine.load_genomes(gs)
        engine.set_error_profile(prof)
        got = engine.replay_bytes(S.Draws(**d))
        want = po.replay(og, H.oracle_profile(prof), po.Draws(**d))
        assert got == want, first_diff(got, want)
        assert engine.replay_bytes(S.Draws(**d), batch_units=37) == want


def test_custom_short_novaseq_native_and_replay(engine):
    """NovaSeq-style empirical profile (simmrd blob layout): per-position alias tables, bins for RL/IS,
    no substitutions (custom_short.rs:522-529)."""
    gs = H.synth_genomes(4, (1, 3), (2500, 9000), seed=31, with_n=True)
    prof = S.CustomShortErrorProfile(H.novaseq_blob())
    assert prof.minimum_genome_size() == 648 and not prof.is_long_read()
    check_native(engine, gs, prof, [1500, 900, 2, 1200], 17)
    og = po.Genomes.from_set(gs)
    d = H.make_pe_draws(gs, prof, 300, seed=5)
    d["read_length"] = np.clip(d["read_length"], 100, 151).astype(np.uint16)
    bo = np.zeros(301, np.uint64)
    bo[1:] = np.cumsum(2 * d["read_length"].astype(np.int64))
    d["base_off"] = bo
    d["phred_u"] = d["phred_u"] * 8.0          # spans [0, weight_sum) of the un-normalised densities
    got = engine.replay_bytes(S.Draws(**d))
    want = po.replay(og, H.oracle_profile(prof), po.Draws(**d))
    assert got == want, first_diff(got, want)


@pytest.mark.parametrize("kind", ["minimal", "perfect", "custom"])
@pytest.mark.parametrize("compat", [False, True])
def test_long_reads_native(engine, kind, compat):
    """simulate.rs:323-406, 464-523 with gamma lengths; compat = start drawn in [0, RL) (simulate.rs:484)."""
    gs = H.synth_genomes(3, (1, 3), (21000, 90000), seed=41, with_n=True)
    prof = {"minimal": S.MinimalLongErrorProfile(30, compat_start=compat), "perfect": S.PerfectLongErrorProfile(compat_start=compat),
            "custom": S.CustomLongErrorProfile(H.ont_blob(), compat_start=compat)}[kind]
    assert prof.is_long_read() and prof.minimum_genome_size() == 20000
    data = check_native(engine, gs, prof, [40, 25, 31], 1234)
    recs = H.parse_fastq(data)
    assert len(recs) == 96 and all(len(s) == len(q) for _, s, q in recs)


def test_long_reads_replay(engine):
    gs = H.synth_genomes(2, (1, 2), (21000, 60000), seed=43)
    og = po.Genomes.from_set(gs)
    for prof in (S.MinimalLongErrorProfile(20), S.PerfectLongErrorProfile()):
        d = H.make_long_draws(gs, prof, 30, seed=9)
        engine.load_genomes(gs)
        engine.set_error_profile(prof)
        got = engine.replay_bytes(S.Draws(**d))
        want = po.replay(og, H.oracle_profile(prof), po.Draws(**d))
        assert got == want, first_diff(got, want)


def test_too_small_contig_is_an_error(engine):
    gs = H.synth_genomes(1, (1, 1), (300, 300), seed=1)
    engine.load_genomes(gs)
    engine.set_error_profile(S.PerfectShortErrorProfile(150, 150))
    with pytest.raises(S.SimmrError) as e:
        engine.simulate_bytes([10], 1)
    assert e.value.code == S._native.E_TOO_SMALL


def test_medium_scale_minimal_short(engine):
    """~1.2 M records through several batches, compared byte for byte."""
    gs = H.synth_genomes(6, (1, 6), (20000, 120000), seed=77)
    check_native(engine, gs, S.MinimalShortErrorProfile(), [200000] * 6, 2024, batch_units=150001)

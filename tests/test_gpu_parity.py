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


@pytest.mark.parametrize("mean_q", [2, 55, 90])
def test_quality_table_shapes(engine, mean_q):
    """Mean qualities that change the sampling tables' shape: mean 2 puts a third of the bases on substitution events
    (dense gap chain, many events per record); mean 55 and 90 need 256 columns (the general column path of the lane
    kernel instead of the cb == 7 one) and make substitutions rare (long "none within 255" chains)."""
    gs = H.synth_genomes(3, (1, 3), (3000, 9000), seed=15)
    prof = S.MinimalShortErrorProfile(150, 150, mean_q)
    t = prof.phred_table()
    assert t["kind"] == 3 and t["cb"] == (7 if mean_q == 2 else 8)
    data = check_native(engine, gs, prof, [1500, 900, 601], 21)
    _, rate = H.short_read_law(t)
    # substitution rate of the run against the tables' law (every record against its source window is checked in
    # test_native_stream_statistics for the oracle; here the two implementations already agree byte for byte)
    assert (rate > 0.25) if mean_q == 2 else (rate < 1e-3)
    gl = H.synth_genomes(2, (1, 1), (30000, 70000), seed=16)
    check_native(engine, gl, S.MinimalLongErrorProfile(mean_q), [12, 9], 22)


def test_single_stream_switch_gives_the_same_bytes(engine, monkeypatch):
    """SIMMR_PREP_SERIAL puts batch preparation and formatting on one stream (bench.py uses it to time the format kernel
    without co-runners); the two-stream default must produce the same bytes over many small batches."""
    gs = H.synth_genomes(4, (1, 3), (3000, 8000), seed=19)
    engine.load_genomes(gs); engine.set_error_profile(S.MinimalShortErrorProfile())
    reads = [1500, 700, 0, 2200]
    two = engine.simulate_bytes(reads, 5, batch_units=101)
    monkeypatch.setenv("SIMMR_PREP_SERIAL", "1")
    one = engine.simulate_bytes(reads, 5, batch_units=101)
    assert one == two == po.simulate(po.Genomes.from_set(gs), H.oracle_profile(S.MinimalShortErrorProfile()), reads, 5)


def test_long_read_piece_boundaries(engine):
    """Long reads are cut into 32 pieces of 16 * ceil(L / 512) bases: compat mode (simulate.rs:484-489)
    produces short and empty long reads, so lanes with empty, partial and full pieces all occur."""
    gs = H.synth_genomes(2, (1, 2), (21000, 40000), seed=23)
    data = check_native(engine, gs, S.MinimalLongErrorProfile(25, compat_start=True), [150, 90], 31)
    lens = sorted(len(s) for _, s, _ in H.parse_fastq(data))
    assert lens[0] < 1024 and lens[-1] > 16000        # a 518-base read: 17 pieces of 32 bases, 15 lanes without a piece


def test_long_profile_too_large_for_shared_memory(engine):
    """A long-read profile with 700 quality positions (358 KB of tables) does not fit next to the long kernel's rings and
    tiles: the run takes the byte path (one CTA per read, a thread per piece) and still matches the oracle."""
    gs = H.synth_genomes(2, (1, 1), (30000, 60000), seed=29)
    check_native(engine, gs, S.CustomLongErrorProfile(H.ont_blob(npos=700)), [9, 7], 41)


def test_batching_and_sharding_do_not_change_bytes(engine):
    gs = H.synth_genomes(4, (1, 3), (3000, 8000), seed=9)
    prof = S.MinimalShortErrorProfile()
    reads = [1500, 700, 0, 2200]
    whole = check_native(engine, gs, prof, reads, 99)
    small = engine.simulate_bytes(reads, 99, batch_units=257)
    assert small == whole
    total = engine.total_units(reads)
    parts = [engine.simulate_bytes(reads, 99, unit_begin=a, unit_end=b) for a, b in
             [(0, 0), (0, 1), (1, total // 3), (total // 3, total // 3), (total // 3, total - 1), (total - 1, total)]]
    assert parts[0] == b"" and parts[3] == b""            # an explicit empty range is empty, also at unit 0 (an empty first shard)
    assert b"".join(parts) == whole
    assert engine.measure(reads, 99, unit_begin=0, unit_end=0) == (0, 0)


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


@pytest.mark.parametrize("k,n_alts", [(5, 20), (3, 4), (7, 6)])
def test_long_reads_kmer_error_injection(engine, k, n_alts):
    """custom-long blobs with an alt-k-mer table: simulate_errors (custom_long.rs:187-232) on the device, between the
    qualities and the (identity) point mutations as simulate.rs:497-503 orders them.  Deletions make the sequence line
    shorter than the quality line.  Genomes with N runs take the byte path (N is a k-mer digit), compat mode adds short
    and empty reads, small batches put seams between the sizing, editing and formatting passes."""
    blob, kmers = H.ont_blob_with_kmers(k=k, n_alts=n_alts)
    prof = S.CustomLongErrorProfile(blob)
    gs = H.synth_genomes(3, (1, 3), (21000, 90000), seed=41, with_n=True)
    data = check_native(engine, gs, prof, [40, 25, 31], 1234)
    recs = H.parse_fastq(data)
    assert len(recs) == 96 and all(len(s) <= len(q) for _, s, q in recs) and any(len(s) < len(q) for _, s, q in recs)
    # every record against the reference loop itself: window from the header, the oracle's picks, orc_simulate_errors
    oprof = H.oracle_profile(prof)
    coff = {gs.contig_id(c): int(gs.contig_off[c]) for c in range(gs.n_contigs)}
    seq = gs.bases.tobytes()
    for u, (h, s, q) in enumerate(recs[:20]):
        hd = h.decode()
        o = coff[hd.split("sid=")[1].split("|")[0].encode()]
        sp, ep = int(hd.split("|sp=")[1].split("|")[0]), int(hd.split("|ep=")[1].split("|")[0])
        window = seq[o + sp:o + ep]
        assert len(q) == len(window)
        edited, picks = po.native_errors(oprof, 1234, u, window)
        assert s == edited == po.simulate_errors(window, k, kmers, picks)
    assert engine.simulate_bytes([40, 25, 31], 1234, batch_units=7) == data
    check_native(engine, gs, S.CustomLongErrorProfile(blob, compat_start=True), [150, 90, 3], 31, batch_units=64)


def test_kmer_errors_reference_known_answer_on_the_device(engine):
    """The reference's own known answer (custom_long.rs:299-343: k = 3, ACC -> CAT, ATC -> ATC, TCG -> TGT turns ACCCG
    into CATGT) through the device: every listed k-mer has a single alternate, so a native run is deterministic and equals
    the reference loop; a contig made of ACCCG repeats shows the literal."""
    kmers = [(H.kmer_code(b"ACC"), [(H.kmer_code(b"CAT"), 1.0)]), (H.kmer_code(b"ATC"), [(H.kmer_code(b"ATC"), 1.0)]),
             (H.kmer_code(b"TCG"), [(H.kmer_code(b"TGT"), 1.0)])]
    quality = [([1.0] * 71, [(i, i) for i in range(70)])]
    blob = H.make_blob(quality, ([1.0], [(20000, 20000)]), None, kmers=kmers, kmer_size=3, rl_mean=19800.0, rl_std=15000.0, is_long=True)
    prof = S.CustomLongErrorProfile(blob)
    rng = np.random.default_rng(5)
    rand = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 30000)]
    rep = np.frombuffer(b"ACCCG" * 6000, np.uint8)
    gs = S.GenomeSet.from_contigs([("kat", [(b"random", rand), (b"repeat", rep)])])
    data = check_native(engine, gs, prof, [60], 3)
    assert po.simulate_errors(b"ACCCG", 3, kmers, [0, 0, 0]) == b"CATGT"
    seen = 0
    for h, s, q in H.parse_fastq(data):
        hd = h.decode()
        sp, ep = int(hd.split("|sp=")[1].split("|")[0]), int(hd.split("|ep=")[1].split("|")[0])
        src = (rep if "sid=repeat" in hd else rand).tobytes()[sp:ep]
        want = po.simulate_errors(src, 3, kmers, [0] * len(src))
        assert s == want and len(s) == len(q)
        if "sid=repeat" in hd and sp % 5 == 0 and len(src) >= 5:
            assert s[:5] == b"CATGT"
            seen += 1
    assert seen > 0


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


def test_malformed_tables_are_refused(engine):
    """Inputs the kernels would index with unchecked: a genome_contig table that runs backwards, a contig larger than
    its span, replay draws with fewer slots than bases."""
    gs = H.synth_genomes(3, (2, 2), (3000, 4000), seed=5)
    bad = S.GenomeSet(gs.bases, gs.contig_off, gs.contig_size, [0, 4, 2, 6], gs.contig_ids, gs.contig_id_off, gs.genome_ids, gs.genome_id_off)
    with pytest.raises(S.SimmrError) as e:
        engine.load_genomes(bad)
    assert e.value.code == S._native.E_INVALID
    sizes = gs.contig_size.copy(); sizes[1] += 1
    bad = S.GenomeSet(gs.bases, gs.contig_off, sizes, gs.genome_contig, gs.contig_ids, gs.contig_id_off, gs.genome_ids, gs.genome_id_off)
    with pytest.raises(S.SimmrError) as e:
        engine.load_genomes(bad)
    assert e.value.code == S._native.E_INVALID
    prof = S.MinimalShortErrorProfile(150, 150, 30)
    engine.load_genomes(gs)
    engine.set_error_profile(prof)
    d = H.make_pe_draws(gs, prof, 50, seed=3)
    d["base_off"] = d["base_off"].copy(); d["base_off"][10:] -= 7          # unit 9 is left with too few slots
    with pytest.raises(S.SimmrError) as e:
        engine.replay_bytes(S.Draws(**d))
    assert e.value.code == S._native.E_INVALID


def test_a_batch_that_overflows_its_buffer_is_refused(engine, monkeypatch):
    """Long-read batches are budgeted at the typical read size, not the largest: should a batch's bytes exceed its
    staging buffer after all, the kernels write nothing and the run fails with SIMMR_E_LIMIT (here the buffer is made
    too small on purpose), with and without a sink, and for the k-mer error walk's scratch copies too."""
    gs = H.synth_genomes(2, (1, 2), (80000, 90000), seed=17)
    engine.load_genomes(gs)
    monkeypatch.setenv("SIMMR_DEBUG_CAP", "100000")
    for prof in (S.MinimalLongErrorProfile(30), S.CustomLongErrorProfile(H.ont_blob_with_kmers(k=5, n_alts=6)[0])):
        engine.set_error_profile(prof)
        with pytest.raises(S.SimmrError) as e:
            engine.simulate([40, 40], 3)
        assert e.value.code == S._native.E_LIMIT
        with pytest.raises(S.SimmrError) as e:
            engine.simulate_bytes([40, 40], 3)
        assert e.value.code == S._native.E_LIMIT
    monkeypatch.delenv("SIMMR_DEBUG_CAP")
    engine.set_error_profile(S.MinimalLongErrorProfile(30))
    assert len(engine.simulate_bytes([40, 40], 3)) > 100000


def test_medium_scale_minimal_short(engine):
    """~1.2 M records through several batches, compared byte for byte."""
    gs = H.synth_genomes(6, (1, 6), (20000, 120000), seed=77)
    check_native(engine, gs, S.MinimalShortErrorProfile(), [200000] * 6, 2024, batch_units=150001)


def test_measure_gives_shard_offsets(engine):
    """simmr_measure (geometry kernel only) = the per-shard byte/record counts of the host-side exchange."""
    gs = H.synth_genomes(4, (1, 3), (3000, 8000), seed=9)
    prof = S.MinimalShortErrorProfile()
    reads = [1500, 700, 0, 2200]
    whole = check_native(engine, gs, prof, reads, 5)
    total = engine.total_units(reads)
    assert engine.measure(reads, 5) == (len(whole), 2 * total)
    at = 0
    from simmr_b200 import sharding
    for k in range(3):
        ub, ue = sharding.shard_range(total, k, 3)
        nb, nr = engine.measure(reads, 5, unit_begin=ub, unit_end=ue)
        assert nr == 2 * (ue - ub)
        assert engine.simulate_bytes(reads, 5, unit_begin=ub, unit_end=ue) == whole[at:at + nb]
        at += nb
    assert at == len(whole)


def _write_fasta(path, contigs):
    with open(path, "wb") as f:
        for cid, seq in contigs:
            f.write(b">" + cid + b"\n")
            s = bytes(seq)
            for i in range(0, len(s), 80):
                f.write(s[i:i + 80] + b"\n")


@pytest.mark.parametrize("mode", ["tsv-custom", "list-uniform", "long"])
def test_cli_binary_end_to_end(engine, tmp_path, mode):
    """The `simmr` binary (flags of cli.rs:93-220, flow of main.rs:20-266) against the oracle: FASTQ bytes
    and <output>.tsv, through FASTA files, a genome file, the contig filter and the abundance profile."""
    import os
    import subprocess
    rng = np.random.default_rng(12)
    n_g = 3
    lens = (25000, 60000) if mode == "long" else (400, 5000)
    paths, uuids, ab = [], [], [0.5, 0.3, 0.2]
    for g in range(n_g):
        contigs = [(f"g{g}_c{c} some text".encode(), np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(rng.integers(*lens)))])
                   for c in range(3)]
        if mode != "long":
            contigs.append((b"tiny", np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 200)]))      # dropped by the filter (main.rs:124)
        p = str(tmp_path / f"g{g}.fna")
        _write_fasta(p, contigs)
        paths.append(p)
        uuids.append(f"uuid-{g:02d}")
    out = str(tmp_path / "reads.fq")
    exe = os.path.join(os.path.dirname(S._native.LIB_PATH), "simmr")
    if mode == "tsv-custom":
        gf = tmp_path / "genomes.tsv"
        gf.write_text("uuid\tabundance\tpath\n" + "".join(f"{u}\t{a}\t{p}\n" for u, a, p in zip(uuids, ab, paths)))
        prof, aprof = S.MinimalShortErrorProfile(100, 220, 25), S.CustomAbundanceProfile(ab, size_adjusted=True)
        cmd = [exe, "--genome-file", str(gf), "--output", out, "--num-reads", "3001", "--seed", "77", "--error-profile", "minimal-short",
               "--read-length", "100", "--insert-size", "220", "--mean-phred-score", "25", "--abundance-profile", "custom", "--size-adjusted"]
        fmt = None
    elif mode == "list-uniform":
        gf = tmp_path / "genomes.txt"
        gf.write_text("".join(p + "\n" for p in paths))
        prof, aprof = S.PerfectShortErrorProfile(150, 150), S.UniformAbundanceProfile()
        fmt = "@{:read_id:}/{:pair:} {:genome_id:}"
        cmd = [exe, "--genome-file", str(gf), "--output", out, "--num-reads=2000", "--seed", "5", "--read-header-format", fmt]
        uuids = None
    else:
        prof, aprof = S.MinimalLongErrorProfile(20), S.ExactAbundanceProfile()
        cmd = [exe] + sum([["--genome", p] for p in paths], []) + ["--output", out, "--num-reads", "25", "--seed", "9", "--error-profile",
                                                                      "minimal-long", "--mean-phred-score", "20", "--abundance-profile", "exact"]
        fmt, uuids = None, None
    open(out, "w").write("stale")                       # existing outputs are replaced (main.rs:191-198)
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    tsv = [ln.split("\t") for ln in open(out + ".tsv").read().splitlines()]
    assert tsv[0] == ["genome_id", "filepath", "num_reads", "abundance"] and [t[1] for t in tsv[1:]] == paths
    got_ids = [t[0] for t in tsv[1:]]
    gs, kept = S.GenomeSet.from_fasta(paths, uuids=uuids if uuids else got_ids, min_size=prof.minimum_genome_size())
    assert kept.all() and (uuids is None or got_ids == uuids)
    reads, abund = aprof.reads_per_genome({"tsv-custom": 3001, "list-uniform": 2000, "long": 25}[mode], gs.genome_sizes)
    assert [int(t[2]) for t in tsv[1:]] == reads.tolist()
    assert [t[3] for t in tsv[1:]] == [po.format_f64(a) for a in abund]
    seed = {"tsv-custom": 77, "list-uniform": 5, "long": 9}[mode]
    want = po.simulate(po.Genomes.from_set(gs), H.oracle_profile(prof), reads, seed, header_format=fmt)
    got = open(out, "rb").read()
    assert got == want, first_diff(got, want)
    bad = subprocess.run([exe, "--genome", paths[0], "--output", out, "--error-profile", "nope"], capture_output=True, text=True)
    assert bad.returncode == 1 and "invalid value" in bad.stderr


def test_cli_output_parts(engine, tmp_path):
    """--output-parts P: chunk j of the run goes to part j mod P, every part written by its own I/O thread from pinned
    buffers the sink retains (SIMMR_SINK_RETAINED / simmr_chunk_release)."""
    import os
    import subprocess
    rng = np.random.default_rng(8)
    paths = []
    for g in range(2):
        p = str(tmp_path / f"g{g}.fna")
        _write_fasta(p, [(f"g{g}_c{c}".encode(), np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(rng.integers(3000, 9000)))]) for c in range(2)])
        paths.append(p)
    out = str(tmp_path / "reads.fq")
    exe = os.path.join(os.path.dirname(S._native.LIB_PATH), "simmr")
    cmd = [exe] + sum([["--genome", p] for p in paths], []) + ["--output", out, "--num-reads", "5000", "--seed", "4", "--error-profile",
                                                                  "minimal-short", "--output-parts", "3", "--batch-units", "100"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert not os.path.exists(out)
    ids = [ln.split("\t")[0] for ln in open(out + ".tsv").read().splitlines()[1:]]
    prof = S.MinimalShortErrorProfile()
    gs, _ = S.GenomeSet.from_fasta(paths, uuids=ids, min_size=prof.minimum_genome_size())
    reads, _ = S.UniformAbundanceProfile().determine_abundances(5000, 2)
    og, oprof = po.Genomes.from_set(gs), H.oracle_profile(prof)
    total = po.total_units(oprof, reads)
    want = [b"", b"", b""]
    for j, ub in enumerate(range(0, total, 100)):
        want[j % 3] += po.simulate(og, oprof, reads, 4, unit_begin=ub, unit_end=min(total, ub + 100))
    for p in range(3):
        got = open(f"{out}.{p:03d}", "rb").read()
        assert got == want[p], first_diff(got, want[p])


def test_cli_gzip_and_split_mates(engine, tmp_path):
    """--gzip --split-mates: the two mates as gzip files whose plain bytes are the records of the interleaved output
    (readme.md:209-211 lists split output as to-do); --gzip alone reads back as the plain file."""
    import gzip
    import os
    import subprocess
    rng = np.random.default_rng(12)
    paths = []
    for g in range(2):
        p = str(tmp_path / f"g{g}.fna")
        _write_fasta(p, [(f"g{g}_c{c}".encode(), np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(rng.integers(3000, 9000)))]) for c in range(2)])
        paths.append(p)
    exe = os.path.join(os.path.dirname(S._native.LIB_PATH), "simmr")
    base = [exe] + sum([["--genome", p] for p in paths], []) + ["--num-reads", "6000", "--seed", "4", "--error-profile", "minimal-short", "--batch-units", "700"]
    out = str(tmp_path / "reads.fq.gz")
    r = subprocess.run(base + ["--output", out, "--gzip", "--split-mates"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ids = [ln.split("\t")[0] for ln in open(out + ".tsv").read().splitlines()[1:]]
    prof = S.MinimalShortErrorProfile()
    gs, _ = S.GenomeSet.from_fasta(paths, uuids=ids, min_size=prof.minimum_genome_size())
    reads, _ = S.UniformAbundanceProfile().determine_abundances(6000, 2)
    want = po.simulate(po.Genomes.from_set(gs), H.oracle_profile(prof), reads, 4)
    lines = want.split(b"\n")[:-1]
    recs = [b"\n".join(lines[i:i + 4]) + b"\n" for i in range(0, len(lines), 4)]
    for tag, mate in (("_R1", recs[0::2]), ("_R2", recs[1::2])):
        got = gzip.decompress(open(str(tmp_path / f"reads{tag}.fq.gz"), "rb").read())
        assert got == b"".join(mate), first_diff(got, b"".join(mate))
    assert not os.path.exists(out)
    r = subprocess.run(base + ["--output", out, "--gzip", "--gzip-level", "6"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ids2 = [ln.split("\t")[0] for ln in open(out + ".tsv").read().splitlines()[1:]]
    gs2, _ = S.GenomeSet.from_fasta(paths, uuids=ids2, min_size=prof.minimum_genome_size())
    want2 = po.simulate(po.Genomes.from_set(gs2), H.oracle_profile(prof), reads, 4)
    got = gzip.decompress(open(out, "rb").read())
    assert got == want2, first_diff(got, want2)


def test_retained_chunks_released_out_of_order(engine):
    """A sink that keeps its chunks (SIMMR_SINK_RETAINED) may give them back in any order and from any thread — the
    command line's part files do.  The ring's buffers are filled round-robin, so the launch thread has to wait for the very
    buffer it is about to fill: with a count of free buffers instead, a run of many chunks overwrote a chunk still in use
    and then waited for ever."""
    import ctypes
    import threading
    gs = H.synth_genomes(2, (1, 2), (3000, 6000), seed=31)
    prof = S.MinimalShortErrorProfile(150, 150, 30)
    engine.load_genomes(gs)
    engine.set_error_profile(prof)
    reads = [3000, 3000]
    lib = S._native.lib()
    import time
    got, held, lock, done, last_put = {}, [], threading.Lock(), threading.Event(), [time.time()]

    def releaser():                      # gives back the newest of the chunks it holds, so that the oldest waits longest;
        while not (done.is_set() and not held):          # when nothing new arrives (the run waits for its chunks) it lets go
            with lock:
                ptr = held.pop() if (len(held) >= 3 or (held and time.time() - last_put[0] > 0.05)) else None
            if ptr is not None:
                lib.simmr_chunk_release(ctypes.c_void_p(ptr))
            else:
                done.wait(0.001)

    def sink(a, first, n):
        got[first] = a.tobytes()
        with lock:
            held.append(a.ctypes.data)
            last_put[0] = time.time()
        return S._native.SINK_RETAINED

    th = threading.Thread(target=releaser, daemon=True)
    th.start()
    runner = threading.Thread(target=lambda: engine.simulate(reads, 9, sink=sink, batch_units=40), daemon=True)
    runner.start()
    runner.join(120)
    finished = not runner.is_alive()
    done.set()
    th.join(10)
    assert finished, "the run did not return: a retained chunk was lost"
    want = po.simulate(po.Genomes.from_set(gs), H.oracle_profile(prof), reads, 9)
    data = b"".join(got[k] for k in sorted(got))
    assert data == want, first_diff(data, want)


def test_full_size_c2_properties(engine):
    """BASELINE config C2 at its full size (100 M reads, 100 genomes, ~0.5 GB): size-independent properties of
    the whole run, and byte parity with the oracle on slices taken far into it (64-bit offsets, batch seams)."""
    import bench
    gs = bench.c2_genomes()
    prof = S.MinimalShortErrorProfile(150, 150, 30)
    reads, abund = S.UniformAbundanceProfile().determine_abundances(100_000_000, gs.n_genomes)
    assert reads.tolist() == [1_000_000] * 100 and abund.tolist() == [1.0] * 100        # uniform.rs:28-29
    engine.load_genomes(gs)
    engine.set_error_profile(prof)
    total = engine.total_units(reads)
    assert total == 50_000_000                                                           # simulate.rs:179: n_reads / 2 pairs
    st = engine.simulate(reads, 42)                                                      # device-resident run
    assert st["units"] == total and st["records"] == 2 * total and st["clamped"] == 0
    mean_len = st["bases"] / st["records"]
    assert abs(mean_len - 149.5) < 0.02                                                  # floor N(150, 15)
    assert st["algorithmic_bytes"] == st["bytes"] + st["bases"]                          # (3L+H+5) = (2L+H+5) + L
    assert engine.measure(reads, 42) == (st["bytes"], st["records"])
    og = po.Genomes.from_set(gs)
    oprof = H.oracle_profile(prof)
    batch = 1_000_003
    for ub in (0, batch - 3, 17 * batch - 1, total - 400):
        ue = min(total, ub + 400)
        got = engine.simulate_bytes(reads, 42, unit_begin=ub, unit_end=ue, batch_units=batch)
        want = po.simulate(og, oprof, reads, 42, unit_begin=ub, unit_end=ue)
        assert got == want, first_diff(got, want)
        recs = H.parse_fastq(got)
        assert recs[0][0].startswith(f"@{ub}|".encode())                                 # read id = global pair index (simulate.rs:85-89)
    # a slice that straddles a device batch seam of a longer run
    got = engine.simulate_bytes(reads, 42, unit_begin=batch - 300, unit_end=batch + 300, batch_units=300)
    want = po.simulate(og, oprof, reads, 42, unit_begin=batch - 300, unit_end=batch + 300)
    assert got == want, first_diff(got, want)


def test_edge_cases(engine):
    """Empty runs, many decimal fields, N-rich genomes, reads too large for a stage, very long reads."""
    gs = H.synth_genomes(3, (1, 2), (3000, 6000), seed=51)
    engine.load_genomes(gs)
    engine.set_error_profile(S.MinimalShortErrorProfile())
    assert engine.simulate_bytes([0, 0, 0], 1) == b"" and engine.simulate_bytes([1, 1, 1], 1) == b""     # n_reads / 2 pairs
    assert engine.measure([0, 1, 0], 1) == (0, 0)
    # five decimal fields: more than the kernel's fixed header program handles, falls back to the token interpreter
    fmt = "@{:read_id:}_{:start_position:}_{:end_position:}_{:read_id:}{:start_position:} {:sequence_id:}"
    check_native(engine, gs, S.MinimalShortErrorProfile(), [300, 200, 100], 4, header_format=fmt)
    check_native(engine, gs, S.MinimalShortErrorProfile(), [300, 200, 100], 4, header_format="{:read_id:}")      # starts and ends with a field
    check_native(engine, gs, S.PerfectShortErrorProfile(), [300, 200, 100], 4, header_format="@")               # one-byte header: word paths off
    # every window touches an N: all records take the byte path
    rng = np.random.default_rng(3)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 20000)].copy()
    seq[::97] = ord("N")
    seq[5000:5100] = np.frombuffer(b"acgtnRYKM-", np.uint8)[rng.integers(0, 10, 100)]
    gn = S.GenomeSet.from_contigs([("nn", [(b"ncontig", seq)])])
    check_native(engine, gn, S.MinimalShortErrorProfile(), [2000], 8)
    check_native(engine, gn, S.CustomShortErrorProfile(H.novaseq_blob()), [2000], 8)
    # short-read profile whose reads do not fit a warp's stage: one CTA per unit straight to global memory
    big = H.synth_genomes(2, (1, 1), (70000, 90000), seed=52)
    check_native(engine, big, S.PerfectShortErrorProfile(20000, 100), [6, 4], 2)
    check_native(engine, big, S.MinimalShortErrorProfile(9000, 12000, 20), [6, 4], 2)
    # long reads up to the u16 limit (perfect_long.rs:40-55 saturates at 65535)
    huge = H.synth_genomes(1, (1, 1), (400000, 400000), seed=53)
    data = check_native(engine, huge, S.MinimalLongErrorProfile(25), [300], 77)
    lens = [len(s) for _, s, _ in H.parse_fastq(data)]
    assert max(lens) == 65535 and min(lens) < 5000


def test_no_write_past_the_batch(engine, monkeypatch):
    """SIMMR_DEBUG_GUARD fills the staging buffer with a pattern and checks the 64 bytes after every batch's
    last byte (compute-sanitizer is not available on the GPU pool)."""
    monkeypatch.setenv("SIMMR_DEBUG_GUARD", "1")
    gs = H.synth_genomes(5, (1, 4), (2500, 30000), seed=61, with_n=True)
    for prof, reads in ((S.MinimalShortErrorProfile(), [3001, 7, 1500, 0, 2222]), (S.PerfectShortErrorProfile(151, 149), [999, 1001, 3, 5, 7]),
                        (S.CustomShortErrorProfile(H.novaseq_blob()), [1200, 800, 400, 200, 100])):
        check_native(engine, gs, prof, reads, 3, batch_units=997)
    gl = H.synth_genomes(2, (1, 2), (21000, 70000), seed=62)
    check_native(engine, gl, S.MinimalLongErrorProfile(20), [33, 21], 5, batch_units=13)


def test_multi_device_context_single_gpu_and_more(engine):
    """simmr_multi: per-device launch threads, one writer thread, batches dealt round-robin and handed to
    the sink in file order.  With one device it must equal the plain context; with every visible device the bytes must
    still be the one-GPU bytes, whatever the batch size."""
    import torch
    gs = H.synth_genomes(4, (1, 3), (3000, 8000), seed=9)
    prof = S.MinimalShortErrorProfile()
    reads = [1500, 700, 0, 2200]
    whole = check_native(engine, gs, prof, reads, 99)
    ndev = torch.cuda.device_count()
    for devices in ([0], list(range(ndev))):
        with S.MultiEngine(devices) as me:
            me.load_genomes(gs)
            me.set_error_profile(prof)
            assert me.total_units(reads) == engine.total_units(reads)
            for bu in (0, 97, 1000):
                order = []
                got = []
                st = me.simulate(reads, 99, batch_units=bu, sink=lambda a, f, n: (order.append(f), got.append(a.tobytes())) and 0)
                assert b"".join(got) == whole and order == sorted(order)
                assert st["bytes"] == len(whole) and st["records"] == 2 * engine.total_units(reads)
            part = []
            me.simulate(reads, 99, unit_begin=100, unit_end=900, batch_units=64, sink=lambda a, f, n: part.append(a.tobytes()) and 0)
            assert b"".join(part) == engine.simulate_bytes(reads, 99, unit_begin=100, unit_end=900)
            st = me.simulate(reads, 99)                                   # device-resident on every device
            assert st["bytes"] == len(whole)
            with pytest.raises(S.SimmrError):
                me.simulate(reads, 99, batch_units=50, sink=lambda a, f, n: 1)       # a failing sink stops the run
            assert b"".join(got) == whole
    with pytest.raises(S.SimmrError):
        S.MultiEngine([0, 0])


def test_sink_runs_on_the_library_writer_thread(engine):
    import threading
    gs = H.synth_genomes(2, (1, 2), (3000, 6000), seed=3)
    engine.load_genomes(gs); engine.set_error_profile(S.PerfectShortErrorProfile())
    seen = set()
    engine.simulate([4000, 4000], 1, batch_units=100, sink=lambda a, f, n: seen.add(threading.get_ident()) and 0)
    assert len(seen) == 1 and threading.get_ident() not in seen


def test_device_sink_sees_the_bytes_in_hbm(engine):
    """simmr_simulate_device: the device-resident run of bench.py's `value`, with a consumer that reads every batch where
    the format kernel left it."""
    gs = H.synth_genomes(3, (1, 3), (3000, 9000), seed=12)
    prof = S.MinimalShortErrorProfile()
    reads = [3000, 1000, 2000]
    whole = check_native(engine, gs, prof, reads, 6)
    got, firsts = [], []
    st = engine.simulate_device(reads, 6, lambda ptr, n, first, units: (got.append(S.copy_from_device(ptr, n).tobytes()), firsts.append((first, units))) and 0,
                                batch_units=700)
    assert b"".join(got) == whole and st["bytes"] == len(whole)
    assert [f for f, _ in firsts] == list(range(0, 3000, 700)) and sum(u for _, u in firsts) == 3000
    with pytest.raises(S.SimmrError):
        engine.simulate_device(reads, 6, lambda *a: 1)

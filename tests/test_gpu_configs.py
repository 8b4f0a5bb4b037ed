"""The five BASELINE.json configurations at their stated sizes, device-resident on one B200:
size-independent properties of the whole run plus byte parity with the CPU oracle on slices taken at the beginning,
across a device batch seam and at the very end of each run (simulate.rs:110-150, 323-406)."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
import simmr_b200 as S
from oracle import pyoracle as po
from test_gpu_parity import first_diff

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def full_run(engine, gs, prof, reads, seed=42):
    engine.load_genomes(gs)
    engine.set_error_profile(prof)
    st = engine.simulate(reads, seed)                         # device-resident: FASTQ bytes stay in HBM staging
    total = engine.total_units(reads)
    assert st["units"] == total
    if prof.err_table()["k"] == 0:
        assert st["algorithmic_bytes"] == st["bytes"] + st["bases"]      # (3L+H+5) = (2L+H+5) + L, BASELINE.md §3
    else:
        assert st["algorithmic_bytes"] > st["bytes"] + st["bases"]       # bases deleted by k-mer errors were read, not written
    assert engine.measure(reads, seed) == (st["bytes"], st["records"])    # the geometry-only pass agrees with the run
    return st, total


def check_slices(engine, gs, prof, reads, slices, seed=42, batch_units=0):
    """[(unit_begin, n_units)]: GPU bytes of the slice (as part of the whole run's index space) == oracle bytes."""
    og, oprof = po.Genomes.from_set(gs), H.oracle_profile(prof)
    out = []
    for ub, n in slices:
        got = engine.simulate_bytes(reads, seed, unit_begin=ub, unit_end=ub + n, batch_units=batch_units)
        want = po.simulate(og, oprof, reads, seed, unit_begin=ub, unit_end=ub + n)
        assert got == want, f"units [{ub}, {ub + n}): " + first_diff(got, want)
        out.append(got)
    return out


def test_c1_perfect_short_5mbp_whole_run(engine):
    """C1: perfect-short PE, 5000 reads from one synthetic 5 Mbp genome — the whole FASTQ against the oracle."""
    gs = H.config_genomes(1, 5_000_000, 5_000_000, (1, 1), seed=1)
    prof = S.PerfectShortErrorProfile(150, 150)
    reads, abund = S.UniformAbundanceProfile().determine_abundances(5000, 1)
    assert reads.tolist() == [5000] and abund.tolist() == [100.0]
    st, total = full_run(engine, gs, prof, reads)
    assert total == 2500 and st["records"] == 5000 and st["bases"] == 750_000 and st["clamped"] == 0
    data = check_slices(engine, gs, prof, reads, [(0, 2500)])[0]
    recs = H.parse_fastq(data)
    assert len(recs) == 5000 and all(q == b"]" * 150 for _, _, q in recs)                 # Q60 + 33 (perfect_short.rs:43)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for i in range(0, 5000, 2):                                                           # RL == IS: mate 2 = RC of mate 1
        assert recs[i + 1][1] == recs[i][1].translate(comp)[::-1]
        hdr = recs[i][0].decode()
        sp = int(hdr.split("|sp=")[1].split("|")[0])
        assert recs[i][1] == gs.bases[sp:sp + 150].tobytes()                              # the slice itself (simulate.rs:261)


def test_c1_on_the_reference_fixture(engine):
    """C1's profile on the reference's own 7 920 bp E. coli fixture (simmr/src/tests/data, committed under
    tests/golden/): loaded through the FASTA path, whole run against the oracle, and the reference's slice / RC
    literals (simulate_tests.rs:36-45) read back out of the records that cover them."""
    path = os.path.join(GOLDEN, "GCF_000005845.2_ASM584v2_genomic.partial.fna")
    prof = S.PerfectShortErrorProfile(60, 100)
    gs, kept = S.GenomeSet.from_fasta([path], uuids=["fixture"], min_size=prof.minimum_genome_size())
    assert kept.all() and gs.n_contigs == 1 and int(gs.contig_size[0]) == 7920
    assert gs.contig_id(0) == b"NC_000913.3 Escherichia coli str. K-12 substr. MG1655, complete genome"
    reads = [6000]
    st, total = full_run(engine, gs, prof, reads, seed=7)
    data = check_slices(engine, gs, prof, reads, [(0, total)], seed=7)[0]
    seq = gs.bases.tobytes()
    assert seq[38:58] == b"TGTGGATTAAAAAAAGAGTG" and seq[8:28] == b"ATTCTGACTGCAACGGGCAA"
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    fwd_hits = rc_hits = 0
    for hdr, s, _ in H.parse_fastq(data):
        h = hdr.decode()
        sp, ep = int(h.split("|sp=")[1].split("|")[0]), int(h.split("|ep=")[1].split("|")[0])
        if h.endswith("rc=f"):
            assert s == seq[sp:ep]
            if sp <= 38 and ep >= 58:                                                     # the record covers [38, 58)
                assert s[38 - sp:58 - sp] == b"TGTGGATTAAAAAAAGAGTG"
                fwd_hits += 1
        else:
            assert s == seq[ep:sp].translate(comp)[::-1]                                  # simulate.rs:283
            if ep <= 78 and sp >= 98:                                                     # covers [78, 98), reverse strand
                assert s[sp - 98:sp - 78] == b"TTACTCACGGCAGGTAACCA"
                rc_hits += 1
    assert fwd_hits > 0 and rc_hits > 0


def _write_fasta(path, contigs):
    with open(path, "wb") as f:
        for cid, seq in contigs:
            f.write(b">" + cid + b"\n")
            n = len(seq) // 80 * 80
            body = np.empty((n // 80, 81), np.uint8)
            body[:, :80] = np.asarray(seq[:n]).reshape(-1, 80)
            body[:, 80] = 10
            body.tofile(f)
            if n < len(seq):
                f.write(np.asarray(seq[n:]).tobytes() + b"\n")


def test_c3_novaseq_500m_reads_tsv_genome_file(engine, tmp_path):
    """C3: NovaSeq-style custom-short blob, 500 M reads over 100 genomes entered through a TSV genome file with uuid and
    abundance (files.rs:56-86) and FASTA files (genome.rs:89-162); custom abundance rounding (custom.rs:28-44)."""
    ref = H.config_genomes(100, 2_000_000, 8_000_000, (1, 20), seed=0x5EED0000)
    ab = H.lognormal_abundances(100)
    paths = []
    for g in range(ref.n_genomes):
        p = str(tmp_path / f"g{g:04d}.fna")
        _write_fasta(p, [(ref.contig_id(c), ref.contig(c)) for c in range(int(ref.genome_contig[g]), int(ref.genome_contig[g + 1]))])
        paths.append(p)
    tsv = tmp_path / "genomes.tsv"
    tsv.write_text("path\tuuid\tabundance\n" + "".join(f"{p}\t{ref.uuid(g)}\t{float(ab[g])!r}\n" for g, p in enumerate(paths)))
    recs = S.parse_genome_file(str(tsv))
    assert [r["filepath"] for r in recs] == paths and [r["uuid"] for r in recs] == [ref.uuid(g) for g in range(100)]
    assert [r["abundance"] for r in recs] == ab.tolist()
    prof = S.CustomShortErrorProfile(H.novaseq_blob())
    gs, kept = S.GenomeSet.from_fasta(paths, uuids=[r["uuid"] for r in recs], min_size=prof.minimum_genome_size())
    assert kept.all() and np.array_equal(gs.bases, ref.bases) and np.array_equal(gs.contig_off, ref.contig_off)
    assert gs.contig_ids == ref.contig_ids and gs.genome_ids == ref.genome_ids
    reads, abund = S.CustomAbundanceProfile([r["abundance"] for r in recs]).reads_per_genome(500_000_000, gs.genome_sizes)
    want_reads, want_ab = po.abundance(po.AB_CUSTOM, 500_000_000, 100, abund_in=ab)
    assert reads.tolist() == list(want_reads) and abund.tolist() == list(want_ab)
    assert reads.tolist() == [int(np.ceil(500_000_000 * a)) for a in ab]                  # custom.rs:40 (sum within [0.99, 1.01])
    st, total = full_run(engine, gs, prof, reads)
    assert total == int((reads // 2).sum()) and st["records"] == 2 * total and st["clamped"] == 0
    assert 141.0 <= st["bases"] / st["records"] <= 151.0
    batch = 5_000_011
    data = check_slices(engine, gs, prof, reads, [(0, 300), (batch - 150, 300), (123_456_789, 300), (total - 300, 300)], batch_units=97)
    qs = np.frombuffer(b"".join(q for _, _, q in H.parse_fastq(data[2])), np.uint8)
    assert 28 <= qs.mean() - 33 <= 38                                                     # blob: mean Q36 decaying to Q30
    assert H.parse_fastq(data[3])[-1][0].startswith(f"@{total - 1}|{ref.uuid(99)}/2".encode())


@pytest.mark.parametrize("which", ["minimal-long", "ont"])
def test_c4_long_reads_10m(engine, which):
    """C4: 10 M long reads with gamma lengths from 50 genomes — (a) minimal-long, (b) an ONT-style custom-long blob
    (empirical Phreds, k-mer error injection, custom_long.rs:187-232)."""
    gs = H.config_genomes(50, 2_000_000, 6_000_000, (1, 4), seed=4, min_contig=100_000)
    prof = S.MinimalLongErrorProfile(30) if which == "minimal-long" else S.CustomLongErrorProfile(H.ont_blob_with_kmers(k=7, n_alts=20)[0])
    reads, _ = S.UniformAbundanceProfile().determine_abundances(10_000_000, 50)
    assert reads.tolist() == [200_000] * 50
    st, total = full_run(engine, gs, prof, reads)
    assert total == 10_000_000 and st["records"] == total
    mean_len = st["bases"] / st["records"]
    # Gamma(1.7777779, 11250) floored and saturated at 65 535: E = 19 822; the synthetic alt-k-mer table
    # deletes about one base in fourteen
    assert (19_000 < mean_len < 20_500) if which == "minimal-long" else (17_000 < mean_len < 19_800)
    check_slices(engine, gs, prof, reads, [(0, 24), (4_999_990, 24), (total - 24, 24)], batch_units=7)


def test_c5_exact_1000_genomes_1b_reads(engine):
    """C5: exact abundance, 1000 genomes (4.4 GB: window offsets of the last genomes exceed 2^32), 1 M reads per genome
    = 1 B reads on one GPU; the sharded view of the same job is checked on the slices' unit ranges."""
    gs = H.config_genomes(1000, 4_000_000, 4_800_000, (1, 6), seed=5)
    assert gs.bases.nbytes > 2**32
    prof = S.MinimalShortErrorProfile(150, 150, 30)
    reads, abund = S.ExactAbundanceProfile().determine_abundances(1_000_000, 1000)
    assert reads.tolist() == [1_000_000] * 1000 and abund.tolist() == [0.1] * 1000      # exact.rs:18-22
    st, total = full_run(engine, gs, prof, reads)
    assert total == 500_000_000 and st["records"] == 1_000_000_000 and st["clamped"] == 0
    assert abs(st["bases"] / st["records"] - 149.5) < 0.02
    first_big = int(np.searchsorted(gs.contig_off[gs.genome_contig[:-1]], 2**32))     # first genome stored wholly above 4 GiB
    assert first_big < 1000
    slices = [(0, 300), (250_000_000 - 150, 300), (first_big * 500_000 - 150, 300), (first_big * 500_000 + 200_000, 300), (total - 300, 300)]
    data = check_slices(engine, gs, prof, reads, slices, batch_units=97)
    assert H.parse_fastq(data[-1])[-1][0].startswith(f"@{total - 1}|{gs.uuid(999)}/2".encode())
    # the same slices as shards of an 8-way split: a shard's bytes do not depend on where the run was cut
    from simmr_b200 import sharding
    ub, ue = sharding.shard_range(total, 7, 8)
    got = engine.simulate_bytes(reads, 42, unit_begin=ue - 300, unit_end=ue)
    assert got == data[-1]


def test_custom_long_replay(engine):
    """Replay (recorded draws, the reference's literal arithmetic) for an is_long blob: alias-sampled Phreds per position
    (custom_short.rs:108-129 semantics), no substitutions (custom_long.rs:238-245)."""
    gs = H.synth_genomes(2, (1, 2), (21000, 60000), seed=47)
    og = po.Genomes.from_set(gs)
    prof = S.CustomLongErrorProfile(H.ont_blob())
    d = H.make_long_draws(gs, prof, 24, seed=11)
    d["phred_u"] = d["phred_u"] * 8.0
    engine.load_genomes(gs)
    engine.set_error_profile(prof)
    got = engine.replay_bytes(S.Draws(**d))
    want = po.replay(og, H.oracle_profile(prof), po.Draws(**d))
    assert got == want, first_diff(got, want)
    assert engine.replay_bytes(S.Draws(**d), batch_units=5) == want


def test_cli_two_gpus_single_file(engine, tmp_path):
    """`simmr --gpus 2`: two device workers, host-side count exchange, one output file byte-identical to the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(5)
    paths = []
    for g in range(3):
        p = str(tmp_path / f"g{g}.fna")
        _write_fasta(p, [(f"g{g}_c{c}".encode(), np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(rng.integers(3000, 9000)))]) for c in range(2)])
        paths.append(p)
    out = str(tmp_path / "reads.fq")
    exe = os.path.join(os.path.dirname(S._native.LIB_PATH), "simmr")
    cmd = [exe] + sum([["--genome", p] for p in paths], []) + ["--output", out, "--num-reads", "9001", "--seed", "3", "--error-profile",
                                                                  "minimal-short", "--gpus", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ids = [ln.split("\t")[0] for ln in open(out + ".tsv").read().splitlines()[1:]]
    prof = S.MinimalShortErrorProfile()
    gs, _ = S.GenomeSet.from_fasta(paths, uuids=ids, min_size=prof.minimum_genome_size())
    reads, _ = S.UniformAbundanceProfile().determine_abundances(9001, 3)
    want = po.simulate(po.Genomes.from_set(gs), H.oracle_profile(prof), reads, 3)
    got = open(out, "rb").read()
    assert got == want, first_diff(got, want)

"""CPU-side checks of libsimmr_b200.so: it loads, exports every symbol include/simmr_b200.h
declares, and its host-only entry points (abundance rounding, sampling tables, FASTA /
genome-file ingest, metadata) agree with the CPU oracle bit for bit.  No compute entry
point is called: without a GPU they must fail loudly, which is asserted."""
import ctypes as C
import gzip
import os
import re

import numpy as np
import pytest

import helpers as H
import simmr_b200 as S
from oracle import pyoracle as po
from simmr_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "simmr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(simmr_[a-z0-9_]+)\s*\(", hdr)) - {"simmr_sink_fn"}
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    lib = N.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.simmr_version()


def test_no_cpu_fallback():
    """Without a usable sm_100 device the compute entry points refuse to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is for CPU-only hosts")
    with pytest.raises(S.SimmrError) as e:
        S.Engine(0)
    assert e.value.code == N.E_CUDA
    # the package must not import the oracle anywhere
    for root, _, files in os.walk(os.path.join(ROOT, "simmr_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "pyoracle" not in text and "simmr_oracle" not in text and "import oracle" not in text, f


PROFILES = [S.PerfectShortErrorProfile(150, 150), S.PerfectShortErrorProfile(20, 20), S.MinimalShortErrorProfile(150, 150, 30),
            S.MinimalShortErrorProfile(100, 400, 17), S.MinimalShortErrorProfile(250, 50, 40), S.MinimalShortErrorProfile(75, 0, 3),
            S.PerfectLongErrorProfile(), S.MinimalLongErrorProfile(30), S.MinimalLongErrorProfile(12, compat_start=True),
            S.CustomShortErrorProfile(H.novaseq_blob()), S.CustomLongErrorProfile(H.ont_blob())]


@pytest.mark.parametrize("prof", PROFILES, ids=lambda p: f"{type(p).__name__}-{p.read_length}-{p.insert_size}-{p.mean_phred_score}")
def test_profile_scalars_and_tables_match_the_oracle(prof):
    """ErrorProfile::minimum_genome_size / is_long_read (base.rs:29-31) and the native sampling
    tables, built independently in host_model.cpp and in the oracle."""
    o = H.oracle_profile(prof)
    assert prof.minimum_genome_size() == o.min_genome_size()
    assert prof.is_long_read() == o.is_long()
    a, b = prof.phred_table(), o.phred_table()
    assert (a["kind"], a["cb"], a["npos"], a["const_q"]) == (b["kind"], b["cb"], b["npos"], b["const_q"])
    assert np.array_equal(a["entries"], b["entries"])
    for which in (0, 1):
        a, b = prof.len_table(which), o.len_table(which)
        assert (a["kind"], a["cb"]) == (b["kind"], b["cb"])
        for k in ("entries", "lo", "hi"):
            assert np.array_equal(a[k], b[k]), (which, k)


def test_reference_profile_constants():
    assert S.PerfectShortErrorProfile(150, 150).minimum_genome_size() == 450       # perfect_short.rs:56-59
    assert S.MinimalShortErrorProfile(100, 400).minimum_genome_size() == 600       # minimal_short.rs:142-145
    assert S.MinimalLongErrorProfile().minimum_genome_size() == 20000              # minimal_long.rs:152-154
    assert S.PerfectLongErrorProfile().is_long_read() and not S.PerfectShortErrorProfile().is_long_read()
    assert S.CustomShortErrorProfile(H.novaseq_blob()).minimum_genome_size() == 648  # (2*149 + 350) as u16, custom_short.rs:535-538
    with pytest.raises(S.SimmrError):
        S.CustomShortErrorProfile(b"\x01\x02\x03").minimum_genome_size()


def test_abundance_counts_bit_identical():
    """a1/a2 of SURVEY §8: uniform.rs:18-95, exact.rs:17-35, custom.rs:20-96."""
    r, a = S.UniformAbundanceProfile().determine_abundances(100, 5)                # abundance_profile_tests.rs:8-30
    assert r.tolist() == [20] * 5 and a.tolist() == [20.0] * 5
    r, a = S.UniformAbundanceProfile().determine_abundances(100_000_000, 100)      # C2
    assert r.tolist() == [1_000_000] * 100 and a.tolist() == [1.0] * 100
    r, a = S.ExactAbundanceProfile().determine_abundances(1_000_000, 1000)         # C5
    assert r.tolist() == [1_000_000] * 1000 and a.tolist() == [0.1] * 1000
    r, _ = S.UniformAbundanceProfile().determine_abundances(1000, 3)
    assert r.tolist() == [334] * 3                                                  # ceil(1000/3)
    rng = np.random.default_rng(5)
    for trial in range(60):
        n = int(rng.integers(1, 40))
        total = int(rng.integers(1, 10 ** int(rng.integers(2, 10))))
        sizes = rng.integers(1000, 9_000_000, n).astype(np.uint64)
        ab = rng.lognormal(0, 1, n)
        if trial % 3 == 0:
            ab = ab / ab.sum()                      # sums to ~1: taken as is (custom.rs:29)
        for adj in (False, True):
            for prof, kind, arg in ((S.UniformAbundanceProfile(adj), po.AB_UNIFORM, None), (S.ExactAbundanceProfile(), po.AB_EXACT, None),
                                    (S.CustomAbundanceProfile(ab, adj), po.AB_CUSTOM, ab)):
                got_r, got_a = prof.reads_per_genome(total, sizes)
                want_r, want_a = po.abundance(kind, total, n, arg, sizes, adj)
                assert np.array_equal(got_r, want_r) and got_a.tobytes() == want_a.tobytes()
    # numpy restatement of custom.rs:28-44 (sequential f64 sum, renormalise outside [0.99, 1.01], ceil)
    ab = np.array([0.5, 0.25, 0.125, 0.0625], np.float64)
    s = 0.0
    for x in ab:
        s += x
    r, a = S.CustomAbundanceProfile(ab).determine_abundances(1001, 4)
    assert s < 0.99 and a.tolist() == (ab / s).tolist() and r.tolist() == [int(np.ceil(1001 * (x / s))) for x in ab]


def _write(tmp_path, name, data, gz=False):
    p = tmp_path / name
    if gz:
        with gzip.open(p, "wb") as f:
            f.write(data)
    else:
        p.write_bytes(data)
    return str(p)


def test_fasta_ingest_matches_oracle_and_reference_fixture(tmp_path):
    """Genome::from_fasta (genome.rs:89-162) with needletail normalize(false): genome_tests.rs:8-20
    on the reference's fixture, plus lower case / IUPAC / CRLF / gzip / blank lines."""
    gs, kept = S.GenomeSet.from_fasta([os.path.join(GOLD, "sample.fna")], uuids=["u0"])
    assert gs.n_genomes == 1 and gs.n_contigs == 2 and kept.all()
    assert [gs.contig_id(c) for c in range(2)] == [b"header1", b"header2"]
    assert gs.contig_size.tolist() == [160, 160] and gs.genome_sizes.tolist() == [320]
    tricky = (b">c1 some description\r\nacgtNNryk\r\nACGU.~-\r\n\r\n>c2\nTTTT\nGG GG\n>empty\n>c4\nA\n")
    plain = _write(tmp_path, "t.fna", tricky)
    for gz in (False, True):          # needletail reads gzip transparently; the oracle reads plain text only
        path = _write(tmp_path, "t.fna.gz", tricky, True) if gz else plain
        gs, _ = S.GenomeSet.from_fasta([path], uuids=["g"])
        og = po.load_fasta(plain)
        assert gs.n_contigs == og["num_seqs"] == 4
        for c in range(4):
            assert gs.contig_id(c) == og["seqs"][c]["id"]
            assert gs.contig(c).tobytes() == og["seqs"][c]["seq"]
        assert gs.contig(0).tobytes() == b"ACGTNNNNNACGT---" and gs.contig_id(0) == b"c1 some description"
        assert gs.contig(1).tobytes() == b"TTTTGGGG" and gs.contig(2).tobytes() == b""
        assert int(gs.genome_sizes[0]) == og["size"]
        cs, _ = S.GenomeSet.from_fasta([path], uuids=["g"], contiguous=True)      # genome.rs:117-136
        oc = po.load_fasta(plain, contiguous=True)
        assert cs.n_contigs == 1 and cs.contig_id(0) == b"whole genome" == oc["seqs"][0]["id"]
        assert cs.contig(0).tobytes() == oc["seqs"][0]["seq"] and int(cs.contig_size[0]) == oc["seqs"][0]["size"]
    with pytest.raises(S.SimmrError):
        S.GenomeSet.from_fasta([_write(tmp_path, "bad.fna", b"ACGT\n")])
    with pytest.raises(S.SimmrError):
        S.GenomeSet.from_fasta([str(tmp_path / "missing.fna")])
    # random uuid: top 64 bits of a v4 UUID printed with {:x} (util.rs:124-129)
    gs, _ = S.GenomeSet.from_fasta([os.path.join(GOLD, "sample.fna")])
    assert re.fullmatch(r"[0-9a-f]{1,16}", gs.uuid(0))


def test_fasta_ingest_compressed_and_parallel(tmp_path, monkeypatch):
    """needletail reads gzip, bzip2 and xz by their magic bytes (genome.rs:93); several files load in parallel and a large
    file is cut at line starts into chunks normalised in parallel (SIMMR_FASTA_CHUNK shrinks the chunks for the test)."""
    import bz2
    import gzip
    import lzma
    rng = np.random.default_rng(4)
    files, texts = [], []
    for g in range(5):
        recs = []
        for c in range(int(rng.integers(1, 5))):
            seq = np.frombuffer(b"ACGTacgtNnRY-.", np.uint8)[rng.integers(0, 14, int(rng.integers(0, 30000)))].tobytes()
            w = int(rng.integers(50, 90))
            recs.append(b">g%d_c%d description %d\n" % (g, c, c) + b"".join(seq[i:i + w] + (b"\r\n" if c % 2 else b"\n") for i in range(0, len(seq), w)))
        texts.append(b"".join(recs))
        files.append(_write(tmp_path, f"g{g}.fna", texts[-1]))
    plain, _ = S.GenomeSet.from_fasta(files, uuids=[f"u{g}" for g in range(5)], threads=1)
    for g in range(5):                                        # against the oracle's single-threaded reader
        og = po.load_fasta(files[g])
        c0 = int(plain.genome_contig[g])
        assert int(plain.genome_contig[g + 1]) - c0 == og["num_seqs"] and int(plain.genome_sizes[g]) == og["size"]
        for c in range(og["num_seqs"]):
            assert plain.contig_id(c0 + c) == og["seqs"][c]["id"] and plain.contig(c0 + c).tobytes() == og["seqs"][c]["seq"]
    same = lambda a, b: (np.array_equal(a.bases, b.bases) and np.array_equal(a.contig_off, b.contig_off) and a.contig_ids == b.contig_ids
                         and np.array_equal(a.genome_contig, b.genome_contig) and np.array_equal(a.genome_sizes, b.genome_sizes))
    monkeypatch.setenv("SIMMR_FASTA_CHUNK", "3000")
    for threads in (2, 8, 0):
        chunked, _ = S.GenomeSet.from_fasta(files, uuids=[f"u{g}" for g in range(5)], threads=threads)
        assert same(plain, chunked)
        cont, _ = S.GenomeSet.from_fasta(files, uuids=None, contiguous=True, threads=threads)
        for g in range(5):
            oc = po.load_fasta(files[g], contiguous=True)
            assert cont.contig(g).tobytes() == oc["seqs"][0]["seq"] and int(cont.contig_size[g]) == oc["seqs"][0]["size"]
    monkeypatch.delenv("SIMMR_FASTA_CHUNK")
    packed = []
    for g, t in enumerate(texts):
        half = len(t) // 2
        data = [gzip.compress(t), gzip.compress(t[:half]) + gzip.compress(t[half:]), bz2.compress(t), bz2.compress(t[:half]) + bz2.compress(t[half:]),
                lzma.compress(t, format=lzma.FORMAT_XZ)][g]
        packed.append(_write(tmp_path, f"p{g}.fna.{'gz gz bz2 bz2 xz'.split()[g]}", data))
    comp, _ = S.GenomeSet.from_fasta(packed, uuids=[f"u{g}" for g in range(5)])
    assert same(plain, comp)
    with pytest.raises(S.SimmrError):
        S.GenomeSet.from_fasta([_write(tmp_path, "cut.fna.gz", gzip.compress(texts[0])[:-20])])
    with pytest.raises(S.SimmrError):
        S.GenomeSet.from_fasta([_write(tmp_path, "cut.fna.bz2", bz2.compress(texts[0])[:-20])])


def test_contig_filter(tmp_path):
    """main.rs:117-162: contigs with size <= minimum are dropped, then empty genomes; Genome.size keeps
    its pre-filter value (main.rs:158-161)."""
    a = _write(tmp_path, "a.fna", b">big\n" + b"ACGT" * 200 + b"\n>small\n" + b"AC" * 100 + b"\n")
    b = _write(tmp_path, "b.fna", b">tiny\nACGT\n")
    c = _write(tmp_path, "c.fna", b">edge\n" + b"A" * 450 + b"\n>ok\n" + b"C" * 451 + b"\n")
    gs, kept = S.GenomeSet.from_fasta([a, b, c], uuids=["a", "b", "c"], min_size=450)
    assert kept.tolist() == [True, False, True]
    assert gs.n_genomes == 2 and [gs.contig_id(i) for i in range(gs.n_contigs)] == [b"big", b"ok"]
    assert gs.genome_sizes.tolist() == [1000, 901] and gs.paths == [a, c] and [gs.uuid(0), gs.uuid(1)] == ["a", "c"]


def test_genome_file_and_metadata(tmp_path):
    """files.rs:56-86 (documented behaviour, readme.md:99-100: see SURVEY §A.4) and files.rs:100-134."""
    simple = _write(tmp_path, "simple.txt", b"/data/a.fna\n/data/b.fna\n\n/data/c d.fna\n")
    tsv = _write(tmp_path, "g.tsv", b"abundance\tfilepath\tnote\tuuid\n0.25\t/x/a.fna\thello\tid-a\n\t/x/b.fna\t\t\n0.5\t/x/c.fna\tz\tid-c\n")
    tsv2 = _write(tmp_path, "g2.tsv", b"path\tgenome_id\n/x/a.fna\tA\n")
    for p in (simple, tsv, tsv2):
        assert S.parse_genome_file(p) == po.parse_genome_file(p)
    assert S.parse_genome_file(simple) == [dict(filepath=f, uuid=None, abundance=None) for f in ("/data/a.fna", "/data/b.fna", "/data/c d.fna")]
    assert S.parse_genome_file(tsv) == [dict(filepath="/x/a.fna", uuid="id-a", abundance=0.25), dict(filepath="/x/b.fna", uuid=None, abundance=None),
                                        dict(filepath="/x/c.fna", uuid="id-c", abundance=0.5)]
    with pytest.raises(S.SimmrError):
        S.parse_genome_file(str(tmp_path / "nope"))
    meta = [("id-a", "/x/a.fna", 1000000, 1.0), ("id-b", "/x/b.fna", 334, 100.0 / 3), ("c", "rel/c.fna", 0, 0.1)]
    out = str(tmp_path / "reads.fq.tsv")
    S.write_metadata(meta, out)
    text = open(out).read()
    assert text == po.metadata_tsv(*zip(*meta))
    assert text.splitlines()[0] == "genome_id\tfilepath\tnum_reads\tabundance"
    assert text.splitlines()[2] == "id-b\t/x/b.fna\t334\t33.333333333333336"


def test_philox_host_copy():
    c = np.array([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], np.uint32)
    k = np.array([0xa4093822, 0x299f31d0], np.uint32)
    o = np.zeros(4, np.uint32)
    N.lib().simmr_debug_philox4x32_10(c.ctypes.data_as(N.u32p), k.ctypes.data_as(N.u32p), o.ctypes.data_as(N.u32p))
    assert o.tolist() == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]             # Random123 kat_vectors


def test_output_helpers_gzip_members_and_mate_split():
    """simmr_gzip_member: concatenated members read back as the plain bytes; simmr_fastq_split_mates: records of an
    interleaved chunk go alternately to the two mates, malformed chunks are refused."""
    import ctypes as C
    import gzip
    L = N.lib()
    rng = np.random.default_rng(3)
    recs = []
    for i in range(200):
        n = int(rng.integers(1, 200))
        seq = bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8))
        q = bytes((33 + rng.integers(0, 60, n)).astype(np.uint8))          # '@' and '+' occur among the qualities
        recs.append(b"@r%d|x/%d m:sp=%d\n" % (i // 2, 1 + i % 2, i) + seq + b"\n+\n" + q + b"\n")
    data = b"".join(recs)
    src = np.frombuffer(data, np.uint8)
    out = b""
    for lo, hi in ((0, 1000), (1000, 1000), (1000, len(data))):             # an empty slice is a member too
        cap = int(L.simmr_gzip_bound(hi - lo))
        dst = np.zeros(cap, np.uint8)
        got = L.simmr_gzip_member(src[lo:].ctypes.data_as(N.u8p), hi - lo, dst.ctypes.data_as(N.u8p), cap, 1)
        assert got > 0
        out += dst[:got].tobytes()
    assert gzip.decompress(out) == data
    small = np.zeros(8, np.uint8)
    assert L.simmr_gzip_member(src.ctypes.data_as(N.u8p), len(data), small.ctypes.data_as(N.u8p), 8, 1) == N.E_LIMIT
    r1, r2 = np.zeros(len(data), np.uint8), np.zeros(len(data), np.uint8)
    n1, n2 = C.c_uint64(), C.c_uint64()
    assert L.simmr_fastq_split_mates(src.ctypes.data_as(N.u8p), len(data), r1.ctypes.data_as(N.u8p), C.byref(n1), r2.ctypes.data_as(N.u8p), C.byref(n2)) == 0
    assert r1[:n1.value].tobytes() == b"".join(recs[0::2]) and r2[:n2.value].tobytes() == b"".join(recs[1::2])
    odd = np.frombuffer(b"".join(recs[:3]), np.uint8)
    assert L.simmr_fastq_split_mates(odd.ctypes.data_as(N.u8p), len(odd), r1.ctypes.data_as(N.u8p), C.byref(n1), r2.ctypes.data_as(N.u8p), C.byref(n2)) == N.E_PARSE
    cut = np.frombuffer(data[:-5], np.uint8)
    assert L.simmr_fastq_split_mates(cut.ctypes.data_as(N.u8p), len(cut), r1.ctypes.data_as(N.u8p), C.byref(n1), r2.ctypes.data_as(N.u8p), C.byref(n2)) == N.E_PARSE

"""ctypes binding of the CPU oracle (libsimmr_oracle.so).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from simmr_b200/.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PERFECT_SHORT, MINIMAL_SHORT, PERFECT_LONG, MINIMAL_LONG, CUSTOM_SHORT, CUSTOM_LONG = range(6)
AB_EXACT, AB_UNIFORM, AB_CUSTOM = range(3)
FLAG_COMPAT_LONG_START = 1

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


class OrcGenomes(C.Structure):
    _fields_ = [("n_genomes", C.c_uint32), ("n_contigs", C.c_uint32),
                ("bases", u8p), ("contig_off", u64p), ("contig_size", u64p),
                ("genome_contig", u32p), ("contig_ids", C.c_char_p), ("contig_id_off", u32p),
                ("genome_ids", C.c_char_p), ("genome_id_off", u32p)]


class OrcProfile(C.Structure):
    _fields_ = [("kind", C.c_int32), ("read_length", C.c_uint16), ("insert_size", C.c_uint16),
                ("mean_phred", C.c_uint8), ("pad_", C.c_uint8 * 3), ("flags", C.c_uint32),
                ("blob", u8p), ("blob_len", C.c_uint64)]


class OrcOpts(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("header_format", C.c_char_p),
                ("unit_begin", C.c_uint64), ("unit_end", C.c_uint64), ("read_id_base", C.c_uint32)]


class OrcDraws(C.Structure):
    _fields_ = [("n_units", C.c_uint64), ("genome", u32p), ("contig", u32p),
                ("read_length", u16p), ("insert_size", u16p), ("start", u64p), ("redraw", u64p),
                ("base_off", u64p), ("phred_z", f32p), ("phred_cand", u32p), ("phred_u", f64p),
                ("phred_inbin", u32p), ("mut_u", f32p), ("mut_k", u8p)]


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("records", "bases", "bytes", "units", "clamped", "mutated")]


def _cpu_stamp():
    """The oracle is built -march=native (it is also the timed CPU baseline): a library built on another machine's CPU
    (this container vs the GPU box) is rebuilt where it runs."""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                import hashlib
                return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force=False):
    if force:
        subprocess.check_call(["make", "-s", "-C", _HERE, "clean"])
    subprocess.check_call(["make", "-s", "-C", _HERE])
    with open(os.path.join(_HERE, ".build_cpu"), "w") as f:
        f.write(_cpu_stamp())


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libsimmr_oracle.so")
        stamp = os.path.join(_HERE, ".build_cpu")
        if not os.path.exists(path):
            build()
        elif not os.path.exists(stamp) or open(stamp).read() != _cpu_stamp():
            build(force=True)
        L = C.CDLL(path)
        L.orc_last_error.restype = C.c_char_p
        L.orc_complement.restype = C.c_uint8
        L.orc_complement.argtypes = [C.c_uint8]
        L.orc_phred_to_probability.restype = C.c_float
        L.orc_phred_to_probability.argtypes = [C.c_uint8]
        L.orc_phred_to_accuracy.restype = C.c_float
        L.orc_phred_to_accuracy.argtypes = [C.c_uint8]
        L.orc_probability_to_phred.restype = C.c_uint8
        L.orc_probability_to_phred.argtypes = [C.c_float]
        L.orc_accuracy_to_phred.restype = C.c_uint8
        L.orc_accuracy_to_phred.argtypes = [C.c_float]
        L.orc_gaussian.restype = C.c_double
        L.orc_gaussian.argtypes = [C.c_double, f64p, C.c_uint64, C.c_double]
        L.orc_simulate_errors.restype = C.c_int64
        L.orc_parse_genome_file.restype = C.c_int64
        L.orc_metadata_tsv.restype = C.c_int64
        L.orc_total_units.restype = C.c_uint64
        L.orc_simulate.restype = C.c_int64
        L.orc_replay.restype = C.c_int64
        L.orc_ref_stream.restype = C.c_int64
        L.orc_native_errors.restype = C.c_int64
        L.orc_fasta_num_seqs.restype = C.c_uint32
        for f in ("orc_fasta_size", "orc_fasta_seq_size", "orc_fasta_seq_len"):
            getattr(L, f).restype = C.c_uint64
        L.orc_fasta_seq.restype = u8p
        L.orc_fasta_seq_id.restype = C.POINTER(C.c_char)
        _LIB = L
    return _LIB


def err():
    return lib().orc_last_error().decode()


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class Genomes:
    """Holds the numpy arrays alive and exposes the C struct."""

    def __init__(self, bases, contig_off, contig_size, genome_contig, contig_ids, contig_id_off,
                 genome_ids, genome_id_off):
        self.bases = np.ascontiguousarray(bases, np.uint8)
        self.contig_off = np.ascontiguousarray(contig_off, np.uint64)
        self.contig_size = np.ascontiguousarray(contig_size, np.uint64)
        self.genome_contig = np.ascontiguousarray(genome_contig, np.uint32)
        self.contig_ids = bytes(contig_ids)
        self.contig_id_off = np.ascontiguousarray(contig_id_off, np.uint32)
        self.genome_ids = bytes(genome_ids)
        self.genome_id_off = np.ascontiguousarray(genome_id_off, np.uint32)
        self.c = OrcGenomes(len(self.genome_contig) - 1, len(self.contig_size),
                            _p(self.bases, u8p), _p(self.contig_off, u64p), _p(self.contig_size, u64p),
                            _p(self.genome_contig, u32p), self.contig_ids, _p(self.contig_id_off, u32p),
                            self.genome_ids, _p(self.genome_id_off, u32p))

    @classmethod
    def from_set(cls, gs):
        """gs: any object with the flat-array fields (e.g. simmr_b200.GenomeSet)."""
        return cls(gs.bases, gs.contig_off, gs.contig_size, gs.genome_contig, gs.contig_ids,
                   gs.contig_id_off, gs.genome_ids, gs.genome_id_off)


class Profile:
    def __init__(self, kind, read_length=150, insert_size=150, mean_phred=30, flags=0, blob=None):
        self._blob = np.frombuffer(blob, np.uint8).copy() if blob is not None else None
        self.c = OrcProfile(kind, read_length, insert_size, mean_phred, (C.c_uint8 * 3)(), flags,
                            _p(self._blob, u8p), 0 if blob is None else len(blob))
        self.kind = kind

    def min_genome_size(self):
        out = C.c_uint16()
        if lib().orc_profile_min_genome_size(C.byref(self.c), C.byref(out)) != 0:
            raise RuntimeError(err())
        return out.value

    def is_long(self):
        out = C.c_int()
        if lib().orc_profile_is_long(C.byref(self.c), C.byref(out)) != 0:
            raise RuntimeError(err())
        return bool(out.value)

    def phred_table(self):
        kind, cb, npos, cq = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        ent = np.zeros(1 << 20, np.uint32)
        n = lib().orc_profile_phred_table(C.byref(self.c), C.byref(kind), C.byref(cb), C.byref(npos), C.byref(cq),
                                          _p(ent, u32p), C.c_uint64(len(ent)))
        if n < 0:
            raise RuntimeError(err())
        return dict(kind=kind.value, cb=cb.value, npos=npos.value, const_q=cq.value, entries=ent[:n].copy())

    def len_table(self, which):
        kind, cb = C.c_uint32(), C.c_uint32()
        ent, lo, hi = (np.zeros(1 << 13, np.uint32) for _ in range(3))
        n = lib().orc_profile_len_table(C.byref(self.c), which, C.byref(kind), C.byref(cb), _p(ent, u32p),
                                        _p(lo, u32p), _p(hi, u32p), C.c_uint64(len(ent)))
        if n < 0:
            raise RuntimeError(err())
        m = max(n, 1)
        return dict(kind=kind.value, cb=cb.value, entries=ent[:n].copy(), lo=lo[:m].copy(), hi=hi[:m].copy())


    def err_table(self):
        k, n = C.c_uint32(), C.c_uint32()
        words = np.zeros(1 << 24, np.uint32)
        m = lib().orc_profile_err_table(C.byref(self.c), C.byref(k), C.byref(n), _p(words, u32p), C.c_uint64(len(words)))
        if m < 0:
            raise RuntimeError(err())
        return dict(k=k.value, n_ref=n.value, words=words[:m].copy())


def total_units(profile, reads_per_genome):
    r = np.ascontiguousarray(reads_per_genome, np.uint64)
    return lib().orc_total_units(C.byref(profile.c), _p(r, u64p), len(r))


def _opts(seed, header_format, unit_begin, unit_end, read_id_base):
    hf = header_format.encode() if isinstance(header_format, str) else header_format
    return OrcOpts(seed, hf, unit_begin, unit_end, read_id_base)


def simulate(genomes, profile, reads_per_genome, seed, header_format=None, unit_begin=0, unit_end=0,
             read_id_base=0, cap=None, want_stats=False):
    """Native (Philox) mode FASTQ bytes for units [unit_begin, unit_end)."""
    r = np.ascontiguousarray(reads_per_genome, np.uint64)
    o = _opts(seed, header_format, unit_begin, unit_end, read_id_base)
    st = OrcStats()
    if cap is None:
        n = unit_end - unit_begin if (unit_begin or unit_end) else total_units(profile, r)
        per = 140000 * 2 + 4096 if profile.is_long() else 2 * (2 * 70000 // 100 + 2048)
        cap = max(1 << 16, int(n) * per)
    buf = np.empty(cap, np.uint8)
    n = lib().orc_simulate(C.byref(genomes.c), C.byref(profile.c), _p(r, u64p), C.byref(o),
                           _p(buf, u8p), C.c_uint64(cap), C.byref(st))
    if n < 0:
        if -n > cap and not err().strip():
            return simulate(genomes, profile, reads_per_genome, seed, header_format, unit_begin, unit_end,
                            read_id_base, cap=-n, want_stats=want_stats)
        raise RuntimeError(err() or f"oracle buffer too small ({-n})")
    out = buf[:n].tobytes()
    return (out, st) if want_stats else out


def bench_pe(genomes, profile, reads_per_genome, seed, unit_begin, n_units, mode, threads):
    """orc_bench_pe -> dict(seconds, bases, records, bytes); mode 0 native draw schedule, 1 the reference's stream."""
    r = np.ascontiguousarray(reads_per_genome, np.uint64)
    sec, nb, nr, ny = C.c_double(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    if lib().orc_bench_pe(C.byref(genomes.c), C.byref(profile.c), _p(r, u64p), C.c_uint64(seed), C.c_uint64(unit_begin),
                          C.c_uint64(n_units), int(mode), int(threads), C.byref(sec), C.byref(nb), C.byref(nr), C.byref(ny)) != 0:
        raise RuntimeError(err())
    return dict(seconds=sec.value, bases=nb.value, records=nr.value, bytes=ny.value)


def ref_stream(genomes, profile, n_pairs, seed):
    """FASTQ of the first n_pairs pairs of genome 0 through the reference's own stream plumbing (ref_stream.inc)."""
    cap = int(n_pairs) * 2 * (2 * 70000 // 100 + 2048) + 4096
    buf = np.empty(cap, np.uint8)
    st = OrcStats()
    n = lib().orc_ref_stream(C.byref(genomes.c), C.byref(profile.c), C.c_uint64(n_pairs), C.c_uint64(seed), _p(buf, u8p),
                             C.c_uint64(cap), C.byref(st))
    if n < 0:
        raise RuntimeError(err())
    return buf[:n].tobytes(), st


def native_errors(profile, seed, unit, seq):
    """k-mer error injection of native mode on `seq` as read `unit` of a run seeded `seed` -> (new sequence, choices):
    `choices` is what orc_simulate_errors (the reference loop with explicit alias picks) needs to give the same bytes."""
    a = np.frombuffer(bytes(seq), np.uint8)
    out = np.empty(len(a) + 16, np.uint8)
    ch = np.empty(len(a) + 16, np.uint32)
    nch = C.c_uint64()
    n = lib().orc_native_errors(C.byref(profile.c), C.c_uint64(seed), C.c_uint64(unit), _p(a, u8p), C.c_uint64(len(a)),
                                _p(out, u8p), C.c_uint64(len(out)), _p(ch, u32p), C.c_uint64(len(ch)), C.byref(nch))
    if n < 0:
        raise RuntimeError(err())
    return out[:n].tobytes(), ch[:nch.value].copy()


def native_error_first_picks(profile, seed, unit0, n_units, seq):
    """The first pick of each of n_units consecutive reads over the same sequence (for distribution tests)."""
    a = np.frombuffer(bytes(seq), np.uint8)
    ch = np.empty(n_units, np.uint32)
    if lib().orc_native_errors(C.byref(profile.c), C.c_uint64(seed), C.c_uint64(unit0), _p(a, u8p), C.c_uint64(len(a)),
                               None, C.c_uint64(n_units), _p(ch, u32p), C.c_uint64(n_units), None) < 0:
        raise RuntimeError(err())
    return ch


def simulate_errors(seq, k, kmers, choices):
    """orc_simulate_errors: custom_long.rs:187-232 with the alias picks supplied; kmers = [(ref_code, [(alt_code, w)])]."""
    a = np.frombuffer(bytes(seq), np.uint8)
    ref = np.array([r for r, _ in kmers], np.uint32)
    off = np.zeros(len(kmers) + 1, np.uint32)
    off[1:] = np.cumsum([len(al) for _, al in kmers])
    alts = np.array([c for _, al in kmers for c, _ in al], np.uint32)
    ch = np.ascontiguousarray(choices, np.uint32)
    out = np.empty(len(a) + 16, np.uint8)
    n = lib().orc_simulate_errors(_p(a, u8p), C.c_uint64(len(a)), int(k), _p(ref, u32p), _p(off, u32p), _p(alts, u32p),
                                  len(kmers), _p(ch, u32p), C.c_uint64(len(ch)), _p(out, u8p), C.c_uint64(len(out)))
    if n < 0:
        raise RuntimeError(err())
    return out[:n].tobytes()


def native_geometry(genomes, profile, reads_per_genome, seed, unit_begin=0, unit_end=0):
    r = np.ascontiguousarray(reads_per_genome, np.uint64)
    n = (unit_end - unit_begin) if (unit_begin or unit_end) else total_units(profile, r)
    o = _opts(seed, None, unit_begin, unit_end, 0)
    g, c, rl, is_ = (np.zeros(n, np.uint32) for _ in range(4))
    st, lo = (np.zeros(n, np.uint64) for _ in range(2))
    if lib().orc_native_geometry(C.byref(genomes.c), C.byref(profile.c), _p(r, u64p), C.byref(o), _p(g, u32p),
                                 _p(c, u32p), _p(rl, u32p), _p(is_, u32p), _p(st, u64p), _p(lo, u64p)) != 0:
        raise RuntimeError(err())
    return dict(genome=g, contig=c, rl=rl, is_=is_, start=st, lo_or_end=lo)


class Draws:
    """Replay draws as numpy arrays."""
    FIELDS = [("genome", np.uint32, u32p), ("contig", np.uint32, u32p), ("read_length", np.uint16, u16p),
              ("insert_size", np.uint16, u16p), ("start", np.uint64, u64p), ("redraw", np.uint64, u64p),
              ("base_off", np.uint64, u64p), ("phred_z", np.float32, f32p), ("phred_cand", np.uint32, u32p),
              ("phred_u", np.float64, f64p), ("phred_inbin", np.uint32, u32p), ("mut_u", np.float32, f32p),
              ("mut_k", np.uint8, u8p)]

    def __init__(self, **kw):
        self.n_units = len(kw["genome"])
        for name, dt, _ in self.FIELDS:
            v = kw.get(name)
            setattr(self, name, None if v is None else np.ascontiguousarray(v, dt))
        self.c = OrcDraws(self.n_units, *[_p(getattr(self, n), t) for n, _, t in self.FIELDS])


def replay(genomes, profile, draws, header_format=None, unit_begin=0, unit_end=0, read_id_base=0, cap=None,
           want_stats=False):
    o = _opts(0, header_format, unit_begin, unit_end, read_id_base)
    st = OrcStats()
    if cap is None:
        nb = int(draws.base_off[-1]) if draws.base_off is not None else 0
        cap = max(1 << 16, 2 * nb + int(draws.n_units) * 4096)
    buf = np.empty(cap, np.uint8)
    n = lib().orc_replay(C.byref(genomes.c), C.byref(profile.c), C.byref(draws.c), C.byref(o), _p(buf, u8p),
                         C.c_uint64(cap), C.byref(st))
    if n < 0:
        raise RuntimeError(err() or f"oracle buffer too small ({-n})")
    out = buf[:n].tobytes()
    return (out, st) if want_stats else out


def abundance(kind, total_reads, n, abund_in=None, genome_sizes=None, size_adjusted=False):
    a = None if abund_in is None else np.ascontiguousarray(abund_in, np.float64)
    s = None if genome_sizes is None else np.ascontiguousarray(genome_sizes, np.uint64)
    reads, ab = np.zeros(n, np.uint64), np.zeros(n, np.float64)
    if lib().orc_abundance(kind, int(size_adjusted), C.c_uint64(total_reads), n, _p(a, f64p), _p(s, u64p),
                           _p(reads, u64p), _p(ab, f64p)) != 0:
        raise RuntimeError(err())
    return reads, ab


def load_fasta(path, contiguous=False):
    h = C.c_void_p()
    e = C.create_string_buffer(512)
    if lib().orc_fasta_load(path.encode(), int(contiguous), C.byref(h), e, 512) != 0:
        raise RuntimeError(e.value.decode())
    L = lib()
    n = L.orc_fasta_num_seqs(h)
    seqs = []
    for i in range(n):
        ln = C.c_uint32()
        idp = L.orc_fasta_seq_id(h, i, C.byref(ln))
        sid = C.string_at(idp, ln.value)
        plen = L.orc_fasta_seq_len(h, i)
        seq = C.string_at(L.orc_fasta_seq(h, i), plen)
        seqs.append(dict(id=sid, seq=seq, size=L.orc_fasta_seq_size(h, i)))
    size = L.orc_fasta_size(h)
    L.orc_fasta_free(h)
    return dict(seqs=seqs, size=size, num_seqs=n)


def parse_genome_file(path):
    buf = C.create_string_buffer(1 << 22)
    n = lib().orc_parse_genome_file(path.encode(), buf, C.c_uint64(len(buf)))
    if n < 0:
        raise RuntimeError(err())
    recs = []
    for line in buf.value.decode().split("\n"):
        if not line:
            continue
        p, u, a = line.split("\t")
        recs.append(dict(filepath=p, uuid=u or None, abundance=None if a == "nan" else float(a)))
    return recs


def metadata_tsv(uuids, paths, num_reads, abundances):
    n = len(uuids)
    U = (C.c_char_p * n)(*[u.encode() for u in uuids])
    P = (C.c_char_p * n)(*[p.encode() for p in paths])
    r = np.ascontiguousarray(num_reads, np.uint64)
    a = np.ascontiguousarray(abundances, np.float64)
    buf = C.create_string_buffer(1 << 22)
    k = lib().orc_metadata_tsv(n, U, P, _p(r, u64p), _p(a, f64p), buf, C.c_uint64(len(buf)))
    if k < 0:
        raise RuntimeError(err())
    return buf.value.decode()


def format_f64(v):
    b = C.create_string_buffer(512)
    lib().orc_format_f64(C.c_double(v), b, 512)
    return b.value.decode()


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)

"""Host-side mirror of the reference's interface for the read-simulation path.

Names, argument meaning and error behaviour follow the reference (paths are
relative to /root/reference/):

  error_profiles/{perfect_short,minimal_short,perfect_long,minimal_long,
                  custom_short,custom_long}.rs  ->  *ErrorProfile classes
  abundance_profiles/{uniform,exact,custom}.rs  ->  *AbundanceProfile classes
  genome.rs Genome::from_fasta, main.rs:117-162 ->  Genome / GenomeSet
  simulate.rs simulate_pe_reads / simulate_long_reads (+ fastq.rs
  write_to_fastq)                               ->  Engine.simulate_*, write_to_fastq
  files.rs parse_genome_file / write_metadata   ->  same names

Everything that computes goes through libsimmr_b200.so (include/simmr_b200.h);
this module only marshals numpy buffers.  No CPU fallback exists.
"""
import ctypes as C
import os

import numpy as np

from . import _native as N

DEFAULT_HEADER_FORMAT = ("@{:read_id:}|{:genome_id:}/{:pair:} metadata:sid={:sequence_id:}|sp={:start_position:}"
                         "|ep={:end_position:}|rc={:reverse_complement:}")   # cli.rs:196


class SimmrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _check(rc, ctx=None):
    if rc < 0:
        raise SimmrError(rc, N.lib().simmr_last_error(ctx).decode(errors="replace"))
    return rc


# --------------------------------------------------------------------------
# error profiles — trait ErrorProfile (error_profiles/base.rs:6-32) as data
# --------------------------------------------------------------------------
class ErrorProfile:
    kind = None

    def __init__(self, read_length=150, insert_size=150, mean_phred_score=30, blob=None, flags=0):
        self.read_length, self.insert_size, self.mean_phred_score = read_length, insert_size, mean_phred_score
        self.flags = flags
        self._blob = None if blob is None else np.frombuffer(bytes(blob), np.uint8).copy()

    def desc(self):
        return N.EProfileDesc(self.kind, self.read_length, self.insert_size, self.mean_phred_score,
                              (C.c_uint8 * 3)(), self.flags, _p(self._blob, N.u8p),
                              0 if self._blob is None else len(self._blob))

    def minimum_genome_size(self):
        """ErrorProfile::minimum_genome_size (base.rs:29)."""
        out = C.c_uint16()
        d = self.desc()
        _check(N.lib().simmr_eprofile_minimum_genome_size(C.byref(d), C.byref(out)))
        return out.value

    def is_long_read(self):
        """ErrorProfile::is_long_read (base.rs:31)."""
        out = C.c_int()
        d = self.desc()
        _check(N.lib().simmr_eprofile_is_long_read(C.byref(d), C.byref(out)))
        return bool(out.value)

    def phred_table(self):
        kind, cb, npos, cq = (C.c_uint32() for _ in range(4))
        ent = np.zeros(1 << 20, np.uint32)
        d = self.desc()
        n = _check(N.lib().simmr_debug_phred_table(C.byref(d), C.byref(kind), C.byref(cb), C.byref(npos), C.byref(cq),
                                                   _p(ent, N.u32p), len(ent)))
        return dict(kind=kind.value, cb=cb.value, npos=npos.value, const_q=cq.value, entries=ent[:n].copy())

    def len_table(self, which):
        kind, cb = C.c_uint32(), C.c_uint32()
        ent, lo, hi = (np.zeros(1 << 13, np.uint32) for _ in range(3))
        d = self.desc()
        n = _check(N.lib().simmr_debug_len_table(C.byref(d), which, C.byref(kind), C.byref(cb), _p(ent, N.u32p),
                                                 _p(lo, N.u32p), _p(hi, N.u32p), len(ent)))
        m = max(n, 1)
        return dict(kind=kind.value, cb=cb.value, entries=ent[:n].copy(), lo=lo[:m].copy(), hi=hi[:m].copy())


    def err_table(self):
        """Alt-k-mer sampling tables of a custom-long blob, flattened; k == 0 when the blob has none."""
        k, n = C.c_uint32(), C.c_uint32()
        words = np.zeros(1 << 24, np.uint32)
        d = self.desc()
        m = _check(N.lib().simmr_debug_err_table(C.byref(d), C.byref(k), C.byref(n), _p(words, N.u32p), len(words)))
        return dict(k=k.value, n_ref=n.value, words=words[:m].copy())


class PerfectShortErrorProfile(ErrorProfile):      # perfect_short.rs
    kind = N.EP_PERFECT_SHORT

    def __init__(self, read_length=150, insert_size=150):
        super().__init__(read_length, insert_size)


class MinimalShortErrorProfile(ErrorProfile):      # minimal_short.rs; sigmas fixed by cli.rs:239-240
    kind = N.EP_MINIMAL_SHORT

    def __init__(self, read_length=150, insert_size=150, mean_phred_score=30):
        super().__init__(read_length, insert_size, mean_phred_score)


class PerfectLongErrorProfile(ErrorProfile):       # perfect_long.rs
    kind = N.EP_PERFECT_LONG

    def __init__(self, compat_start=False):
        super().__init__(flags=N.FLAG_COMPAT_LONG_START if compat_start else 0)


class MinimalLongErrorProfile(ErrorProfile):       # minimal_long.rs
    kind = N.EP_MINIMAL_LONG

    def __init__(self, mean_phred_score=30, compat_start=False):
        super().__init__(mean_phred_score=mean_phred_score, flags=N.FLAG_COMPAT_LONG_START if compat_start else 0)


class CustomShortErrorProfile(ErrorProfile):       # custom_short.rs, blob = encoding.rs:268-281
    kind = N.EP_CUSTOM_SHORT

    def __init__(self, blob):
        super().__init__(blob=blob)


class CustomLongErrorProfile(ErrorProfile):        # custom_long.rs / SURVEY §A.5
    kind = N.EP_CUSTOM_LONG

    def __init__(self, blob, compat_start=False):
        super().__init__(blob=blob, flags=N.FLAG_COMPAT_LONG_START if compat_start else 0)


# --------------------------------------------------------------------------
# abundance profiles — trait AbundanceProfile (abundance_profiles/base.rs:10-69)
# --------------------------------------------------------------------------
class AbundanceProfile:
    kind = None

    def __init__(self, size_adjusted=False, abundances=None):
        self.size_adjusted = bool(size_adjusted)
        self.abundances = None if abundances is None else np.ascontiguousarray(abundances, np.float64)

    def is_size_aware(self):
        return self.size_adjusted and self.kind != N.AP_EXACT

    def _counts(self, total_reads, n, sizes, adjusted):
        d = N.AProfileDesc(self.kind, int(adjusted), _p(self.abundances, N.f64p))
        s = None if sizes is None else np.ascontiguousarray(sizes, np.uint64)
        reads, ab = np.zeros(n, np.uint64), np.zeros(n, np.float64)
        _check(N.lib().simmr_abundance_counts(C.byref(d), total_reads, n, _p(s, N.u64p), _p(reads, N.u64p), _p(ab, N.f64p)))
        return reads, ab

    def determine_abundances(self, total_reads, num_genomes):
        """-> (reads_per_genome, abundance) arrays; base.rs:27-37."""
        return self._counts(total_reads, num_genomes, None, False)

    def reads_per_genome(self, total_reads, genome_sizes):
        """determine_abundances followed by adjust_for_size when size aware (simulate.rs:121-132)."""
        return self._counts(total_reads, len(genome_sizes), genome_sizes, self.is_size_aware())


class UniformAbundanceProfile(AbundanceProfile):   # uniform.rs
    kind = N.AP_UNIFORM


class ExactAbundanceProfile(AbundanceProfile):     # exact.rs
    kind = N.AP_EXACT

    def __init__(self):
        super().__init__(False)


class CustomAbundanceProfile(AbundanceProfile):    # custom.rs
    kind = N.AP_CUSTOM

    def __init__(self, abundances, size_adjusted=False):
        super().__init__(size_adjusted, abundances)


# --------------------------------------------------------------------------
# genomes — genome.rs, main.rs:117-162
# --------------------------------------------------------------------------
class GenomeSet:
    """Flat, device-shaped genome set (simmr_genome_set)."""

    def __init__(self, bases, contig_off, contig_size, genome_contig, contig_ids, contig_id_off, genome_ids,
                 genome_id_off, genome_sizes=None, paths=None):
        self.bases = np.ascontiguousarray(bases, np.uint8)
        self.contig_off = np.ascontiguousarray(contig_off, np.uint64)
        self.contig_size = np.ascontiguousarray(contig_size, np.uint64)
        self.genome_contig = np.ascontiguousarray(genome_contig, np.uint32)
        self.contig_ids = bytes(contig_ids)
        self.contig_id_off = np.ascontiguousarray(contig_id_off, np.uint32)
        self.genome_ids = bytes(genome_ids)
        self.genome_id_off = np.ascontiguousarray(genome_id_off, np.uint32)
        self.n_genomes, self.n_contigs = len(self.genome_contig) - 1, len(self.contig_size)
        self.genome_sizes = (np.ascontiguousarray(genome_sizes, np.uint64) if genome_sizes is not None else
                             np.array([self.contig_size[self.genome_contig[g]:self.genome_contig[g + 1]].sum()
                                       for g in range(self.n_genomes)], np.uint64))
        self.paths = list(paths) if paths is not None else [""] * self.n_genomes
        self._cid = np.frombuffer(self.contig_ids + b"\0", np.uint8).copy()
        self._gid = np.frombuffer(self.genome_ids + b"\0", np.uint8).copy()

    def as_c(self):
        return N.GenomeSetC(self.n_genomes, self.n_contigs, _p(self.bases, N.u8p), _p(self.contig_off, N.u64p),
                            _p(self.contig_size, N.u64p), _p(self.genome_contig, N.u32p),
                            self._cid.ctypes.data, _p(self.contig_id_off, N.u32p),
                            self._gid.ctypes.data, _p(self.genome_id_off, N.u32p))

    def uuid(self, g):
        return self.genome_ids[self.genome_id_off[g]:self.genome_id_off[g + 1]].decode(errors="replace")

    def contig_id(self, c):
        return self.contig_ids[self.contig_id_off[c]:self.contig_id_off[c + 1]]

    def contig(self, c):
        return self.bases[int(self.contig_off[c]):int(self.contig_off[c + 1])]

    @classmethod
    def from_contigs(cls, genomes):
        """genomes: list of (uuid, [(contig_id bytes, bases bytes/ndarray), ...])."""
        bases, coff, csize, gcon, cids, cidoff, gids, gidoff = [], [0], [], [0], b"", [0], b"", [0]
        for uuid, contigs in genomes:
            for cid, seq in contigs:
                a = np.frombuffer(seq, np.uint8) if isinstance(seq, (bytes, bytearray)) else np.asarray(seq, np.uint8)
                bases.append(a)
                coff.append(coff[-1] + len(a))
                csize.append(len(a))
                cids += cid if isinstance(cid, bytes) else cid.encode()
                cidoff.append(len(cids))
            gcon.append(len(csize))
            gids += uuid if isinstance(uuid, bytes) else uuid.encode()
            gidoff.append(len(gids))
        allb = np.concatenate(bases) if bases else np.zeros(0, np.uint8)
        return cls(allb, coff, csize, gcon, cids, cidoff, gids, gidoff)

    @classmethod
    def from_fasta(cls, paths, uuids=None, contiguous=False, min_size=None, threads=0):
        """Genome::from_fasta per path (genome.rs:89-162) then, when `min_size` is
        given, the contig filter of main.rs:117-162.  Returns (set, kept mask)."""
        L = N.lib()
        b = C.c_void_p()
        _check(L.simmr_builder_create(C.byref(b)))
        try:
            n = len(paths)
            P = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
            U = (C.c_char_p * n)(*[None if uuids is None or uuids[i] is None else uuids[i].encode() for i in range(n)])
            rc = L.simmr_builder_add_fastas(b, P, U, n, int(contiguous), int(threads))      # parallel over files and chunks
            if rc < 0:
                raise SimmrError(rc, L.simmr_builder_error(b).decode(errors="replace"))
            sizes = [L.simmr_builder_genome_size(b, g) for g in range(len(paths))]
            kept = np.ones(len(paths), np.uint8)
            if min_size is not None and not contiguous:        # main.rs:117 `if !args.contiguous`
                _check(L.simmr_builder_filter(b, min_size, _p(kept, N.u8p)))
            v = N.GenomeSetC()
            _check(L.simmr_builder_view(b, C.byref(v)))
            ng, nc = v.n_genomes, v.n_contigs
            arr = lambda ptr, n, dt: np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)
            coff = arr(v.contig_off, nc + 1, np.uint64)
            cidoff = arr(v.contig_id_off, nc + 1, np.uint32)
            gidoff = arr(v.genome_id_off, ng + 1, np.uint32)
            gs = cls(arr(v.bases, int(coff[-1]), np.uint8), coff, arr(v.contig_size, nc, np.uint64),
                     arr(v.genome_contig, ng + 1, np.uint32), C.string_at(v.contig_ids, int(cidoff[-1])), cidoff,
                     C.string_at(v.genome_ids, int(gidoff[-1])), gidoff,
                     genome_sizes=[s for s, k in zip(sizes, kept) if k], paths=[p for p, k in zip(paths, kept) if k])
            return gs, kept.astype(bool)
        finally:
            L.simmr_builder_destroy(b)


# --------------------------------------------------------------------------
# files.rs
# --------------------------------------------------------------------------
def parse_genome_file(path):
    """files.rs:56-86 -> list of dict(filepath, uuid, abundance)."""
    buf = C.create_string_buffer(1 << 24)
    n = N.lib().simmr_parse_genome_file(os.fsencode(path), buf, len(buf))
    _check(n)
    recs = []
    for line in buf.value.decode().split("\n"):
        if line:
            p, u, a = line.split("\t")
            recs.append(dict(filepath=p, uuid=u or None, abundance=None if a == "nan" else float(a)))
    return recs


def write_metadata(metadata, output):
    """files.rs:100-134; metadata = [(uuid, path, num_reads, abundance)]."""
    n = len(metadata)
    U = (C.c_char_p * n)(*[m[0].encode() for m in metadata])
    P = (C.c_char_p * n)(*[m[1].encode() for m in metadata])
    r = np.array([m[2] for m in metadata], np.uint64)
    a = np.array([m[3] for m in metadata], np.float64)
    _check(N.lib().simmr_write_metadata(os.fsencode(output), n, U, P, _p(r, N.u64p), _p(a, N.f64p)))


# --------------------------------------------------------------------------
# replay draws
# --------------------------------------------------------------------------
class Draws:
    FIELDS = [("genome", np.uint32, N.u32p), ("contig", np.uint32, N.u32p), ("read_length", np.uint16, N.u16p),
              ("insert_size", np.uint16, N.u16p), ("start", np.uint64, N.u64p), ("redraw", np.uint64, N.u64p),
              ("base_off", np.uint64, N.u64p), ("phred_z", np.float32, N.f32p), ("phred_cand", np.uint32, N.u32p),
              ("phred_u", np.float64, N.f64p), ("phred_inbin", np.uint32, N.u32p), ("mut_u", np.float32, N.f32p),
              ("mut_k", np.uint8, N.u8p)]

    def __init__(self, **kw):
        self.n_units = len(kw["genome"])
        for name, dt, _ in self.FIELDS:
            v = kw.get(name)
            setattr(self, name, None if v is None else np.ascontiguousarray(v, dt))

    def as_c(self):
        return N.DrawsC(self.n_units, *[_p(getattr(self, n), t) for n, _, t in self.FIELDS])


# --------------------------------------------------------------------------
# the hot path
# --------------------------------------------------------------------------
class Engine:
    """One B200 context (simmr_ctx): device-resident genomes + profile."""

    def __init__(self, device=0):
        self._ctx = C.c_void_p()
        _check(N.lib().simmr_ctx_create(device, C.byref(self._ctx)))
        self.genomes = None
        self.profile = None

    def close(self):
        if self._ctx:
            N.lib().simmr_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
        _check(N.lib().simmr_device_info(self._ctx, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)), self._ctx)
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), hbm_bytes=mem.value)

    def load_genomes(self, gs):
        c = gs.as_c()
        _check(N.lib().simmr_genomes_load(self._ctx, C.byref(c)), self._ctx)
        self.genomes = gs

    def set_error_profile(self, prof):
        d = prof.desc()
        _check(N.lib().simmr_error_profile_set(self._ctx, C.byref(d)), self._ctx)
        self.profile = prof

    def total_units(self, reads_per_genome):
        r = np.ascontiguousarray(reads_per_genome, np.uint64)
        return N.lib().simmr_total_units(self._ctx, _p(r, N.u64p))

    def stats(self):
        s = N.Stats()
        _check(N.lib().simmr_stats_get(self._ctx, C.byref(s)), self._ctx)
        return {k: getattr(s, k) for k, _ in N.Stats._fields_}

    def _opts(self, seed, header_format, unit_begin, unit_end, read_id_base, batch_units):
        hf = header_format.encode() if isinstance(header_format, str) else header_format
        explicit = unit_end is not None          # (0, 0) is then an empty range, not "everything" (an empty first shard)
        return N.RunOpts(0 if seed is None else seed, int(seed is not None), read_id_base, hf, unit_begin if explicit else 0,
                         unit_end if explicit else 0, batch_units, N.RUN_EXPLICIT_RANGE if explicit else 0, 0)

    def _run(self, fn, first, opts, sink):
        keep = None
        if sink is None:
            cb = None
        else:
            def tramp(_user, data, length, first_unit, n_units):
                try:
                    r = sink(np.ctypeslib.as_array(data, shape=(length,)) if length else np.zeros(0, np.uint8),
                             first_unit, n_units)
                    return int(r or 0)
                except Exception as e:   # surface Python errors as a sink failure
                    tramp.err = e
                    return 1
            tramp.err = None
            keep = N.SINK_FN(tramp)
            cb = C.cast(keep, C.c_void_p)
        rc = fn(self._ctx, first, C.byref(opts), cb, None)
        if sink is not None and keep is not None and tramp.err is not None:
            raise tramp.err
        self._check_rc(rc)

    def _check_rc(self, rc):
        return _check(rc, self._ctx)

    def simulate(self, reads_per_genome, seed=None, sink=None, header_format=None, unit_begin=0, unit_end=None,
                 read_id_base=0, batch_units=0):
        """simmr_simulate: native Philox mode.  sink(chunk ndarray, first_unit, n_units);
        sink=None keeps the FASTQ bytes in device staging (device-side throughput)."""
        r = np.ascontiguousarray(reads_per_genome, np.uint64)
        o = self._opts(seed, header_format, unit_begin, unit_end, read_id_base, batch_units)
        self._run(N.lib().simmr_simulate, _p(r, N.u64p), o, sink)
        return self.stats()

    def simulate_device(self, reads_per_genome, seed, device_sink, header_format=None, unit_begin=0, unit_end=None,
                        read_id_base=0, batch_units=0):
        """simmr_simulate_device: device_sink(device_pointer int, n_bytes, first_unit, n_units) sees every batch's FASTQ
        bytes in HBM (no copy to the host); the pointer is valid until the callback returns."""
        r = np.ascontiguousarray(reads_per_genome, np.uint64)
        o = self._opts(seed, header_format, unit_begin, unit_end, read_id_base, batch_units)
        err = []

        def tramp(_user, ptr, length, first_unit, n_units):
            try:
                return int(device_sink(ptr or 0, length, first_unit, n_units) or 0)
            except Exception as e:
                err.append(e)
                return 1
        cb = N.DEVICE_SINK_FN(tramp)
        rc = N.lib().simmr_simulate_device(self._ctx, _p(r, N.u64p), C.byref(o), C.cast(cb, C.c_void_p), None)
        if err:
            raise err[0]
        self._check_rc(rc)
        return self.stats()

    def measure(self, reads_per_genome, seed, header_format=None, unit_begin=0, unit_end=None, read_id_base=0):
        """simmr_measure: (FASTQ bytes, records) a unit range will produce; geometry kernel only."""
        r = np.ascontiguousarray(reads_per_genome, np.uint64)
        o = self._opts(seed, header_format, unit_begin, unit_end, read_id_base, 0)
        nb, nr = C.c_uint64(), C.c_uint64()
        _check(N.lib().simmr_measure(self._ctx, _p(r, N.u64p), C.byref(o), C.byref(nb), C.byref(nr)), self._ctx)
        return nb.value, nr.value

    def simulate_bytes(self, reads_per_genome, seed=None, **kw):
        chunks = []
        self.simulate(reads_per_genome, seed, sink=lambda a, f, n: chunks.append(a.tobytes()), **kw)
        return b"".join(chunks)

    def replay(self, draws, sink=None, header_format=None, unit_begin=0, unit_end=None, read_id_base=0, batch_units=0):
        d = draws.as_c()
        o = self._opts(0, header_format, unit_begin, unit_end, read_id_base, batch_units)
        self._run(N.lib().simmr_replay, C.byref(d), o, sink)
        return self.stats()

    def replay_bytes(self, draws, **kw):
        chunks = []
        self.replay(draws, sink=lambda a, f, n: chunks.append(a.tobytes()), **kw)
        return b"".join(chunks)

    # ---- reference-shaped entry points (simulate.rs:110-150, 323-406) ----
    def _simulate_reads(self, num_reads, genomes, error_profile, abundance_profile, seed, want_long, **kw):
        if error_profile.is_long_read() != want_long:
            raise SimmrError(N.E_INVALID, "error profile does not match the read type")
        if genomes is not self.genomes:
            self.load_genomes(genomes)
        if error_profile is not self.profile:
            self.set_error_profile(error_profile)
        reads, abund = abundance_profile.reads_per_genome(num_reads, genomes.genome_sizes)
        out = kw.pop("output", None)
        if out is None:
            data = self.simulate_bytes(reads, seed, **kw)
        else:
            data = None
            write_to_fastq(self, reads, out, seed=seed, **kw)
        meta = [(genomes.uuid(g), genomes.paths[g], int(reads[g]), float(abund[g])) for g in range(genomes.n_genomes)]
        return meta, data

    def simulate_pe_reads(self, num_reads, genomes, error_profile, abundance_profile, seed=None, **kw):
        """simulate.rs:110-150.  Returns ([(uuid, path, n_reads, abundance)], FASTQ bytes)."""
        return self._simulate_reads(num_reads, genomes, error_profile, abundance_profile, seed, False, **kw)

    def simulate_long_reads(self, num_reads, genomes, error_profile, abundance_profile, seed=None, **kw):
        """simulate.rs:323-406."""
        return self._simulate_reads(num_reads, genomes, error_profile, abundance_profile, seed, True, **kw)


class MultiEngine(Engine):
    """simmr_multi: one context per listed device behind one handle.  A run's batches are dealt round-robin to the
    devices; a single writer thread inside the library hands them to the sink in file order, so the bytes are those of a
    one-GPU run."""

    def __init__(self, devices):
        self._ctx = C.c_void_p()
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        rc = N.lib().simmr_multi_create(arr, len(self.devices), C.byref(self._ctx))
        if rc < 0:
            raise SimmrError(rc, N.lib().simmr_multi_last_error(None).decode(errors="replace"))
        self.genomes = None
        self.profile = None

    def _mcheck(self, rc):
        if rc < 0:
            raise SimmrError(rc, N.lib().simmr_multi_last_error(self._ctx).decode(errors="replace"))
        return rc

    _check_rc = _mcheck

    def close(self):
        if self._ctx:
            N.lib().simmr_multi_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def load_genomes(self, gs):
        c = gs.as_c()
        self._mcheck(N.lib().simmr_multi_genomes_load(self._ctx, C.byref(c)))
        self.genomes = gs

    def set_error_profile(self, prof):
        d = prof.desc()
        self._mcheck(N.lib().simmr_multi_error_profile_set(self._ctx, C.byref(d)))
        self.profile = prof

    def total_units(self, reads_per_genome):
        r = np.ascontiguousarray(reads_per_genome, np.uint64)
        return N.lib().simmr_multi_total_units(self._ctx, _p(r, N.u64p))

    def stats(self):
        s = N.Stats()
        self._mcheck(N.lib().simmr_multi_stats_get(self._ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in N.Stats._fields_}

    def simulate(self, reads_per_genome, seed=None, sink=None, header_format=None, unit_begin=0, unit_end=None,
                 read_id_base=0, batch_units=0):
        r = np.ascontiguousarray(reads_per_genome, np.uint64)
        o = self._opts(seed, header_format, unit_begin, unit_end, read_id_base, batch_units)
        self._run(N.lib().simmr_multi_simulate, _p(r, N.u64p), o, sink)
        return self.stats()

    def measure(self, *a, **k):
        raise NotImplementedError("a multi-device run needs no shard offsets: its writer thread emits batches in file order")

    def replay(self, *a, **k):
        raise NotImplementedError("replay is a single-device (test) path")

    def simulate_device(self, *a, **k):
        raise NotImplementedError("a device consumer binds to one device: open an Engine on it")

    def device_info(self):
        raise NotImplementedError("per-device information: open an Engine on the device")


def copy_from_device(ptr, nbytes):
    """The bytes at a device pointer as handed to a device sink (simmr_simulate_device) -> numpy array."""
    from cuda.bindings import runtime as cudart
    out = np.empty(nbytes, np.uint8)
    if nbytes:
        err, = cudart.cudaMemcpy(out.ctypes.data, ptr, nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        if int(err) != 0:
            raise SimmrError(N.E_CUDA, f"cudaMemcpy from the device sink's pointer failed ({err})")
    return out


def write_to_fastq(engine, reads_per_genome, output, seed=None, header_format=None, append=False, **kw):
    """fastq.rs:14-124 fused with the simulation: streams FASTQ chunks to `output`."""
    with open(output, "ab" if append else "wb") as f:
        engine.simulate(reads_per_genome, seed, sink=lambda a, fu, n: f.write(a.data) and 0,
                        header_format=header_format, **kw)
    return engine.stats()

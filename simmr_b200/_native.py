"""ctypes declarations for libsimmr_b200.so (include/simmr_b200.h).

The library is the product; this module only loads it and fails loudly when it
is missing.  There is no Python or CPU fallback for the compute entry points.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIMMR_B200_LIB") or os.path.join(_HERE, "libsimmr_b200.so")   # the override loads another build of the same library (kernel A/B runs)

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)

OK, E_INVALID, E_CUDA, E_IO, E_PARSE, E_STATE, E_TOO_SMALL, E_SINK, E_LIMIT = 0, -1, -2, -3, -4, -5, -6, -7, -8
EP_PERFECT_SHORT, EP_MINIMAL_SHORT, EP_PERFECT_LONG, EP_MINIMAL_LONG, EP_CUSTOM_SHORT, EP_CUSTOM_LONG = range(6)
AP_EXACT, AP_UNIFORM, AP_CUSTOM = range(3)
FLAG_COMPAT_LONG_START = 1


class GenomeSetC(C.Structure):
    _fields_ = [("n_genomes", C.c_uint32), ("n_contigs", C.c_uint32), ("bases", u8p),
                ("contig_off", u64p), ("contig_size", u64p), ("genome_contig", u32p),
                ("contig_ids", C.c_void_p), ("contig_id_off", u32p),
                ("genome_ids", C.c_void_p), ("genome_id_off", u32p)]


class EProfileDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("read_length", C.c_uint16), ("insert_size", C.c_uint16),
                ("mean_phred", C.c_uint8), ("pad_", C.c_uint8 * 3), ("flags", C.c_uint32),
                ("blob", u8p), ("blob_len", C.c_uint64)]


class AProfileDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("size_adjusted", C.c_int32), ("abundances", f64p)]


class RunOpts(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("has_seed", C.c_int32), ("read_id_base", C.c_uint32),
                ("header_format", C.c_char_p), ("unit_begin", C.c_uint64), ("unit_end", C.c_uint64),
                ("batch_units", C.c_uint64), ("flags", C.c_uint32), ("pad_", C.c_uint32)]


RUN_EXPLICIT_RANGE = 1
SINK_RETAINED = 2           # a sink's return value: it keeps the chunk and gives it back with simmr_chunk_release


class Stats(C.Structure):
    _fields_ = [("units", C.c_uint64), ("records", C.c_uint64), ("bases", C.c_uint64), ("bytes", C.c_uint64),
                ("clamped", C.c_uint64), ("algorithmic_bytes", C.c_uint64), ("device_ms", C.c_double),
                ("format_ms", C.c_double), ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64)]


class DrawsC(C.Structure):
    _fields_ = [("n_units", C.c_uint64), ("genome", u32p), ("contig", u32p), ("read_length", u16p),
                ("insert_size", u16p), ("start", u64p), ("redraw", u64p), ("base_off", u64p),
                ("phred_z", f32p), ("phred_cand", u32p), ("phred_u", f64p), ("phred_inbin", u32p),
                ("mut_u", f32p), ("mut_k", u8p)]


SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, u8p, C.c_uint64, C.c_uint64, C.c_uint64)
DEVICE_SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64)

# every symbol include/simmr_b200.h declares
EXPORTS = [
    "simmr_ctx_create", "simmr_ctx_destroy", "simmr_last_error", "simmr_device_info",
    "simmr_genomes_load", "simmr_error_profile_set", "simmr_eprofile_minimum_genome_size",
    "simmr_eprofile_is_long_read", "simmr_abundance_counts", "simmr_simulate", "simmr_simulate_device", "simmr_replay",
    "simmr_total_units", "simmr_measure", "simmr_stats_get", "simmr_builder_create", "simmr_builder_destroy",
    "simmr_builder_add_fasta", "simmr_builder_add_fastas", "simmr_builder_filter", "simmr_builder_view", "simmr_builder_genome_size",
    "simmr_builder_error", "simmr_parse_genome_file", "simmr_write_metadata", "simmr_debug_phred_table",
    "simmr_debug_len_table", "simmr_debug_err_table", "simmr_debug_philox4x32_10", "simmr_version",
    "simmr_multi_create", "simmr_multi_destroy", "simmr_multi_genomes_load", "simmr_multi_error_profile_set",
    "simmr_gzip_member", "simmr_gzip_bound", "simmr_fastq_split_mates",
    "simmr_chunk_release", "simmr_multi_total_units", "simmr_multi_simulate", "simmr_multi_stats_get", "simmr_multi_last_error",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C simmr_b200/csrc`).  simmr_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.simmr_last_error.restype = C.c_char_p
        L.simmr_last_error.argtypes = [C.c_void_p]
        L.simmr_version.restype = C.c_char_p
        L.simmr_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.simmr_ctx_destroy.argtypes = [C.c_void_p]
        L.simmr_ctx_destroy.restype = None
        L.simmr_device_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), u64p]
        L.simmr_genomes_load.argtypes = [C.c_void_p, C.POINTER(GenomeSetC)]
        L.simmr_error_profile_set.argtypes = [C.c_void_p, C.POINTER(EProfileDesc)]
        L.simmr_eprofile_minimum_genome_size.argtypes = [C.POINTER(EProfileDesc), u16p]
        L.simmr_eprofile_is_long_read.argtypes = [C.POINTER(EProfileDesc), C.POINTER(C.c_int)]
        L.simmr_abundance_counts.argtypes = [C.POINTER(AProfileDesc), C.c_uint64, C.c_uint32, u64p, u64p, f64p]
        L.simmr_simulate.argtypes = [C.c_void_p, u64p, C.POINTER(RunOpts), C.c_void_p, C.c_void_p]
        L.simmr_simulate_device.argtypes = [C.c_void_p, u64p, C.POINTER(RunOpts), C.c_void_p, C.c_void_p]
        L.simmr_replay.argtypes = [C.c_void_p, C.POINTER(DrawsC), C.POINTER(RunOpts), C.c_void_p, C.c_void_p]
        L.simmr_measure.argtypes = [C.c_void_p, u64p, C.POINTER(RunOpts), u64p, u64p]
        L.simmr_total_units.argtypes = [C.c_void_p, u64p]
        L.simmr_total_units.restype = C.c_uint64
        L.simmr_stats_get.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.simmr_gzip_member.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64, C.c_int]
        L.simmr_gzip_member.restype = C.c_int64
        L.simmr_gzip_bound.argtypes = [C.c_uint64]
        L.simmr_gzip_bound.restype = C.c_uint64
        L.simmr_fastq_split_mates.argtypes = [u8p, C.c_uint64, u8p, u64p, u8p, u64p]
        L.simmr_chunk_release.argtypes = [C.c_void_p]
        L.simmr_chunk_release.restype = None
        L.simmr_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        L.simmr_multi_destroy.argtypes = [C.c_void_p]
        L.simmr_multi_destroy.restype = None
        L.simmr_multi_genomes_load.argtypes = [C.c_void_p, C.POINTER(GenomeSetC)]
        L.simmr_multi_error_profile_set.argtypes = [C.c_void_p, C.POINTER(EProfileDesc)]
        L.simmr_multi_total_units.argtypes = [C.c_void_p, u64p]
        L.simmr_multi_total_units.restype = C.c_uint64
        L.simmr_multi_simulate.argtypes = [C.c_void_p, u64p, C.POINTER(RunOpts), C.c_void_p, C.c_void_p]
        L.simmr_multi_stats_get.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.simmr_multi_last_error.argtypes = [C.c_void_p]
        L.simmr_multi_last_error.restype = C.c_char_p
        L.simmr_builder_create.argtypes = [C.POINTER(C.c_void_p)]
        L.simmr_builder_destroy.argtypes = [C.c_void_p]
        L.simmr_builder_destroy.restype = None
        L.simmr_builder_add_fasta.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        L.simmr_builder_add_fastas.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_uint32, C.c_int, C.c_int]
        L.simmr_builder_filter.argtypes = [C.c_void_p, C.c_uint16, u8p]
        L.simmr_builder_view.argtypes = [C.c_void_p, C.POINTER(GenomeSetC)]
        L.simmr_builder_genome_size.argtypes = [C.c_void_p, C.c_uint32]
        L.simmr_builder_genome_size.restype = C.c_uint64
        L.simmr_builder_error.argtypes = [C.c_void_p]
        L.simmr_builder_error.restype = C.c_char_p
        L.simmr_parse_genome_file.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
        L.simmr_parse_genome_file.restype = C.c_int64
        L.simmr_write_metadata.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), u64p, f64p]
        L.simmr_debug_phred_table.argtypes = [C.POINTER(EProfileDesc), u32p, u32p, u32p, u32p, u32p, C.c_uint64]  # kind, cb, npos, const_q
        L.simmr_debug_len_table.argtypes = [C.POINTER(EProfileDesc), C.c_int, u32p, u32p, u32p, u32p, u32p, C.c_uint64]
        L.simmr_debug_err_table.argtypes = [C.POINTER(EProfileDesc), u32p, u32p, u32p, C.c_uint64]
        L.simmr_debug_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.simmr_debug_philox4x32_10.restype = None
        _lib = L
    return _lib

"""simmr_b200 — B200-native (sm_100a) replacement for the read-simulation hot
path of genomicsoup/simmr.  The product is libsimmr_b200.so behind the C ABI in
include/simmr_b200.h; this package is the thin host mirror of the reference's
interface used by tests, bench.py and Python callers."""
from . import _native
from .api import (DEFAULT_HEADER_FORMAT, SimmrError, ErrorProfile, PerfectShortErrorProfile,
                  MinimalShortErrorProfile, PerfectLongErrorProfile, MinimalLongErrorProfile,
                  CustomShortErrorProfile, CustomLongErrorProfile, AbundanceProfile, UniformAbundanceProfile,
                  ExactAbundanceProfile, CustomAbundanceProfile, GenomeSet, Draws, Engine, MultiEngine, parse_genome_file,
                  write_metadata, write_to_fastq, copy_from_device)

__all__ = [n for n in dir() if not n.startswith("_")]

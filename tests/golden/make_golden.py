#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference tree (run in the authoring container;
/root/reference does not exist on the GPU box, the tests only read the committed output).

What it writes
  sample.fna, GCF_000005845.2_ASM584v2_genomic.partial.fna
      the two FASTA fixtures of the reference's unit tests
      (simmr/src/tests/data/, used by genome_tests.rs:9 and simulate_tests.rs:22)
  reference_kats.json
      every known-answer vector the reference's own tests hold for the read-simulation
      path, each with the file:line it was taken from.  The script checks
      that every literal really occurs in the cited file before writing it, so a vector
      cannot drift from the reference unnoticed.

The reference is Rust and cannot be executed here (no rustc/cargo), so there are no
"outputs of the reference run here"; the vectors are the literals of its test-suite.
"""
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

KATS = {
    "complement": {  # simmr/src/util.rs:15-27 complement(), per byte
        "cite": ["simmr/src/tests/util_tests.rs:8-50", "shared/src/util.rs:107-138"],
        "cases": [["aacctg", "ttggac"], ["TAGCNNNN", "ATCGNNNN"], ["CaTTagG", "GtAAtcC"]],
    },
    "encode_quality_score": {  # util.rs:46-57
        "cite": ["simmr/src/tests/util_tests.rs:69-81"],
        "cases": [[0, "!"], [1, "\""], [10, "+"], [41, "J"]],
        "vector": [[0, 1, 10, 41], "!\"+J"],
    },
    "phred_f32": {  # util.rs:69-111, exact f32 equality in the reference's asserts
        "cite": ["simmr/src/tests/util_tests.rs:84-109"],
        "probability_to_phred": [[0.1, 10], [0.001, 30], [0.000001, 60]],
        "phred_to_probability": [[10, 0.1], [30, 0.001], [60, 0.000001]],
        "phred_to_accuracy": [[10, 0.9], [30, 0.999], [60, 0.999999]],
        "accuracy_to_phred": [[0.9, 10], [0.999, 30], [0.999999, 60]],
    },
    "genome_from_fasta": {  # genome.rs:89-162
        "cite": ["simmr/src/tests/genome_tests.rs:8-20"],
        "file": "sample.fna", "num_seqs": 2, "size": 320,
        "ids": ["header1", "header2"], "sizes": [160, 160],
    },
    "perfect_short": {  # perfect_short.rs:22-44
        "cite": ["simmr/src/tests/error_profile_tests.rs:8-21"],
        "read_length": 150, "insert_size": 150, "quality_len": 150, "quality": 60,
    },
    "uniform_abundance": {  # uniform.rs:18-35
        "cite": ["simmr/src/tests/abundance_profile_tests.rs:8-30"],
        "total_reads": 100, "genomes": 5, "reads": [20] * 5, "abundance": [20.0] * 5,
    },
    "slices": {  # simulate.rs:261-262 slicing + util.rs:31-37 reverse_complement (tests are #[ignore]d as
        # stale for the RNG; the slice/RC literals are still valid known answers on the fixture)
        "cite": ["simmr/src/tests/simulate_tests.rs:36-45", "simmr/src/tests/simulate_tests.rs:82-91"],
        "file": "GCF_000005845.2_ASM584v2_genomic.partial.fna",
        "forward": [[38, 58, "TGTGGATTAAAAAAAGAGTG"], [8, 28, "ATTCTGACTGCAACGGGCAA"]],
        "reverse_complement": [[78, 98, "TTACTCACGGCAGGTAACCA"], [48, 68, "TGCTATCAGACACTCTTTTT"]],
        "quality": 60,
    },
    "gaussian": {  # custom_short.rs:49-61 / custom_long.rs
        "cite": ["simmr/src/error_profiles/custom_short.rs:552-561", "simmr/src/error_profiles/custom_long.rs:263-272"],
        "x": 4.0, "xs": [9.0, 8.0, 7.0, 6.0, 5.0, 4.0, 3.0, 2.0, 1.0, 0.0], "bandwidth": 0.1,
        "value": 0.3989422804014327,
    },
    "kmer_splice": {  # custom_long.rs:187-232 simulate_errors
        "cite": ["simmr/src/error_profiles/custom_long.rs:299-343"],
        "k": 3, "map": {"ACC": ["CAT"], "ATC": ["ATC"], "TCG": ["TGT"]}, "input": "ACCCG", "output": "CATGT",
    },
    "two_bit_codec": {  # shared/src/encoding.rs two_bit_encode_kmer / two_bit_decode_kmer
        "cite": ["shared/src/encoding.rs:288-314"],
        "cases": [["ACGT", 0xE4], ["AAAAA", 0x0], ["TTATC", 0x1CF], ["GCGCATCT", 0xDC66]],
    },
    "bytes_to_string": {"cite": ["shared/src/util.rs:140-143"], "cases": [["foo", "foo"]]},
    "default_header_format": {  # cli.rs:196
        "cite": ["simmr/src/cli.rs:196"],
        "value": "@{:read_id:}|{:genome_id:}/{:pair:} metadata:sid={:sequence_id:}|sp={:start_position:}"
                 "|ep={:end_position:}|rc={:reverse_complement:}",
    },
}

# literals that must be findable in the cited files
PROBES = {
    "simmr/src/tests/util_tests.rs": ["aacctg", "ttggac", "TAGCNNNN", "ATCGNNNN", "CaTTagG", "GtAAtcC", "b'J'",
                                      "0.000001", "0.999999"],
    "simmr/src/tests/genome_tests.rs": ["header1", "header2", "320", "160"],
    "simmr/src/tests/error_profile_tests.rs": ["150", "*q == 60"],
    "simmr/src/tests/abundance_profile_tests.rs": ["determine_abundances(100, 5)", "20.0"],
    "simmr/src/tests/simulate_tests.rs": ["TGTGGATTAAAAAAAGAGTG", "TTACTCACGGCAGGTAACCA", "ATTCTGACTGCAACGGGCAA",
                                          "TGCTATCAGACACTCTTTTT"],
    "simmr/src/error_profiles/custom_short.rs": ["0.3989422804014327"],
    "simmr/src/error_profiles/custom_long.rs": ["0.3989422804014327", "\"ACCCG\"", "\"CATGT\""],
    "shared/src/encoding.rs": ["0x000000E4", "0x000001CF", "0x0000DC66", "GCGCATCT"],
    "simmr/src/cli.rs": [KATS["default_header_format"]["value"]],
}


def main():
    if not os.path.isdir(REF):
        sys.exit("the reference tree is not mounted; tests/golden/ is committed, nothing to do")
    for rel, needles in PROBES.items():
        text = open(os.path.join(REF, rel), encoding="utf-8").read()
        for n in needles:
            if n not in text:
                sys.exit(f"literal {n!r} not found in {rel}: the golden table is out of date")
    for f in ("sample.fna", "GCF_000005845.2_ASM584v2_genomic.partial.fna"):
        shutil.copyfile(os.path.join(REF, "simmr/src/tests/data", f), os.path.join(HERE, f))
    with open(os.path.join(HERE, "reference_kats.json"), "w") as fh:
        json.dump(KATS, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print("wrote", os.path.join(HERE, "reference_kats.json"))


if __name__ == "__main__":
    main()

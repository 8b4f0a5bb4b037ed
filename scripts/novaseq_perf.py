"""Device-side probe for the NovaSeq-style custom profile (C3 shape, smaller)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers as H
import simmr_b200 as S
rng = np.random.default_rng(0)
gl = [(f"{g:016x}", [(f"g{g:04d}_c000".encode(), np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 5_000_000, dtype=np.uint8)])]) for g in range(20)]
gs = S.GenomeSet.from_contigs(gl)
ab = rng.lognormal(0, 1, 20); ab /= ab.sum()
with S.Engine(0) as eng:
    eng.load_genomes(gs)
    eng.set_error_profile(S.CustomShortErrorProfile(H.novaseq_blob()))
    reads, _ = S.CustomAbundanceProfile(ab).determine_abundances(int(16e6), 20)
    for _ in range(3):
        st = eng.simulate(reads, 42)
    print(f"NovaSeq custom-short: {st['records']} rec {st['bases']} bases device {st['device_ms']:.2f} ms format {st['format_ms']:.2f} ms -> "
          f"{st['bases'] / st['device_ms'] / 1e6:.1f} Gbases/s, alg {st['algorithmic_bytes'] / st['format_ms'] / 1e6:.0f} GB/s fmt-only")

"""Device-side throughput probe for long reads (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simmr_b200 as S
rng = np.random.default_rng(0)
gl = [(f"{g:016x}", [(f"g{g:04d}_c000".encode(), np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 3_000_000, dtype=np.uint8)])]) for g in range(8)]
gs = S.GenomeSet.from_contigs(gl)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 40000
with S.Engine(0) as eng:
    eng.load_genomes(gs)
    for prof in (S.MinimalLongErrorProfile(30), S.PerfectLongErrorProfile()):
        eng.set_error_profile(prof)
        reads, _ = S.UniformAbundanceProfile().determine_abundances(n, 8)
        for _ in range(3):
            st = eng.simulate(reads, 42)
        print(f"{type(prof).__name__}: {st['records']} rec {st['bases']} bases device {st['device_ms']:.2f} ms format {st['format_ms']:.2f} ms -> "
              f"{st['bases'] / st['device_ms'] / 1e6:.1f} Gbases/s, alg {st['algorithmic_bytes'] / st['format_ms'] / 1e6:.0f} GB/s fmt-only")

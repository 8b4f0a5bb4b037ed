"""Kernel A/B probe on the C2 bench job (development aid; bench.py is the contract): the device-resident run without a
consumer, format-kernel milliseconds and algorithmic GB/s in the timed region and with every kernel on one stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
import simmr_b200 as S
gs = bench.c2_genomes()
reads, _ = S.UniformAbundanceProfile().determine_abundances(bench.READS_PER_GPU, bench.N_GENOMES)
with S.Engine(0) as eng:
    eng.load_genomes(gs)
    eng.set_error_profile(S.MinimalShortErrorProfile(150, 150, 30))
    for _ in range(2):
        eng.simulate(reads, 42)
    f = a = 0.0
    for _ in range(3):
        st = eng.simulate(reads, 42); f += st["format_ms"]; a += st["algorithmic_bytes"]
    os.environ["SIMMR_PREP_SERIAL"] = "1"
    st = eng.simulate(reads, 42)
    print(f"{os.environ.get('SIMMR_B200_LIB', 'in-tree'):45s} format {f / 3:.3f} ms/step  timed {a / f / 1e6:.0f} GB/s frac {a / f / 1e6 / 6552.3:.4f}   alone {st['format_ms']:.3f} ms frac {st['algorithmic_bytes'] / st['format_ms'] / 1e6 / 6552.3:.4f}")

"""Device-side probe of the ONT-style custom-long profile with k-mer error injection (C4b shape): 10 M reads, 50 genomes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers as H
import simmr_b200 as S
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
gs = H.config_genomes(50, 2_000_000, 6_000_000, (1, 4), seed=4, min_contig=100_000)
with S.Engine(0) as eng:
    eng.load_genomes(gs)
    for name, prof in (("ONT k=7, 20 alternates", S.CustomLongErrorProfile(H.ont_blob_with_kmers(k=7, n_alts=20)[0])), ("minimal-long", S.MinimalLongErrorProfile(30))):
        eng.set_error_profile(prof)
        reads, _ = S.UniformAbundanceProfile().determine_abundances(n, 50)
        for _ in range(2):
            st = eng.simulate(reads, 42)
        print(f"{name}: {st['records']} reads {st['bases']} bases device {st['device_ms']:.1f} ms format {st['format_ms']:.1f} ms -> {st['bases'] / st['device_ms'] / 1e6:.1f} Gbases/s, launches {st['kernel_launches']}")

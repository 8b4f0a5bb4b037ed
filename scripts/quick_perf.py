"""Quick device-side throughput probe (development aid; bench.py is the contract)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import simmr_b200 as S

def genomes(n, mbp, seed=0):
    rng = np.random.default_rng(seed)
    gl = []
    for g in range(n):
        seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(mbp * 1e6), dtype=np.uint8)]
        gl.append((f"{g:016x}", [(f"g{g:04d}_c000".encode(), seq)]))
    return S.GenomeSet.from_contigs(gl)

def main():
    ng = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    reads_total = int(float(sys.argv[2])) if len(sys.argv) > 2 else 8_000_000
    gs = genomes(ng, 5.0)
    with S.Engine(0) as eng:
        print(eng.device_info())
        t = time.time(); eng.load_genomes(gs); print("load s", time.time() - t)
        which = os.environ.get('QP_PROF', 'both')
        profs = [p for n, p in (('perfect', S.PerfectShortErrorProfile()), ('minimal', S.MinimalShortErrorProfile())) if which in ('both', n)]
        for prof in profs:
            eng.set_error_profile(prof)
            reads, _ = S.UniformAbundanceProfile().determine_abundances(reads_total, ng)
            for it in range(int(os.environ.get('QP_ITERS', '3'))):
                st = eng.simulate(reads, 42)
            gb = st["bases"] / st["device_ms"] / 1e6
            print(f"{type(prof).__name__}: {st['records']} rec {st['bases']} bases device {st['device_ms']:.2f} ms format {st['format_ms']:.2f} ms "
                  f"-> {gb:.1f} Gbases/s, alg {st['algorithmic_bytes'] / st['device_ms'] / 1e6:.0f} GB/s, fmt-only {st['algorithmic_bytes'] / st['format_ms'] / 1e6:.0f} GB/s launches {st['kernel_launches']}")
            if os.environ.get('QP_NOSINK'): continue
            t = time.time(); n = [0]
            st = eng.simulate(reads, 42, unit_begin=0, unit_end=min(2_000_000, eng.total_units(reads)), sink=lambda a, f, k: n.__setitem__(0, n[0] + len(a)))
            dt = time.time() - t
            print(f"  e2e sink: {n[0] / 1e9:.2f} GB in {dt:.2f} s = {n[0] / dt / 1e9:.2f} GB/s host")
main()

"""End-to-end (host buffers in, FASTQ bytes in pinned host memory out) rate on a C2-shaped job, smaller."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import simmr_b200 as S
gs = bench.c2_genomes()
reads, _ = S.UniformAbundanceProfile().determine_abundances(40_000_000, 100)
with S.Engine(0) as eng:
    eng.load_genomes(gs); eng.set_error_profile(S.MinimalShortErrorProfile())
    eng.simulate(reads, 42, unit_end=2_000_000, sink=lambda a, f, n: 0)
    for rep in range(2):
        t = time.perf_counter(); sunk = [0]
        eng.load_genomes(gs)
        st = eng.simulate(reads, 42, sink=lambda a, f, n: sunk.__setitem__(0, sunk[0] + a.shape[0]))
        dt = time.perf_counter() - t
        print(f"e2e: {st['bases'] / dt / 1e9:.2f} Gbases/s, {sunk[0] / dt / 1e9:.1f} GB/s to host, {dt:.2f} s")

"""Runs the five BASELINE.json configs once on one B200 and writes profiles/<tag>_configs.json:
device-side rate, algorithmic GB/s of the format kernel, and the parity / property check of each.
C2 is the bench workload (bench.py is the contract for it); the others are parity-test cases.
usage: run_configs.py [tag]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                                   # noqa: E402
import helpers as H                            # noqa: E402
import simmr_b200 as S                         # noqa: E402
from oracle import pyoracle as po              # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


genomes = H.config_genomes          # synthetic sets of the named sizes (simmr_b200/synth.py)


def run(eng, gs, prof, reads, seed=42, iters=3, slices=((0, 300),)):
    eng.load_genomes(gs)
    eng.set_error_profile(prof)
    for _ in range(iters):
        st = eng.simulate(reads, seed)
    total = eng.total_units(reads)
    og, oprof = po.Genomes.from_set(gs), H.oracle_profile(prof)
    ok = True
    for ub, n in slices:
        ub = min(ub, max(0, total - n))
        ue = min(total, ub + n)
        ok &= eng.simulate_bytes(reads, seed, unit_begin=ub, unit_end=ue) == po.simulate(og, oprof, reads, seed, unit_begin=ub, unit_end=ue)
    ok &= eng.measure(reads, seed) == (st["bytes"], st["records"])
    return {"records": st["records"], "bases": st["bases"], "fastq_bytes": st["bytes"], "device_ms": round(st["device_ms"], 3), "format_ms": round(st["format_ms"], 3),
            "gbases_per_s": round(st["bases"] / st["device_ms"] / 1e6, 1), "records_per_s": round(st["records"] / st["device_ms"] * 1e3),
            "format_kernel_alg_gbs": round(st["algorithmic_bytes"] / st["format_ms"] / 1e6, 1),
            "format_kernel_roofline_frac": round(st["algorithmic_bytes"] / st["format_ms"] / 1e6 / PEAK, 3),
            "clamped": st["clamped"], "oracle_parity_on_slices_and_measure": bool(ok)}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = {"peak_gbs": PEAK, "note": "one B200, device-side (FASTQ bytes left in HBM staging), seed 42; slices of every run are byte-compared with the CPU oracle"}
    with S.Engine(0) as eng:
        t = time.time()
        g1 = genomes(1, 5_000_000, 5_000_000, (1, 1), seed=1)
        prof = S.PerfectShortErrorProfile(150, 150)
        reads, _ = S.UniformAbundanceProfile().determine_abundances(5000, 1)
        eng.load_genomes(g1); eng.set_error_profile(prof)
        whole = eng.simulate_bytes(reads, 42)
        r = run(eng, g1, prof, reads, slices=((0, 2500),))
        r["whole_run_identical_to_oracle"] = whole == po.simulate(po.Genomes.from_set(g1), H.oracle_profile(prof), reads, 42)
        out["C1 perfect-short PE, 5000 reads, one 5 Mbp genome"] = r
        print("C1", r, flush=True)

        g100 = bench.c2_genomes()
        reads, _ = S.UniformAbundanceProfile().determine_abundances(100_000_000, 100)
        r = run(eng, g100, S.MinimalShortErrorProfile(150, 150, 30), reads, slices=((0, 300), (25_000_000, 300), (49_999_000, 300)))
        out["C2 minimal-short PE 150, 100M reads, 100 genomes, uniform"] = r
        print("C2", r, flush=True)

        ab = np.random.default_rng(3).lognormal(0, 1, 100); ab /= ab.sum()
        reads, _ = S.CustomAbundanceProfile(ab).reads_per_genome(500_000_000, g100.genome_sizes)
        r = run(eng, g100, S.CustomShortErrorProfile(H.novaseq_blob()), reads, slices=((0, 300), (123_456_789, 300), (249_000_000, 300)))
        out["C3 NovaSeq custom-short, 500M reads, uuid+abundance genome file (custom abundance)"] = r
        print("C3", r, flush=True)

        g50 = genomes(50, 2_000_000, 6_000_000, (1, 4), seed=4, min_contig=100_000)
        reads, _ = S.UniformAbundanceProfile().determine_abundances(10_000_000, 50)
        r = run(eng, g50, S.MinimalLongErrorProfile(30), reads, iters=2, slices=((0, 20), (5_000_000, 20)))
        out["C4a minimal-long, 10M reads, 50 genomes"] = r
        print("C4a", r, flush=True)
        r = run(eng, g50, S.CustomLongErrorProfile(H.ont_blob()), reads, iters=2, slices=((0, 20), (7_000_000, 20)))
        out["C4b' ONT-style custom-long without an alt-k-mer table (empirical Phreds, gamma lengths), 10M reads"] = r
        print("C4b'", r, flush=True)
        r = run(eng, g50, S.CustomLongErrorProfile(H.ont_blob_with_kmers(k=7, n_alts=20)[0]), reads, iters=2, slices=((0, 20), (7_000_000, 20)))
        out["C4b ONT-style custom-long with a k = 7 alt-k-mer table of 20 alternates per k-mer (k-mer error injection on device), 10M reads"] = r
        print("C4b", r, flush=True)

        g1000 = genomes(1000, 4_000_000, 4_800_000, (1, 6), seed=5)
        reads, _ = S.ExactAbundanceProfile().determine_abundances(1_000_000, 1000)
        r = run(eng, g1000, S.MinimalShortErrorProfile(150, 150, 30), reads, iters=2, slices=((0, 300), (499_999_000, 300)))
        r["genome_bytes"] = int(g1000.bases.nbytes)
        out["C5 exact abundance, 1000 genomes (~4 GB), 1B reads, minimal-short (one GPU's view of the whole job)"] = r
        print("C5", r, flush=True)
        out["wall_s"] = round(time.time() - t, 1)
    path = os.path.join(ROOT, "gpurun_out", f"{tag}_configs.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()

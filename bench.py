#!/usr/bin/env python
"""bench.py — headline benchmark of the simmr read-simulation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], "C2"): minimal-short PE 150 bp, 100 M reads from
a 100-genome synthetic metagenome (2-8 Mbp each, 1-20 contigs), uniform abundance.
A step is one pass of the whole hot path (sizes -> scan -> fused gather / RC / Phred
/ substitution / FASTQ format) over one 100 M-read job per GPU.  With N GPUs the job
is N x 100 M reads sharded by global unit (pair) index, genomes replicated, no
data-path collective ("weak" scaling).

value : device-side Gbases/s, genomes resident in HBM, FASTQ bytes left in HBM staging.
e2e   : the same job through simmr_genomes_load + simmr_simulate with HOST buffers:
        genomes copied host->device and every FASTQ byte copied device->pinned host
        inside the timed region.
roofline : algorithmic bytes (sum of 3L+H+5 per record) of the format
        kernel / its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline : the CPU oracle (C++ restatement of the reference; the Rust reference
        cannot be built in this image) on a bounded sample of the same job.
--impl reference : that same oracle, all host threads, printed as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "simulated_bases_per_second"
UNIT = "Gbases/s"
N_GENOMES = 100
READS_PER_GPU = 100_000_000
SEED = 42


def c2_genomes():
    """100 genomes, sizes uniform in [2, 8] Mbp, 1-20 contigs each >= 50 kbp, i.i.d. ACGT,
    ids g{genome:04}_c{contig:03}, uuids %016x of splitmix64(genome)."""
    import simmr_b200 as S
    from helpers import splitmix64
    rng = np.random.default_rng(0x5EED0000)
    sizes = rng.integers(2_000_000, 8_000_001, N_GENOMES)
    total = int(sizes.sum())
    bases = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, total, dtype=np.uint8)]
    coff, csize, gcon, cids, cidoff, gids, gidoff = [0], [], [0], b"", [0], b"", [0]
    for g in range(N_GENOMES):
        nc = int(rng.integers(1, 21))
        cuts = np.sort(rng.integers(0, int(sizes[g]) - nc * 50_000 + 1, nc - 1)) if nc > 1 else np.zeros(0, np.int64)
        lens = np.diff(np.concatenate([[0], cuts, [int(sizes[g]) - nc * 50_000]])) + 50_000
        for c, ln in enumerate(lens):
            coff.append(coff[-1] + int(ln))
            csize.append(int(ln))
            cids += f"g{g:04d}_c{c:03d}".encode()
            cidoff.append(len(cids))
        gcon.append(len(csize))
        gids += f"{splitmix64(g):016x}".encode()
        gidoff.append(len(gids))
    assert coff[-1] == total
    return S.GenomeSet(bases, coff, csize, gcon, cids, cidoff, gids, gidoff)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                  "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        while not self.stop_flag.is_set():
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))
        p.terminate()

    def summary(self, t0=None, t1=None):
        """Samples taken inside [t0, t1] (the timed region); all samples if that leaves none."""
        rows = [r for t, r in self.rows if t0 is None or t0 <= t <= t1] or [r for _, r in self.rows]
        sm = [float(r[1]) for r in rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
            if any(len(r) >= 8 and r[col].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def oracle_rate(og, prof, reads, unit_begin, n_units, threads, mode):
    """Gbases/s of the CPU oracle over n_units pairs on `threads` std::threads inside the oracle library (buffers
    allocated and touched before its clock starts).  mode 1 = the reference's own stream plumbing (ChaCha12 StdRng
    re-seeded per pair, ziggurat normal and f32 powf per base: oracle/ref_stream.inc); mode 0 = the oracle's restatement
    of the GPU's draw schedule."""
    from oracle import pyoracle as po
    r = po.bench_pe(og, prof, reads, SEED, unit_begin, n_units, mode, threads)
    return r["bases"] / r["seconds"] / 1e9, r["bases"], r["seconds"]


def oracle_inputs(gs):
    from oracle import pyoracle as po
    import simmr_b200 as S
    from helpers import oracle_profile
    return po.Genomes.from_set(gs), oracle_profile(S.MinimalShortErrorProfile(150, 150, 30))


REF_ALGO = ("C++ restatement of simmr's CPU path with the reference's own arithmetic: ChaCha12 StdRng re-seeded 7x per pair, "
            "one ziggurat normal and one f32 powf per base, FASTQ formatted into memory (oracle/ref_stream.inc; the Rust "
            "reference cannot be built in this image)")


def sized_sample(og, prof, reads, threads, mode, seconds):
    """Pairs that take about `seconds` on `threads` threads: short runs, grown until one lasts a good part of the target
    (thread start-up and a busy host make the first short runs poor predictors)."""
    n = 4_000 * threads
    for _ in range(8):
        _, _, dt = oracle_rate(og, prof, reads, 0, n, threads, mode)
        if dt >= 0.4 * seconds or n >= 20_000_000:
            break
        n = int(min(20_000_000, n * min(8.0, 0.8 * seconds / max(dt, 1e-3))))
    return n


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path on all host threads.  The reference is Rust and
    cannot be compiled in this image, so its algorithm is timed from the oracle's restatement of it (kind "port")."""
    if rank != 0:
        return
    gs = c2_genomes()
    og, prof = oracle_inputs(gs)
    reads = np.full(N_GENOMES, READS_PER_GPU * world // N_GENOMES, np.uint64)
    cores = os.cpu_count() or 1
    sample = sized_sample(og, prof, reads, cores, 1, 6.0)                  # ~6 s of work per step
    vals, times = [], []
    for it in range(args.warmup + args.steps):
        r, nb, dt = oracle_rate(og, prof, reads, (it * sample) % 40_000_000, sample, cores, 1)
        if it >= args.warmup:
            vals.append(r)
            times.append(dt)
    v = float(np.mean(vals))
    desc = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "algorithm": REF_ALGO,
            "sample": f"{sample} pairs ({2 * sample} records) per step of the C2 job, {cores} std::threads"}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times)) * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                      "data": "synthetic", "config": workload_config(world), "cpu_baseline": desc,
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline(gs, reads):
    """Both CPU algorithms on a bounded sample (about 20 s in all): `value` is the reference's arithmetic on all cores —
    what --impl reference prints — with the native-schedule port and the single-thread rates beside it."""
    og, prof = oracle_inputs(gs)
    cores = os.cpu_count() or 1
    out = {}
    for name, mode, secs in (("reference_arithmetic", 1, 6.0), ("native_schedule", 0, 3.0)):
        n_all = sized_sample(og, prof, reads, cores, mode, secs)
        r_all, _, d_all = oracle_rate(og, prof, reads, 0, n_all, cores, mode)
        n_1 = sized_sample(og, prof, reads, 1, mode, secs / 2)
        r_1, _, d_1 = oracle_rate(og, prof, reads, 0, n_1, 1, mode)
        out[name] = {"all_cores": r_all, "one_thread": r_1, "pairs_all_cores": n_all, "pairs_one_thread": n_1,
                     "seconds_all_cores": round(d_all, 2), "seconds_one_thread": round(d_1, 2)}
    ra = out["reference_arithmetic"]
    return {"value": ra["all_cores"], "unit": UNIT, "cores": cores, "kind": "port", "algorithm": REF_ALGO,
            "sample": f"first {ra['pairs_all_cores']} pairs of the C2 job on {cores} std::threads ({ra['seconds_all_cores']} s); "
                      f"1 thread: {ra['one_thread']:.4f} {UNIT} on {ra['pairs_one_thread']} pairs",
            "single_thread_value": ra["one_thread"],
            "native_schedule_port": {"value": out["native_schedule"]["all_cores"], "single_thread_value": out["native_schedule"]["one_thread"],
                                     "note": "the oracle's restatement of the GPU's own draw schedule (tables, MWC chains, sparse "
                                             "substitutions) — cheaper than the reference's arithmetic, reported for comparison"}}


def workload_config(world):
    return {"workload": "C2: minimal-short PE 150 bp, 100M reads per GPU from a 100-genome synthetic metagenome "
                        "(2-8 Mbp, 1-20 contigs), uniform abundance, seed 42",
            "reads_per_gpu": READS_PER_GPU, "genomes": N_GENOMES, "sharding": f"units x{world}, genomes replicated",
            "l2": "inputs (~500 MB genomes) and outputs (~38 GB FASTQ per step) both exceed the 126 MB L2: no flush needed"}


_T0 = time.perf_counter()


def note(msg):
    """progress on stderr (the JSON line on stdout stays alone)"""
    print(f"[bench {time.perf_counter() - _T0:7.1f} s] {msg}", file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="simmr_b200")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--reads-per-gpu", type=int, default=READS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-to-file", action="store_true")
    ap.add_argument("--extras-seconds", type=float, default=240.0, help="time limit for the sub-records after the headline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import simmr_b200 as S
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    rpg = args.reads_per_gpu
    note("building the C2 genome set")
    gs = c2_genomes()
    reads, _ = S.UniformAbundanceProfile().determine_abundances(rpg * world, N_GENOMES)   # a1: ceil(total/n)
    eng = S.Engine(local)
    eng.load_genomes(gs)
    eng.set_error_profile(S.MinimalShortErrorProfile(150, 150, 30))
    total_units = eng.total_units(reads)
    ub, ue = rank * total_units // world, (rank + 1) * total_units // world

    # `value`: the device-resident run through simmr_simulate_device — every batch is handed to a consumer on the device
    # (here: it reads the batch's first and last byte back, '@' and '\n'), FASTQ bytes never leave HBM
    from cuda.bindings import runtime as cudart
    probe = np.zeros(2, np.uint8)
    seen = [0, 0]

    def consumer(ptr, n, first, units):
        if n:
            cudart.cudaMemcpy(probe.ctypes.data, ptr, 1, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
            cudart.cudaMemcpy(probe.ctypes.data + 1, ptr + n - 1, 1, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
            if probe[0] != ord("@") or probe[1] != ord("\n"):
                return 1
        seen[0] += n
        seen[1] += 1
        return 0

    note("genomes resident; device-resident steps")
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        eng.simulate_device(reads, SEED, consumer, unit_begin=ub, unit_end=ue)
    barrier()
    t_timed0 = time.perf_counter()
    dev_ms = fmt_ms = 0.0
    bases = records = alg = launches = nbytes = 0
    seen[:] = [0, 0]
    for _ in range(args.steps):
        st = eng.simulate_device(reads, SEED, consumer, unit_begin=ub, unit_end=ue)
        dev_ms += st["device_ms"]; fmt_ms += st["format_ms"]
        bases += st["bases"]; records += st["records"]; alg += st["algorithmic_bytes"]; launches += st["kernel_launches"]; nbytes += st["bytes"]
    barrier()
    t_timed1 = time.perf_counter()
    sampler.stop_flag.set()
    assert seen[0] == nbytes                     # the consumer was shown every byte of every step
    n_format_launches = seen[1]
    t = torch.tensor([dev_ms, float(bases), float(records)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        max_ms, all_bases, all_records = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        max_ms, all_bases, all_records = dev_ms, float(bases), float(records)
    value = all_bases / (max_ms / 1e3) / 1e9

    # the format kernel by itself: the library normally prepares batch b+1 (geometry, sizes, offsets, size-sorted tiles) on a
    # second stream beside batch b's format kernel, which shares the SMs; one extra untimed step with everything on one
    # stream gives the kernel's duration without co-runners (reported next to the figure of the timed region)
    note("timed region done; one step with every kernel on one stream")
    os.environ["SIMMR_PREP_SERIAL"] = "1"
    st = eng.simulate(reads, SEED, unit_begin=ub, unit_end=ue)
    del os.environ["SIMMR_PREP_SERIAL"]
    alone_ms, alone_alg = st["format_ms"], st["algorithmic_bytes"]
    alone_launches = max(1, n_format_launches // max(1, args.steps))

    # end to end: host genome buffers -> device, FASTQ bytes -> pinned host -> the sink (library writer thread), inside the timed region
    note("end-to-end steps (host buffers)")
    e2e_bases = e2e_d2h = 0
    if args.e2e_steps:      # untimed: the pinned output ring is allocated on the first call that has a sink
        eng.simulate(reads, SEED, unit_begin=ub, unit_end=min(ue, ub + 2_000_000), sink=lambda a, f, n: 0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        eng.load_genomes(gs)
        sunk = [0]
        st = eng.simulate(reads, SEED, unit_begin=ub, unit_end=ue, sink=lambda a, f, n: sunk.__setitem__(0, sunk[0] + a.shape[0]))
        assert sunk[0] == st["bytes"]
        e2e_bases += st["bases"]; e2e_d2h += st["d2h_bytes"]
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    tb = torch.tensor([float(e2e_bases)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX); dist.all_reduce(tb, op=dist.ReduceOp.SUM)
    e2e_value = float(tb[0]) / float(te[0]) / 1e9 if args.e2e_steps else None

    out = None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = alg / (fmt_ms / 1e3) / 1e9
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get("format_kernel_dram_bytes_per_launch"), tj.get("source")
        n_launch = max(1, n_format_launches)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u8", "data": "synthetic", "config": workload_config(world),
               "reads_per_second": all_records / (max_ms / 1e3), "clocks": sampler.summary(t_timed0, t_timed1), "gpu_launches": launches,
               "value_note": "device-resident: every batch's FASTQ bytes are handed to a consumer on the device (simmr_simulate_device); e2e is the same job with host buffers",
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(gs.bases.nbytes),
                       "d2h_bytes_per_step": int(e2e_d2h // max(1, args.e2e_steps)), "steps": args.e2e_steps},
               "roofline": {"bound": "hbm", "kernel": "k_format_lanes", "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                            "bytes": "algorithmic: sum over records of 3L + H + 5 (L source bases at one byte each, H + 2L + 5 written; SURVEY.md §8d) — the kernel "
                                     "reads the 2-bit genome copy, so its real DRAM traffic (`traffic`, an offline ncu capture) is lower",
                            "algorithmic_bytes_per_launch": alg / n_launch,
                            "launch_ms": fmt_ms / n_launch, "launches": n_launch, "kernel_share_of_step": fmt_ms / dev_ms,
                            "whole_step_frac": (alg / (dev_ms / 1e3) / 1e9) / peak,
                            "note": "timed region: the next batch's small kernels run beside this kernel on a second stream",
                            "alone": {"achieved": alone_alg / (alone_ms / 1e3) / 1e9, "frac": alone_alg / (alone_ms / 1e3) / 1e9 / peak,
                                      "launch_ms": alone_ms / alone_launches, "launches": alone_launches,
                                      "note": "one extra untimed step with every kernel on one stream"}}}
    # The headline is complete here.  The sub-records that follow (strong scaling, FASTQ to file, CPU baseline) run under a
    # time limit: if one of them stalls, the line is printed without it rather than not at all.
    def give_up():
        note("the sub-records exceeded their time limit: printing the line without the missing ones")
        if out is not None:
            print(json.dumps(out), flush=True)
        os._exit(0)

    watchdog = threading.Timer(args.extras_seconds, give_up)
    watchdog.daemon = True
    watchdog.start()
    note("strong-scaling record (C5)")
    copy_peak = d2h_copy_rate(torch, dist, barrier) if args.e2e_steps else None
    if out is not None and copy_peak is not None:      # what the box gives a plain device -> pinned-host copy on all ranks at once
        out["e2e"]["d2h_copy_gbs_all_ranks"] = copy_peak
        out["e2e"]["d2h_gbs_achieved"] = out["e2e"]["d2h_bytes_per_step"] * world * args.e2e_steps / (float(te[0]) * 1e9)
    strong = strong_scaling(S, eng, torch, dist, barrier, rank, world) if not args.no_strong else None
    if out is not None and strong is not None:
        out["strong"] = strong
    if rank == 0 and world == 1:           # reported at N = 1 only; the scaling runs stay short
        if not args.no_to_file:
            note("FASTQ to file")
            out["to_file"] = fastq_to_file(S, gs, local)
        if not args.no_cpu_baseline:
            note("CPU baseline")
            out["cpu_baseline"] = cpu_baseline(gs, reads)
    watchdog.cancel()
    if out is not None:
        print(json.dumps(out), flush=True)
    note("done")
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def d2h_copy_rate(torch, dist, barrier):
    """Aggregate GB/s of plain cudaMemcpyAsync device -> pinned host, 1 GiB buffers, every rank at the same time: the
    ceiling the end-to-end figure (every FASTQ byte crosses PCIe) can reach on this box."""
    try:
        n = 1 << 30
        dev = torch.empty(n, dtype=torch.uint8, device="cuda")
        host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        host.copy_(dev, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            host.copy_(dev, non_blocking=True)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        nb = torch.tensor([4.0 * n], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX); dist.all_reduce(nb, op=dist.ReduceOp.SUM)
        del dev, host
        return float(nb[0]) / float(dt[0]) / 1e9
    except Exception as e:
        note(f"copy-rate probe failed: {e!r}")
        return None


def strong_scaling(S, eng, torch, dist, barrier, rank, world):
    """BASELINE config 5 as a fixed job: exact abundance, 1000 synthetic genomes (4.4 GB), 1 M reads each = 1 B reads,
    sharded by unit index over the ranks (genomes replicated, no data-path collective); device time = max over ranks.
    Rank 0 then runs the whole job alone, which gives the efficiency T1 / (N * TN) inside one bench run."""
    from simmr_b200.synth import synthetic_genomes
    try:
        g5 = synthetic_genomes(1000, 4_000_000, 4_800_000, (1, 6), seed=5)
        reads, _ = S.ExactAbundanceProfile().determine_abundances(1_000_000, 1000)
        eng.load_genomes(g5)
        eng.set_error_profile(S.MinimalShortErrorProfile(150, 150, 30))
        total = eng.total_units(reads)
        ub, ue = rank * total // world, (rank + 1) * total // world
        eng.simulate(reads, SEED, unit_begin=ub, unit_end=ue)              # warm-up
        barrier()
        st = eng.simulate(reads, SEED, unit_begin=ub, unit_end=ue)
        barrier()
        t = torch.tensor([st["device_ms"], float(st["bases"])], dtype=torch.float64, device="cuda")
        if dist is not None:
            tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            ms_n, bases = float(tm[0]), float(ts[1])
        else:
            ms_n, bases = float(t[0]), float(t[1])
        ms_1 = ms_n
        if world > 1 and rank == 0:
            ms_1 = eng.simulate(reads, SEED)["device_ms"]
        barrier()
        return {"workload": "C5: exact abundance, 1000 synthetic genomes (4.4 GB), 1 M reads per genome = 1 B reads (500 M pairs), minimal-short, fixed job sharded over the ranks",
                "scaling": "strong", "n_gpus": world, "device_ms": ms_n, "value": bases / (ms_n / 1e3) / 1e9, "unit": UNIT,
                "device_ms_one_gpu": ms_1, "efficiency_vs_one_gpu": ms_1 / (world * ms_n), "genome_bytes": int(g5.bases.nbytes)}
    except Exception as e:                                                  # the headline must not die with the sub-record
        return {"error": repr(e)}


def fastq_to_file(S, gs, device):
    """FASTQ to a file on tmpfs through the `simmr` command (FASTA files in, FASTQ out): 20 M reads of the C2 job, one
    output file and eight part files; rates are FASTQ bytes over the command's own simulate-and-write phase and over
    its whole wall time (CUDA context, FASTA ingest and genome upload included)."""
    import shutil
    import tempfile
    exe = os.path.join(ROOT, "simmr_b200", "simmr")
    root = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="simmr_bench_", dir=root)
    try:
        lines = []
        for g in range(gs.n_genomes):
            p = os.path.join(d, f"g{g:04d}.fna")
            with open(p, "wb") as f:
                for c in range(int(gs.genome_contig[g]), int(gs.genome_contig[g + 1])):
                    seq = gs.contig(c)
                    n = len(seq) // 80 * 80
                    body = np.empty((n // 80, 81), np.uint8)
                    body[:, :80] = seq[:n].reshape(-1, 80)
                    body[:, 80] = 10
                    f.write(b">" + gs.contig_id(c) + b"\n")
                    body.tofile(f)
                    if n < len(seq):
                        f.write(seq[n:].tobytes() + b"\n")
            lines.append(f"{p}\t{gs.uuid(g)}\n")
        open(os.path.join(d, "genomes.tsv"), "w").write("path\tuuid\n" + "".join(lines))
        out = {"filesystem": "tmpfs (/dev/shm)" if root else "default temporary directory", "reads": 20_000_000}
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(device))
        for name, extra in (("one_file", []), ("eight_parts", ["--output-parts", "8"]), ("gzip_split_mates", ["--gzip", "--split-mates"])):
            o = os.path.join(d, "reads.fq")
            t0 = time.perf_counter()
            r = subprocess.run([exe, "--genome-file", os.path.join(d, "genomes.tsv"), "--output", o, "--num-reads", "20000000", "--seed", "42",
                                "--error-profile", "minimal-short"] + extra, capture_output=True, text=True, env=env, timeout=90)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                out[name] = {"error": r.stderr[-300:]}
                continue
            files = ([o] if not extra else [f"{o}.{k:03d}" for k in range(8)] if name == "eight_parts"
                     else [os.path.join(d, "reads_R1.fq"), os.path.join(d, "reads_R2.fq")])
            on_disk = sum(os.path.getsize(p) for p in files)
            phase = [ln for ln in r.stderr.splitlines() if "FASTQ bytes (" in ln]
            secs = float(phase[0].split(") in ")[1].split(" s")[0]) if phase else None
            nbases = int(phase[0].split(" records, ")[1].split(" bases")[0]) if phase else None
            size = int(phase[0].split("Wrote ")[1].split(" FASTQ")[0]) if phase else on_disk
            out[name] = {"fastq_bytes": size, "file_bytes": on_disk, "wall_s": round(wall, 3), "gb_per_s_wall": size / wall / 1e9,
                         "write_phase_s": secs, "gb_per_s_write_phase": size / secs / 1e9 if secs else None,
                         "gbases_per_s_write_phase": nbases / secs / 1e9 if secs else None}
            for p in files:
                os.remove(p)
        return out
    except Exception as e:
        return {"error": repr(e)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()

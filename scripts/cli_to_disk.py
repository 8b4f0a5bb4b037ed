"""FASTQ-to-file rate of the simmr command line (FASTA in, FASTQ + .tsv out): the whole reference flow of main.rs:20-266
through the C ABI, on tmpfs and on the box's local disk.  Writes gpurun_out/<tag>_cli_to_file.json.
usage: cli_to_disk.py [tag] [reads]"""
import json, os, subprocess, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
reads = int(float(sys.argv[2])) if len(sys.argv) > 2 else 40_000_000
work = "/tmp/simmr_cli"
os.makedirs(work, exist_ok=True)
rng = np.random.default_rng(5)
rows = ["path\tuuid"]
for g in range(20):
    path = f"{work}/g{g:02d}.fna"
    with open(path, "wb") as f:
        for c in range(3):
            n = int(rng.integers(1_000_000, 2_500_000))
            s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n, dtype=np.uint8)]
            f.write(f">g{g:04d}_c{c:03d}\n".encode())
            full = n // 80                                   # 80-column lines
            f.write(np.concatenate([s[:full * 80].reshape(full, 80), np.full((full, 1), 10, np.uint8)], axis=1).tobytes())
            if n % 80:
                f.write(s[full * 80:].tobytes() + b"\n")
    rows.append(f"{path}\t{g:016x}")
open(f"{work}/genomes.tsv", "w").write("\n".join(rows) + "\n")
out = {"reads": reads, "profile": "minimal-short", "genomes": 20}
for name, target in (("tmpfs", "/dev/shm/simmr_cli_out.fq"), ("local_disk", f"{work}/out.fq")):
    for f in (target, target + ".tsv"):
        if os.path.exists(f):
            os.remove(f)
    t = time.perf_counter()
    r = subprocess.run([os.path.join(ROOT, "simmr_b200", "simmr"), "--genome-file", f"{work}/genomes.tsv", "--output", target,
                        "--num-reads", str(reads), "--error-profile", "minimal-short", "--seed", "42"], capture_output=True, text=True)
    dt = time.perf_counter() - t
    if r.returncode != 0:
        out[name] = {"error": r.stderr[-300:]}
        continue
    size = os.path.getsize(target)
    out[name] = {"wall_s": round(dt, 2), "fastq_bytes": size, "gb_per_s_to_file": round(size / dt / 1e9, 2),
                 "gbases_per_s": round(reads * 149.5 / dt / 1e9, 2), "note": "wall clock of the whole command: FASTA parse, load, simulate, write"}
    os.remove(target)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_cli_to_file.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

"""Writes the judged profile summaries under profiles/ from gpurun_out/ captures.
usage: profile_report.py <tag> <launches.csv> <full.ncu-rep> <kernel mangled substring>"""
import collections, csv, io, json, os, subprocess, sys
tag, launches, rep, ksub = sys.argv[1:5]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
os.makedirs(out, exist_ok=True)

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(list)
for r in rows[hi + 1:]:
    v = float(r[mv].replace(',', ''))
    v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3, 's': v * 1e6}[r[mu]]
    agg[r[kn].split('(')[0].replace('void ', '')].append(v)
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out, f"{tag}_launches.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {sum(len(v) for v in agg.values())} launches of the bench command\n")
    f.write("# (per-launch times under ncu are serialised and cold-cache: compare shares, not absolutes)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k:40s} launches={len(v):4d} total_ms={sum(v) / 1e3:9.3f} avg_us={sum(v) / len(v):9.1f} share={100 * sum(v) / tot:5.1f}%\n")

def run(*a):
    return subprocess.run(a, capture_output=True, text=True).stdout
summ = run(sys.executable, os.path.join(root, "scripts", "ncu_summary.py"), rep, "12")
lines = run(sys.executable, os.path.join(root, "scripts", "ncu_lines.py"), rep, ksub, "14")
with open(os.path.join(out, f"{tag}_format_kernel_ncu_full.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on, one launch of the format kernel inside the bench command\n")
    f.write(summ + "\n# instructions and stall samples by source line (inlined code is attributed to its call site)\n" + lines)
raw = run("ncu", "-i", rep, "--page", "raw", "--csv")
r = list(csv.reader(io.StringIO(raw)))
h, units, row = r[0], r[1], r[2]
def val(name):
    v, u = float(row[h.index(name)].replace(',', '')), units[h.index(name)]
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
rd, wr = val('dram__bytes_read.sum'), val('dram__bytes_write.sum')
json.dump({"format_kernel_dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "source": f"profiles/{tag}_format_kernel_ncu_full.txt", "kernel": row[h.index('Kernel Name')]},
          open(os.path.join(out, "traffic.json"), "w"), indent=1)
print(open(os.path.join(out, f"{tag}_launches.txt")).read())

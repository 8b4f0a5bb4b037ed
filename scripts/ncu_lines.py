"""Per-source-line instruction and stall-sample totals for one kernel of an .ncu-rep,
using nvdisasm line info from the built library.  usage: ncu_lines.py rep kernel_mangled_substr [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "simmr_b200", "libsimmr_b200.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("engine")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
line_of, cur, infn = {}, None, False
for ln in dis:
    if ln.startswith(".text."):
        infn = ksub in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File ".*?", line (\d+)(.*)', ln)
    if m:
        if 'inlined at' in ln and cur is not None and ln.strip().startswith('//## File') and 'inlined' in m.group(2):
            pass
        cur = int(m.group(1))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[start[0]]
end = start[1] - 1 if len(start) > 1 else len(rows)
ins = [r for r in rows[start[0] + 1:end] if len(r) > 10]
ci, csamp, cth = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
base = int(ins[0][0], 16)
ex, sm, th = collections.Counter(), collections.Counter(), collections.Counter()
for r in ins:
    off = int(r[0], 16) - base
    l = line_of.get(off)
    ex[l] += int(r[ci]); sm[l] += int(r[csamp]); th[l] += int(r[cth])
te, ts = sum(ex.values()), sum(sm.values())
srcl = open(os.path.join(root, "simmr_b200", "csrc", "device_code.cuh")).read().split("\n")
print(f"total warp-instr {te}  samples {ts}")
print("--- by instructions executed")
for l, v in ex.most_common(top):
    txt = srcl[l - 1].strip()[:100] if l else '?'
    print(f"{l!s:>5} instr {100 * v / te:5.1f}%  samp {100 * sm[l] / ts:5.1f}%  act {th[l] / max(1, v):4.1f}  {txt}")

"""Static size of the innermost loops of a kernel that contain LDS.64 (the pair loop of the format kernel).
usage: sass_loop.py <mangled kernel substring>"""
import re, subprocess, sys, os, collections
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(root, "simmr_b200", "libsimmr_b200.so")], capture_output=True, text=True).stdout
ksub = sys.argv[1]
lines, on = [], False
for ln in out.split("\n"):
    if "Function :" in ln:
        on = ksub in ln
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(lines)}
loops = []
for i, (a, t) in enumerate(lines):
    m = re.search(r"BRA(?:\.U)?(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr:
            loops.append((addr[tgt], i))
for lo, hi in sorted(loops, key=lambda x: x[1] - x[0]):
    body = [t for _, t in lines[lo:hi + 1]]
    n64 = sum("LDS.64" in t for t in body)
    if n64 == 0 or any(l2 > lo and h2 < hi for l2, h2 in loops if (l2, h2) != (lo, hi) and sum("LDS.64" in t for _, t in lines[l2:h2 + 1])):
        continue
    ops = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for t in body)
    print(f"loop {lines[lo][0]:#x}-{lines[hi][0]:#x}: {len(body)} instr, LDS.64={n64} -> {len(body) * 8 / n64:.1f} per pair;", dict(ops.most_common(12)))

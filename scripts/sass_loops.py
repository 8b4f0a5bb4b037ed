"""Opcode mix of the loops of a kernel (by static SASS): usage sass_loops.py <lib.so> <mangled kernel substring> [min LDS]"""
import re, subprocess, sys, collections
lib, ksub = sys.argv[1], sys.argv[2]
minlds = int(sys.argv[3]) if len(sys.argv) > 3 else 8
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
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
ALU = {"LOP3", "ISETP", "SHF", "SEL", "PRMT", "VIADD", "IADD3", "LEA", "IABS", "IMNMX", "VIMNMX", "PLOP3", "BREV", "FLO", "POPC", "IADD", "MOV", "BMSK", "SGXT", "R2P", "P2R", "CS2R", "VABSDIFF"}
for i, (a, t) in enumerate(lines):
    m = re.search(r"BRA(?:\.U)?(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*(0x[0-9a-f]+)?", t)
    m2 = re.search(r"(0x[0-9a-f]+)", t) if "BRA" in t else None
    if m2:
        tgt = int(m2.group(1), 16)
        if tgt <= a and tgt in addr:
            body = [x for _, x in lines[addr[tgt]:i + 1]]
            ops = collections.Counter((x.split()[1] if x.startswith("@") else x.split()[0]).split(".")[0] for x in body)
            if ops.get("LDS", 0) >= minlds:
                alu = sum(v for k, v in ops.items() if k in ALU)
                fma = sum(v for k, v in ops.items() if k in ("IMAD", "FFMA", "FMUL", "FADD"))
                print(f"loop {tgt:#x}-{a:#x}: {len(body)} instr  ALU {alu}  IMAD {fma}  LDS {ops.get('LDS',0)} STS {ops.get('STS',0)} LDG {ops.get('LDG',0)} |", dict(ops.most_common(14)))

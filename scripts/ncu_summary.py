"""Summarise an .ncu-rep: key raw metrics, SASS opcode mix, hottest instructions, stall reasons."""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__grid_size', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warp_latency_issue_stalled_barrier.ratio',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for row in rows[2:]:
    print('===', row[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr:
            print(f'  {w:75s} {row[hdr.index(w)]:>18s} {units[hdr.index(w)]}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[start[0]]
end = start[1] - 1 if len(start) > 1 else len(rows)
ins = [r for r in rows[start[0] + 1:end] if len(r) > 10]
ci, cs, ct, csamp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[ci]) for r in ins if r[ci].isdigit())
tsamp = sum(int(r[csamp]) for r in ins if r[csamp].isdigit())
print('total warp instr', tot, 'sass lines', len(ins), 'samples', tsamp)
op = collections.Counter()
for r in ins:
    if r[ci].isdigit():
        t = r[cs].split()
        o = t[1] if t[0].startswith('@') else t[0]
        op[o.split('.')[0]] += int(r[ci])
print('opcode mix:', ', '.join(f'{k} {100 * v / tot:.1f}%' for k, v in op.most_common(18)))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.Counter()
for r in ins:
    for i in stall_cols:
        if r[i].isdigit():
            agg[hdr[i]] += int(r[i])
ts = sum(agg.values())
print('stall samples:', ', '.join(f'{k[6:]} {100 * v / ts:.1f}%' for k, v in agg.most_common(10)))
print(f'--- top {top} instructions by stall samples')
order = sorted(range(len(ins)), key=lambda i: -int(ins[i][csamp]) if ins[i][csamp].isdigit() else 0)[:top]
for i in sorted(order):
    r = ins[i]
    why = max(stall_cols, key=lambda c: int(r[c]) if r[c].isdigit() else 0)
    print(f'{i:5d} samp {int(r[csamp]):6d} ({100 * int(r[csamp]) / tsamp:4.1f}%) exec {int(r[ci]):9d} {hdr[why][6:]:14s} {r[cs][:80]}')

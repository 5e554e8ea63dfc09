"""Dev tool: one-screen summary of an .ncu-rep (per captured kernel): duration, DRAM bytes, tensor-pipe %, issue %, registers,
stall mix. usage: ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); h = r[0]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg']
print("report:", rep)
for row in r[2:]:
    print("kernel:", row[h.index('Kernel Name')])
    for k in keys:
        if k in h: print(f"  {k:75s} {row[h.index(k)]:>18s} {r[1][h.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, x in enumerate(rows) if x and x[0] == 'Address']
for n, hi in enumerate(his):
    hh = rows[hi]; end = his[n + 1] - 1 if n + 1 < len(his) else len(rows)
    data = [x for x in rows[hi + 1:end] if len(x) == len(hh)]
    ci = {k: i for i, k in enumerate(hh)}
    stalls = [k for k in hh if k.startswith('stall_') and 'Not Issued' not in k]
    tot = {k: sum(int(x[ci[k]]) for x in data if x[ci[k]].isdigit()) for k in stalls}
    ts = sum(tot.values()) or 1
    print(f"stall mix (kernel {n}, {len(data)} SASS instructions, {ts} samples):",
          {k[6:]: round(v / ts, 3) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]})

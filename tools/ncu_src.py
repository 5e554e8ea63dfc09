"""Dev tool: summarise an ncu report's source page: stall mix per kernel and the top-N stalled SASS lines.
usage: ncu_src.py report.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h = r[0]
keys = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active']
for row in r[2:]:
    print(row[h.index('Kernel Name')][:90])
    for k in keys:
        if k in h: print('  ', k, row[h.index(k)], r[1][h.index(k)])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, x in enumerate(rows) if x and x[0] == 'Address']
for n, hi in enumerate(his[:1]):
    hh = rows[hi]; end = his[n + 1] - 1 if n + 1 < len(his) else len(rows)
    data = [x for x in rows[hi + 1:end] if len(x) == len(hh)]
    ci = {k: i for i, k in enumerate(hh)}
    stalls = [k for k in hh if k.startswith('stall_') and 'Not Issued' not in k]
    tot = {k: sum(int(x[ci[k]]) for x in data if x[ci[k]].isdigit()) for k in stalls}
    ts = sum(tot.values())
    print('instrs', len(data), 'samples', ts, {k[6:]: round(v / ts, 3) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]})
    samp = ci['# Samples']
    top = sorted(data, key=lambda x: -int(x[samp]) if x[samp].isdigit() else 0)[:topn]
    for x in top:
        st = {k[6:]: int(x[ci[k]]) for k in stalls if x[ci[k]].isdigit() and int(x[ci[k]]) > 0}
        st = sorted(st.items(), key=lambda y: -y[1])[:2]
        print('   ', x[0][-5:], x[samp].rjust(6), x[ci['Instructions Executed']].rjust(9), x[1][:66].ljust(66), st)

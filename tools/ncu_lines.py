"""Dev tool: stall samples of an ncu report aggregated per CUDA source line (needs -lineinfo and the matching .o).
usage: ncu_lines.py report.ncu-rep build/obj.o kernel-name-substring [topN]"""
import csv, subprocess, sys, io, re, glob, os, collections
rep, obj, pat = sys.argv[1:4]; topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); h = r[0]
for k in ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg']:
    if k in h: print('  ', k, r[2][h.index(k)], r[1][h.index(k)])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, x in enumerate(rows) if x and x[0] == 'Address'][0]
hh = rows[hi]; data = [x for x in rows[hi + 1:] if len(x) == len(hh)]
ci = {k: i for i, k in enumerate(hh)}
base = int(data[0][0], 16)
tmp = os.path.join('/tmp', 'ncu_lines_x'); os.makedirs(tmp, exist_ok=True)
subprocess.run(f"cd {tmp} && rm -f *.cubin && cuobjdump -xelf all {os.path.abspath(obj)} >/dev/null 2>&1", shell=True)
cub = glob.glob(tmp + '/*.cubin')[0]
out = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout
m = {}; line = None; fn = None; func = None
for l in out.splitlines():
    mm = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if mm: fn = os.path.basename(mm.group(1)); line = int(mm.group(2)); continue
    mf = re.match(r'\s*\.text\.(\S+):', l)
    if mf: func = mf.group(1); continue
    ma = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if ma and func and pat in func: m[int(ma.group(1), 16)] = (fn, line)
by = collections.Counter(); samp = ci['# Samples']
stalls = [k for k in hh if k.startswith('stall_') and 'Not Issued' not in k]
bys = collections.defaultdict(collections.Counter)
for x in data:
    if x[samp].isdigit():
        key = m.get(int(x[0], 16) - base, (None, None))
        by[key] += int(x[samp])
        for k in stalls:
            if x[ci[k]].isdigit(): bys[key][k[6:]] += int(x[ci[k]])
tot = sum(by.values()); print('samples', tot)
srcdir = os.path.join(os.path.dirname(os.path.abspath(obj)), '..', 'csrc')
for (f, ln), c in by.most_common(topn):
    try: text = open(os.path.join(srcdir, f)).read().splitlines()[ln - 1].strip()[:90]
    except Exception: text = ''
    print(f'{c:6d} {100 * c / tot:5.1f}%  {f}:{ln}  {dict(bys[(f, ln)].most_common(2))}  {text}')

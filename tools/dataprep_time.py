"""Row f-4 (one-shot data preparation): every K9 kernel timed alone with CUDA events on a column of c4's size
(50 M float64 rows = 400 MB, larger than the 126 MB L2, so successive launches read from HBM), achieved GB/s =
ALGORITHMIC bytes / launch time next to the measured copy bandwidth (MEASURED_PEAKS.json); then the whole chain
(percent -> range -> bin -> vocabulary -> tokens) end to end from a HOST column (H2D inside the timed region). The CPU
baseline beside it is bench.py --config prep's cpu_baseline (round-1 artefacts r01_dataprep_v1..v6 still carry one).
DATAPREP_ONCE=1 runs each kernel once (ncu target).          usage: dataprep_time.py [out.json] [rows]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tradeaid_b200 import lib, data_utils as du

dev = "cuda"
ONCE = os.environ.get("DATAPREP_ONCE", "0") == "1"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0


def timeit(fn, n=5):
    if ONCE:
        fn(); torch.cuda.synchronize()
        return float("nan")
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


out = []
def report(kernel, nbytes, ms, note=""):
    gbs = nbytes / ms / 1e6 if ms == ms else None
    out.append(dict(kernel=kernel, rows=N, algorithmic_bytes=nbytes, ms=None if ms != ms else round(ms, 4),
                    achieved_GBs=None if gbs is None else round(gbs, 1), peak_GBs=peak,
                    frac=None if gbs is None else round(gbs / peak, 3), note=note))
    print(json.dumps(out[-1]), flush=True)


g = torch.Generator(device=dev).manual_seed(1)
steps = torch.randn(N, dtype=torch.float64, device=dev, generator=g) * (2.0 / N ** 0.5)
prices = (100.0 * torch.exp(torch.cumsum(steps, 0))).mul(100).round().div(100)
del steps
st = torch.empty(lib.PREP_STATUS_WORDS, dtype=torch.int64, device=dev); lib.prep_status_reset(st)
starts = torch.zeros(1, dtype=torch.int64, device=dev)
pct = torch.empty_like(prices)
report("prep_percent_kernel", N * 16, timeit(lambda: lib.prep_percent_changes(prices, pct, starts, 2, st)))
thr = torch.tensor(du._log10_thresholds(), dtype=torch.float64, device=dev)
fac = torch.tensor(du._decade_factors(2), dtype=torch.float64, device=dev)
rng_out = torch.empty_like(prices)
report("prep_range_kernel", N * 16, timeit(lambda: lib.prep_range(prices, rng_out, True, 1, 10.0, 100.0, 99.9, thr, fac, st)))
ranks = torch.tensor([N // 1000, N // 1000 + 1, N - N // 1000 - 1, N - N // 1000], dtype=torch.int64, device=dev)
osv = torch.empty(4, dtype=torch.float64, device=dev)
ws = torch.empty(lib.prep_select_workspace_bytes(), dtype=torch.uint8, device=dev)
report("prep_select_hist_kernel x8 (+pick)", N * 8 * 8, timeit(lambda: lib.prep_order_statistics(pct, ranks, osv, ws, st)),
       "4 order statistics, 8 radix passes over the column")
lo, hi = osv.tolist()[0], osv.tolist()[3]
mm = torch.empty(3, dtype=torch.float64, device=dev)
mws = torch.empty(lib.prep_minmax_workspace_bytes(N), dtype=torch.uint8, device=dev)
report("prep_minmax_partial_kernel (+final)", N * 8, timeit(lambda: lib.prep_filtered_minmax(pct, lo, hi, mm, mws)))
pos, neg = du._group_boundaries(10, 2.2, max(abs(mm[0].item()), abs(mm[1].item())))
pos_d, neg_d = torch.tensor(pos, dtype=torch.float64, device=dev), torch.tensor(neg, dtype=torch.float64, device=dev)
groups = torch.empty_like(pct)
report("prep_bin_kernel", N * 16, timeit(lambda: lib.prep_bin_assign(pct, groups, pos_d, neg_d, True)), "10 groups per sign")
cap = 1 << 17
table = torch.empty(cap, dtype=torch.int64, device=dev); counters = torch.empty(2, dtype=torch.int64, device=dev)
report("prep_vocab_insert_kernel", N * 8, timeit(lambda: lib.prep_vocab_insert(rng_out, table, counters, st)))
V = int(counters[0])
keys = torch.empty(V, dtype=torch.int64, device=dev); slots = torch.empty(V, dtype=torch.int32, device=dev)
voc = torch.empty(V, dtype=torch.float64, device=dev); tos = torch.empty(cap, dtype=torch.int32, device=dev)
def build():
    counters[1] = 0
    lib.prep_vocab_build(table, V, keys, slots, voc, tos, counters)
report("prep_vocab_compact + rank kernels", cap * 8 + V * 24, timeit(build), f"V={V}, capacity {cap}")
tokens = torch.empty(N, dtype=torch.int32, device=dev)
report("prep_vocab_lookup_kernel", N * 12, timeit(lambda: lib.prep_vocab_lookup(rng_out, table, tos, tokens)))
assert st.tolist() == [lib.PREP_NO_ROW] * lib.PREP_STATUS_WORDS, st.tolist()

if not ONCE:
    # the whole chain through the drop-in API, from a pinned HOST column to tokens in a row of the token store
    params = ["f", 1, None, True, None, 2, 10, None, False, "price", 0.1, 2.2]
    host_col = prices.cpu().pin_memory()
    store_row = torch.empty(N, dtype=torch.int32, device=dev)
    def chain():
        col = host_col.to(dev, non_blocking=True)
        return du.prepare_modality(col, params, tokens_out=store_row)
    chain(); torch.cuda.synchronize()
    t0 = time.perf_counter(); tok, vocab = chain(); torch.cuda.synchronize(); t1 = time.perf_counter()
    e2e_s = t1 - t0
    summary = dict(chain="percent(2) -> bin(10 groups, 0.1 %, 2.2) -> vocabulary -> int32 tokens", rows=N,
                   e2e_seconds=round(e2e_s, 4), e2e_rows_per_s=round(N / e2e_s), h2d_bytes=N * 8, vocabulary_size=len(vocab),
                   cpu_baseline="see bench.py --config prep (the oracle is imported by tests/, smoke() and bench.py only)")
    print(json.dumps(summary), flush=True)
    out.append(summary)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)

"""Dev tool: time the HBM-bound row kernels at c4 sizes. Prints achieved GB/s (algorithmic bytes)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tradeaid_b200 import lib
dev = "cuda"
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for G, rows, C in [(8, 16384, 1024), (1, 16384, 1024), (2, 4096, 256)]:
    x = torch.randn(G, rows, C, device=dev); gamma = torch.randn(G, C, device=dev); beta = torch.randn(G, C, device=dev)
    y = torch.empty(G, rows, C, device=dev, dtype=torch.bfloat16); mean = torch.empty(G * rows, device=dev); rstd = torch.empty(G * rows, device=dev)
    ms = timeit(lambda: lib.layernorm_fwd(x, gamma, beta, y, mean, rstd))
    print(json.dumps(dict(k="ln_fwd", G=G, rows=rows, C=C, ms=round(ms, 4), gbs=round(G * rows * C * 6 / ms / 1e6))))
    dy = torch.randn(G, rows, C, device=dev).to(torch.bfloat16); dx = torch.randn(G, rows, C, device=dev); dxb = torch.empty(G, rows, C, device=dev, dtype=torch.bfloat16)
    dg = torch.empty(G, C, device=dev); db = torch.empty(G, C, device=dev)
    ws = torch.empty(lib.layernorm_bwd_workspace(G, rows, C), device=dev)
    ms = timeit(lambda: lib.layernorm_bwd(dy, x, gamma, mean, rstd, dx, dx, dxb, dg, db, ws))
    print(json.dumps(dict(k="ln_bwd", G=G, rows=rows, C=C, ms=round(ms, 4), gbs=round(G * rows * C * 16 / ms / 1e6))))

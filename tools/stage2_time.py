"""Dev tool: time the q/k/v stage-2 GEMMs (block-diagonal packing, engine strides) at c4 shapes. Env toggles of gemm.cu
(TA_GEMM_NST, TA_GEMM_2CTA_MINK, TA_GEMM_M_MAJOR) apply. usage: stage2_time.py [BT]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tradeaid_b200 import lib
BT = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
M, H, hs, h2, gsz = 8, 16, 64, 32, 4
C, NG = H * hs, 3 * H // gsz
dev = "cuda"
hid = torch.randn(M, BT, 3 * H * h2, device=dev).to(torch.bfloat16)
qkv = torch.empty(M, BT, 3 * C, device=dev, dtype=torch.bfloat16)
dhid = torch.empty_like(hid)
w = torch.randn(M, NG, gsz * hs, gsz * h2, device=dev).to(torch.bfloat16)
hg = hid.as_strided((M, NG, BT, gsz * h2), (BT * 3 * H * h2, gsz * h2, 3 * H * h2, 1))
dg = dhid.as_strided((M, NG, BT, gsz * h2), (BT * 3 * H * h2, gsz * h2, 3 * H * h2, 1))
qg = qkv.as_strided((M, NG, BT, gsz * hs), (BT * 3 * C, gsz * hs, 3 * C, 1))
def t(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
fwd = t(lambda: lib.gemm(hg, w, out_bf16=qg))
bwd = t(lambda: lib.gemm(qg, w, b_major=1, aux=hg, aux_mode=lib.AUX_TANH_BWD, out_bf16=dg))
fb = (hid.numel() + qkv.numel()) * 2
bb = (qkv.numel() + 2 * hid.numel()) * 2
print(json.dumps(dict(env={k: v for k, v in os.environ.items() if k.startswith("TA_GEMM")}, fwd_ms=round(fwd, 4),
                      fwd_GBs=round(fb / fwd / 1e6, 1), bwd_ms=round(bwd, 4), bwd_GBs=round(bb / bwd / 1e6, 1))))

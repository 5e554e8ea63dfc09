"""Dev tool: attention forward only (for ncu).  usage: attn_fwd_only.py G B T H hs S [iters]"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tradeaid_b200 import lib
G, B, T, H, hs, S = map(int, sys.argv[1:7])
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 5
C = H * hs
torch.manual_seed(0)
q = torch.randn(G, B * T, C, device="cuda").to(torch.bfloat16)
kv = [torch.randn(G, B * T, 2 * C, device="cuda").to(torch.bfloat16) for _ in range(S)]
o = torch.zeros_like(q); o_src = torch.zeros(S, G, B * T, C, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(S, G, B, H, T, device="cuda")
ks, vs = kv, [t[..., hs:] for t in kv]
def fwd(): lib.attention_fwd(q, ks, vs, o, lse, B=B, T=T, H=H, hs=hs, head_stride_kv=2 * hs, o_src=(o_src if S > 1 and not os.environ.get("NO_OSRC") else None))
for _ in range(2): fwd()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): fwd()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(json.dumps(dict(shape=sys.argv[1:7], ms=ms, tflops=2.0 * G * B * H * S * T * T * hs * 2 / 2 / ms / 1e9)))

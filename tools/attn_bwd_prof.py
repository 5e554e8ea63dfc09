"""Dev tool: per-kernel durations of one attention backward call (torch.profiler / CUPTI). usage: attn_bwd_prof.py G B T H hs S"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from tradeaid_b200 import lib
G, B, T, H, hs, S = map(int, sys.argv[1:7])
C = H * hs
torch.manual_seed(0)
q = torch.randn(G, B * T, C, device="cuda").to(torch.bfloat16)
kv = [torch.randn(G, B * T, 2 * C, device="cuda").to(torch.bfloat16) for _ in range(S)]
d_o = torch.randn(G, B * T, C, device="cuda").to(torch.bfloat16)
o = torch.zeros_like(q); o_src = torch.zeros(S, G, B * T, C, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(S, G, B, H, T, device="cuda")
dq = torch.zeros_like(q); dkv = [torch.zeros_like(t) for t in kv]
ks, vs = kv, [t[..., hs:] for t in kv]
lib.attention_fwd(q, ks, vs, o, lse, B=B, T=T, H=H, hs=hs, head_stride_kv=2 * hs, o_src=o_src if S > 1 else None)
def bwd(): lib.attention_bwd(q, ks, vs, o, d_o, lse, dq, dkv, [t[..., hs:] for t in dkv], B=B, T=T, H=H, hs=hs, head_stride_kv=2 * hs, o_src=o_src if S > 1 else None)
for _ in range(3): bwd()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10): bwd()
    torch.cuda.synchronize()
for e in prof.key_averages():
    print(f"{e.key[:70]:70s} n={e.count:3d} avg={e.device_time_total / max(e.count, 1):9.1f} us")

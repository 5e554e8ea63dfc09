"""Dev tool: dump the clock64 timeline of CTA 0 of the attention backward. usage: attn_bwd_trace.py G B T H hs S"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
buf = torch.zeros(4 * 1024, dtype=torch.int64, device="cuda")
os.environ["TA_ATTN_TRACE"] = hex(buf.data_ptr())
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
bwd(); torch.cuda.synchronize(); buf.zero_(); bwd(); torch.cuda.synchronize()
t = buf.cpu().view(4, 1024)
ev = []
names = {0: {1: "L kv slot free", 2: "L q slot0 free", 3: "L q slot1 free"}, 1: {1: "I operands ready", 2: "I S/dP h0 issued", 3: "I S/dP h1 issued", 4: "I pds_full seen", 5: "I grads issued", 6: "I S h0 issued", 7: "I S h1 issued"},
         2: {14: "M0 prefetch issued", 11: "M0 loads issued", 12: "M0 g_done(last) seen", 13: "M0 kv readout done", 7: "M0 prologue start", 8: "M0 bar1 passed", 9: "M0 readouts done", 10: "M0 D stored", 1: "M0 prologue done", 2: "M0 h0 avail", 3: "M0 h1 avail", 4: "M0 h0 math done", 5: "M0 g_done seen", 6: "M0 pds arrive", 15: "M0 S avail", 16: "M0 pair top"},
         3: {4: "G pds_full seen", 5: "G grads issued"}}
for role in range(4):
    for x in t[role].tolist():
        if x: ev.append((x >> 8, role, x & 255))
ev.sort()
t0 = ev[0][0]
skip = int(sys.argv[7]) if len(sys.argv) > 7 else 60
for c, role, tag in ev[skip:skip + 110]:
    print(f"{c - t0:8d}  {'      ' * role}{names[role].get(tag, tag)}")

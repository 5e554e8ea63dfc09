"""Where the 5e-3 logits error of the bf16 path comes from, on the CPU emulation of the engine (tests/emu_ops.py), c2-like shape
(2 modalities with cross-attention, C=256, H=8, L=4, T=64, B=8, random init): bf16 everywhere vs fp32 GEMM operands with bf16
attention vs split (hi+lo bf16) GEMM operands in the K-concatenated 3-term form the existing tcgen05 kernel can run.
Round-1 result: 5.3e-3 / 2.7e-4 / A-only 3.7e-3, B-only 3.8e-3, both 2.7e-4 (normwise, vs the fp32 oracle).
usage: python tests/soak/precision_split_experiment.py      (test infrastructure: imports the oracle)"""
import os
import sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import emu_ops
from oracle import restatement as R
from tradeaid_b200.engine import Engine
from tradeaid_b200.params import Dims
torch.manual_seed(0)
V, cross, T, C, H, L, B = [512, 256], [True, True], 64, 256, 8, 4, 8
cfg = R.Cfg(T, C, H, L, V, cross)
sd = R.random_state_dict(cfg, seed=3)
g = torch.Generator().manual_seed(1)
idx = [torch.randint(0, v, (B, T), generator=g) for v in V]
tgt = [torch.randint(0, v, (B, T), generator=g) for v in V]
ref_logits, ref_losses = R.forward(sd, cfg, idx, tgt)
def run(act_dtype, attn_bf16):
    orig = emu_ops.attention_fwd
    if attn_bf16:
        def attn(q, ks, vs, o, lse, **kw):
            bf = torch.bfloat16
            # operands q/k/v rounded to bf16 as the kernel reads them, P rounded to bf16 before P V, output kept fp32
            q2 = q.to(bf); ks2 = [k.to(bf) for k in ks]; vs2 = [v.to(bf) for v in vs]
            return orig(q2, ks2, vs2, o, lse, **kw)
        emu_ops.attention_fwd = attn
    try:
        eng = Engine(Dims(V, cross, T, C, H, L), "cpu", ops=emu_ops, act_dtype=act_dtype)
        nv = eng.layout.named_views(eng.master)
        for k, v in sd.items(): nv[k].copy_(v)
        logits, losses = eng.forward(torch.stack(idx), torch.stack(tgt), save=False)
    finally:
        emu_ops.attention_fwd = orig
    return [float((a.view_as(b) - b).norm() / b.norm()) for a, b in zip(logits, ref_logits)], [abs(float(a) - float(b)) / float(b) for a, b in zip(losses, ref_losses)]
print("bf16 everywhere (as built)      :", run(torch.bfloat16, False))
print("fp32 GEMM operands, bf16 attention:", run(torch.float32, True))
print("fp32 everything (wiring)          :", run(torch.float32, False))

def run_gemm_variant(split_a, split_b):
    """fp32 storage; every GEMM operand rounded to bf16 (1 term) or to hi+lo bf16 pairs (2 terms), products summed in fp32:
    the K-concatenation form [A_hi|A_lo|A_hi] x [B_hi|B_hi|B_lo] when both are split (lo*lo dropped)."""
    og, oa = emu_ops.gemm, emu_ops.attention_fwd
    bf = torch.bfloat16
    def parts(t, split):
        hi = t.to(bf).float()
        return (hi, (t - hi).to(bf).float()) if split else (hi, None)
    def gemm(A, B, **kw):
        a_hi, a_lo = parts(A.float(), split_a); b_hi, b_lo = parts(B.float(), split_b)
        am, bm = kw.get("a_major", 0), kw.get("b_major", 0)
        cat_dim_a = -1 if am == 0 else -2
        cat_dim_b = -1 if bm == 0 else -2
        As, Bs = [a_hi], [b_hi]
        if a_lo is not None: As.append(a_lo); Bs.append(b_hi)
        if b_lo is not None: As.append(a_hi); Bs.append(b_lo)
        return og(torch.cat(As, cat_dim_a), torch.cat(Bs, cat_dim_b), **kw)
    def attn(q, ks, vs, o, lse, **kw):
        return oa(q.to(bf), [k.to(bf) for k in ks], [v.to(bf) for v in vs], o, lse, **kw)
    emu_ops.gemm, emu_ops.attention_fwd = gemm, attn
    try:
        eng = Engine(Dims(V, cross, T, C, H, L), "cpu", ops=emu_ops, act_dtype=torch.float32)
        nv = eng.layout.named_views(eng.master)
        for k, v in sd.items(): nv[k].copy_(v)
        logits, losses = eng.forward(torch.stack(idx), torch.stack(tgt), save=False)
    finally:
        emu_ops.gemm, emu_ops.attention_fwd = og, oa
    return [round(float((a.view_as(b) - b).norm() / b.norm()), 6) for a, b in zip(logits, ref_logits)]
for sa, sb in [(False, False), (True, False), (False, True), (True, True)]:
    print(f"GEMM operands: A {'hi+lo' if sa else 'bf16 '}  B {'hi+lo' if sb else 'bf16 '}  bf16 attention ->", run_gemm_variant(sa, sb))

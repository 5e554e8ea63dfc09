"""Parity at the headline configuration and at the north star's stated tolerance.

* c4 widths: 8 modalities, cross-attention on the first four (7 key/value sources per querier, /root/reference/model.py:136-159),
  n_embd 1024, 16 heads, block_size 256, vocabulary 1024 per modality (grouped output heads), CTA-pair GEMMs and the
  block-diagonal stage 2 all at once — against the CPU oracle (oracle/restatement.py, pinned to the reference by
  tests/test_oracle.py), at depth 2 and at the full depth 12.
* precision="bf16x3" (fp32 activations, three-term split GEMM operands on the same tcgen05 kernel): logits within 2e-3
  normwise of the fp32 reference, on the reference-made goldens and on the c2 / c3 / c4 shapes.
* precision="bf16" (throughput mode): the bound is the measured error + margin, with torch.autocast(bf16) over the oracle as
  the control — single-pass bf16 operands cannot meet 2e-3, neither can torch's own autocast on the reference.
Every measured figure is appended to gpurun_out/parity_r02.jsonl (copied to profiles/ by the builder)."""
import json
import os

import pytest
import torch

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu

GOLDEN_CASES = ["demo", "m3", "m3_short", "hs32", "nocross"]
# bounds: logits normwise, loss relative, gradient normwise (all parameters together)
# bf16 = measured on B200 in round 2 (logits 5.5e-3 c2 ... 7.6e-3 c4 at depth 12; gradients 4.5e-2 ... 6.9e-2, growing with
# depth) plus a margin; torch.autocast(bf16) over the oracle, the control, gives 6.7-7.9e-3 on the logits of the same cases.
# bf16x3: the north star's 2e-3 on the logits with a wide margin (exact fp32 attention forward + split GEMM operands); its
# gradients still pass through the bf16 tensor-core attention backward and bf16 LayerNorm / bias-gradient inputs.
# measured in round 2 (profiles/r02_parity_v2.jsonl): logits 1.1e-5 (c2) ... 2.0e-5 (full c4), gradients 2.6e-3 ... 4.7e-3
BOUNDS = {"bf16": (9e-3, 2e-3, 7e-2), "bf16x3": (2e-4, 2e-5, 1e-2)}


def _record(**kw):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "parity_r02.jsonl"), "a") as f:
        f.write(json.dumps(kw) + "\n")


def _engine(V, cross, T, C, H, L, sd, precision):
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.params import Dims
    eng = Engine(Dims(V, cross, T, C, H, L), "cuda", precision=precision)
    nv = eng.layout.named_views(eng.master)
    for k, v in sd.items():
        nv[k].copy_(v)
    return eng


def _compare(eng, idx, tgt, ref_logits, ref_losses, ref_grads, tag, precision, bounds=None, check=True):
    lb, sb, gb = bounds or BOUNDS[precision]
    logits, losses = eng.forward(torch.stack(idx).cuda(), torch.stack(tgt).cuda(), save=True)
    eng.backward()
    torch.cuda.synchronize()
    rels = [float((a.cpu() - b).norm() / b.norm()) for a, b in zip(logits, ref_logits)]
    lrel = [abs(float(a) - float(b)) / abs(float(b)) for a, b in zip(losses.cpu(), ref_losses)]
    gv = eng.layout.named_views(eng.grad)
    num = den = 0.0
    worst = (0.0, None)
    for k, ref in ref_grads.items():
        if not eng.layout.is_active(k):
            continue
        e, r = float((gv[k].cpu() - ref).pow(2).sum()), float(ref.pow(2).sum())
        num += e
        den += r
        if r > 0 and (e / r) ** 0.5 > worst[0]:
            worst = ((e / r) ** 0.5, k)
    grel = (num / den) ** 0.5
    _record(test=tag, precision=precision, logits_normwise_max=max(rels), loss_rel_max=max(lrel), grad_normwise=grel,
            worst_grad=worst, bounds=[lb, sb, gb])
    if check:
        assert max(rels) <= lb, (tag, precision, rels)
        assert max(lrel) <= sb, (tag, precision, lrel)
        assert grel <= gb, (tag, precision, grel, worst)
    return max(rels), max(lrel), grel


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_bf16x3_meets_2e3_on_reference_goldens(name):
    """Fixtures made by the reference itself (tests/golden/make_golden.py): logits <= 2e-3 normwise in the bf16x3 mode."""
    g = load_golden(f"model_{name}.pt")
    c = g["cfg"]
    eng = _engine(c["V"], c["cross"], c["T"], c["C"], c["H"], c["L"], g["state_dict"], "bf16x3")
    _compare(eng, g["idx"], g["tgt"], g["logits"], g["losses"], g["grads"], f"golden_{name}", "bf16x3")
    # inference mode reuses buffers but runs the same kernels
    logits2, _ = eng.forward(torch.stack(g["idx"]).cuda(), None, save=False)
    for a, b in zip(logits2, g["logits"]):
        assert float((a.cpu() - b).norm() / b.norm()) <= 2e-3


SHAPES = {
    "c2": (2, [True, False], [1024, 1024], 8, 64, 256, 8, 4),                       # c2 at full depth, reduced batch
    "c3": (4, [True, False, False, True], [1024] * 4, 2, 128, 512, 8, 6),           # c3 at full depth
    "c4_L2": (8, [True] * 4 + [False] * 4, [1024] * 8, 2, 256, 1024, 16, 2),        # c4 widths, two layers
}


def _oracle_case(M, cross, V, B, T, C, H, L, autocast=False):
    from oracle import restatement as R
    cfg = R.Cfg(T, C, H, L, V, cross)
    sd = R.random_state_dict(cfg, seed=3)
    g = torch.Generator().manual_seed(5)
    idx = [torch.randint(0, V[i], (B, T), generator=g) for i in range(M)]
    tgt = [torch.randint(0, V[i], (B, T), generator=g) for i in range(M)]
    ref = R.forward_backward(sd, cfg, idx, tgt)
    control = None
    if autocast:      # the control: torch's own bf16 autocast over the same fp32 restatement (logits and gradients)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            al, _, ag = R.forward_backward(sd, cfg, idx, tgt)
        num = sum(float((ag[k].float() - r).pow(2).sum()) for k, r in ref[2].items())
        den = sum(float(r.pow(2).sum()) for r in ref[2].values())
        control = {"logits": max(float((a.float() - b).norm() / b.norm()) for a, b in zip(al, ref[0])),
                   "grads": (num / den) ** 0.5}
    return sd, idx, tgt, ref, control


@pytest.mark.parametrize("shape", list(SHAPES))
def test_both_precisions_against_oracle(shape):
    M, cross, V, B, T, C, H, L = SHAPES[shape]
    sd, idx, tgt, (rl, rs, rg), control = _oracle_case(M, cross, V, B, T, C, H, L, autocast=True)
    out = {}
    for precision in ("bf16", "bf16x3"):          # both are measured and recorded before anything is asserted
        eng = _engine(V, cross, T, C, H, L, sd, precision)
        out[precision] = _compare(eng, idx, tgt, rl, rs, rg, shape, precision, check=False)
        eng.release_buffers()
        del eng
        torch.cuda.empty_cache()
    _record(test=shape, control="torch.autocast(cpu, bf16) over the oracle", control_values=control, ours=out)
    for precision, (lg, ls, gr) in out.items():
        lb, sb, gb = BOUNDS[precision]
        assert lg <= lb and ls <= sb and gr <= gb, (shape, precision, lg, ls, gr)
    assert out["bf16x3"][0] < 0.02 * out["bf16"][0] and out["bf16x3"][2] < 0.2 * out["bf16"][2]
    # the single-pass bf16 mode is in the same class as torch's own autocast on the same model
    assert out["bf16"][0] <= 1.5 * control["logits"] + 1e-3, (out, control)
    assert out["bf16"][2] <= 2.0 * control["grads"] + 1e-2, (out, control)


def test_c4_full_depth_against_oracle():
    """BASELINE.json configs[3] as named: 8 modalities, n_embd 1024, 16 heads, 12 layers, block_size 256, vocabulary 1024 —
    the whole 1.89 G-parameter model, batch 2 (the oracle runs it on the host in well under a minute)."""
    M, cross, V, B, T, C, H, L = 8, [True] * 4 + [False] * 4, [1024] * 8, 2, 256, 1024, 16, 12
    sd, idx, tgt, (rl, rs, rg), _ = _oracle_case(M, cross, V, B, T, C, H, L)
    out = {}
    for precision in ("bf16", "bf16x3"):
        eng = _engine(V, cross, T, C, H, L, sd, precision)
        out[precision] = _compare(eng, idx, tgt, rl, rs, rg, "c4_full", precision, check=False)
        eng.release_buffers()
        del eng
        torch.cuda.empty_cache()
    # depth 12: the bf16 rounding of the stream gradient accumulates over 12 layers (measured 6.9e-2)
    bounds = {"bf16": (9e-3, 2e-3, 9e-2), "bf16x3": BOUNDS["bf16x3"]}
    for precision, (lg, ls, gr) in out.items():
        lb, sb, gb = bounds[precision]
        assert lg <= lb and ls <= sb and gr <= gb, ("c4_full", precision, lg, ls, gr)


@pytest.mark.parametrize("G,B,T,H,hs,S", [(1, 2, 256, 16, 64, 7), (1, 3, 200, 16, 64, 7)])
def test_attention_seven_sources_sixteen_heads(G, B, T, H, hs, S):
    """The c4 cross-attention launch shape at kernel level: 7 key/value sources, 16 heads of 64, T = 256 (forward incl. the
    per-source outputs and log-sum-exps, and backward) against fp32 torch."""
    from tradeaid_b200 import lib
    from test_kernels_gpu import _ref_attention
    torch.manual_seed(11)
    C = H * hs
    q = torch.randn(G, B * T, C, device="cuda").to(torch.bfloat16)
    kv = [torch.randn(G, B * T, H, 2 * hs, device="cuda").to(torch.bfloat16) for _ in range(S)]
    d_o = torch.randn(G, B * T, C, device="cuda").to(torch.bfloat16)
    o = torch.zeros(G, B * T, C, device="cuda", dtype=torch.bfloat16)
    o_src = torch.zeros(S, G, B * T, C, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(S, G, B, H, T, device="cuda")
    k_views = [t.view(G, B * T, 2 * C) for t in kv]
    v_views = [t.view(G, B * T, 2 * C)[..., hs:] for t in kv]
    lib.attention_fwd(q, k_views, v_views, o, lse, B=B, T=T, H=H, hs=hs, head_stride_kv=2 * hs, o_src=o_src)
    dq = torch.full_like(q, float("nan"))
    dkv = [torch.full((G, B * T, 2 * C), float("nan"), device="cuda", dtype=torch.bfloat16) for _ in range(S)]
    lib.attention_bwd(q, k_views, v_views, o, d_o, lse, dq, dkv, [t[..., hs:] for t in dkv], B=B, T=T, H=H, hs=hs,
                      head_stride_kv=2 * hs, o_src=o_src)
    torch.cuda.synchronize()
    qf = q.float().requires_grad_(True)
    kvf = [t.float().requires_grad_(True) for t in kv]
    ro, ros, rl = _ref_attention(qf, [t[..., :hs] for t in kvf], [t[..., hs:] for t in kvf], B, T, H, hs)
    assert (o.float() - ro).abs().max().item() <= 1.5e-2 * ro.abs().max().item()
    assert ((o.float() - ro).norm() / ro.norm()).item() <= 6e-3
    assert (o_src.float() - ros).abs().max().item() <= 1.5e-2 * ros.abs().max().item()
    assert (lse - rl).abs().max().item() <= 2e-3
    ro.backward(d_o.float())
    tol = 2e-2
    assert (dq.float() - qf.grad).abs().max().item() <= tol * qf.grad.abs().max().item()
    for s in range(S):
        ref = kvf[s].grad.view(G, B * T, 2 * C)
        assert (dkv[s].float() - ref).abs().max().item() <= tol * ref.abs().max().item()


@pytest.mark.parametrize("G,B,T,H,hs,S", [(2, 2, 100, 3, 24, 1), (1, 2, 256, 4, 64, 3), (1, 1, 600, 2, 64, 2), (1, 3, 4, 4, 8, 1)])
def test_attention_fwd_f32_is_exact(G, B, T, H, hs, S):
    """ta_attention_fwd_f32 (the bf16x3 mode's attention) against fp32 torch: outputs, per-source outputs and log-sum-exps to
    fp32 round-off, on strided views like the engine's packed q|k|v and [k|v]-per-head buffers."""
    from tradeaid_b200 import lib
    from test_kernels_gpu import _ref_attention
    torch.manual_seed(12)
    C = H * hs
    qkv = torch.randn(G, B * T, 3 * C, device="cuda")
    kv = [torch.randn(G, B * T, H, 2 * hs, device="cuda") for _ in range(S)]
    q = qkv[..., :C]
    o = torch.zeros(G, B * T, C, device="cuda")
    o_src = torch.zeros(S, G, B * T, C, device="cuda")
    lse = torch.zeros(S, G, B, H, T, device="cuda")
    k_views = [t.view(G, B * T, 2 * C) for t in kv]
    v_views = [t.view(G, B * T, 2 * C)[..., hs:] for t in kv]
    lib.attention_fwd_f32(q, k_views, v_views, o, lse, B=B, T=T, H=H, hs=hs, head_stride_kv=2 * hs, o_src=o_src)
    ro, ros, rl = _ref_attention(q.contiguous(), [t[..., :hs] for t in kv], [t[..., hs:] for t in kv], B, T, H, hs)
    assert (o - ro).abs().max().item() <= 2e-5 * max(1.0, ro.abs().max().item())
    assert (o_src - ros).abs().max().item() <= 2e-5 * max(1.0, ros.abs().max().item())
    assert (lse - rl).abs().max().item() <= 2e-5


@pytest.mark.parametrize("G,R,C,rows", [(3, 37, 40, False), (2, 64, 28, True), (1, 5, 1024, False)])
def test_split_and_convert_kernels(G, R, C, rows):
    """ta_split_bf16 / ta_convert against torch: hi = bf16(x), lo = bf16(x - hi), slots side by side or stacked; hi + lo
    reproduces x to 2^-16 relative."""
    from tradeaid_b200 import lib
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(2, G, R, C + 8, device="cuda", generator=g)[..., :C]          # strided source, two group dims
    shape = (2, G, 3 * R, C + 8) if rows else (2, G, R, 3 * C + 8)
    dst = torch.full(shape, 7.0, device="cuda", dtype=torch.bfloat16)
    view = dst[..., :C] if rows else dst[..., :3 * C]
    lib.split_bf16(x, view, lib.SPLIT_A, 3, rows)
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    want = torch.cat([hi, lo, hi], -2 if rows else -1)
    assert torch.equal(view, want)
    assert torch.all(dst[..., (C if rows else 3 * C):] == 7.0)                     # the pitch padding is untouched
    assert ((hi.float() + lo.float() - x).abs() <= 2.0 ** -16 * x.abs() + 1e-30).all()
    lib.split_bf16(x, view, lib.SPLIT_B, 3, rows)
    assert torch.equal(view, torch.cat([hi, hi, lo], -2 if rows else -1))
    y = torch.zeros(2, G, R, C, device="cuda", dtype=torch.bfloat16)
    lib.convert(x, y)
    assert torch.equal(y, hi)
    z = torch.zeros(2, G, R, C + 3, device="cuda")[..., :C]
    lib.convert(y, z)
    assert torch.equal(z, hi.float())


def test_layernorm_fwd_f32_matches_torch():
    from tradeaid_b200 import lib
    G, rows, C = 3, 50, 256
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(G, rows, C, device="cuda", generator=g) * 3 + 1
    w, b = torch.randn(G, C, device="cuda", generator=g), torch.randn(G, C, device="cuda", generator=g)
    y = torch.zeros(G, rows, C, device="cuda")
    mean, rstd = torch.zeros(G * rows, device="cuda"), torch.zeros(G * rows, device="cuda")
    lib.layernorm_fwd_f32(x, w, b, y, mean, rstd)
    for i in range(G):
        ref = torch.nn.functional.layer_norm(x[i], (C,), w[i], b[i])
        assert (y[i] - ref).abs().max().item() <= 2e-5

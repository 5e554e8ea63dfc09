"""Host logic of the bf16x3 precision mode (tradeaid_b200/precision.py) on the CPU emulation of the kernel namespace: the
engine's unchanged forward / backward program over fp32 activations and three-term split GEMM operands must land within the
north star's 2e-3 logits bound on the fixtures made by the reference itself — and well inside the single-pass bf16 error."""
import pytest
import torch

import emu_ops
from conftest import load_golden
from tradeaid_b200.engine import Engine
from tradeaid_b200.params import Dims
from tradeaid_b200.precision import SplitPrecisionOps

CASES = ["demo", "m3", "m3_short", "hs32", "nocross"]


def build(name, precision):
    g = load_golden(f"model_{name}.pt")
    c = g["cfg"]
    eng = Engine(Dims(c["V"], c["cross"], c["T"], c["C"], c["H"], c["L"]), "cpu", ops=emu_ops, precision=precision)
    nv = eng.layout.named_views(eng.master)
    for k, v in g["state_dict"].items():
        nv[k].copy_(v)
    return g, eng


@pytest.mark.parametrize("name", CASES)
def test_bf16x3_forward_backward_on_reference_goldens(name):
    g, eng = build(name, "bf16x3")
    assert isinstance(eng.ops, SplitPrecisionOps) and eng.act_dtype == torch.float32
    idx, tgt = torch.stack(g["idx"]), torch.stack(g["tgt"])
    logits, losses = eng.forward(idx, tgt, save=True)
    eng.backward()
    for a, b in zip(logits, g["logits"]):
        assert float((a - b).norm() / b.norm()) <= 1e-3
    for a, b in zip(losses, g["losses"]):
        assert abs(float(a) - float(b)) <= 1e-4 * abs(float(b))
    gv = eng.layout.named_views(eng.grad)
    num = sum(float((gv[k] - r).pow(2).sum()) for k, r in g["grads"].items() if eng.layout.is_active(k))
    den = sum(float(r.pow(2).sum()) for k, r in g["grads"].items() if eng.layout.is_active(k))
    assert (num / den) ** 0.5 <= 1e-2
    logits2, _ = eng.forward(idx, None, save=False)
    for a, b in zip(logits2, logits):
        assert torch.equal(a, b)


def test_weight_splits_follow_the_master_weights():
    """The splits of weight operands are kept between launches and re-made whenever the operand copies are refreshed
    (after an optimizer step): a forward after changing the master weights must see the new values."""
    g, eng = build("m3", "bf16x3")
    idx = torch.stack(g["idx"])
    a, _ = eng.forward(idx, None, save=False)
    a = [t.clone() for t in a]
    n_kept = len(eng.ops._weight_splits)
    assert n_kept > 0
    eng.master.mul_(1.5)
    eng.shadows_stale = True
    b, _ = eng.forward(idx, None, save=False)
    assert len(eng.ops._weight_splits) == n_kept          # refreshed in place, no new buffers
    assert any(not torch.allclose(x, y) for x, y in zip(a, b))
    ref = Engine(eng.d, "cpu", ops=emu_ops, precision="bf16x3")
    ref.master.copy_(eng.master)
    c, _ = ref.forward(idx, None, save=False)
    for x, y in zip(b, c):
        assert torch.equal(x, y)


def test_unknown_precision_is_rejected():
    with pytest.raises(ValueError):
        Engine(Dims([8, 8], [True, False], 4, 32, 4, 1), "cpu", ops=emu_ops, precision="fp8")


def test_kv_cache_generate_path_in_bf16x3():
    """Prompt pass + decode steps through the K/V cache in the bf16x3 mode equal the full forward of the same mode."""
    g, eng = build("m3", "bf16x3")
    from tradeaid_b200.engine import KVCache
    idx = torch.stack(g["idx"])
    M, B, T = idx.shape
    full, _ = eng.forward(idx, None, save=False)
    full = [t.clone() for t in full]
    cache = KVCache(eng, B, T)
    n0 = max(1, T // 2)
    eng.forward(idx[:, :, :n0].contiguous(), None, save=False, cache=cache)
    for t in range(n0, T):
        step, _ = eng.forward(idx[:, :, t:t + 1].contiguous(), None, save=False, cache=cache)
        for i in range(M):
            assert torch.allclose(step[i][:, 0], full[i][:, t], atol=2e-3, rtol=2e-3)

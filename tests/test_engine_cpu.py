"""Host-logic tests of the engine on CPU: the forward/backward PROGRAM (buffer wiring, grouping, packed layouts, the
hand-written backward) driven through tests/emu_ops.py and compared with fixtures made by the reference itself.
float32 activations: must agree with the reference to fp32 round-off (proves the wiring exactly);
bfloat16 activations: must agree within the bf16 budget of SURVEY.md appendix D."""
import pytest
import torch

import emu_ops
from conftest import load_golden
from tradeaid_b200.engine import Engine
from tradeaid_b200.params import Dims

CASES = ["demo", "m3", "m3_short", "hs32", "nocross"]


def build(name, act_dtype):
    g = load_golden(f"model_{name}.pt")
    c = g["cfg"]
    dims = Dims(c["V"], c["cross"], c["T"], c["C"], c["H"], c["L"])
    eng = Engine(dims, "cpu", ops=emu_ops, act_dtype=act_dtype)
    nv = eng.layout.named_views(eng.master)
    assert set(nv) == set(g["state_dict"])
    for k, v in g["state_dict"].items():
        nv[k].copy_(v)
    return g, eng


@pytest.mark.parametrize("name", CASES)
def test_forward_backward_exact_wiring(name):
    g, eng = build(name, torch.float32)
    idx, tgt = torch.stack(g["idx"]), torch.stack(g["tgt"])
    logits, losses = eng.forward(idx, tgt, save=True)
    for a, b in zip(logits, g["logits"]):
        assert (a - b).abs().max().item() <= 2e-5, (a - b).abs().max()
    for a, b in zip(losses, g["losses"]):
        assert abs(float(a) - float(b)) <= 2e-6 * max(1.0, abs(float(b)))
    eng.backward()
    gv = eng.layout.named_views(eng.grad)
    for k, ref in g["grads"].items():
        if not eng.layout.is_active(k):
            assert gv[k].abs().max().item() == 0.0
            continue
        sc = ref.abs().max().item() + 1e-8
        assert (gv[k] - ref).abs().max().item() <= 1e-4 * sc + 1e-7, k
    # inference mode (no saved activations, buffers reused across layers) gives the same logits
    logits2, none = eng.forward(idx, None, save=False)
    assert none is None
    for a, b in zip(logits2, g["logits"]):
        assert (a - b).abs().max().item() <= 2e-5


@pytest.mark.parametrize("name", CASES)
def test_forward_backward_bf16_budget(name):
    g, eng = build(name, torch.bfloat16)
    idx, tgt = torch.stack(g["idx"]), torch.stack(g["tgt"])
    logits, losses = eng.forward(idx, tgt, save=True)
    for a, b in zip(logits, g["logits"]):
        rel = (a - b).norm() / b.norm()
        assert rel <= 1.5e-2, rel          # normwise; SURVEY.md D measured ~6e-3 at init for bf16 operands
    for a, b in zip(losses, g["losses"]):
        assert abs(float(a) - float(b)) <= 2e-3 * abs(float(b))
    eng.backward()
    gv = eng.layout.named_views(eng.grad)
    num = den = 0.0
    for k, ref in g["grads"].items():
        if eng.layout.is_active(k):
            num += float((gv[k] - ref).pow(2).sum())
            den += float(ref.pow(2).sum())
    assert (num / den) ** 0.5 <= 3e-2


def test_runs_helper():
    from tradeaid_b200.engine import _runs
    assert _runs([0, 3]) == [(0, 3, 2)]
    assert _runs([0, 1, 2, 3]) == [(0, 1, 4)]
    assert _runs([1]) == [(1, 1, 1)]
    assert _runs([0, 1, 3, 7]) == [(0, 1, 2), (3, 4, 2)]
    assert _runs([0, 2, 4, 5]) == [(0, 2, 3), (5, 1, 1)]


def test_stage2_blockdiag_equals_per_head_groups(monkeypatch):
    """engine.py: stage 2 of the q/k/v MLPs as block-diagonal problems of gsz heads (the packing the CUDA path uses for
    hs in {32, 64}) must give exactly what one problem per (q|k|v, head) gives — the padding is exact zeros."""
    from oracle import restatement as R
    V, cross, T, C, H, L, B = [24, 10, 20], [True, False, True], 8, 256, 8, 2, 2
    sd = R.random_state_dict(R.Cfg(T, C, H, L, V, cross), seed=4)
    g = torch.Generator().manual_seed(9)
    idx = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in V])
    tgt = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in V])
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("TRADEAID_STAGE2_BLOCKDIAG", flag)
        eng = Engine(Dims(V, cross, T, C, H, L), "cpu", ops=emu_ops, act_dtype=torch.float32)
        assert (eng._bd is not None) == (flag == "1")
        nv = eng.layout.named_views(eng.master)
        for k, v in sd.items():
            nv[k].copy_(v)
        logits, losses = eng.forward(idx, tgt, save=True)
        eng.backward()
        outs[flag] = (torch.cat([x.flatten() for x in logits]).clone(), eng.grad.clone())
    assert torch.equal(outs["1"][0], outs["0"][0]) and torch.equal(outs["1"][1], outs["0"][1])
    # and against the reference restatement
    ref_logits, _, ref_grads = R.forward_backward(sd, R.Cfg(T, C, H, L, V, cross), list(idx), list(tgt))
    assert (outs["1"][0] - torch.cat([x.flatten() for x in ref_logits])).abs().max().item() <= 2e-5


def test_kv_cache_decode_equals_full_forward():
    """engine.KVCache host logic (prompt fill, per-position decode, cross-attention caches per source run) in fp32: every
    step must reproduce the full forward's logits at that position to round-off."""
    from oracle import restatement as R
    from tradeaid_b200.engine import KVCache
    V, cross, T, C, H, L, B = [24, 10, 20, 12], [True, False, True, False], 12, 32, 2, 2, 3
    eng = Engine(Dims(V, cross, T, C, H, L), "cpu", ops=emu_ops, act_dtype=torch.float32)
    sd = R.random_state_dict(R.Cfg(T, C, H, L, V, cross), seed=4)
    nv = eng.layout.named_views(eng.master)
    for k, v in sd.items():
        nv[k].copy_(v)
    g = torch.Generator().manual_seed(9)
    idx = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in V])
    ref_logits, _ = R.forward(sd, R.Cfg(T, C, H, L, V, cross), list(idx), None)
    full, _ = eng.forward(idx, None, save=False)
    full = [x.clone() for x in full]
    for a, b in zip(full, ref_logits):
        assert (a - b).abs().max().item() <= 2e-5
    cache = KVCache(eng, B)
    n0 = 5
    lg, _ = eng.forward(idx[:, :, :n0].contiguous(), None, save=False, cache=cache)
    assert max((a - b[:, :n0]).abs().max().item() for a, b in zip(lg, full)) <= 1e-6
    for t in range(n0, T):
        lg, _ = eng.forward(idx[:, :, t:t + 1].contiguous(), None, save=False, cache=cache)
        assert cache.length == t + 1
        assert max((a[:, 0] - b[:, t]).abs().max().item() for a, b in zip(lg, full)) <= 1e-6
    with pytest.raises(ValueError):
        eng.forward(idx[:, :, :1].contiguous(), None, save=False, cache=cache)      # cache is full


def test_grouped_output_heads_equal_per_modality_heads(monkeypatch):
    """engine.py: with one vocabulary size for every modality the output-head kinds are a constant stride apart in the
    flat buffer and go through one grouped launch per GEMM; logits, losses and gradients must equal the per-modality form
    (and the reference restatement)."""
    from oracle import restatement as R
    V, cross, T, C, H, L, B = [32, 32, 32], [True, False, True], 8, 32, 2, 2, 3
    cfg = R.Cfg(T, C, H, L, V, cross)
    sd = R.random_state_dict(cfg, seed=4)
    g = torch.Generator().manual_seed(9)
    idx = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in V])
    tgt = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in V])
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("TRADEAID_GROUPED_HEADS", flag)
        eng = Engine(Dims(V, cross, T, C, H, L), "cpu", ops=emu_ops, act_dtype=torch.float32)
        assert (eng._heads_grouped is not None) == (flag == "1")
        nv = eng.layout.named_views(eng.master)
        for k, v in sd.items():
            nv[k].copy_(v)
        logits, losses = eng.forward(idx, tgt, save=True)
        eng.backward()
        outs[flag] = (torch.cat([x.flatten() for x in logits]).clone(), losses.clone(), eng.grad.clone())
        eng.backward([1.0, 0.5, 2.0])      # weighted losses take the per-modality path in both forms
        outs[flag + "w"] = eng.grad.clone()
    assert all(torch.equal(a, b) for a, b in zip(outs["1"], outs["0"])) and torch.equal(outs["1w"], outs["0w"])
    ref_logits, _, ref_grads = R.forward_backward(sd, cfg, list(idx), list(tgt))
    assert (outs["1"][0] - torch.cat([x.flatten() for x in ref_logits])).abs().max().item() <= 2e-5
    # vocabularies of different sizes (or odd hidden widths) keep the per-modality launches
    assert Engine(Dims([32, 40], [False, False], T, C, H, 1), "cpu", ops=emu_ops)._heads_grouped is None
    assert Engine(Dims([30, 30], [False, False], T, C, H, 1), "cpu", ops=emu_ops)._heads_grouped is None


@pytest.mark.parametrize("name", ["m3", "m3_short"])
def test_per_layer_streams_match_reference(name):
    """The engine's residual stream after every block (saved for backward as x3 / x2) against forward hooks on the
    reference's blocks (tests/golden/layer_activations.pt), fp32 activations."""
    g, eng = build(name, torch.float32)
    acts = load_golden("layer_activations.pt")[name]
    idx, tgt = torch.stack(g["idx"]), torch.stack(g["tgt"])
    eng.forward(idx, tgt, save=True)
    B, T = idx.shape[1], idx.shape[2]
    for l, Sl in enumerate(eng._saved["layers"]):
        x = Sl["x3"] if eng.d.Xc else Sl["x2"]
        for i in range(eng.d.M):
            assert (x[i].view(B, T, -1) - acts[l][i]).abs().max().item() <= 2e-5, (l, i)


def test_decode_step_with_device_side_length_matches_the_full_forward(monkeypatch):
    """The decode step in its graph-replayable form (positions-so-far read from cache.len_dev by the embedding, cache-append
    and decode-attention ops) on the CPU emulation: every position's logits equal the full forward's."""
    from tradeaid_b200.engine import KVCache
    g, eng = build("m3", torch.float32)
    monkeypatch.setenv("TRADEAID_DECODE_DYN", "1")
    idx = torch.stack(g["idx"])
    M, B, T = idx.shape
    full, _ = eng.forward(idx, None, save=False)
    full = [t.clone() for t in full]
    cache = KVCache(eng, B, T)
    eng.forward(idx[:, :, :3].contiguous(), None, save=False, cache=cache)
    assert int(cache.len_dev) == 3
    for t in range(3, T):
        step = eng.decode_step(idx[:, :, t:t + 1].contiguous(), cache)
        assert cache.length == t + 1 and int(cache.len_dev) == t + 1
        for i in range(M):
            assert (step[i][:, 0] - full[i][:, t]).abs().max().item() <= 2e-5


@pytest.mark.parametrize("name", ["m3", "nocross"])
def test_activation_recompute_gives_the_same_gradients_with_fewer_buffers(name):
    """Engine(recompute=True): forward keeps one stream copy per block, backward reruns each block's forward into one shared
    set of buffers first. Same program, same order: losses and every gradient are bit-identical to the saving run."""
    g, eng = build(name, torch.bfloat16)
    idx, tgt = torch.stack(g["idx"]), torch.stack(g["tgt"])
    _, losses = eng.forward(idx, tgt, save=True)
    eng.backward()
    ref_grad, ref_losses, ref_bytes = eng.grad.clone(), losses.clone(), eng.activation_bytes()
    g2, eng2 = build(name, torch.bfloat16)
    eng2.recompute = True
    _, losses2 = eng2.forward(idx, tgt, save=True)
    assert eng2._saved["recompute"] and set(eng2._saved["layers"][0]) == {"x0"}
    eng2.backward()
    assert torch.equal(losses2, ref_losses)
    assert torch.equal(eng2.grad, ref_grad)
    if eng.d.n_layer > 1:
        assert eng2.activation_bytes() < ref_bytes
    # a second step reuses the shared buffers
    _, losses3 = eng2.forward(idx, tgt, save=True)
    eng2.backward()
    assert torch.equal(losses3, ref_losses) and torch.equal(eng2.grad, ref_grad)

"""Parity of the CUDA engine (through the C ABI) with fixtures made by the reference itself, and with the CPU oracle
restatement on seeded inputs at sizes the oracle finishes in seconds."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

CASES = ["demo", "m3", "m3_short", "hs32", "nocross"]


def build(name):
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.params import Dims
    g = load_golden(f"model_{name}.pt")
    c = g["cfg"]
    dims = Dims(c["V"], c["cross"], c["T"], c["C"], c["H"], c["L"])
    eng = Engine(dims, "cuda")
    nv = eng.layout.named_views(eng.master)
    for k, v in g["state_dict"].items():
        nv[k].copy_(v)
    return g, eng


@pytest.mark.parametrize("name", CASES)
def test_golden_forward_backward(name):
    g, eng = build(name)
    idx, tgt = torch.stack(g["idx"]).cuda(), torch.stack(g["tgt"]).cuda()
    logits, losses = eng.forward(idx, tgt, save=True)
    eng.backward()
    torch.cuda.synchronize()
    for a, b in zip(logits, g["logits"]):
        rel = (a.cpu() - b).norm() / b.norm()
        assert rel <= 1.5e-2, rel     # bf16 GEMM operands vs the fp32 reference, normwise
    for a, b in zip(losses.cpu(), g["losses"]):
        assert abs(float(a) - float(b)) <= 2e-3 * abs(float(b))
    gv = eng.layout.named_views(eng.grad)
    num = den = 0.0
    for k, ref in g["grads"].items():
        if eng.layout.is_active(k):
            num += float((gv[k].cpu() - ref).pow(2).sum())
            den += float(ref.pow(2).sum())
    assert (num / den) ** 0.5 <= 3e-2
    logits2, _ = eng.forward(idx, None, save=False)
    for a, b in zip(logits2, logits):
        assert torch.equal(a, b)      # inference mode reuses buffers but runs the same kernels: bit-identical


@pytest.mark.parametrize("M,cross,V,B,T,C,H,L", [
    (2, [True, False], [1024, 1024], 8, 64, 256, 8, 2),       # c2 shapes, reduced batch / depth
    (4, [True, False, False, True], [512, 300, 64, 1000], 2, 128, 512, 8, 1),   # c3 shapes
    (3, [True, True, True], [100, 100, 100], 1, 256, 128, 2, 1),   # T = 256 (two query tiles), hs = 64
])
def test_against_oracle(M, cross, V, B, T, C, H, L):
    from oracle import restatement as R
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.params import Dims
    cfg = R.Cfg(T, C, H, L, V, cross)
    sd = R.random_state_dict(cfg, seed=3)
    g = torch.Generator().manual_seed(5)
    idx = [torch.randint(0, V[i], (B, T), generator=g) for i in range(M)]
    tgt = [torch.randint(0, V[i], (B, T), generator=g) for i in range(M)]
    ref_logits, ref_losses, ref_grads = R.forward_backward(sd, cfg, idx, tgt)
    eng = Engine(Dims(V, cross, T, C, H, L), "cuda")
    nv = eng.layout.named_views(eng.master)
    for k, v in sd.items():
        nv[k].copy_(v)
    logits, losses = eng.forward(torch.stack(idx).cuda(), torch.stack(tgt).cuda(), save=True)
    eng.backward()
    for a, b in zip(logits, ref_logits):
        rel = (a.cpu() - b).norm() / b.norm()
        assert rel <= 1.5e-2, rel
    for a, b in zip(losses.cpu(), ref_losses):
        assert abs(float(a) - float(b)) <= 2e-3 * abs(float(b))
    gv = eng.layout.named_views(eng.grad)
    num = den = 0.0
    for k, ref in ref_grads.items():
        num += float((gv[k].cpu() - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
    # bf16 rounding of the stream gradient before dW = dY^T X: with random targets the true gradient is a small
    # residue of large per-row terms, so 2^-9 relative rounding shows up as a few percent normwise. The CPU emulation
    # with the same rounding points (tests/emu_ops.py) gives the same figure to 3 digits (0.0351 vs 0.0352 on c2 shapes).
    assert (num / den) ** 0.5 <= 5e-2, (num / den) ** 0.5
    # against the emulation with identical rounding points the CUDA path must agree much more tightly
    import emu_ops
    emu = Engine(Dims(V, cross, T, C, H, L), "cpu", ops=emu_ops, act_dtype=torch.bfloat16)
    ev = emu.layout.named_views(emu.master)
    for k, v in sd.items():
        ev[k].copy_(v)
    elog, elos = emu.forward(torch.stack(idx), torch.stack(tgt), save=True)
    emu.backward()
    for a, b in zip(logits, elog):
        assert ((a.cpu() - b).norm() / b.norm()).item() <= 5e-3   # same rounding points, different summation order


def _dropout_masks(eng, B, T, p):
    """The keep / (1 - p) multipliers the engine's kernels used in its LAST training forward, read back through the mask
    kernels (same seeds, same device step counter) and keyed like oracle.restatement.forward(masks=...)."""
    from tradeaid_b200 import lib
    d = eng.d
    M, C, H, L = d.M, d.n_embd, d.n_head, d.n_layer
    BT = B * T
    masks = {}
    ones = torch.ones(M, BT, C, device="cuda", dtype=torch.bfloat16)
    out = torch.empty_like(ones)
    for l in range(L):
        for site, key in ((eng.SITE_SA_PROJ, "sa_proj"), (eng.SITE_FFN, "ffn")):
            lib.dropout_bf16(ones, out, p, eng._seed(site, l), eng._drop_step)
            for i in range(M):
                masks[(key, l, i)] = ((out[i] > 0).float() / (1 - p)).view(B, T, C).cpu()
        am = torch.empty(1, M, B, H, T, T, device="cuda", dtype=torch.uint8)
        lib.dropout_attn_mask(am, 1, M, B, H, T, p, eng._seed(eng.SITE_SA_ATTN, l), eng._drop_step)
        for i in range(M):
            masks[("sa_attn", l, i)] = (am[0, i].float() / (1 - p)).cpu()
        if d.Xc:
            K = M - 1
            for xi, i in enumerate(d.xc):
                cm = torch.empty(K, 1, B, H, T, T, device="cuda", dtype=torch.uint8)
                lib.dropout_attn_mask(cm, K, 1, B, H, T, p, eng._seed(eng.SITE_CA_ATTN, l, xi), eng._drop_step)
                for s in range(K):
                    masks[("ca_attn", l, i, s)] = (cm[s, 0].float() / (1 - p)).cpu()
            for view, pos in eng._sub(ones, d.xc):      # one GEMM launch per contiguous run of enabled modalities
                n = view.shape[0]
                lib.dropout_bf16(ones[:n], out[:n], p, eng._seed(eng.SITE_CA_PROJ, l, pos[0]), eng._drop_step)
                for k, pp in enumerate(pos):
                    masks[("ca_proj", l, d.xc[pp])] = ((out[k] > 0).float() / (1 - p)).view(B, T, C).cpu()
    return masks


@pytest.mark.parametrize("M,cross,V,B,T,C,H,L", [
    (3, [True, False, True], [100, 64, 80], 2, 64, 128, 4, 2),
    (2, [True, True], [120, 90], 1, 256, 128, 2, 1),
])
def test_training_step_with_dropout_against_oracle(M, cross, V, B, T, C, H, L):
    """Dropout at the reference's five sites (model.py:69, 91, 116, 151, 171): the engine's forward/backward with p = 0.2
    against the oracle given the SAME masks. (torch's own Philox stream cannot be reproduced, see csrc/philox.cuh.)"""
    from oracle import restatement as R
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.params import Dims
    p = 0.2
    cfg = R.Cfg(T, C, H, L, V, cross)
    sd = R.random_state_dict(cfg, seed=4)
    g = torch.Generator().manual_seed(6)
    idx = [torch.randint(0, V[i], (B, T), generator=g) for i in range(M)]
    tgt = [torch.randint(0, V[i], (B, T), generator=g) for i in range(M)]
    eng = Engine(Dims(V, cross, T, C, H, L, dropout=p), "cuda")
    nv = eng.layout.named_views(eng.master)
    for k, v in sd.items():
        nv[k].copy_(v)
    logits, losses = eng.forward(torch.stack(idx).cuda(), torch.stack(tgt).cuda(), save=True, train=True)
    eng.backward()
    masks = _dropout_masks(eng, B, T, p)
    ref_logits, ref_losses, ref_grads = R.forward_backward(sd, cfg, idx, tgt, masks=masks)
    for a, b in zip(logits, ref_logits):
        rel = (a.cpu() - b).norm() / b.norm()
        assert rel <= 2e-2, rel
    for a, b in zip(losses.cpu(), ref_losses):
        assert abs(float(a) - float(b)) <= 3e-3 * abs(float(b))
    gv = eng.layout.named_views(eng.grad)
    num = den = 0.0
    for k, ref in ref_grads.items():
        num += float((gv[k].cpu() - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
    assert (num / den) ** 0.5 <= 6e-2, (num / den) ** 0.5
    # a second training forward draws different masks (device step counter), eval mode none at all
    l1 = [x.clone() for x in logits]
    logits2, _ = eng.forward(torch.stack(idx).cuda(), torch.stack(tgt).cuda(), save=True, train=True)
    assert not torch.equal(logits2[0], l1[0])
    e1, _ = eng.forward(torch.stack(idx).cuda(), None, save=False)
    e1 = [x.clone() for x in e1]
    e2, _ = eng.forward(torch.stack(idx).cuda(), None, save=False)
    assert all(torch.equal(a, b) for a, b in zip(e1, e2))


@pytest.mark.parametrize("C,H", [(256, 8), (256, 4)])
def test_stage2_blockdiag_is_bit_identical_to_per_head_groups(C, H, monkeypatch):
    """engine.py: q/k/v stage 2 packed gsz heads per block-diagonal problem vs one grouped problem per (q|k|v, head):
    the extra operands are exact zeros, so logits, losses and every gradient must be bit-identical."""
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.params import Dims
    V, cross, T, L, B = [100, 64, 80], [True, False, True], 64, 2, 4
    g = torch.Generator().manual_seed(8)
    idx = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in V]).cuda()
    tgt = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in V]).cuda()
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("TRADEAID_STAGE2_BLOCKDIAG", flag)
        eng = Engine(Dims(V, cross, T, C, H, L), "cuda")
        assert (eng._bd is not None) == (flag == "1")
        eng.layout.init_(eng.master, generator=torch.Generator(device="cuda").manual_seed(0))
        logits, losses = eng.forward(idx, tgt, save=True)
        eng.backward()
        out[flag] = ([x.clone() for x in logits], losses.clone(), eng.grad.clone())
    assert all(torch.equal(a, b) for a, b in zip(out["1"][0], out["0"][0]))
    assert torch.equal(out["1"][1], out["0"][1])
    assert torch.equal(out["1"][2], out["0"][2])


@pytest.mark.parametrize("M,cross,V,B,T,C,H,L,n0,scale", [
    (3, [True, False, True], [100, 64, 80], 3, 48, 128, 4, 2, 17, 3.0),
    # c5 widths (C=768, H=12), three query tiles. Weights at their init scale: at 3x both the decode path and the full
    # forward sit 2-5e-2 from the fp32 oracle (bf16 noise amplified; tests/soak/kv_diag.py), so they are 2-4e-2 apart
    (4, [True, False, False, True], [64, 64, 64, 64], 2, 300, 768, 12, 1, 290, 1.0),
    (1, [False], [50], 2, 16, 64, 2, 2, 1, 3.0),
])
def test_kv_cache_decode_matches_full_forward(M, cross, V, B, T, C, H, L, n0, scale):
    """engine.KVCache: prompt of n0 positions through the normal forward, then one position at a time against the cache;
    every step's logits must match the full forward over the whole sequence (same weights, same bf16 rounding points up
    to the attention kernel's summation order)."""
    from tradeaid_b200.engine import Engine, KVCache
    from tradeaid_b200.params import Dims
    eng = Engine(Dims(V, cross, T, C, H, L), "cuda")
    eng.layout.init_(eng.master, generator=torch.Generator(device="cuda").manual_seed(2))
    eng.master.mul_(scale)      # 3x: sharper attention than the N(0, 0.02) init gives
    g = torch.Generator().manual_seed(3)
    idx = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in V]).cuda()
    full, _ = eng.forward(idx, None, save=False)
    full = [x.clone() for x in full]
    cache = KVCache(eng, B)
    lg, _ = eng.forward(idx[:, :, :n0].contiguous(), None, save=False, cache=cache)
    for a, b in zip(lg, full):
        assert ((a - b[:, :n0]).norm() / b[:, :n0].norm()).item() <= 1e-3
    steps = list(range(n0, min(T, n0 + 12)))
    for t in steps:
        lg, _ = eng.forward(idx[:, :, t:t + 1].contiguous(), None, save=False, cache=cache)
        assert cache.length == t + 1
        for a, b in zip(lg, full):
            rel = ((a[:, 0] - b[:, t]).norm() / b[:, t].norm()).item()
            # the full forward rounds the un-normalised probabilities to bf16 before P V (tcgen05 operand), the decode
            # kernel keeps them in fp32
            assert rel <= 1e-2, (t, rel)
    with pytest.raises(ValueError):
        eng.forward(idx[:, :, :2].contiguous(), None, save=False, cache=cache)      # decode takes one position at a time


def test_grouped_output_heads_are_bit_identical(monkeypatch):
    """Output heads of all modalities in one grouped launch per GEMM (equal vocabulary sizes) vs one launch per modality."""
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.params import Dims
    V, cross, T, C, H, L, B = [256, 256, 256], [True, False, True], 64, 128, 4, 1, 4
    g = torch.Generator().manual_seed(8)
    idx = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in V]).cuda()
    tgt = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in V]).cuda()
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("TRADEAID_GROUPED_HEADS", flag)
        eng = Engine(Dims(V, cross, T, C, H, L), "cuda")
        assert (eng._heads_grouped is not None) == (flag == "1")
        eng.layout.init_(eng.master, generator=torch.Generator(device="cuda").manual_seed(0))
        logits, losses = eng.forward(idx, tgt, save=True)
        eng.backward()
        out[flag] = ([x.clone() for x in logits], losses.clone(), eng.grad.clone())
    assert all(torch.equal(a, b) for a, b in zip(out["1"][0], out["0"][0]))
    assert torch.equal(out["1"][1], out["0"][1]) and torch.equal(out["1"][2], out["0"][2])


@pytest.mark.parametrize("name", ["m3", "m3_short"])
def test_per_layer_streams_match_reference(name):
    """The CUDA engine's residual stream after every block against forward hooks on the reference's blocks
    (tests/golden/layer_activations.pt): bf16 GEMM operands, fp32 stream — normwise within the bf16 budget per layer."""
    g, eng = build(name)
    acts = load_golden("layer_activations.pt")[name]
    idx, tgt = torch.stack(g["idx"]).cuda(), torch.stack(g["tgt"]).cuda()
    eng.forward(idx, tgt, save=True)
    B, T = idx.shape[1], idx.shape[2]
    for l, Sl in enumerate(eng._saved["layers"]):
        x = Sl["x3"] if eng.d.Xc else Sl["x2"]
        for i in range(eng.d.M):
            ref = acts[l][i]
            rel = ((x[i].view(B, T, -1).cpu() - ref).norm() / ref.norm()).item()
            assert rel <= 1e-2, (l, i, rel)


def test_decode_step_graph_replay_is_bit_identical_to_plain_launches(monkeypatch):
    """Engine.decode_step: eager first step, captured second step, replays afterwards (the cached length read from device
    memory by the embedding / cache-append / decode-attention kernels) against the plain-launch decode path, bit for bit."""
    from tradeaid_b200.engine import KVCache
    g, eng = build("m3")
    idx = torch.stack(g["idx"]).cuda()
    M, B, T = idx.shape
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("TRADEAID_DECODE_GRAPH", mode)
        cache = KVCache(eng, B, T)
        eng.forward(idx[:, :, :2].contiguous(), None, save=False, cache=cache)
        got = []
        for t in range(2, T):
            step = eng.decode_step(idx[:, :, t:t + 1].contiguous(), cache)
            got.append(torch.stack([s.clone() for s in step]) if len({s.shape for s in step}) == 1 else [s.clone() for s in step])
        outs[mode] = got
        assert cache.length == T and int(cache.len_dev) == T
        assert (cache._graph is not None) == (mode == "1")
    for a, b in zip(outs["0"], outs["1"]):
        if isinstance(a, list):
            assert all(torch.equal(x, y) for x, y in zip(a, b))
        else:
            assert torch.equal(a, b)


@pytest.mark.parametrize("p", [0.0, 0.2])
def test_activation_recompute_is_bit_identical(p):
    """Engine(recompute=True) keeps one stream copy per block and reruns each block's forward inside backward(): the kernels are
    deterministic and the dropout masks are functions of (seed, step, coordinates), so losses and every gradient equal the
    saving run bit for bit - with and without dropout, eagerly and through the trainer's CUDA graph."""
    from oracle import restatement as R
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.params import Dims
    from tradeaid_b200.trainer import Trainer
    M, cross, V, B, T, C, H, L = 3, [True, False, True], [100, 64, 80], 2, 64, 128, 4, 3
    cfg = R.Cfg(T, C, H, L, V, cross)
    sd = R.random_state_dict(cfg, seed=4)
    g = torch.Generator().manual_seed(6)
    idx = torch.stack([torch.randint(0, V[i], (B, T), generator=g) for i in range(M)]).cuda()
    tgt = torch.stack([torch.randint(0, V[i], (B, T), generator=g) for i in range(M)]).cuda()
    out = []
    for rc in (False, True):
        eng = Engine(Dims(V, cross, T, C, H, L, dropout=p), "cuda", recompute=rc)
        nv = eng.layout.named_views(eng.master)
        for k, v in sd.items():
            nv[k].copy_(v)
        _, losses = eng.forward(idx, tgt, save=True, train=p > 0)
        eng.backward()
        torch.cuda.synchronize()
        out.append((losses.clone(), eng.grad.clone(), eng.activation_bytes()))
        if rc and p == 0.0:      # three optimizer steps as CUDA-graph replays equal three eager steps of the saving engine
            tr = Trainer(eng, lr=1e-3)
            tr.capture(idx.clone(), tgt.clone())
            got = [tr.replay().clone() for _ in range(3)]
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert out[1][2] < out[0][2]
    if p > 0.0:
        return
    eng = Engine(Dims(V, cross, T, C, H, L), "cuda")
    nv = eng.layout.named_views(eng.master)
    for k, v in sd.items():
        nv[k].copy_(v)
    tr = Trainer(eng, lr=1e-3)
    want = [tr.step(idx, tgt).clone() for _ in range(3)]
    for a, b in zip(got, want):
        assert torch.equal(a, b)

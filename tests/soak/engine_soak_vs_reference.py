"""Offline differential soak (BUILD CONTAINER ONLY: imports /root/reference): the engine's forward / backward PROGRAM
(engine.py over the CPU emulation of the kernel namespace, float32 storage = exact wiring) and the oracle restatement
against the reference model.py itself on RANDOM configurations - 1..4 modalities with random cross-attention flags, odd
vocabularies (down to 2 and 3: hidden width 1), heads 1..8, head sizes 8..64, 1..3 layers, windows shorter than block_size,
batch 1..3. Logits, losses and every parameter gradient (autograd of the reference) must agree to float32 round-off;
inactive parameters (single modality with the cross flag) must stay zero; then three AdamW steps of the
reference loop (torch.optim.AdamW) against the engine-owned Trainer (fused AdamW): loss curve and parameters.
usage: cd /tmp && python /root/repo/tests/soak/engine_soak_vs_reference.py [trials] [seed]"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, "/root/reference")
os.chdir("/tmp")
import config_utils  # noqa: E402
import emu_ops  # noqa: E402
from oracle import restatement as R  # noqa: E402
from tradeaid_b200.engine import Engine  # noqa: E402
from tradeaid_b200.params import Dims  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
worst = dict(logits=0.0, loss=0.0, grad=0.0, oracle=0.0)
for trial in range(trials):
    M = rng.choice([1, 2, 2, 3, 4])
    H = rng.choice([1, 2, 3, 4, 8])
    hs = rng.choice([8, 16, 32, 64])            # the engine's supported head sizes (params.Dims)
    if (H * hs) % 16:
        H += 1                                   # n_embd % 16 == 0
    C, L = H * hs, rng.choice([1, 2, 3])
    T = rng.choice([2, 4, 8, 16])
    Tr = rng.choice([T, T, rng.randint(1, T)])
    B = rng.choice([1, 2, 3])
    V = [rng.choice([2, 3, 5, 16, 57, 70]) for _ in range(M)]
    cross = [rng.random() < 0.6 for _ in range(M)]
    config_utils._config_cache = dict(device="cpu", batch_size=B, block_size=T, eval_iters=2, n_embd=C, n_head=H, n_layer=L,
                                      dropout=0.0, fixed_values=[], output_file_name="", project_file_path="/tmp/")
    for name in ("model",):
        sys.modules.pop(name, None)          # model.py reads the configuration at import (model.py:25-27)
    import model as ref_model
    torch.manual_seed(rng.randint(0, 10 ** 6))
    params = [["synthetic", i + 1, None, False, None, None, None, None, c, f"m{i}", None, None] for i, c in enumerate(cross)]
    m = ref_model.MultimodalTransformer(M, V, params)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for k, p in m.named_parameters():
            p.add_((0.1 if ("ln" in k or "norm" in k) else 0.02) * torch.randn(p.shape, generator=g))
    m.eval()
    idx = [torch.randint(0, V[i], (B, Tr), generator=g) for i in range(M)]
    tgt = [torch.randint(0, V[i], (B, Tr), generator=g) for i in range(M)]
    logits, losses = m(idx, tgt)
    m.zero_grad()
    sum(losses).backward()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items() if not k.endswith(".tril")}
    # oracle
    ocfg = R.Cfg(T, C, H, L, V, cross)
    o_logits, o_losses = R.forward(sd, ocfg, idx, tgt)
    for a, b in zip(o_logits, logits):
        worst["oracle"] = max(worst["oracle"], float((a - b).abs().max()))
        assert (a - b).abs().max().item() <= 5e-5, (trial, "oracle logits")
    # engine program on the emulation
    eng = Engine(Dims(V, cross, T, C, H, L), "cpu", ops=emu_ops, act_dtype=torch.float32)
    nv = eng.layout.named_views(eng.master)
    assert set(nv) == set(sd), (trial, set(nv) ^ set(sd))
    for k, v in sd.items():
        nv[k].copy_(v)
    e_logits, e_losses = eng.forward(torch.stack(idx), torch.stack(tgt), save=True)
    cfg_s = dict(M=M, V=V, cross=cross, B=B, T=T, Tr=Tr, C=C, H=H, L=L)
    for a, b in zip(e_logits, logits):
        d = float((a.view_as(b) - b).abs().max())
        worst["logits"] = max(worst["logits"], d)
        assert d <= 5e-5, (trial, cfg_s, "logits", d)
    for a, b in zip(e_losses, losses):
        d = abs(float(a) - float(b))
        worst["loss"] = max(worst["loss"], d)
        assert d <= 5e-6 * max(1.0, abs(float(b))), (trial, cfg_s, "loss", d)
    eng.backward()
    gv = eng.layout.named_views(eng.grad)
    grad_scale = {}
    for k, p in m.named_parameters():
        ref_g = p.grad if p.grad is not None else torch.zeros_like(p)
        grad_scale[k] = ref_g.abs().max().item()
        if not eng.layout.is_active(k):
            assert gv[k].abs().max().item() == 0.0 and ref_g.abs().max().item() == 0.0, (trial, cfg_s, k)
            continue
        sc = ref_g.abs().max().item() + 1e-8
        d = (gv[k] - ref_g).abs().max().item() / sc
        worst["grad"] = max(worst["grad"], d)
        assert d <= 3e-4 + 1e-6 / sc, (trial, cfg_s, k, d)
    # three AdamW steps: the reference loop of main.py:641-650 with torch.optim.AdamW against the engine-owned Trainer
    from tradeaid_b200.trainer import Trainer
    lr = rng.choice([1e-3, 3e-3])
    opt = torch.optim.AdamW(m.parameters(), lr=lr)
    tr = Trainer(eng, lr=lr)
    for _ in range(3):
        lg, ls = m(idx, tgt)
        opt.zero_grad(set_to_none=True)
        sum(ls).backward()
        opt.step()
        e_ls = tr.step(torch.stack(idx), torch.stack(tgt))
        d = max(abs(float(a) - float(b)) for a, b in zip(e_ls, ls))
        worst["curve"] = max(worst.get("curve", 0.0), d)
        assert d <= 2e-5, (trial, cfg_s, "loss curve", d)
    # Adam's m / sqrt(v) turns float32 round-off of near-zero gradient elements (dead ReLU units ...) into O(lr) differences
    # of single elements, so the UPDATE of every tensor is compared normwise
    after = dict(m.named_parameters())
    for k, v in nv.items():
        if eng.layout.is_active(k) and grad_scale[k] > 1e-7:     # (a 1-position window: q/k gradients are pure round-off)
            upd_ref = after[k].detach() - sd[k]
            d = float((v - after[k].detach()).norm() / (upd_ref.norm() + 1e-12))
            worst["param"] = max(worst.get("param", 0.0), d)
            assert d <= 2e-2, (trial, cfg_s, k, d)
print(f"engine soak: {trials} random configurations agree with the reference (max abs logits {worst['logits']:.1e}, loss {worst['loss']:.1e}, "
      f"max relative gradient {worst['grad']:.1e}; oracle logits {worst['oracle']:.1e}; 3 AdamW steps: losses {worst.get('curve', 0):.1e}, parameter updates {worst.get('param', 0):.1e} normwise)")

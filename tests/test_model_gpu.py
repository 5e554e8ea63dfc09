"""The drop-in surface on the GPU: nn.Module forward/backward with a stock torch.optim.AdamW (the reference's loop,
main.py:641-650), the engine-owned Trainer (fused AdamW, CUDA graph), the feeder, and the loss-curve criterion."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _model(name):
    from tradeaid_b200 import model as tm
    g = load_golden(f"model_{name}.pt")
    c = g["cfg"]
    tm.configure(c["T"], c["C"], c["H"], c["L"])
    params = [["f", i, None, False, None, None, None, None, c["cross"][i], f"m{i}"] for i in range(c["M"])]
    m = tm.MultimodalTransformer(c["M"], c["V"], params)
    m.load_state_dict(g["state_dict"], strict=True)
    return g, m.to("cuda")


@pytest.mark.parametrize("name", ["demo", "m3", "hs32", "nocross"])
def test_reference_loop_with_stock_adamw(name):
    """main.py's loop body, unmodified: logits, losses = m(xb, yb); sum(losses).backward(); torch AdamW step."""
    g, m = _model(name)
    m.train()
    idx = [t.cuda() for t in g["idx"]]
    tgt = [t.cuda() for t in g["tgt"]]
    opt = torch.optim.AdamW(m.parameters(), lr=g["cfg"]["lr"])
    curve = []
    for _ in range(g["cfg"]["steps"]):
        logits, losses = m(idx, tgt)
        total = sum(losses)
        opt.zero_grad(set_to_none=True)
        total.backward()
        opt.step()
        curve.append(total.item())
    for a, b in zip(curve, g["curve"]):
        assert abs(a - b) <= 1e-2, (curve, g["curve"])      # north-star loss-curve tolerance: 1e-2 absolute
    sd = m.state_dict()
    tot = cnt = 0.0
    for k, ref in g["state_dict_after"].items():
        # Adam normalises every element's step to ~lr = 1e-3 whatever the gradient size, so an element whose gradient is
        # below the bf16 noise floor may move the other way: bound = 2 * steps * lr; the bulk must agree far better.
        d = (sd[k].cpu() - ref).abs()
        assert d.max().item() <= 2 * g["cfg"]["steps"] * g["cfg"]["lr"] + 1e-4, k
        tot += float(d.sum()); cnt += d.numel()
    assert tot / cnt <= 2e-4, tot / cnt
    assert list(sd.keys()) == g["state_dict_keys"]


def test_forward_without_targets_and_eval_nograd():
    g, m = _model("m3")
    m.eval()
    idx = [t.cuda() for t in g["idx"]]
    logits, none = m(idx)
    assert none is None
    for a, b in zip(logits, g["logits"]):
        assert a.shape == b.shape and ((a.cpu() - b).norm() / b.norm()).item() <= 1.5e-2
    with torch.no_grad():
        logits2, losses = m(idx, [t.cuda() for t in g["tgt"]])
    for a, b in zip(losses, g["losses"]):
        assert abs(a.item() - b.item()) <= 2e-3 * abs(b.item())
    for a, b in zip(logits, logits2):
        assert torch.equal(a, b)


def test_generate_bookkeeping():
    g, m = _model("m3")
    m.eval()
    torch.manual_seed(0)
    idx = [t[:, :5].cuda() for t in g["idx"]]
    out = m.generate(idx, max_new_tokens=6, modality_to_generate=2)
    assert [o.shape[1] for o in out] == [11, 11, 11]
    for i in (0, 1):   # other modalities are padded with their own last token (model.py:435-440)
        assert torch.equal(out[i][:, :5], idx[i]) and (out[i][:, 5:] == idx[i][:, -1:]).all()
    assert torch.equal(out[2][:, :5], idx[2]) and int(out[2].max()) < g["cfg"]["V"][2]


def test_generate_with_kv_cache_matches_recompute():
    """generate(): the K/V-cache path (growing phase) against the reference's recompute-everything form with the same
    torch seed — the sampled tokens agree (logits equal up to bf16 summation order; a rare flip of a near-tie sample is
    tolerated), and past block_size both fall back to the sliding-window full forward."""
    g, m = _model("m3")
    m.eval()
    T = g["cfg"]["T"]          # block_size 8
    idx = [t[:, :3].cuda() for t in g["idx"]]
    torch.manual_seed(5)
    a = m.generate(idx, max_new_tokens=9, modality_to_generate=0, use_cache=True)      # 3 -> 12 > block_size: slides at the end
    torch.manual_seed(5)
    b = m.generate(idx, max_new_tokens=9, modality_to_generate=0, use_cache=False)
    assert [x.shape for x in a] == [x.shape for x in b]
    same = (a[0] == b[0]).float().mean().item()
    assert same >= 0.9, same
    for i in (1, 2):
        assert torch.equal(a[i], b[i])
    assert T == 8


def test_trainer_matches_golden_curve_and_graph_replay_is_identical():
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.params import Dims
    from tradeaid_b200.trainer import Trainer
    g = load_golden("model_m3.pt")
    c = g["cfg"]

    def make():
        eng = Engine(Dims(c["V"], c["cross"], c["T"], c["C"], c["H"], c["L"]), "cuda")
        nv = eng.layout.named_views(eng.master)
        for k, v in g["state_dict"].items():
            nv[k].copy_(v)
        return eng, Trainer(eng, lr=c["lr"])

    idx, tgt = torch.stack(g["idx"]).cuda(), torch.stack(g["tgt"]).cuda()
    eng, tr = make()
    curve = [float(tr.step(idx, tgt).sum()) for _ in range(c["steps"])]
    for a, b in zip(curve, g["curve"]):
        assert abs(a - b) <= 1e-2
    eng2, tr2 = make()
    tr2.capture(idx, tgt)
    curve2 = [float(tr2.replay().sum()) for _ in range(c["steps"])]
    assert curve2 == curve                                   # same kernels, same order: bit-identical
    assert torch.equal(eng2.master, eng.master)
    assert tr2.launches_per_step > 50


def test_feeder_bit_exact_on_device():
    from tradeaid_b200 import training_utils as tu
    g = load_golden("get_batch.pt")
    full, n_val = g["full"], g["n_val"]
    tu.configure(device="cuda", batch_size=g["batch_size"], block_size=g["block_size"], eval_iters=2)
    tu._stores.clear()
    tu.all_train_sets = [d[:-n_val].tolist() for d in full]
    tu.all_val_sets = [d[-n_val:].clone() for d in full]
    tu.all_vocabularies = [[float(k) for k in range(50)]] * 3
    tu.all_modality_params = [["s", i, None, False, None, None, None, None, False, f"m{i}"] for i in range(3)]
    tu.file_lengths = g["file_lengths"]
    tu.num_modalities = 3
    tu.is_percents = False
    torch.manual_seed(g["seed"])
    for call in g["calls"]:
        xb, yb = tu.get_batch(call["split"], 1)
        for a, b in zip(xb, call["xb"]):
            assert a.is_cuda and a.dtype == torch.int64 and torch.equal(a.cpu(), b)
        for a, b in zip(yb, call["yb"]):
            assert torch.equal(a.cpu(), b)
    tu._stores.clear()


def test_estimate_loss_matches_oracle(capsys):
    """training_utils.estimate_loss (training_utils.py:387-520) on the device: mean summed loss and directional counts of
    eval_iters batches per split, against the oracle restatement fed the same batches (same torch seed -> same indices)."""
    from oracle import restatement as R
    from tradeaid_b200 import model as tm, training_utils as tu
    g, m = _model("m3")
    c = g["cfg"]
    V, T, B, iters = c["V"], c["T"], 5, 3
    gen = torch.Generator().manual_seed(12)
    full = [torch.randint(0, v, (400,), generator=gen) for v in V]
    n_val = 80
    tu.configure(device="cuda", batch_size=B, block_size=T, eval_iters=iters)
    tu._stores.clear()
    tu.all_train_sets = [d[:-n_val].tolist() for d in full]
    tu.all_val_sets = [d[-n_val:].clone() for d in full]
    tu.all_vocabularies = [[0.25 * k - 3.0 for k in range(V[0])], [float(k - 1) for k in range(V[1])], [f"s{k}" for k in range(V[2])]]
    tu.all_modality_params = [["s", i, None, i == 1, None, None, None, None, c["cross"][i], f"m{i}"] for i in range(3)]
    tu.file_lengths = [400]
    tu.num_modalities = 3
    tu.is_percents = False
    tu.m = m
    m.train()
    torch.manual_seed(21)
    out = tu.estimate_loss(7, 100)
    assert m.training                                   # estimate_loss restores train mode (training_utils.py:519)
    printed = capsys.readouterr().out
    # the oracle on the same batches
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items() if not k.endswith(".tril")}
    cfg = R.Cfg(T, c["C"], c["H"], c["L"], V, c["cross"])
    torch.manual_seed(21)
    for split, data in (("train", [d[:-n_val] for d in full]), ("val", [d[-n_val:] for d in full])):
        tot, wins, losses = [], [0, 0, 0], [0, 0, 0]
        for _ in range(iters):
            ix = R.batch_starting_indices(len(data[0]), T, B, split, [400], False)
            xb, yb = R.gather_batch(data, ix, T)
            logits, ls = R.forward(sd, cfg, xb, yb)
            tot.append(float(sum(ls)))
            w, l, _, _ = R.directional_metrics(logits, xb, yb, tu.all_vocabularies, tu.all_modality_params)
            wins = [a + b for a, b in zip(wins, w)]
            losses = [a + b for a, b in zip(losses, l)]
        assert abs(out[split] - sum(tot) / iters) <= 2e-3 * abs(sum(tot) / iters)
        disp = "Train Set" if split == "train" else "Val Set"
        block = printed.split(f"DIRECTIONAL METRICS - {disp}")[1]
        for i in (0, 1):      # arg-max of bf16-path logits vs fp32 logits can flip on a near-tie: allow one sample
            import re
            mm = re.search(rf"m{i}\s+(\d+)/(\d+)", block)
            assert mm and int(mm.group(2)) == wins[i] + losses[i] == iters * B
            assert abs(int(mm.group(1)) - wins[i]) <= 1
        assert "m2: No data processed (non-numeric)" in block
    tu._stores.clear()
    tu.m = None


def test_200_step_loss_curve_vs_reference_golden():
    """North-star criterion against the reference itself: tests/golden/curve200.pt holds the 200 losses of the UNMODIFIED
    reference (its own get_batch, model, torch AdamW; lr 1e-3). The engine's trainer fed the same index stream (same
    torch seed) must stay within 1e-2 absolute over the whole curve."""
    from oracle import restatement as R
    from tradeaid_b200 import synthetic as syn
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.feeder import TokenStore, generate_batch_starting_indices
    from tradeaid_b200.params import Dims
    from tradeaid_b200.trainer import Trainer
    g = load_golden("curve200.pt")
    c = g["cfg"]
    sd = R.random_state_dict(R.Cfg(c["T"], c["C"], c["H"], c["L"], c["V"], c["cross"]), seed=0)
    for k in list(sd):
        if ".heads." in k and k.endswith(".2.weight") and "sa_layers" in k:
            sd[k] = sd[k] / 10
    eng = Engine(Dims(c["V"], c["cross"], c["T"], c["C"], c["H"], c["L"]), "cuda")
    nv = eng.layout.named_views(eng.master)
    for k, v in sd.items():
        nv[k].copy_(v)
    tr = Trainer(eng, lr=c["lr"])
    store = TokenStore([syn.token_stream("price", c["N"], 64, 1000), syn.token_stream("indicator", c["N"], 64, 1001)], "cuda")
    torch.manual_seed(c["seed"])
    worst = 0.0
    for step in range(c["steps"]):
        ix = generate_batch_starting_indices(c["N"], c["T"], c["B"], "train", [c["N"]], False)
        xb, yb = store.gather(ix, c["T"])
        worst = max(worst, abs(float(tr.step(xb, yb).sum()) - g["curve"][step]))
    assert worst <= 1e-2, worst


def test_200_step_loss_curve_vs_oracle():
    """North-star criterion: the 200-step loss curve matches the fp32 reference within 1e-2 absolute (identical
    batches, identical initial weights, lr 1e-3)."""
    from oracle import restatement as R
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.params import Dims
    from tradeaid_b200.trainer import Trainer
    V, cross, T, C, H, L, B = [64, 64], [True, False], 32, 128, 4, 2, 16
    cfg = R.Cfg(T, C, H, L, V, cross)
    sd = R.random_state_dict(cfg, seed=0)
    for k in list(sd):      # reference-like scale for the second q/k/v matrices
        if ".heads." in k and k.endswith(".2.weight") and "sa_layers" in k:
            sd[k] = sd[k] / 10
    # SURVEY.md appendix E streams (a price-like reflected random walk and its EMA indicator): a learnable signal, as in
    # the real workload. (On unlearnable noise the 200-step Adam trajectory is chaotic: even an fp32 run that differs
    # from the reference only in summation order drifts by 1.4e-2 — measured with tests/emu_ops.py in float32.)
    from tradeaid_b200 import synthetic as syn
    data = [syn.token_stream("price", 4096, 64, 1000), syn.token_stream("indicator", 4096, 64, 1001)]
    eng = Engine(Dims(V, cross, T, C, H, L), "cuda")
    nv = eng.layout.named_views(eng.master)
    for k, v in sd.items():
        nv[k].copy_(v)
    tr = Trainer(eng, lr=1e-3)
    params = {k: v.clone() for k, v in sd.items()}
    m1 = {k: torch.zeros_like(v) for k, v in params.items()}
    m2 = {k: torch.zeros_like(v) for k, v in params.items()}
    torch.manual_seed(2)
    worst = 0.0
    for step in range(1, 201):
        ix = R.batch_starting_indices(4096, T, B, "train", [4096], False)
        xb, yb = R.gather_batch(data, ix, T)
        _, losses, grads = R.forward_backward(params, cfg, xb, yb)
        R.adamw_step(params, grads, m1, m2, step, 1e-3)
        ours = tr.step(torch.stack(xb).cuda(), torch.stack(yb).cuda())
        worst = max(worst, abs(float(ours.sum()) - float(sum(losses))))
    assert worst <= 1e-2, worst

"""CPU tests of the drop-in surface: state-dict compatibility with the reference, feeder host logic (bit-exact indices
and windows through the emulated gather), metrics, and that the C-ABI library exports every declared symbol."""
import ctypes
import os
import re

import pytest
import torch

import emu_ops
from conftest import ROOT, load_golden


def _model(name):
    from tradeaid_b200 import model as tm
    g = load_golden(f"model_{name}.pt")
    c = g["cfg"]
    tm.configure(c["T"], c["C"], c["H"], c["L"])
    params = [["f", i, None, False, None, None, None, None, c["cross"][i], f"m{i}"] for i in range(c["M"])]
    return g, tm.MultimodalTransformer(c["M"], c["V"], params)


@pytest.mark.parametrize("name", ["demo", "m3", "nocross"])
def test_state_dict_is_reference_compatible(name):
    g, m = _model(name)
    sd = m.state_dict()
    assert list(sd.keys()) == g["state_dict_keys"]          # same keys, same order, incl. the tril buffers
    T = g["cfg"]["T"]
    for k, v in sd.items():
        if k.endswith(".tril"):
            assert torch.equal(v, torch.tril(torch.ones(T, T)))
        else:
            assert v.shape == g["state_dict"][k].shape
    # load a reference checkpoint (with tril buffers), strict
    ref_sd = dict(g["state_dict"])
    for k in g["state_dict_keys"]:
        if k.endswith(".tril"):
            ref_sd[k] = torch.tril(torch.ones(T, T))
    m.load_state_dict(ref_sd, strict=True)
    for k, v in g["state_dict"].items():
        assert torch.equal(m.state_dict()[k], v)
    # parameters are views into ONE flat buffer
    flat = m._flat
    for _, p in m.named_parameters():
        assert p.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
    assert [n for n, _ in m.named_parameters()] == [k for k in g["state_dict_keys"] if not k.endswith(".tril")]


def test_init_statistics():
    _, m = _model("m3")
    sd = m.state_dict()
    w = sd["blocks.0.ffwd_layers.0.net.0.weight"]
    assert abs(float(w.std()) - 0.02) < 2e-3 and abs(float(w.mean())) < 2e-3
    assert float(sd["blocks.0.ffwd_layers.0.net.0.bias"].abs().max()) == 0.0
    assert torch.equal(sd["blocks.0.ln1_layers.0.weight"], torch.ones(32))
    assert float(sd["blocks.0.ln1_layers.0.bias"].abs().max()) == 0.0


def test_unsupported_shapes_are_refused_at_construction():
    """Shapes the kernels cannot take are refused with a ValueError when the dimensions are declared, not by a TMA descriptor
    error in the middle of a step: head sizes outside {8, 16, ..., 64}, n_embd not divisible by n_head or by 16 (the
    projection MLP's hidden width n_embd/2 is a bf16 row pitch; found by tests/soak/engine_soak_vs_reference.py)."""
    from tradeaid_b200.params import Dims
    for C, H in ((24, 3), (40, 5), (8, 1), (30, 4), (32, 8), (256, 2)):
        with pytest.raises(ValueError):
            Dims([16, 16], [True, False], 8, C, H, 1)
    for C, H in ((16, 2), (48, 3), (64, 1), (1024, 16)):
        Dims([16, 16], [True, False], 8, C, H, 1)


def test_module_under_the_stock_loop_on_the_emulation():
    """The nn.Module's host logic without a GPU (model._ops_override = the CPU emulation): forward through the autograd
    bridge, `sum(losses).backward()` deposits the oracle's gradients into `.grad`, a stock torch AdamW step moves the flat
    buffer, and generate() gives the same tokens with and without the K/V cache. (tests/soak/stock_loop_soak_vs_reference.py
    runs the same loop against the reference model itself, offline.)"""
    from oracle import restatement as R
    from tradeaid_b200 import model as tm
    g = load_golden("model_m3.pt")
    c = g["cfg"]
    tm.configure(block_size=c["T"], n_embd=c["C"], n_head=c["H"], n_layer=c["L"], dropout=0.0)
    params = [["s", i + 1, None, False, None, None, None, None, c["cross"][i], f"m{i}"] for i in range(c["M"])]
    tm._ops_override = emu_ops
    try:
        m = tm.MultimodalTransformer(c["M"], c["V"], params)
        m.load_state_dict({k: v for k, v in g["state_dict"].items()}, strict=False)
        m = m.to("cpu")
        logits, losses = m(g["idx"], g["tgt"])
        for a, b in zip(logits, g["logits"]):
            assert (a - b).abs().max().item() <= 2e-5
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
        opt.zero_grad(set_to_none=True)
        sum(losses).backward()
        for k, p in m.named_parameters():
            ref = g["grads"][k]
            if p.grad is None:
                assert ref.abs().max().item() == 0.0, k
            else:
                assert (p.grad - ref).abs().max().item() <= 1e-4 * (ref.abs().max().item() + 1e-8) + 1e-7, k
        # gradient accumulation: a second backward without zero_grad adds to .grad, as torch does
        first = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        _, losses_b = m(g["idx"], g["tgt"])
        (2.0 * losses_b[0] + sum(losses_b[1:])).backward()
        sd_here = {k: v.detach().clone() for k, v in m.state_dict().items() if not k.endswith(".tril")}
        ocfg = R.Cfg(c["T"], c["C"], c["H"], c["L"], c["V"], c["cross"])
        _, ref_losses = R.forward({k: v.requires_grad_(True) for k, v in sd_here.items()}, ocfg, g["idx"], g["tgt"])
        (2.0 * ref_losses[0] + sum(ref_losses[1:])).backward()
        for k, p in m.named_parameters():
            if k in first:
                want_g = first[k] + (sd_here[k].grad if sd_here[k].grad is not None else 0)
                assert (p.grad - want_g).abs().max().item() <= 2e-4 * (want_g.abs().max().item() + 1e-8) + 1e-7, k
        opt.zero_grad(set_to_none=True)
        _, losses_c = m(g["idx"], g["tgt"])
        sum(losses_c).backward()                                        # back to one step's gradients for the optimizer
        before = m._flat.clone()
        opt.step()
        assert not torch.equal(before, m._flat)                       # the stock optimizer updated the engine's flat buffer
        # ... and the next forward sees the new weights - in EVAL mode too (estimate_loss right after optimizer.step();
        # the operand copies used to be refreshed only by training-mode forwards: evaluation ran one step behind)
        m.eval()
        with torch.no_grad():
            lg_eval, _ = m(g["idx"], g["tgt"])
        sd_now = {k: v.detach().clone() for k, v in m.state_dict().items() if not k.endswith(".tril")}
        want, _ = R.forward(sd_now, R.Cfg(c["T"], c["C"], c["H"], c["L"], c["V"], c["cross"]), g["idx"], g["tgt"])
        for a, b in zip(lg_eval, want):
            assert (a - b).abs().max().item() <= 2e-5
        assert (lg_eval[0] - logits[0]).abs().max().item() > 1e-4
        m.train()
        lg2, _ = m(g["idx"], g["tgt"])
        assert (lg2[0] - lg_eval[0]).abs().max().item() <= 2e-5
        m.eval()
        prompt = [t[:, :2].clone() for t in g["idx"]]
        outs = []
        for use_cache in (True, False):
            torch.manual_seed(5)
            outs.append(m.generate([p.clone() for p in prompt], max_new_tokens=c["T"] + 2, modality_to_generate=1, use_cache=use_cache))
        assert all(torch.equal(a, b) for a, b in zip(*outs)) and outs[0][1].shape[1] == 2 + c["T"] + 2
    finally:
        tm._ops_override = None


def test_cpu_forward_fails_loudly():
    _, m = _model("demo")
    with pytest.raises(RuntimeError, match="CUDA"):
        m([torch.zeros(2, 4, dtype=torch.long)] * 2)


def test_indices_bit_exact_product_side():
    from tradeaid_b200.feeder import generate_batch_starting_indices
    for c in load_golden("indices.pt"):
        torch.manual_seed(c["seed"])
        assert torch.equal(generate_batch_starting_indices(*c["args"]), c["ix"])
        assert torch.equal(generate_batch_starting_indices(*c["args"]), c["ix2"])
    with pytest.raises(ValueError, match="cannot be equal to or greater"):
        generate_batch_starting_indices(8, 8, 4, "train", [8], False)
    with pytest.raises(TypeError, match="'batch_size' must be a positive integer"):
        generate_batch_starting_indices(80, 8, 0, "train", [80], False)
    with pytest.raises(ValueError, match="No valid starting positions"):
        generate_batch_starting_indices(20, 8, 4, "train", [5, 5, 5, 5], False)


def test_indices_randomised_against_oracle():
    """600 random configurations (files, split, is_percents, sizes down to the error cases): the product feeder and the
    oracle restatement give the same indices, leave the torch CPU generator at the same position and raise the same
    errors. (tests/soak/index_soak_vs_reference.py runs the same comparison against the reference itself, offline.)"""
    import random
    from oracle import restatement as R
    from tradeaid_b200.feeder import generate_batch_starting_indices as ours
    rng = random.Random(5)
    seen = {"ok": 0, "err": 0}
    for trial in range(600):
        lens = [rng.randint(5, 400) for _ in range(rng.choice([1, 1, 2, 3, 5, 8]))]
        split = rng.choice(["train", "val"])
        frac = rng.uniform(0.05, 0.95)
        data_size = max(2, int(sum(lens) * (frac if split == "train" else 1 - frac)))
        T, B, pct, seed = rng.choice([2, 4, 8, 16, 64]), rng.choice([1, 3, 16]), rng.random() < 0.5, rng.randint(0, 10 ** 6)
        res = []
        for fn in (ours, R.batch_starting_indices):
            torch.manual_seed(seed)
            try:
                out = torch.as_tensor(fn(data_size, T, B, split, lens, pct)).tolist()
                res.append(("ok", out, int(torch.randint(0, 2 ** 31 - 1, (1,)))))
            except (ValueError, TypeError, RuntimeError) as e:
                res.append(("err", type(e).__name__, str(e)))
        assert res[0] == res[1], (trial, data_size, T, B, split, lens, pct, res)
        seen[res[0][0]] += 1
    assert seen["ok"] > 400 and seen["err"] > 10


def test_get_batch_bit_exact_through_module_globals():
    from tradeaid_b200 import training_utils as tu
    g = load_golden("get_batch.pt")
    full, n_val = g["full"], g["n_val"]
    tu.configure(device="cpu", batch_size=g["batch_size"], block_size=g["block_size"], eval_iters=2)
    tu._ops_override = emu_ops
    tu._stores.clear()
    tu.all_train_sets = [d[:-n_val].tolist() for d in full]     # python lists, as main.py provides them
    tu.all_val_sets = [d[-n_val:].clone() for d in full]
    tu.all_vocabularies = [[float(k) for k in range(50)]] * 3
    tu.all_modality_params = [["s", i, None, False, None, None, None, None, False, f"m{i}"] for i in range(3)]
    tu.file_lengths = g["file_lengths"]
    tu.num_modalities = 3
    tu.is_percents = False
    torch.manual_seed(g["seed"])
    try:
        for call in g["calls"]:
            xb, yb = tu.get_batch(call["split"], 1)
            for a, b in zip(xb, call["xb"]):
                assert a.dtype == torch.int64 and torch.equal(a, b)
            for a, b in zip(yb, call["yb"]):
                assert torch.equal(a, b)
    finally:
        tu._ops_override = None
        tu._stores.clear()


def test_get_batch_parity_mode_shards_concatenate_to_the_reference_batch():
    """Data-parallel parity mode: with batch_shard = (rank, world) every rank draws the SAME indices as the
    single-process reference (same RNG consumption) and gathers its slice of the rows; the slices concatenate to the
    reference's batch (the golden get_batch trace) for every world size that divides the batch."""
    from tradeaid_b200 import training_utils as tu
    g = load_golden("get_batch.pt")
    full, n_val, B = g["full"], g["n_val"], g["batch_size"]
    tu.configure(device="cpu", batch_size=B, block_size=g["block_size"], eval_iters=2)
    tu._ops_override = emu_ops
    tu.all_train_sets = [d[:-n_val].tolist() for d in full]
    tu.all_val_sets = [d[-n_val:].clone() for d in full]
    tu.all_vocabularies = [[float(k) for k in range(50)]] * 3
    tu.all_modality_params = [["s", i, None, False, None, None, None, None, False, f"m{i}"] for i in range(3)]
    tu.file_lengths = g["file_lengths"]
    tu.num_modalities = 3
    tu.is_percents = False
    try:
        worlds = [w for w in (1, 2, 4) if B % w == 0]
        assert len(worlds) >= 2
        for world in worlds:
            parts = []
            for rank in range(world):
                tu._stores.clear()
                tu.batch_shard = (rank, world)
                torch.manual_seed(g["seed"])
                parts.append([tu.get_batch(call["split"], 1) for call in g["calls"]])
                after = int(torch.randint(0, 2 ** 31 - 1, (1,)))
                if rank == 0:
                    first_after = after
                assert after == first_after                       # every rank leaves the host generator at the same place
            for c, call in enumerate(g["calls"]):
                for i in range(3):
                    assert torch.equal(torch.cat([parts[r][c][0][i] for r in range(world)]), call["xb"][i])
                    assert torch.equal(torch.cat([parts[r][c][1][i] for r in range(world)]), call["yb"][i])
        tu.batch_shard = (0, B + 1)
        with pytest.raises(ValueError, match="must be divisible"):
            tu.get_batch("train", 0)
    finally:
        tu.batch_shard = None
        tu._ops_override = None
        tu._stores.clear()


def test_directional_metrics_match_reference():
    from tradeaid_b200 import training_utils as tu
    g = load_golden("metrics.pt")
    with pytest.raises(RuntimeError):       # CPU tensors and no kernel namespace: the product refuses, it has no CPU path
        tu.calculate_evaluation_metrics(g["logits"], g["xb"], g["yb"], 3, g["vocabs"], g["params"], None)
    tu._ops_override = emu_ops              # host logic (eligibility rules, accumulators) over the CPU emulation of the kernel
    try:
        w, l, c, p = tu.calculate_evaluation_metrics(g["logits"], g["xb"], g["yb"], 3, g["vocabs"], g["params"], None)
    finally:
        tu._ops_override = None
    rw, rl, rc, rp = g["result"]
    assert w == rw and l == rl and p == rp
    for a, b in zip(c, rc):
        assert abs(a - b) <= 1e-5


def test_library_exports_every_declared_symbol():
    """The C ABI: libtradeaid_b200.so must load and export every function include/tradeaid_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "tradeaid_b200.h")).read()
    names = set(re.findall(r"\b(ta_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    path = os.path.join(ROOT, "trade-aid-multimodal-transformer_b200", "libtradeaid_b200.so")
    assert os.path.exists(path), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(path)
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    lib.ta_abi_version.restype = ctypes.c_int32
    m = re.search(r"#define TA_ABI_VERSION (\d+)", hdr)
    assert lib.ta_abi_version() == int(m.group(1))


def test_device_augmentation_host_logic_matches_reference():
    """TokenStore.augment_modality (word pre-draw from Python's `random`, eligible/accepted matching, generator
    re-positioning) over the CPU emulation of ta_augment_tokens, against the reference's own outputs."""
    import random
    from tradeaid_b200.feeder import TokenStore
    g = load_golden("augmentation.pt")
    for c in g["cases"]:
        for rounds, key_out, key_tail in ((1, "after1", "tail1"), (3, "after3", "tail3")):
            store = TokenStore([c["data"], c["data"]], "cpu", ops=emu_ops)
            random.seed(c["seed"])
            for _ in range(rounds):
                store.augment_modality(1, c["size"], c["V"])
            assert store.store[1].tolist() == c[key_out], (c["name"], rounds)
            assert store.store[0].tolist() == c["data"]
            assert [random.getrandbits(32) for _ in range(4)] == c[key_tail]
    store = TokenStore([[1, 2, 3]], "cpu", ops=emu_ops)
    for size, exc, msg in g["errors"]:
        with pytest.raises({"TypeError": TypeError, "ValueError": ValueError}[exc]) as ei:
            store.augment_modality(0, size, 10)
        assert str(ei.value) == msg
    store.augment_modality(0, None, 10)
    assert store.store[0].tolist() == [1, 2, 3]


def test_get_batch_with_the_has_header_quirk_matches_oracle():
    """get_batch as shipped (all_modality_params[r][2] = has_header = True -> rand_size 1, training_utils.py:353-360):
    three training calls against the oracle's in-place host walk + gather with the same `random` / torch seeds."""
    import random
    from oracle import restatement as R
    from tradeaid_b200 import training_utils as tu
    gen = torch.Generator().manual_seed(5)
    M, N, V, T, B = 2, 300, 40, 8, 6
    data = [torch.randint(0, V, (N,), generator=gen).tolist() for _ in range(M)]
    tu.configure(device="cpu", batch_size=B, block_size=T, eval_iters=1)
    tu._ops_override = emu_ops
    tu._stores.clear()
    try:
        tu.all_train_sets = [list(d) for d in data]
        tu.all_val_sets = [torch.tensor(d[:50]) for d in data]
        tu.all_vocabularies = [[float(k) for k in range(V)]] * M
        tu.all_modality_params = [["s", i, True if i == 0 else None, False, None, None, None, None, False, f"m{i}"] for i in range(M)]
        tu.file_lengths = [N]
        tu.num_modalities = M
        tu.is_percents = False
        ref = [list(d) for d in data]
        random.seed(3); torch.manual_seed(4)
        ours = [tu.get_batch("train", 1) for _ in range(3)]
        tail = random.getrandbits(32)
        random.seed(3); torch.manual_seed(4)
        for xb, yb in ours:
            ref[0] = R.add_rand_to_data_points(ref[0], True, V)
            ix = R.batch_starting_indices(N, T, B, "train", [N], False)
            rx, ry = R.gather_batch([torch.tensor(d) for d in ref], ix, T)
            assert all(torch.equal(a, b) for a, b in zip(xb, rx)) and all(torch.equal(a, b) for a, b in zip(yb, ry))
        assert tail == random.getrandbits(32)
        assert ref[0] != data[0]
    finally:
        tu._ops_override = None
        tu._stores.clear()

"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden.py
The reference never seeds anything — seeds are set here, around its calls.
Config is injected the way SURVEY.md §1 describes: config_utils._config_cache is set before model construction."""
import os
import random
import sys

import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
os.chdir("/tmp")  # the reference writes relative paths; keep /root/reference untouched

import config_utils  # noqa: E402


def set_cfg(block_size, n_embd, n_head, n_layer, batch_size, dropout=0.0):
    config_utils._config_cache = dict(device="cpu", batch_size=batch_size, block_size=block_size, eval_iters=2,
                                      n_embd=n_embd, n_head=n_head, n_layer=n_layer, dropout=dropout, fixed_values=[],
                                      output_file_name="", project_file_path="/tmp/")


def params_list(cross, names=None, is_pct=None):
    out = []
    for i, c in enumerate(cross):
        out.append(["synthetic", i + 1, None, bool(is_pct[i]) if is_pct else False, None, None, None, None, c,
                    names[i] if names else f"mod{i}", None, None])
    return out


def model_case(name, M, V, cross, B, T, C, H, L, T_run=None, steps=3, lr=1e-3):
    set_cfg(T, C, H, L, B)
    import model as ref_model
    torch.manual_seed(0)
    m = ref_model.MultimodalTransformer(M, V, params_list(cross))
    # make LayerNorm affine params and biases non-trivial so parity exercises them
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if "ln" in k or "norm" in k:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif k.endswith(".bias"):
                p.add_(0.02 * torch.randn(p.shape, generator=g))
    m.eval()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    Tr = T_run or T
    gi = torch.Generator().manual_seed(11)
    idx = [torch.randint(0, V[i], (B, Tr), generator=gi) for i in range(M)]
    tgt = [torch.randint(0, V[i], (B, Tr), generator=gi) for i in range(M)]
    logits, losses = m(idx, tgt)
    m.zero_grad()
    sum(losses).backward()
    grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in m.named_parameters()}
    logits_nt, none = m(idx)
    assert none is None
    # k AdamW steps on the same batch (main.py:464, 646-650)
    opt = torch.optim.AdamW(m.parameters(), lr=lr)
    curve = []
    for _ in range(steps):
        lg, ls = m(idx, tgt)
        tot = sum(ls)
        opt.zero_grad(set_to_none=True)
        tot.backward()
        opt.step()
        curve.append(float(tot))
    sd_after = {k: v.clone() for k, v in m.state_dict().items() if not k.endswith(".tril")}
    keys_with_tril = list(sd0.keys())
    sd0 = {k: v for k, v in sd0.items() if not k.endswith(".tril")}  # tril is derivable; keep fixtures small
    torch.save(dict(cfg=dict(M=M, V=V, cross=cross, B=B, T=T, C=C, H=H, L=L, T_run=Tr, lr=lr, steps=steps),
                    state_dict=sd0, state_dict_keys=keys_with_tril, idx=idx, tgt=tgt,
                    logits=[t.detach() for t in logits], losses=[t.detach() for t in losses], grads=grads,
                    curve=curve, state_dict_after=sd_after), os.path.join(OUT, f"model_{name}.pt"))
    print(name, "losses", [round(float(x), 5) for x in losses], "curve", [round(c, 5) for c in curve])


def index_cases():
    import training_utils as tu
    cases = []
    spec = [
        # data_size, block, batch, split, file_lengths, is_percents
        (900, 16, 32, "train", [1000], False),
        (100, 16, 32, "val", [1000], True),
        (900, 8, 64, "train", [300, 250, 450], False),
        (900, 8, 64, "train", [300, 250, 450], True),
        (100, 8, 64, "val", [300, 250, 450], False),
        (500, 8, 64, "val", [300, 250, 450], True),
        (620, 4, 33, "train", [5, 300, 3, 400, 200], False),
        (288, 4, 33, "val", [5, 300, 3, 400, 200], True),
        (90, 4, 4, "train", [100], False),
    ]
    for n, (ds, T, B, split, fl, pct) in enumerate(spec):
        torch.manual_seed(100 + n)
        ix = tu.generate_batch_starting_indices(ds, T, B, split, fl, pct)
        ix2 = tu.generate_batch_starting_indices(ds, T, B, split, fl, pct)  # second draw from the same stream
        cases.append(dict(args=(ds, T, B, split, fl, pct), seed=100 + n, ix=ix, ix2=ix2))
    torch.save(cases, os.path.join(OUT, "indices.pt"))
    print("indices:", len(cases), "cases")


def get_batch_case():
    """get_batch end to end through the reference's module globals (main.py:387-396), quirk off ([2] = None)."""
    import training_utils as tu
    set_cfg(8, 32, 4, 1, 6)
    M, N = 3, 400
    g = torch.Generator().manual_seed(5)
    full = [torch.randint(0, 50, (N,), generator=g) for _ in range(M)]
    n_val = 40
    tu.all_train_sets = [d[:-n_val].tolist() for d in full]
    tu.all_val_sets = [d[-n_val:].clone() for d in full]
    tu.all_vocabularies = [[float(k) for k in range(50)]] * M
    tu.all_modality_params = params_list([True, False, False])
    tu.file_lengths = [150, 250]
    tu.num_modalities = M
    tu.is_percents = False
    out = {}
    torch.manual_seed(42)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for split in ("train", "val", "train"):
            xb, yb = tu.get_batch(split, 1)
            out.setdefault("calls", []).append(dict(split=split, xb=xb, yb=yb))
    torch.save(dict(full=full, n_val=n_val, file_lengths=[150, 250], block_size=8, batch_size=6, seed=42, **out),
               os.path.join(OUT, "get_batch.pt"))
    print("get_batch: 3 calls")


def metrics_case():
    import training_utils as tu
    M, B, T = 3, 16, 6
    V = [21, 9, 5]
    g = torch.Generator().manual_seed(3)
    logits = [torch.randn(B, T, V[i], generator=g) for i in range(M)]
    xb = [torch.randint(0, V[i], (B, T), generator=g) for i in range(M)]
    yb = [torch.randint(0, V[i], (B, T), generator=g) for i in range(M)]
    vocabs = [[float(k) * 0.5 - 3 for k in range(V[0])], [float(k - 4) for k in range(V[1])], ["a", "b", "c", "d", "e"]]
    params = params_list([True, False, False], is_pct=[False, True, False])
    res = tu.calculate_evaluation_metrics(logits, xb, yb, M, vocabs, params, None)
    torch.save(dict(logits=logits, xb=xb, yb=yb, vocabs=vocabs, params=params, result=res),
               os.path.join(OUT, "metrics.pt"))
    print("metrics:", res)


def curve200_case():
    """North-star criterion "the 200-step loss curve matches within 1e-2 absolute": the UNMODIFIED reference trained for
    200 AdamW steps (lr 1e-3, main.py:646-650 loop body) through its own get_batch on SURVEY appendix-E streams; stores
    the summed loss of every step. Initial weights = oracle.restatement.random_state_dict(seed 0) with the second
    q/k/v matrices at a reference-like scale, loaded into the reference model (so the fixture needs no weights)."""
    sys.path.append(os.path.dirname(os.path.dirname(OUT)))      # repo root (for oracle/), after the reference on the path
    from oracle import restatement as R
    import importlib.util
    spec = importlib.util.spec_from_file_location("syn", os.path.join(os.path.dirname(os.path.dirname(OUT)),
                                                                     "trade-aid-multimodal-transformer_b200", "synthetic.py"))
    syn = importlib.util.module_from_spec(spec); spec.loader.exec_module(syn)
    import training_utils as tu
    V, cross, T, C, H, L, B, N, steps, lr = [64, 64], [True, False], 32, 128, 4, 2, 16, 4096, 200, 1e-3
    set_cfg(T, C, H, L, B)
    import model as ref_model
    m = ref_model.MultimodalTransformer(2, V, params_list(cross))
    sd = R.random_state_dict(R.Cfg(T, C, H, L, V, cross), seed=0)
    for k in list(sd):
        if ".heads." in k and k.endswith(".2.weight") and "sa_layers" in k:
            sd[k] = sd[k] / 10
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith(".tril") for k in missing), (missing, unexpected)
    data = [syn.token_stream("price", N, 64, 1000), syn.token_stream("indicator", N, 64, 1001)]
    tu.all_train_sets = [d.tolist() for d in data]
    tu.all_val_sets = [d[:64].clone() for d in data]
    tu.all_vocabularies = [[float(k) for k in range(64)]] * 2
    tu.all_modality_params = params_list(cross)
    tu.file_lengths = [N]
    tu.num_modalities = 2
    tu.is_percents = False
    opt = torch.optim.AdamW(m.parameters(), lr=lr)
    m.train()
    torch.manual_seed(2)
    curve = []
    for _ in range(steps):
        xb, yb = tu.get_batch("train", 1)
        _, losses = m(xb, yb)
        tot = sum(losses)
        opt.zero_grad(set_to_none=True)
        tot.backward()
        opt.step()
        curve.append(float(tot))
    torch.save(dict(cfg=dict(V=V, cross=cross, T=T, C=C, H=H, L=L, B=B, N=N, steps=steps, lr=lr, seed=2), curve=curve),
               os.path.join(OUT, "curve200.pt"))
    print("curve200:", [round(c, 4) for c in curve[::25]], round(curve[-1], 4))


def layer_activations_case():
    """Per-block activations of the reference (forward hooks on m.blocks[l]: the list of M residual streams each
    MultimodalBlock returns, model.py:214-244) for the weights and inputs of model_m3.pt / model_m3_short.pt — localises a
    parity failure to a layer."""
    import model as ref_model
    out = {}
    for name in ("m3", "m3_short"):
        g = torch.load(os.path.join(OUT, f"model_{name}.pt"))
        c = g["cfg"]
        set_cfg(c["T"], c["C"], c["H"], c["L"], c["B"])
        m = ref_model.MultimodalTransformer(c["M"], c["V"], params_list(c["cross"]))
        missing, unexpected = m.load_state_dict(g["state_dict"], strict=False)
        assert not unexpected and all(k.endswith(".tril") for k in missing)
        m.eval()
        acts = []
        hooks = [blk.register_forward_hook(lambda mod, inp, res: acts.append([t.detach().clone() for t in res])) for blk in m.blocks]
        with torch.no_grad():
            logits, _ = m(g["idx"])
        for h in hooks:
            h.remove()
        assert all((a - b).abs().max() == 0 for a, b in zip(logits, g["logits"]))     # same run as the stored logits
        out[name] = acts
    torch.save(out, os.path.join(OUT, "layer_activations.pt"))
    print("layer activations:", {k: (len(v), len(v[0]), tuple(v[0][0].shape)) for k, v in out.items()})


def augmentation_case():
    """data_utils.add_rand_to_data_points (the `[2]` quirk of get_batch, training_utils.py:353-360) on seeded lists:
    outputs after one and after three cumulative applications, and the next outputs of Python's `random` (pins how many
    Mersenne-Twister words the function consumed)."""
    from data_utils import add_rand_to_data_points
    cases = []
    for name, N, V, size, seed in (("s1", 1000, 50, 1, 11), ("s1_true", 777, 9, True, 12), ("s2", 500, 12, 2, 13),
                                   ("s3", 2000, 300, 3, 14), ("edge_band", 64, 3, 1, 15), ("tiny", 1, 10, 1, 16)):
        g = torch.Generator().manual_seed(seed)
        data = torch.randint(0, V, (N,), generator=g).tolist()
        random.seed(seed)
        work = list(data)
        after1 = list(add_rand_to_data_points(work, size, V))
        tail1 = [random.getrandbits(32) for _ in range(4)]
        random.seed(seed)
        work = list(data)
        for _ in range(3):
            work = add_rand_to_data_points(work, size, V)
        tail3 = [random.getrandbits(32) for _ in range(4)]
        cases.append(dict(name=name, data=data, V=V, size=size, seed=seed, after1=after1, tail1=tail1, after3=list(work),
                          tail3=tail3))
    errors = []
    for size in (False, 0, 4, 1.5):
        try:
            add_rand_to_data_points([1, 2, 3], size, 10)
            errors.append((size, None, None))
        except (TypeError, ValueError) as e:
            errors.append((size, type(e).__name__, str(e)))
    torch.save(dict(cases=cases, errors=errors), os.path.join(OUT, "augmentation.pt"))
    print("augmentation:", len(cases), "cases", errors)


if __name__ == "__main__":
    random.seed(0)
    model_case("demo", M=2, V=[57, 3], cross=[True, False], B=4, T=4, C=32, H=4, L=2)
    model_case("m3", M=3, V=[57, 3, 20], cross=[True, False, True], B=3, T=8, C=32, H=4, L=2)
    model_case("m3_short", M=3, V=[24, 10, 20], cross=[False, True, True], B=2, T=8, C=32, H=2, L=1, T_run=5)
    model_case("hs32", M=2, V=[40, 33], cross=[True, True], B=2, T=16, C=64, H=2, L=1)
    model_case("nocross", M=1, V=[30], cross=[True], B=2, T=8, C=32, H=4, L=1)
    index_cases()
    get_batch_case()
    metrics_case()
    augmentation_case()
    curve200_case()
    layer_activations_case()

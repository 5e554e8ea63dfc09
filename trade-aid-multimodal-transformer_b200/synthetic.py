"""Synthetic workloads for the configurations BASELINE.json names: shapes, vocabularies, token
streams and seeds. There is no network for real market data; bench.py and the parity tests use these."""
from __future__ import annotations

import torch

# name: (kinds per modality, cross flags, vocab, rows, n_embd, n_head, n_layer, block_size, batch per GPU)
CONFIGS = {
    "c1": dict(kinds=["price", "volume"], cross=[True, False], vocab=[57, 3], rows=100, n_embd=32, n_head=4, n_layer=2,
               block_size=4, batch=4, note="examples/demo_config.yaml shapes (reference CPU-runnable case)"),
    "c2": dict(kinds=["price", "volume"], cross=[True, False], vocab=[1024, 1024], rows=1_000_000, n_embd=256, n_head=8,
               n_layer=4, block_size=64, batch=64, note="2 modalities, 1M rows, 1xB200"),
    "c3": dict(kinds=["price", "volume", "time", "indicator"], cross=[True, False, False, True], vocab=[1024] * 4,
               rows=1_000_000, n_embd=512, n_head=8, n_layer=6, block_size=128, batch=64,
               note="4 modalities with selective cross-attention, 8xB200 DP"),
    "c4": dict(kinds=["price"] * 4 + ["volume"] * 2 + ["time", "indicator"], cross=[True] * 4 + [False] * 4,
               vocab=[1024] * 8, rows=50_000_000, n_embd=1024, n_head=16, n_layer=12, block_size=256, batch=64,
               note="8 modalities, 50M rows, vocab 1k/modality, 1/2/4/8xB200 DP"),
    "c5": dict(kinds=["price", "volume", "time", "indicator"], cross=[True, False, False, True], vocab=[1024] * 4,
               rows=1_000_000, n_embd=768, n_head=12, n_layer=6, block_size=1024, batch=256,
               note="long-context inference: forward/generate on 1xB200"),
}


def token_stream(kind, n, vocab, seed, device="cpu", price=None):
    """int64 [n] token ids in [0, vocab). price-like: reflected random walk; volume-like: uniform; time-like: a counter;
    indicator-like: an EMA of the price stream, re-quantised."""
    g = torch.Generator(device=device).manual_seed(seed)
    if kind == "price":
        steps = torch.randint(-3, 4, (n,), generator=g, device=device)
        walk = torch.cumsum(steps, 0) + vocab // 2
        period = 2 * (vocab - 1)
        w = torch.remainder(walk, period)
        return torch.where(w < vocab, w, period - w)
    if kind == "volume":
        return torch.randint(0, vocab, (n,), generator=g, device=device)
    if kind == "time":
        return torch.arange(n, device=device) % vocab
    if kind == "indicator":
        base = (price if price is not None else token_stream("price", n, vocab, seed + 7, device)).double()
        # EMA(alpha = 0.1) as a causal convolution truncated where 0.9^k < 1e-6
        k = torch.arange(132, device=device, dtype=torch.float64)
        w = 0.1 * 0.9 ** k
        padded = torch.cat([base[:1].expand(131), base])
        ema = torch.nn.functional.conv1d(padded.view(1, 1, -1), w.flip(0).view(1, 1, -1)).view(-1)
        return ema.round().long().clamp_(0, vocab - 1)
    raise ValueError(kind)


def make_streams(cfg, device="cpu", rows=None):
    n = rows or cfg["rows"]
    out, price = [], None
    for i, (kind, v) in enumerate(zip(cfg["kinds"], cfg["vocab"])):
        s = token_stream(kind, n, v, 1000 + i, device, price)
        if kind == "price" and price is None:
            price = s
        out.append(s)
    return out


def modality_params(cfg):
    """Legacy per-modality lists as schema.to_legacy_list lays them out (schema.py:213-248); index 2 = None keeps the
    reference's `[2]` augmentation quirk off on both sides, index 8 = cross_attention, 9 = name."""
    return [["synthetic", i + 1, None, False, None, None, None, None, c, f"{k}{i}", None, None]
            for i, (k, c) in enumerate(zip(cfg["kinds"], cfg["cross"]))]


def train_flops_per_step(cfg, batch=None):
    """Algorithmic FLOPs of forward+backward for one step (causal-half attention, 2 FLOPs per MAC,
    backward = 2x forward, recompute not credited)."""
    B = batch or cfg["batch"]
    T, C, H, L = cfg["block_size"], cfg["n_embd"], cfg["n_head"], cfg["n_layer"]
    M = len(cfg["vocab"])
    Xc = sum(cfg["cross"]) if M > 1 else 0
    K = M - 1
    BT = B * T
    A = 2 * B * T * T * C
    fwd = L * (M * ((21 + 3 / H) * BT * C * C + A) + Xc * ((4 + 4 * K) * BT * C * C + K * A))
    fwd += sum(2 * BT * (C * (v // 2) + (v // 2) * v) for v in cfg["vocab"])
    return 3 * fwd

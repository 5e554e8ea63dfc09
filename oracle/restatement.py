"""oracle/restatement.py — CPU restatement of the trade-AId hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module; the
product package (tradeaid_b200) never does. It restates, in plain fp32/fp64 PyTorch-on-CPU (the reference's own
language and numerics library), the algorithm of:

  /root/reference/model.py            (forward + loss; backward through torch autograd on this restatement)
  /root/reference/training_utils.py   (generate_batch_starting_indices, get_batch, calculate_evaluation_metrics)
  /root/reference/main.py:464,646-650 (the AdamW step: torch.optim.AdamW defaults)

Pinning: the reference's own tests hold NO golden vector for this path, so the restatement is pinned
against outputs of the reference itself: tests/golden/make_golden.py imports /root/reference (in the build
container), runs it on seeded inputs and stores inputs + outputs under tests/golden/; tests/test_oracle.py checks
this file against those fixtures (fp32: <= 1e-6 relative; indices: bit-exact).

The formulation is the *packed* one the CUDA engine uses (all heads of a projection in one matrix), so it doubles as
the layout contract: see `pack_*` below.
"""
from __future__ import annotations

import math
import numbers

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------- configuration
class Cfg:
    """Hyper-parameters the reference reads from config_utils at module construction (model.py:25-27)."""

    def __init__(self, block_size, n_embd, n_head, n_layer, vocab_sizes, cross, dropout=0.0):
        self.block_size = int(block_size)
        self.n_embd = int(n_embd)
        self.n_head = int(n_head)
        self.n_layer = int(n_layer)
        self.vocab_sizes = [int(v) for v in vocab_sizes]
        self.cross = [bool(c) for c in cross]  # all_modality_params[i][8], model.py:196
        self.dropout = float(dropout)
        self.M = len(self.vocab_sizes)
        self.hs = self.n_embd // self.n_head
        self.h2 = self.hs // 2

    def others(self, i):
        """KV source modalities of querier i, in the reference's order (model.py:198, 236)."""
        return [j for j in range(self.M) if j != i]

    def cross_enabled(self, i):
        # model.py:235-238: enabled flag AND at least one other modality
        return self.cross[i] and self.M > 1


# --------------------------------------------------------------------------------------------- packing
def pack_self_attention(sd, cfg, l, i, dtype=torch.float32):
    """Packed weights of blocks.{l}.sa_layers.{i} (model.py:30-92).

    W1 [3*H*h2, C], b1 [3*H*h2]: rows ordered (n in (query,key,value), head h, r)   <- heads.h.n.0.{weight,bias}
    W2 [3, H, hs, h2]                                                               <- heads.h.n.2.weight
    """
    p = f"blocks.{l}.sa_layers.{i}"
    names = ("query", "key", "value")
    W1 = torch.cat([sd[f"{p}.heads.{h}.{n}.0.weight"] for n in names for h in range(cfg.n_head)], 0).to(dtype)
    b1 = torch.cat([sd[f"{p}.heads.{h}.{n}.0.bias"] for n in names for h in range(cfg.n_head)], 0).to(dtype)
    W2 = torch.stack([torch.stack([sd[f"{p}.heads.{h}.{n}.2.weight"] for h in range(cfg.n_head)]) for n in names]).to(dtype)
    return W1, b1, W2


def pack_cross_attention(sd, cfg, l, i, dtype=torch.float32):
    """Packed weights of blocks.{l}.cross_attention_layers.{i} (model.py:95-159).

    Wq  [C, C]         rows ordered (head h, d)                 <- heads.h.query.weight [hs, C]
    Wkv [K, H, 2hs, C] per source position kv_i; rows [0:hs] = key, [hs:2hs] = value (split at model.py:144)
    """
    p = f"blocks.{l}.cross_attention_layers.{i}"
    H = cfg.n_head
    Wq = torch.cat([sd[f"{p}.heads.{h}.query.weight"] for h in range(H)], 0).to(dtype)
    Wkv = torch.stack([torch.stack([sd[f"{p}.heads.{h}.kv_projections.{s}.weight"] for h in range(H)])
                       for s in range(len(cfg.others(i)))]).to(dtype)
    return Wq, Wkv


# --------------------------------------------------------------------------------------------- forward
def _causal_attention(q, k, v, drop=None):
    """softmax(mask(q k^T * hs^-0.5)) v for [B, H, T, hs] tensors (model.py:65-72 and :147-154). `drop`: optional
    [B, H, T, T] multiplier (keep / (1 - p)) applied to the probabilities, i.e. nn.Dropout with a given mask
    (model.py:69, 151)."""
    T = q.shape[-2]
    aff = q @ k.transpose(-2, -1) * (k.shape[-1] ** -0.5)
    mask = torch.ones(T, T, dtype=torch.bool, device=q.device).tril()
    aff = aff.masked_fill(~mask, float("-inf"))
    wei = F.softmax(aff, dim=-1)
    if drop is not None:
        wei = wei * drop
    return wei @ v


def _ln(x, sd, key):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"].to(x.dtype), sd[key + ".bias"].to(x.dtype), 1e-5)


def _proj_mlp(o, sd, p, dtype):
    """Linear(C->C/2) -> tanh -> Linear(C/2->C) (model.py:82-86, 102-106)."""
    h = torch.tanh(o @ sd[f"{p}.proj.0.weight"].to(dtype).T + sd[f"{p}.proj.0.bias"].to(dtype))
    return h @ sd[f"{p}.proj.2.weight"].to(dtype).T + sd[f"{p}.proj.2.bias"].to(dtype)


def forward(sd, cfg: Cfg, idx_list, targets_list=None, dtype=torch.float32, return_hidden=False, masks=None):
    """Restates MultimodalTransformer.forward (model.py:380-402). Dropout: off (eval mode / p = 0) unless `masks` gives the
    multipliers keep / (1 - p) of the five nn.Dropout sites explicitly (the reference draws them from torch's RNG stream):
    ("sa_attn", l, i) and ("ca_attn", l, i, s): [B, H, T, T]; ("sa_proj", l, i), ("ffn", l, i), ("ca_proj", l, i): [B, T, C].

    sd: state dict with the reference's keys. Returns (logits_list, losses_list | None)."""
    M, C, H, hs, h2 = cfg.M, cfg.n_embd, cfg.n_head, cfg.hs, cfg.h2
    xs = []
    per_layer = []      # return_hidden="layers": the M residual streams after every block (model.py:214-244)
    for i in range(M):  # model.py:300-319
        idx = idx_list[i]
        B, T = idx.shape
        tok = sd[f"pre_block.token_embedding_tables.{i}.weight"].to(dtype)[idx]
        pos = sd["pre_block.position_embedding_table.weight"].to(dtype)[:T]
        xs.append(tok + pos)
    for l in range(cfg.n_layer):  # model.py:214-244
        att = []
        for i in range(M):
            x = xs[i]
            B, T, _ = x.shape
            W1, b1, W2 = pack_self_attention(sd, cfg, l, i, dtype)
            hid = torch.tanh(_ln(x, sd, f"blocks.{l}.ln1_layers.{i}") @ W1.T + b1).view(B, T, 3, H, h2)
            qkv = torch.einsum("btnhr,nhdr->nbhtd", hid, W2)  # [3, B, H, T, hs]
            mk = masks or {}
            o = _causal_attention(qkv[0], qkv[1], qkv[2], mk.get(("sa_attn", l, i))).transpose(1, 2).reshape(B, T, C)  # head-major concat
            pr = _proj_mlp(o, sd, f"blocks.{l}.sa_layers.{i}", dtype)
            x = x + (pr * mk[("sa_proj", l, i)] if ("sa_proj", l, i) in mk else pr)
            p = f"blocks.{l}.ffwd_layers.{i}.net"
            hdn = torch.relu(_ln(x, sd, f"blocks.{l}.ln2_layers.{i}") @ sd[f"{p}.0.weight"].to(dtype).T + sd[f"{p}.0.bias"].to(dtype))
            ff = hdn @ sd[f"{p}.2.weight"].to(dtype).T + sd[f"{p}.2.bias"].to(dtype)
            x = x + (ff * mk[("ffn", l, i)] if ("ffn", l, i) in mk else ff)
            att.append(x)
        new = []
        for i in range(M):
            x = att[i]
            if cfg.cross_enabled(i):
                B, T, _ = x.shape
                Wq, Wkv = pack_cross_attention(sd, cfg, l, i, dtype)
                q = (_ln(x, sd, f"blocks.{l}.ln_cross_layers.{i}") @ Wq.T).view(B, T, H, hs).transpose(1, 2)
                o = 0
                for s, j in enumerate(cfg.others(i)):
                    kv = torch.einsum("btc,hdc->bhtd", att[j], Wkv[s])  # un-normalised stream of source j
                    o = o + _causal_attention(q, kv[..., :hs], kv[..., hs:], (masks or {}).get(("ca_attn", l, i, s)))
                o = o.transpose(1, 2).reshape(B, T, C)
                pr = _proj_mlp(o, sd, f"blocks.{l}.cross_attention_layers.{i}", dtype)
                x = x + (pr * masks[("ca_proj", l, i)] if masks and ("ca_proj", l, i) in masks else pr)
            new.append(x)
        xs = new
        if return_hidden == "layers":
            per_layer.append(list(xs))
    logits_list = []
    for i in range(M):  # model.py:339-352
        p = f"post_block.soft_score_layers.{i}"
        h = torch.tanh(_ln(xs[i], sd, f"post_block.fin_norm_layers.{i}") @ sd[f"{p}.0.weight"].to(dtype).T + sd[f"{p}.0.bias"].to(dtype))
        logits_list.append(h @ sd[f"{p}.2.weight"].to(dtype).T + sd[f"{p}.2.bias"].to(dtype))
    if return_hidden == "layers":
        return logits_list, per_layer
    if return_hidden:
        return logits_list, xs
    if targets_list is None:
        return logits_list, None
    losses = []
    for i in range(M):  # model.py:394-399
        B, T, V = logits_list[i].shape
        losses.append(F.cross_entropy(logits_list[i].reshape(B * T, V), targets_list[i].reshape(B * T)))
    return logits_list, losses


def forward_backward(sd, cfg, idx_list, targets_list, dtype=torch.float32, masks=None):
    """Forward + autograd backward of sum(losses) (main.py:646-649). Returns (logits, losses, grads by key)."""
    params = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in sd.items() if not k.endswith(".tril")}
    logits, losses = forward(params, cfg, idx_list, targets_list, dtype, masks=masks)
    sum(losses).backward()
    # parameters the forward never touches keep grad None (e.g. cross-attention modules of a single-modality model,
    # model.py:235-238) and torch.optim.AdamW skips them entirely — no weight decay either.
    grads = {k: p.grad for k, p in params.items()}
    return [t.detach() for t in logits], [t.detach() for t in losses], grads


# --------------------------------------------------------------------------------------------- optimizer
def adamw_step(params, grads, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
    """One torch.optim.AdamW step with torch-2.11 defaults, in place (main.py:464).
    `step` is the 1-based step count AFTER increment."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    for k in params:
        p, g = params[k], grads[k]
        if g is None:
            continue
        p.mul_(1.0 - lr * weight_decay)
        exp_avg[k].lerp_(g, 1.0 - beta1)
        exp_avg_sq[k].mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        denom = (exp_avg_sq[k].sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(exp_avg[k], denom, value=-(lr / bc1))


# --------------------------------------------------------------------------------------------- feeder
def split_file_lengths(data_size, split, file_lengths):
    """Per-file lengths of the current split (training_utils.py:109-135)."""
    out, acc = [], 0
    n = len(file_lengths)
    for f in range(n):
        this = file_lengths[f] if split == "train" else file_lengths[n - 1 - f]
        acc += this
        if acc <= data_size:
            out.append(this)
        else:
            out.append(data_size - (acc - this))
        if acc >= data_size:
            if split == "val":
                out.reverse()
            break
    return out


def batch_starting_indices(data_size, block_size, batch_size, split, file_lengths, is_percents, generator=None):
    """Restates generate_batch_starting_indices (training_utils.py:33-181). Consumes the SAME torch CPU generator
    draws as the reference (one torch.randint of `batch_size` int64 values), so with equal seeds it is bit-exact."""
    if not isinstance(data_size, int) or data_size <= 0:
        raise TypeError("'data_size' must be a positive integer.")
    if not isinstance(block_size, int) or block_size <= 0:
        raise TypeError("'block_size' must be a positive integer.")
    if block_size >= data_size:
        raise ValueError("'block_size' cannot be equal to or greater than 'data_size'.")
    if not isinstance(batch_size, int) or batch_size <= 0:
        raise TypeError("'batch_size' must be a positive integer.")
    if not isinstance(split, str) or split not in ("train", "val"):
        raise ValueError("'split' must be 'train' or 'val'.")
    if not isinstance(file_lengths, list) or not len(file_lengths) >= 1:
        raise TypeError("'file_lengths' must be a list containing at least 1 element.")
    if not isinstance(is_percents, bool):
        raise TypeError("'is_percents' must be a boolean.")
    xy = block_size + 1
    off = 1 if is_percents else 0
    if len(file_lengths) == 1:
        return torch.randint(off, data_size - xy + 1, (batch_size,), generator=generator)
    lens = split_file_lengths(data_size, split, file_lengths)
    valid = torch.tensor([max(0, n - xy - off + 1) for n in lens], dtype=torch.long)
    total = int(valid.sum())
    if total <= 0:
        raise ValueError("No valid starting positions available for the given block size and file lengths.")
    r = torch.randint(total, (batch_size,), generator=generator)
    cum = torch.cumsum(valid, 0)
    k = torch.searchsorted(cum, r, right=True)
    starts = torch.cumsum(torch.tensor(lens, dtype=torch.long), 0) - torch.tensor(lens, dtype=torch.long)
    return starts[k] + (r - (cum[k] - valid[k])) + off


def gather_batch(data_list, ix, block_size):
    """x = d[i:i+T], y = d[i+1:i+T+1] for every modality with the same ix (training_utils.py:374-382)."""
    ar = torch.arange(block_size)
    xs = [d[ix[:, None] + ar[None, :]] for d in data_list]
    ys = [d[ix[:, None] + ar[None, :] + 1] for d in data_list]
    return xs, ys


# --------------------------------------------------------------------------------------------- augmentation
def add_rand_to_data_points(numeric_data, rand_size, vocab_size, rng=None):
    """Restates data_utils.add_rand_to_data_points (data_utils.py:293-358) for a list of ints: every element v with
    max_r < v < vocab_size - max_r (max_r = rand_size) gets `random.choice([0, 1, -1, ..., rand_size, -rand_size])`
    added, IN PLACE and in index order; elements outside that band consume no random number. `rng` is a
    random.Random (default: the module-level generator the reference uses). get_batch passes all_modality_params[r][2]
    as rand_size (training_utils.py:353), i.e. the has_header bool of the legacy list: True acts as 1, False fails the
    range check below exactly like the reference (bool is an int)."""
    import random as _random
    rng = rng or _random
    if not isinstance(numeric_data, list):
        raise TypeError("numeric_data must be a list or a tensor.")
    if not numeric_data:
        raise ValueError("numeric_data cannot be empty.")
    if not isinstance(rand_size, (int, type(None))):
        raise TypeError("rand_size must be an integer or null.")
    if rand_size is not None and (rand_size < 1 or rand_size > 3):
        raise ValueError("rand_size must be an integer between 1 and 3, or null.")
    if not isinstance(vocab_size, int) or vocab_size <= 0:
        raise TypeError("vocab_size must be a positive integer.")
    if rand_size is None:
        return numeric_data
    rand_list = [0]
    for r in range(rand_size):
        rand_list.extend([r + 1, -(r + 1)])
    hi = max(rand_list)
    for n in range(len(numeric_data)):
        if hi < numeric_data[n] < vocab_size - hi:
            numeric_data[n] += rng.choice(rand_list)
    return numeric_data


# --------------------------------------------------------------------------------------------- metrics
def _sign(cur, prev, is_pct):
    """_get_direction_sign (training_utils.py:184-212)."""
    if is_pct:
        return 1 if cur > 0 else (-1 if cur < 0 else 0)
    if not isinstance(prev, numbers.Number):
        return None
    ch = cur - prev
    return 1 if ch > 0 else (-1 if ch < 0 else 0)


def directional_metrics(logits_list, xb_list, yb_list, vocabularies, modality_params):
    """Restates calculate_evaluation_metrics (training_utils.py:215-330): wins, losses, certainty, processed."""
    M = len(logits_list)
    wins, losses, cert, proc = [0] * M, [0] * M, [0.0] * M, [0] * M
    for i in range(M):
        vocab = vocabularies[i]
        is_pct = modality_params[i][3]
        numeric = all(isinstance(v, numbers.Number) for v in vocab)
        min_len = 1 if is_pct else 2
        y = yb_list[i]
        if not (numeric and y.ndim >= 2 and y.shape[1] >= min_len and y.shape[0] > 0):
            continue
        proc[i] = 1
        last = logits_list[i][:, -1, :].float()
        probs = F.softmax(last, dim=-1)
        pred = last.argmax(-1)
        for j in range(last.shape[0]):
            prev = None if is_pct else vocab[int(xb_list[i][j, -1])]
            ps = _sign(vocab[int(pred[j])], prev, is_pct)
            as_ = _sign(vocab[int(y[j, -1])], prev, is_pct)
            if ps is None or as_ is None:
                continue
            if ps == as_:
                wins[i] += 1
            else:
                losses[i] += 1
            cert[i] += sum(float(probs[j, t]) for t, tv in enumerate(vocab) if _sign(tv, prev, is_pct) == ps)
    return wins, losses, cert, proc


# --------------------------------------------------------------------------------------------- helpers
def random_state_dict(cfg: Cfg, seed=0, dtype=torch.float32):
    """A state dict with the reference's key set and N(0, 0.02) init (model.py:372-378), *not* bit-equal to the
    reference's construction-order RNG stream — parity tests load reference-made state dicts instead."""
    g = torch.Generator().manual_seed(seed)
    C, H, hs, h2, T = cfg.n_embd, cfg.n_head, cfg.hs, cfg.h2, cfg.block_size
    sd = {}

    def w(*shape):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * 0.02).to(dtype)

    def ln(key):
        sd[key + ".weight"] = (1.0 + 0.1 * torch.randn(C, generator=g)).to(dtype)
        sd[key + ".bias"] = (0.1 * torch.randn(C, generator=g)).to(dtype)

    for i, V in enumerate(cfg.vocab_sizes):
        sd[f"pre_block.token_embedding_tables.{i}.weight"] = w(V, C)
    sd["pre_block.position_embedding_table.weight"] = w(T, C)
    for l in range(cfg.n_layer):
        for i in range(cfg.M):
            p = f"blocks.{l}.sa_layers.{i}"
            for h in range(H):
                for n in ("key", "query", "value"):
                    sd[f"{p}.heads.{h}.{n}.0.weight"] = w(h2, C)
                    sd[f"{p}.heads.{h}.{n}.0.bias"] = w(h2)
                    sd[f"{p}.heads.{h}.{n}.2.weight"] = w(hs, h2) * 10
            sd[f"{p}.proj.0.weight"], sd[f"{p}.proj.0.bias"] = w(C // 2, C), w(C // 2)
            sd[f"{p}.proj.2.weight"], sd[f"{p}.proj.2.bias"] = w(C, C // 2), w(C)
            p = f"blocks.{l}.ffwd_layers.{i}.net"
            sd[f"{p}.0.weight"], sd[f"{p}.0.bias"] = w(4 * C, C), w(4 * C)
            sd[f"{p}.2.weight"], sd[f"{p}.2.bias"] = w(C, 4 * C), w(C)
            ln(f"blocks.{l}.ln1_layers.{i}")
            ln(f"blocks.{l}.ln2_layers.{i}")
            if cfg.cross[i]:
                p = f"blocks.{l}.cross_attention_layers.{i}"
                for h in range(H):
                    sd[f"{p}.heads.{h}.query.weight"] = w(hs, C)
                    for s in range(cfg.M - 1):
                        sd[f"{p}.heads.{h}.kv_projections.{s}.weight"] = w(2 * hs, C)
                sd[f"{p}.proj.0.weight"], sd[f"{p}.proj.0.bias"] = w(C // 2, C), w(C // 2)
                sd[f"{p}.proj.2.weight"], sd[f"{p}.proj.2.bias"] = w(C, C // 2), w(C)
                ln(f"blocks.{l}.ln_cross_layers.{i}")
    for i, V in enumerate(cfg.vocab_sizes):
        ln(f"post_block.fin_norm_layers.{i}")
        p = f"post_block.soft_score_layers.{i}"
        sd[f"{p}.0.weight"], sd[f"{p}.0.bias"] = w(V // 2, C), w(V // 2)
        sd[f"{p}.2.weight"], sd[f"{p}.2.bias"] = w(V, V // 2), w(V)
    return sd

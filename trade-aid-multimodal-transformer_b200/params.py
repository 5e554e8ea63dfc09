"""Parameter layout of the engine: ONE flat fp32 master buffer whose slices are, at the same time,

  * the packed matrices the kernels consume (all heads of a projection in one matrix, all modalities and layers of a
    weight kind in one [L, M, ...] block — one grouped-GEMM launch per kind and layer), and
  * contiguous views with exactly the names and shapes of the reference's state dict (SURVEY.md A.3;
    /root/reference/model.py:30-370), so `state_dict()` / `load_state_dict()` / `parameters()` stay drop-in.

Packing follows SURVEY.md A.2 (verified equal to the reference in fp64): W1 rows ordered (q|k|v, head, r), W2 blocks
[3, H, hs, hs/2], cross Wq rows (head, d), cross Wkv stored BY SOURCE modality: for source j the 2*hs-row blocks of every
querier i != j are stacked so that the K/V projection of a_j for all its queriers is one GEMM.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field

import torch

ALIGN = 64  # elements; keeps every kind 256-byte aligned in fp32 and 128-byte aligned in the bf16 shadow


def r8(n):
    return (n + 7) // 8 * 8


@dataclass
class Dims:
    vocab_sizes: list
    cross: list            # all_modality_params[i][8] truthiness (model.py:196)
    block_size: int
    n_embd: int
    n_head: int
    n_layer: int
    dropout: float = 0.0
    M: int = field(init=False)

    def __post_init__(self):
        self.vocab_sizes = [int(v) for v in self.vocab_sizes]
        self.cross = [bool(c) for c in self.cross]
        self.M = len(self.vocab_sizes)
        assert len(self.cross) == self.M
        C, H = self.n_embd, self.n_head
        if C % H != 0:
            raise ValueError(f"n_embd ({C}) must be divisible by n_head ({H})")
        self.hs = C // H
        self.h2 = self.hs // 2
        self.h2p = r8(self.h2)
        if C % 16 != 0:
            raise ValueError("tradeaid_b200 requires n_embd % 16 == 0 (the projection MLP's hidden width n_embd/2 is a bf16 "
                             "row pitch: TMA needs 16 bytes)")
        if self.hs % 8 != 0 or self.hs > 64:
            raise ValueError(f"tradeaid_b200 supports head_size in {{8,16,...,64}}, got {self.hs}")
        # modalities whose cross-attention modules exist in the state dict (model.py:194-212)
        self.xc_decl = [i for i in range(self.M) if self.cross[i]]
        # modalities whose cross-attention actually runs (model.py:235-238: needs another modality)
        self.xc = self.xc_decl if self.M > 1 else []
        self.Xc = len(self.xc)
        # queriers of each source j, in increasing querier order
        self.queriers = [[i for i in self.xc if i != j] for j in range(self.M)]
        self.qoff = [0]
        for j in range(self.M):
            self.qoff.append(self.qoff[-1] + len(self.queriers[j]))
        self.voff = [0]
        for v in self.vocab_sizes:
            self.voff.append(self.voff[-1] + v)

    def others(self, i):
        return [j for j in range(self.M) if j != i]


class Layout:
    """kind name -> (offset, shape) in the flat buffer; reference parameter name -> (kind, index)."""

    def __init__(self, d: Dims):
        self.d = d
        self.kinds = OrderedDict()
        self.names = OrderedDict()   # reference name -> (kind, index tuple)
        self._off = 0
        L, M, C, H, hs, h2, T = d.n_layer, d.M, d.n_embd, d.n_head, d.hs, d.h2, d.block_size
        add = self._add
        add("tok", (d.voff[-1], C))
        add("pos", (T, C))
        for k in ("ln1_w", "ln1_b", "ln2_w", "ln2_b"):
            add(k, (L, M, C))
        add("sa_w1", (L, M, 3, H, h2, C))
        add("sa_b1", (L, M, 3, H, h2))
        add("sa_w2", (L, M, 3, H, hs, h2))
        add("sa_pw1", (L, M, C // 2, C)); add("sa_pb1", (L, M, C // 2))
        add("sa_pw2", (L, M, C, C // 2)); add("sa_pb2", (L, M, C))
        add("ff_w1", (L, M, 4 * C, C)); add("ff_b1", (L, M, 4 * C))
        add("ff_w2", (L, M, C, 4 * C)); add("ff_b2", (L, M, C))
        add("lnf_w", (M, C)); add("lnf_b", (M, C))
        for i, V in enumerate(d.vocab_sizes):
            add(f"hd_w1_{i}", (V // 2, C)); add(f"hd_b1_{i}", (V // 2,))
            add(f"hd_w2_{i}", (V, V // 2)); add(f"hd_b2_{i}", (V,))
        active_cross = d.Xc > 0
        if active_cross:
            self._add_cross(d.Xc, d.qoff[-1])
        self.n_active = self._off          # AdamW runs over [0, n_active)
        if d.xc_decl and not active_cross:
            # single-modality model with the cross flag set: modules exist, never run, never get a gradient
            self._add_cross(len(d.xc_decl), 0)
        self.total = self._off
        self._build_names()

    def _add(self, name, shape):
        n = 1
        for s in shape:
            n *= s
        self.kinds[name] = (self._off, tuple(shape))
        self._off = (self._off + n + ALIGN - 1) // ALIGN * ALIGN

    def _add_cross(self, nx, nkv):
        d = self.d
        L, C, H, hs = d.n_layer, d.n_embd, d.n_head, d.hs
        add = self._add
        add("lnc_w", (L, nx, C)); add("lnc_b", (L, nx, C))
        add("ca_wq", (L, nx, H, hs, C))
        add("ca_wkv", (L, nkv, H, 2 * hs, C))
        add("ca_pw1", (L, nx, C // 2, C)); add("ca_pb1", (L, nx, C // 2))
        add("ca_pw2", (L, nx, C, C // 2)); add("ca_pb2", (L, nx, C))

    def _build_names(self):
        """Reference state-dict names in the reference's own registration order (parameters only)."""
        d = self.d
        N = self.names
        for i in range(d.M):
            N[f"pre_block.token_embedding_tables.{i}.weight"] = ("tok", (slice(d.voff[i], d.voff[i + 1]),))
        N["pre_block.position_embedding_table.weight"] = ("pos", ())
        nidx = {"query": 0, "key": 1, "value": 2}
        xdecl = d.xc_decl
        for l in range(d.n_layer):
            b = f"blocks.{l}"
            for i in range(d.M):
                p = f"{b}.sa_layers.{i}"
                for h in range(d.n_head):
                    for n in ("key", "query", "value"):
                        N[f"{p}.heads.{h}.{n}.0.weight"] = ("sa_w1", (l, i, nidx[n], h))
                        N[f"{p}.heads.{h}.{n}.0.bias"] = ("sa_b1", (l, i, nidx[n], h))
                        N[f"{p}.heads.{h}.{n}.2.weight"] = ("sa_w2", (l, i, nidx[n], h))
                N[f"{p}.proj.0.weight"] = ("sa_pw1", (l, i)); N[f"{p}.proj.0.bias"] = ("sa_pb1", (l, i))
                N[f"{p}.proj.2.weight"] = ("sa_pw2", (l, i)); N[f"{p}.proj.2.bias"] = ("sa_pb2", (l, i))
            for i in range(d.M):
                p = f"{b}.ffwd_layers.{i}.net"
                N[f"{p}.0.weight"] = ("ff_w1", (l, i)); N[f"{p}.0.bias"] = ("ff_b1", (l, i))
                N[f"{p}.2.weight"] = ("ff_w2", (l, i)); N[f"{p}.2.bias"] = ("ff_b2", (l, i))
            for i in range(d.M):
                N[f"{b}.ln1_layers.{i}.weight"] = ("ln1_w", (l, i)); N[f"{b}.ln1_layers.{i}.bias"] = ("ln1_b", (l, i))
            for i in range(d.M):
                N[f"{b}.ln2_layers.{i}.weight"] = ("ln2_w", (l, i)); N[f"{b}.ln2_layers.{i}.bias"] = ("ln2_b", (l, i))
            for xi, i in enumerate(xdecl):
                p = f"{b}.cross_attention_layers.{i}"
                srcs = d.others(i)
                for h in range(d.n_head):
                    N[f"{p}.heads.{h}.query.weight"] = ("ca_wq", (l, xi, h))
                    for s, j in enumerate(srcs):
                        N[f"{p}.heads.{h}.kv_projections.{s}.weight"] = ("ca_wkv", (l, d.qoff[j] + d.queriers[j].index(i), h))
                N[f"{p}.proj.0.weight"] = ("ca_pw1", (l, xi)); N[f"{p}.proj.0.bias"] = ("ca_pb1", (l, xi))
                N[f"{p}.proj.2.weight"] = ("ca_pw2", (l, xi)); N[f"{p}.proj.2.bias"] = ("ca_pb2", (l, xi))
            for xi, i in enumerate(xdecl):
                N[f"{b}.ln_cross_layers.{i}.weight"] = ("lnc_w", (l, xi)); N[f"{b}.ln_cross_layers.{i}.bias"] = ("lnc_b", (l, xi))
        for i in range(d.M):
            N[f"post_block.fin_norm_layers.{i}.weight"] = ("lnf_w", (i,)); N[f"post_block.fin_norm_layers.{i}.bias"] = ("lnf_b", (i,))
        for i in range(d.M):
            p = f"post_block.soft_score_layers.{i}"
            N[f"{p}.0.weight"] = (f"hd_w1_{i}", ()); N[f"{p}.0.bias"] = (f"hd_b1_{i}", ())
            N[f"{p}.2.weight"] = (f"hd_w2_{i}", ()); N[f"{p}.2.bias"] = (f"hd_b2_{i}", ())

    # ------------------------------------------------------------------------------------------ views
    def view(self, flat, kind):
        off, shape = self.kinds[kind]
        n = 1
        for s in shape:
            n *= s
        return flat[off:off + n].view(shape)

    def views(self, flat):
        return {k: self.view(flat, k) for k in self.kinds}

    def named_views(self, flat):
        """OrderedDict reference-name -> contiguous view of `flat` with the reference's shape."""
        kv = self.views(flat)
        out = OrderedDict()
        for name, (kind, idx) in self.names.items():
            t = kv[kind][idx] if idx else kv[kind]
            assert t.is_contiguous(), name
            out[name] = t
        return out

    GLOBAL_HEAD_PREFIXES = ("lnf_", "hd_")     # output heads + final LayerNorm: first gradients backward produces
    GLOBAL_EMBED_KINDS = ("tok", "pos")        # embedding tables: the last

    def grad_ranges(self, stage, layers=()):
        """Element ranges [(start, stop), ...] of the flat gradient buffer (below n_active) written by a backward stage:
        'heads' (output heads + final LayerNorm), 'embed' (token + position tables) or 'layers' with `layers` = a run of
        consecutive layer indices — every per-layer kind is laid out [L, ...], so the run is one slice per kind."""
        out = []
        for kind, (off, shape) in self.kinds.items():
            if off >= self.n_active:
                continue
            n = 1
            for s_ in shape:
                n *= s_
            is_head = kind.startswith(self.GLOBAL_HEAD_PREFIXES)
            is_embed = kind in self.GLOBAL_EMBED_KINDS
            if stage == "heads" and is_head or stage == "embed" and is_embed:
                out.append((off, off + n))
            elif stage == "layers" and not is_head and not is_embed and n:
                per = n // self.d.n_layer
                lo, hi = min(layers), max(layers) + 1
                assert list(range(lo, hi)) == sorted(layers), "layers must be consecutive"
                out.append((off + lo * per, off + hi * per))
        out = [r for r in out if r[1] > r[0]]
        out.sort()
        merged = []
        for a, b in out:
            if merged and merged[-1][1] == a:
                merged[-1] = (merged[-1][0], b)
            else:
                merged.append((a, b))
        return merged

    def is_active(self, name):
        kind, _ = self.names[name]
        return self.kinds[kind][0] < self.n_active

    def tril_names(self):
        """Names of the persistent causal-mask buffers of the reference (model.py:55, 133), in state-dict order info:
        returns {module prefix: None}; the model synthesises them on state_dict() and drops them on load."""
        d = self.d
        out = []
        for l in range(d.n_layer):
            for i in range(d.M):
                out += [f"blocks.{l}.sa_layers.{i}.heads.{h}.tril" for h in range(d.n_head)]
            for i in d.xc_decl:
                out += [f"blocks.{l}.cross_attention_layers.{i}.heads.{h}.tril" for h in range(d.n_head)]
        return out

    def init_(self, flat, generator=None):
        """Reference initialisation statistics (model.py:372-378): N(0, 0.02) for Linear/Embedding weights, zero
        Linear biases, LayerNorm weight 1 / bias 0. (The reference's exact RNG stream depends on module construction
        order; bit-equal weights come from load_state_dict of a reference checkpoint.)"""
        flat.zero_()
        for k, t in self.views(flat).items():
            cls = kind_class(k)
            if cls == "ln_weight":
                t.fill_(1.0)
            elif cls == "weight":
                t.normal_(0.0, 0.02, generator=generator)
        return flat


def kind_class(kind):
    """'weight' (Linear / Embedding matrices), 'bias' (Linear and LayerNorm biases) or 'ln_weight'."""
    if kind.startswith("ln"):
        return "ln_weight" if kind.endswith("_w") else "bias"
    stem = kind.split("_")[1] if "_" in kind else kind
    if stem.startswith(("b", "pb")):
        return "bias"
    return "weight"

"""Drop-in replacement for /root/reference/model.py's public surface:

    MultimodalTransformer(num_modalities, vocab_sizes, all_modality_params)      (model.py:358)
        .forward(idx_list, targets_list=None) -> (logits_list, losses_list | None) (model.py:380-402)
        .generate(idx_list, max_new_tokens=1, modality_to_generate=0)             (model.py:404-446)
    an nn.Module whose parameters()/state_dict()/load_state_dict() carry the reference's names and shapes, including
    the persistent per-head `tril` buffers, so main.py (main.py:463-483, 635, 646-650) runs unchanged.

Hyper-parameters are ambient, as in the reference: read from the reference's `config_utils` getters at construction
(model.py:25-27) — or from `configure(...)` when used standalone.

Every parameter is a VIEW into the engine's flat fp32 master buffer; forward/backward are sequences of sm_100a kernel
launches (tradeaid_b200.engine). `sum(losses).backward()` works through a torch.autograd.Function whose backward runs the
engine's backward program and deposits `.grad` views on the parameters, so a stock torch.optim.AdamW can step them.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .params import Dims, Layout

_explicit_cfg = None
_precision = None      # None: the TRADEAID_PRECISION environment variable, default "bf16" (engine.Engine.PRECISIONS)


def set_precision(precision):
    """"bf16": single-pass bf16 GEMM operands (throughput mode). "bf16x3": fp32 activations and three-term split operands on
    the same kernels — logits within 2e-3 of the reference's fp32 forward (precision.py). Applies to models moved to the
    device afterwards."""
    global _precision
    from .engine import Engine
    if precision is not None and precision not in Engine.PRECISIONS:
        raise ValueError(f"precision must be one of {Engine.PRECISIONS}")
    _precision = precision


def configure(block_size, n_embd, n_head, n_layer, dropout=0.0):
    """Standalone alternative to the reference's config_utils cache."""
    global _explicit_cfg
    _explicit_cfg = dict(block_size=int(block_size), n_embd=int(n_embd), n_head=int(n_head), n_layer=int(n_layer),
                         dropout=float(dropout))


def _hyper():
    if _explicit_cfg is not None:
        return dict(_explicit_cfg)
    try:
        from config_utils import _get_block_size, _get_dropout, _get_n_embd, _get_n_head, _get_n_layer
    except ImportError as e:
        raise RuntimeError("no configuration: call tradeaid_b200.model.configure(...) or put the reference's "
                           "config_utils on sys.path") from e
    return dict(block_size=_get_block_size(), n_embd=_get_n_embd(), n_head=_get_n_head(), n_layer=_get_n_layer(),
                dropout=_get_dropout())


class _Node(nn.Module):
    """Empty container used to reproduce the reference's dotted parameter names."""


class _Bridge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, model, idx, tgt):
        logits, losses = model._engine.forward(idx, tgt, save=True, train=model.training)
        ctx.model = model
        ctx.gen = model._engine._fwd_gen       # the engine keeps ONE set of saved activations: stamp it
        return losses.clone()

    @staticmethod
    def backward(ctx, dlosses):
        model = ctx.model
        eng = model._engine
        if eng is None or eng._saved is None or eng._saved.get("gen") != ctx.gen:
            raise RuntimeError("tradeaid_b200: backward() of a forward whose saved activations are gone - the engine keeps the "
                               "activations of the LATEST grad-enabled forward only (another forward(idx, targets) with "
                               "gradients enabled ran in between); call backward() before the next training forward")
        ones = model._ones_like(dlosses)
        # sum(losses).backward() (main.py:646) hands in exact ones: compared on the device, no read-back on that path
        w = None if (dlosses is ones or bool(torch.equal(dlosses, ones))) else dlosses.tolist()
        # torch accumulates into an existing .grad. The .grad tensors ARE views of the engine's gradient buffer, which the
        # backward program overwrites, so gradients still held by the parameters (a second backward without zero_grad: gradient
        # accumulation) are set aside first and added back. main.py calls optimizer.zero_grad(set_to_none=True) before every
        # backward (main.py:648): nothing is held then and nothing is copied.
        held = model._held_grads()
        eng.backward(w)
        for view, old in held:
            view.add_(old)
        model._deposit_grads()
        return None, None, None, None


_ops_override = None   # tests inject a CPU emulation of the kernel namespace (tests/emu_ops.py); the product leaves this None


class MultimodalTransformer(nn.Module):
    def __init__(self, num_modalities, vocab_sizes, all_modality_params):
        super().__init__()
        self.num_modalities = num_modalities
        self.vocab_sizes = vocab_sizes
        self.all_modality_params = all_modality_params
        hp = _hyper()
        cross = [bool(all_modality_params[i][8]) for i in range(num_modalities)]
        self._dims = Dims(list(vocab_sizes)[:num_modalities], cross, hp["block_size"], hp["n_embd"], hp["n_head"],
                          hp["n_layer"], hp["dropout"])
        self._layout = Layout(self._dims)
        flat = torch.zeros(self._layout.total, dtype=torch.float32)
        self._layout.init_(flat)
        self._flat = flat
        self._engine = None
        self._anchor = None
        self._param_list = []
        for name, view in self._layout.named_views(flat).items():
            node = self
            parts = name.split(".")
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, _Node())
                node = node._modules[p]
            par = nn.Parameter(view)
            node.register_parameter(parts[-1], par)
            self._param_list.append((name, par))
        self._tril_names = self._layout.tril_names()
        self._register_state_dict_hook(self._add_tril)
        self._register_load_state_dict_pre_hook(self._drop_tril)

    # ------------------------------------------------------------------------------------------ state dict
    def _add_tril(self, module, state_dict, prefix, local_metadata):
        """The reference keeps a persistent [T, T] lower-triangular buffer per head (model.py:55, 133). The engine never
        needs one; they are synthesised here, at the reference's positions, for checkpoint compatibility."""
        T = self._dims.block_size
        dev = next(iter(state_dict.values())).device if state_dict else "cpu"
        tril = torch.tril(torch.ones(T, T, device=dev))
        out = OrderedDict()
        done = set()
        for k, v in state_dict.items():
            # `tril` is registered before the head's sub-modules: emit it when the first key of a head shows up
            if ".heads." in k:
                head = k[len(prefix):].split(".heads.")[0] + ".heads." + k[len(prefix):].split(".heads.")[1].split(".")[0]
                if head not in done:
                    done.add(head)
                    out[prefix + head + ".tril"] = tril
            out[k] = v
        state_dict.clear()
        state_dict.update(out)
        return state_dict

    def _drop_tril(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        for k in [k for k in state_dict if k.startswith(prefix) and k.endswith(".tril")]:
            del state_dict[k]
        self._dirty = True

    # ------------------------------------------------------------------------------------------ device placement
    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self._reflatten()
        return self

    def _reflatten(self):
        """nn.Module.to()/cuda() moves every parameter separately (main.py:463); gather them back into one flat
        buffer on the new device and re-point each parameter at its view."""
        dev = self._param_list[0][1].device
        dtype = self._param_list[0][1].dtype
        if dtype != torch.float32:
            raise TypeError("tradeaid_b200 keeps fp32 master parameters (bf16 operand copies are internal)")
        if dev.type == "cuda":
            from .engine import Engine
            eng = Engine(self._dims, dev, precision=_precision)
            flat = eng.master
        elif _ops_override is not None:      # tests only: the host program over the CPU emulation of the kernel namespace
            from .engine import Engine
            eng = Engine(self._dims, dev, ops=_ops_override, act_dtype=None if _precision == "bf16x3" else torch.float32,
                         precision=_precision)
            flat = eng.master
        else:
            eng, flat = None, torch.zeros(self._layout.total, dtype=torch.float32, device=dev)
        views = self._layout.named_views(flat)
        with torch.no_grad():
            for name, par in self._param_list:
                views[name].copy_(par.data)
                par.data = views[name]
                par.grad = None
        self._flat, self._engine = flat, eng
        self._anchor = torch.zeros(1, device=dev, requires_grad=True) if eng is not None else None
        self._dirty = True

    def _ones_like(self, t):
        o = getattr(self, "_ones", None)
        if o is None or o.shape != t.shape or o.device != t.device or o.dtype != t.dtype:
            o = self._ones = torch.ones_like(t)
        return o

    def _held_grads(self):
        """[(gradient-buffer view, copy of its current content)] for the parameters whose .grad is such a view."""
        if all(par.grad is None for _, par in self._param_list):       # main.py's loop: nothing to look at
            return []
        gviews = self._grad_views
        names = [name for name, par in self._param_list
                 if par.grad is not None and self._layout.is_active(name) and par.grad.data_ptr() == gviews[name].data_ptr()]
        if not names:
            return []
        if len(names) == sum(1 for name, _ in self._param_list if self._layout.is_active(name)):
            whole = self._engine.grad[:self._layout.n_active]          # every parameter holds one: one copy of the flat buffer
            return [(whole, whole.clone())]
        return [(gviews[name], gviews[name].clone()) for name in names]

    def _deposit_grads(self):
        self._dirty = True               # gradients were produced: the parameters are about to change
        gviews = self._grad_views
        for name, par in self._param_list:
            if not self._layout.is_active(name):
                continue  # never used by forward: the reference leaves .grad = None and AdamW skips it
            g = gviews[name]
            if par.grad is None:
                par.grad = g
            elif par.grad.data_ptr() != g.data_ptr():
                par.grad.add_(g)

    # ------------------------------------------------------------------------------------------ forward
    def _require_engine(self):
        if self._engine is None:
            raise RuntimeError("tradeaid_b200.MultimodalTransformer runs on a CUDA device only (sm_100a kernels, no CPU "
                               "fallback): move the model with .to('cuda') first")
        eng = self._engine
        if self._param_list[0][1].data_ptr() != eng.master.data_ptr():
            self._reflatten()
            eng = self._engine
        if not hasattr(self, "_grad_views") or self._grad_views_of is not eng:
            self._grad_views = eng.layout.named_views(eng.grad)
            self._grad_views_of = eng
        return eng

    def forward(self, idx_list, targets_list=None):
        eng = self._require_engine()
        M = self.num_modalities
        idx = torch.stack([t.to(eng.device, torch.int64) for t in idx_list[:M]]).contiguous()
        self._sync_shadows(eng)
        B, T = idx.shape[1], idx.shape[2]
        if targets_list is None:
            logits, _ = eng.forward(idx, None, save=False, train=self.training)
            return [l.clone() for l in logits], None
        tgt = torch.stack([t.to(eng.device, torch.int64) for t in targets_list[:M]]).contiguous()
        if torch.is_grad_enabled():
            losses = _Bridge.apply(self._anchor, self, idx, tgt)
            logits = [l.view(B, T, -1).clone() for l in eng._saved["logits"]]
        else:
            logits, losses = eng.forward(idx, tgt, save=False, train=self.training)
            logits, losses = [l.clone() for l in logits], losses.clone()
        return logits, [losses[i] for i in range(M)]

    @torch.no_grad()
    def generate(self, idx_list, max_new_tokens=1, modality_to_generate=0, use_cache=True):
        """Same sequence bookkeeping and sampling as the reference (model.py:404-446): crop to block_size, logits of the
        last position of the target modality, softmax + torch.multinomial, append; other modalities are padded with their
        last token or cropped. The reference runs the full forward over the window for every new token. Here, while the
        sequences still fit block_size (no cropping: earlier positions keep their activations), the first step fills a
        K/V cache and every later step computes only the new position against it (engine.KVCache, ta_attention_decode);
        once the window starts to slide, positions shift and the full forward is used, as in the reference.
        use_cache=False forces the reference's recompute-everything form."""
        T = self._dims.block_size
        mg = modality_to_generate
        seqs = [idx.clone() for idx in idx_list]
        eng = self._require_engine()
        use_cache = use_cache and not (self.training and self._dims.dropout > 0.0)   # dropout draws need the full pass
        cache = None
        for _ in range(max_new_tokens):
            n = seqs[mg].shape[1]
            M = self.num_modalities
            fits = use_cache and n <= T and all(s.shape[1] == n and s.shape[0] == seqs[mg].shape[0] for s in seqs[:M])
            if fits and (cache is None or cache.length == n - 1):
                from .engine import KVCache
                self._sync_shadows(eng)
                if cache is None:
                    cache = KVCache(eng, seqs[mg].shape[0], T)
                    new = torch.stack([s.to(eng.device, torch.int64) for s in seqs[:M]]).contiguous()
                    logits, _ = eng.forward(new, None, save=False, train=False, cache=cache)
                else:
                    new = torch.stack([s[:, -1:].to(eng.device, torch.int64) for s in seqs[:M]]).contiguous()
                    logits = eng.decode_step(new, cache)       # replayed as a CUDA graph from the second decode step on
                last = logits[mg][:, -1, :]
            else:
                cache = None
                logits_list, _ = self([s[:, -T:] for s in seqs])
                last = logits_list[mg][:, -1, :]
            probs = F.softmax(last, dim=-1)
            nxt = torch.multinomial(probs, num_samples=1)
            seqs[mg] = torch.cat((seqs[mg], nxt.to(seqs[mg].device)), dim=1)
            tl = seqs[mg].shape[1]
            for i in range(self.num_modalities):
                if i == mg:
                    continue
                if seqs[i].shape[1] < tl:
                    seqs[i] = torch.cat((seqs[i], seqs[i][:, -1:]), dim=1)
                elif seqs[i].shape[1] > tl:
                    seqs[i] = seqs[i][:, :tl]
        return seqs

    def _sync_shadows(self, eng):
        if self.training or getattr(self, "_dirty", True):
            eng.shadows_stale = True     # a stock optimizer updates the fp32 views in place between forwards
        # After a training-mode forward an optimizer step may follow, so the NEXT forward must refresh the operand copies
        # even in eval mode (estimate_loss right after optimizer.step(), main.py:603/650); consecutive eval-mode forwards
        # (eval_iters of them) refresh once. _deposit_grads marks the same for a backward taken in eval mode.
        self._dirty = bool(self.training)

"""The forward/backward program of the trade-AId multimodal transformer as a sequence of C-ABI kernel launches.

Restates /root/reference/model.py:380-402 (forward + loss) and its autograd backward over packed buffers:
every modality of a layer goes through ONE grouped launch per op (the reference loops over modalities, heads and
key/value sources in Python: model.py:90, 222-242), activations live in bf16 between GEMMs, the residual stream, the
LayerNorm statistics, softmax, tanh and the loss stay in fp32.

`ops` is the kernel namespace — `tradeaid_b200.lib` (ctypes over libtradeaid_b200.so) in the product. Tests inject a
CPU emulation of the same ops to check this host logic without a GPU; the engine itself never falls back.
"""
from __future__ import annotations

import os

import torch

from .params import Dims, Layout, r8

ACT_NONE, ACT_TANH, ACT_RELU = 0, 1, 2
AUX_TANH_BWD, AUX_RELU_BWD = 1, 2


def _runs(idxs):
    """Split a sorted index list into maximal arithmetic progressions -> list of (start, step, count)."""
    out, k = [], 0
    while k < len(idxs):
        if k + 1 == len(idxs):
            out.append((idxs[k], 1, 1))
            break
        step = idxs[k + 1] - idxs[k]
        n = 2
        while k + n < len(idxs) and idxs[k + n] - idxs[k + n - 1] == step:
            n += 1
        out.append((idxs[k], step, n))
        k += n
    return out


class KVCache:
    """Keys / values of every layer at positions [0, length), for `generate` in the growing phase (model.py:404-446 re-runs
    the whole window for every new token; while the sequence still fits block_size the earlier positions are unchanged —
    causal masks, absolute position embeddings — so only the new position is computed against this cache).
    self_kv[l]: [M, B, capacity, 2C] (K | V of the modality's own stream); cross_kv[l][r]: [n_r, B, capacity, nq_r*2C] per
    run r of source modalities (engine._kv_runs), laid out like the K/V projection buffers of the forward."""

    def __init__(self, eng, B, capacity=None):
        d = eng.d
        self.B = int(B)
        self.capacity = int(capacity or d.block_size)
        if self.capacity > d.block_size:
            raise ValueError("cache capacity exceeds block_size: beyond it the window slides and every position changes")
        self.length = 0
        # the same count on the device: the decode step replayed as a CUDA graph reads it there (Engine.decode_step)
        self.len_dev = torch.zeros(1, dtype=torch.int32, device=eng.device)
        self.dyn = False                 # True while a decode step runs with the device-side length
        self._graph = None
        self._static_idx = None
        self._static_logits = None
        self._dyn_steps = 0
        C, dt, dev = d.n_embd, eng.act_dtype, eng.device
        self.self_kv = [torch.zeros(d.M, self.B, self.capacity, 2 * C, dtype=dt, device=dev) for _ in range(d.n_layer)]
        self.cross_kv = [[torch.zeros(j1 - j0, self.B, self.capacity, nq * 2 * C, dtype=dt, device=dev)
                          for j0, j1, nq in eng._kv_runs()] if d.Xc else [] for _ in range(d.n_layer)]

    def bytes(self):
        return sum(t.numel() * t.element_size() for t in self.self_kv) + \
            sum(t.numel() * t.element_size() for r in self.cross_kv for t in r)


class Engine:
    PRECISIONS = ("bf16", "bf16x3")

    def __init__(self, dims: Dims, device, ops=None, act_dtype=None, precision=None, recompute=None):
        self.d = dims
        # activation recompute (TRADEAID_RECOMPUTE=1): forward(save=True) keeps one fp32 stream copy per block instead of the
        # block's activations, backward() reruns each block's forward first (+1 forward of compute, ~1/10 of the activations)
        self.recompute = bool(int(os.environ.get("TRADEAID_RECOMPUTE", "0"))) if recompute is None else bool(recompute)
        self.layout = Layout(dims)
        self.device = torch.device(device)
        if ops is None:
            if self.device.type != "cuda":
                raise RuntimeError("tradeaid_b200 runs on CUDA (sm_100a) only; there is no CPU path")
            from . import lib as ops  # raises ImportError if libtradeaid_b200.so is missing
        # precision: "bf16" = one bf16 rounding per GEMM operand (throughput mode, ~5e-3 normwise on the logits, like
        # torch.autocast on the reference); "bf16x3" = fp32 activations and three-term split GEMM operands on the same
        # kernels (precision.py: the north star's 2e-3 logits bound against the fp32 reference, 3x the GEMM FLOPs)
        if precision is None:
            precision = os.environ.get("TRADEAID_PRECISION", "bf16")
        if precision not in self.PRECISIONS:
            raise ValueError(f"precision must be one of {self.PRECISIONS}, got {precision!r}")
        self.precision = precision
        if precision == "bf16x3":
            from .precision import SplitPrecisionOps
            ops, act_dtype = SplitPrecisionOps(ops), torch.float32
        # storage type of GEMM operands / saved activations. bf16 in the throughput mode (the CUDA GEMM / attention
        # kernels accept nothing else); fp32 in the bf16x3 mode; the CPU test-suite also runs the host logic with float32
        # over the plain emulation to check the backward wiring exactly.
        self.act_dtype = act_dtype if act_dtype is not None else torch.bfloat16
        act_dtype = self.act_dtype
        self.ops = ops
        n = self.layout.total
        self.master = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.shadow = torch.zeros(n, dtype=act_dtype, device=self.device)
        self.P = self.layout.views(self.master)
        self.G = self.layout.views(self.grad)
        self.S = self.layout.views(self.shadow)
        self._padded = {}      # kind -> zero-padded bf16 (or fp32 for biases) copy when the direct view is not TMA-able
        self._make_padded()
        # q/k/v stage 2 (model.py:36-50: Linear(hs/2 -> hs) per head and projection) is block diagonal: 3H blocks of
        # [hs x hs/2]. As 3H grouped problems with K = hs/2 the GEMM is bound by per-tile latency (c4: 49 152 tiles of
        # 128 x 64 x 32 per launch, 2.4x the HBM time). Packing `gsz` neighbouring blocks on the diagonal of one
        # [gsz*hs x gsz*hs/2] matrix (zeros elsewhere) turns it into 3H/gsz problems with full 256-wide tiles; the
        # zeros add exact zeros to the fp32 accumulators, so the results are bit-identical.
        self._bd = None
        gsz = 256 // dims.hs if dims.hs > 0 else 0
        if (os.environ.get("TRADEAID_STAGE2_BLOCKDIAG", "1") != "0" and dims.h2p == dims.h2 and gsz > 1
                and (3 * dims.n_head) % gsz == 0):
            ng = 3 * dims.n_head // gsz
            self._bd = (gsz, ng, torch.zeros(dims.n_layer, dims.M, ng, gsz * dims.hs, gsz * dims.h2, dtype=act_dtype,
                                             device=self.device))
        if hasattr(ops, "register_weight_storage"):      # bf16x3: splits of weight operands are kept between launches
            for t in [self.shadow, *self._padded.values()] + ([self._bd[2]] if self._bd is not None else []):
                ops.register_weight_storage(t)
        self._bufs = {}
        # output heads (model.py:339-352): with one vocabulary size for every modality (and TMA-able widths) the per-modality
        # weight kinds hd_w1_i, hd_b1_i, hd_w2_i, hd_b2_i sit a constant stride apart in the flat buffer, so all
        # modalities go through one grouped launch per GEMM instead of M launches of a few dozen tiles each
        self._heads_grouped = None
        Vs = dims.vocab_sizes
        if (os.environ.get("TRADEAID_GROUPED_HEADS", "1") != "0" and dims.M > 1 and len(set(Vs)) == 1
                and (Vs[0] // 2) % 8 == 0 and Vs[0] // 2 > 0):
            off = [self.layout.kinds[f"hd_w1_{i}"][0] for i in range(dims.M)]
            stride = off[1] - off[0]
            if all(off[i + 1] - off[i] == stride for i in range(dims.M - 1)):
                V, V2, C, M = Vs[0], Vs[0] // 2, dims.n_embd, dims.M
                o = {k: self.layout.kinds[f"hd_{k}_0"][0] for k in ("w1", "b1", "w2", "b2")}
                shapes = {"w1": ((M, V2, C), (stride, C, 1)), "b1": ((M, V2), (stride, 1)),
                          "w2": ((M, V, V2), (stride, V2, 1)), "b2": ((M, V), (stride, 1))}
                view = lambda flat, k: flat.as_strided(shapes[k][0], shapes[k][1], o[k])  # noqa: E731
                self._heads_grouped = {"P": {k: view(self.master, k) for k in shapes}, "G": {k: view(self.grad, k) for k in shapes},
                                       "S": {k: view(self.shadow, k) for k in ("w1", "w2")}}
        self._voff = torch.tensor(dims.voff, dtype=torch.int64, device=self.device)
        self._saved = None
        self.shadows_stale = True
        # dropout (model.py:69, 91, 116, 151, 171): masks are a function of (seed, coordinates); `seed` is fixed per
        # (site, layer), a device step counter mixed in by the kernels makes every step (and graph replay) different
        self.dropout_base_seed = 0x5DEECE66D
        self._drop_step = None
        self._drop_step_host = 0      # host mirror of the device counter (value the NEXT launches will read)
        self._fwd_gen = 0             # generation stamp of the saved activations (model._Bridge checks it in backward)
        # where backward reports ("part", l) to its stage hook (trainer.py: the SM-margin window of the data-parallel
        # exchange): after the whole cross-attention backward of the layer ("cross") or after its attention kernels ("attn")
        self._dp_window = os.environ.get("TRADEAID_DP_WINDOW", "cross")

    # ------------------------------------------------------------------------------------------ weights
    def _make_padded(self):
        d = self.d
        dev = self.device
        L, M, H = d.n_layer, d.M, d.n_head
        if d.h2p != d.h2:
            self._padded["sa_w1"] = torch.zeros(L, M, 3, H, d.h2p, d.n_embd, dtype=self.act_dtype, device=dev)
            self._padded["sa_b1"] = torch.zeros(L, M, 3, H, d.h2p, dtype=torch.float32, device=dev)
            self._padded["sa_w2"] = torch.zeros(L, M, 3, H, d.hs, d.h2p, dtype=self.act_dtype, device=dev)
        for k, (off, shape) in self.layout.kinds.items():
            if k in self._padded or len(shape) < 2 or k in ("tok", "pos") or k.startswith("ln"):
                continue
            if k.split("_")[1].startswith(("b", "pb")):
                continue
            if shape[-1] % 8 != 0 and shape[-1] > 0:
                self._padded[k] = torch.zeros(*shape[:-1], r8(shape[-1]), dtype=self.act_dtype, device=dev)

    def refresh_shadows(self, flat_cast=True):
        """bf16 GEMM-operand copies of the fp32 master weights. `flat_cast=False` when the fused AdamW has already
        written the flat shadow and only the zero-padded copies (odd shapes) are left to do."""
        ops, d = self.ops, self.d
        if flat_cast:
            ops.cast_bf16(self.master, self.shadow)
        for k, dst in self._padded.items():
            src = self.P[k]
            if k == "sa_w1":
                ops.cast_pad(src.reshape(-1, d.h2, d.n_embd), dst.view(-1, d.h2p, d.n_embd))
            elif k == "sa_b1":
                ops.unpad(src.reshape(-1, 1, d.h2), dst.view(-1, 1, d.h2p)[..., :d.h2])
            else:
                ops.cast_pad(src.reshape(-1, src.shape[-2], src.shape[-1]), dst.view(-1, dst.shape[-2], dst.shape[-1]))
        if self._bd is not None:
            gsz, ng, w2bd = self._bd
            hs, h2 = d.hs, d.h2
            n = d.n_layer * d.M * ng
            src = self.P["sa_w2"].reshape(n, gsz, hs, h2)
            R, Cc = gsz * hs, gsz * h2
            for k in range(gsz):     # block k of every group -> rows [k*hs, (k+1)*hs), columns [k*h2, (k+1)*h2)
                dst = w2bd.as_strided((n, hs, h2), (R * Cc, Cc, 1), w2bd.storage_offset() + k * (hs * Cc + h2))
                ops.cast_pad(src[:, k], dst)
        if hasattr(ops, "refresh_weight_splits"):
            ops.refresh_weight_splits()
        self.shadows_stale = False

    def W(self, kind):
        """bf16 operand view of a weight kind (logical shape, possibly with a padded pitch)."""
        if kind in self._padded and kind not in ("sa_b1",):
            src_shape = self.layout.kinds[kind][1]
            t = self._padded[kind]
            if kind in ("sa_w1",):
                return t
            return t[..., :src_shape[-1]]
        return self.S[kind]

    # ------------------------------------------------------------------------------------------ buffers
    def buf(self, name, shape, dtype):
        key = name
        t = self._bufs.get(key)
        shape = tuple(int(s) for s in shape)
        if t is None or t.shape != shape or t.dtype != dtype:
            t = torch.zeros(shape, dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t

    def release_buffers(self):
        self._bufs.clear()
        self._saved = None

    # ------------------------------------------------------------------------------------------ dropout
    SITE_SA_ATTN, SITE_SA_PROJ, SITE_FFN, SITE_CA_ATTN, SITE_CA_PROJ = range(5)

    def _seed(self, site, layer, sub=0):
        """64-bit seed of one dropout site of one layer (splitmix64 of the coordinates)."""
        z = (self.dropout_base_seed + 0x9E3779B97F4A7C15 * (1 + site + 8 * (layer + 1024 * sub))) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def next_dropout_step(self):
        """Advance the device step counter the kernels mix into every dropout seed (call once per training forward;
        a plain device op, so it can be captured in a CUDA graph)."""
        if self._drop_step is None:
            self._drop_step = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._drop_step.add_(1)
        self._drop_step_host += 1

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, idx, targets=None, save=True, train=False, cache=None):
        """idx, targets: int64 [M, B, T'] on the device (T' <= block_size). Returns (logits list of fp32 [B, T', V_i],
        losses fp32 [M] or None). With save=True the activations needed by backward() are kept. train=True applies
        dropout (p = dims.dropout) at the reference's five sites; eval / p = 0 is the deterministic path.
        cache (a KVCache, inference only): idx holds the positions [cache.length, cache.length + T'). An empty cache is
        filled by a normal forward over the prompt; afterwards T' must be 1 and the new row attends to the cached keys /
        values (ta_attention_decode)."""
        d, ops, P = self.d, self.ops, self.P
        pd = float(d.dropout) if train else 0.0
        if pd > 0.0:
            self.next_dropout_step()
        S_ = self._seed
        if self.shadows_stale:
            self.refresh_shadows()
        M, C, H, hs, L = d.M, d.n_embd, d.n_head, d.hs, d.n_layer
        h2p = d.h2p
        assert idx.dim() == 3 and idx.shape[0] == M and idx.dtype == torch.int64 and idx.is_contiguous()
        _, B, T = idx.shape
        if T > d.block_size:
            raise ValueError(f"sequence length {T} exceeds block_size {d.block_size}")
        BT = B * T
        if hasattr(ops, "set_pdl") and "TA_PDL" not in os.environ:
            # launch-latency territory (c2: 282 launches of a few microseconds each): let the next kernel's CTAs be scheduled
            # while the previous one drains; large problems run one persistent CTA per SM and gain nothing (common.cuh)
            # (not on the cached decode path: there the early-scheduled CTAs of the next small kernel take SM slots from the
            # HBM-bound decode-attention kernel - c5 decode 5.7 -> 7.1 ms per position with it)
            ops.set_pdl(M * BT * C <= (8 << 20) and cache is None)
        bf, f32 = self.act_dtype, torch.float32
        S = {"B": B, "T": T, "layers": [], "idx": idx, "pd": pd, "drop_step": self._drop_step_host}
        t0 = 0
        if cache is not None:
            if save or train or targets is not None:
                raise ValueError("the K/V cache is an inference path: save=False, train=False, no targets")
            t0 = cache.length
            if cache.B != B or t0 + T > cache.capacity:
                raise ValueError(f"cache holds {t0} of {cache.capacity} positions for batch {cache.B}; got batch {B}, {T} new")
            if t0 > 0 and T != 1:
                raise ValueError("decode steps take one new position at a time")
        decode = cache is not None and t0 > 0

        def nb(name, l, shape, dtype):
            return self.buf(f"{name}.{l}" if save else name, shape, dtype)

        dyn = decode and cache.dyn        # positions-so-far read from cache.len_dev by the kernels (graph replay)
        x = nb("x0", 0, (M, BT, C), f32)
        if dyn:
            ops.embed_fwd_at(idx.view(M, BT), P["tok"], self._voff, P["pos"], x, cache.len_dev)
        else:
            ops.embed_fwd(idx.view(M, BT), P["tok"], self._voff, P["pos"][t0:t0 + T] if cache is not None else P["pos"], x, T)
        recompute = bool(self.recompute and save and cache is None)
        if recompute:
            # activation recompute: only each block's input is kept (a copy per layer, 2 x 4 bytes per stream element);
            # everything else lives in one shared set of buffers that backward() refills layer by layer
            def nbl(name, l, shape, dtype):
                return self.buf("rc." + name, shape, dtype)
        else:
            nbl = nb
        for l in range(L):
            if recompute and l > 0:
                xin = self.buf(f"xin.{l}", (M, BT, C), f32)
                xin.copy_(x)
                x = xin
            xi = x
            x, Sl = self._layer_forward(l, x, B, T, nbl, pd, cache, t0)
            S["layers"].append({"x0": xi} if recompute else Sl)
        if recompute:
            xin = self.buf(f"xin.{L}", (M, BT, C), f32)
            xin.copy_(x)
            x = xin
        S["recompute"] = recompute
        # output heads   model.py:339-352
        hf = nb("hf", 0, (M, BT, C), bf)
        stf = nb("stf", 0, (2, M * BT), f32)
        ops.layernorm_fwd(x, P["lnf_w"], P["lnf_b"], hf, stf[0], stf[1])
        logits, hhs = [], []
        hg = self._heads_grouped
        if hg is not None:
            V = d.vocab_sizes[0]
            hd_all = nb("hd_all", 0, (M, BT, V // 2), bf)
            lg_all = nb("logits_all", 0, (M, BT, V), f32)
            ops.gemm(hf, hg["S"]["w1"], bias=hg["P"]["b1"], act=ACT_TANH, out_bf16=hd_all)
            ops.gemm(hd_all, hg["S"]["w2"], bias=hg["P"]["b2"], out_f32=lg_all)
            logits, hhs = [lg_all[i] for i in range(M)], [hd_all[i] for i in range(M)]
            S.update(hd_all=hd_all)
        for i, V in enumerate(d.vocab_sizes if hg is None else []):
            V2 = V // 2
            hd = nb(f"hd{i}", 0, (BT, r8(max(V2, 1))), bf)[:, :V2]
            lg = nb(f"logits{i}", 0, (BT, V), f32)
            if V2 > 0:
                ops.gemm(hf[i], self.W(f"hd_w1_{i}"), bias=P[f"hd_b1_{i}"], act=ACT_TANH, out_bf16=hd)
                ops.gemm(hd, self.W(f"hd_w2_{i}"), bias=P[f"hd_b2_{i}"], out_f32=lg)
            else:  # V == 1: Linear(V//2 = 0 -> 1) is just its bias
                lg.copy_(P[f"hd_b2_{i}"].expand(BT, V))
            logits.append(lg)
            hhs.append(hd)
        S.update(x_final=x, hf=hf, stf=stf, hd=hhs, logits=logits)
        losses = None
        if targets is not None:
            assert targets.shape == idx.shape and targets.dtype == torch.int64 and targets.is_contiguous()
            losses = nb("losses", 0, (M,), f32)
            dls = []
            dl_all = nb("dlogits_all", 0, (M, BT, d.vocab_sizes[0]), bf) if (save and hg is not None) else None
            S.update(dl_all=dl_all)
            for i, V in enumerate(d.vocab_sizes):
                row_loss = nb(f"rowloss{i}", 0, (BT,), f32)
                if dl_all is not None:
                    dl = dl_all[i]
                else:
                    dl = nb(f"dlogits{i}", 0, (BT, r8(V)), bf)[:, :V] if save else None
                ops.cross_entropy(logits[i], targets[i].view(BT), row_loss, losses[i:i + 1], dl, 1.0)
                dls.append(dl)
            S["dlogits"] = dls
        if save:       # a forward that does not save uses its own (unsuffixed) buffers and leaves the saved set alone, so
            self._fwd_gen += 1          # forward(save) -> evaluation forward -> backward() works as with torch's autograd
            S["gen"] = self._fwd_gen
            self._saved = S
        if cache is not None:
            cache.length = t0 + T
            if dyn:
                cache.len_dev.add_(1)        # part of the step (and of its graph): the next replay sees the new count
            else:
                cache.len_dev.fill_(cache.length)
        return [lg.view(B, T, -1) for lg in logits], losses

    def _layer_forward(self, l, x, B, T, nb, pd=0.0, cache=None, t0=0):
        """One transformer block (model.py:214-244) on the residual stream x [M, B*T, C]: returns (stream out, the block's saved
        activations). `nb(name, l, shape, dtype)` names the buffers: per layer when they are kept for backward(), shared when
        not (inference, or activation recompute - backward() then calls this again for layer l right before it needs them)."""
        d, ops, P = self.d, self.ops, self.P
        S_ = self._seed
        M, C, H, hs = d.M, d.n_embd, d.n_head, d.hs
        h2p = d.h2p
        BT = B * T
        bf, f32 = self.act_dtype, torch.float32
        decode = cache is not None and t0 > 0
        dyn = decode and cache.dyn
        Sl = {}
        Sl["x0"] = x
        h1 = nb("h1", l, (M, BT, C), bf)
        st1 = nb("st1", l, (2, M * BT), f32)
        ops.layernorm_fwd(x, P["ln1_w"][l], P["ln1_b"][l], h1, st1[0], st1[1])
        # q/k/v tanh-MLPs, stage 1: all heads and q|k|v in one GEMM (N = 3*H*h2p)   model.py:36-50, 62-63, 71
        hid = nb("hid", l, (M, BT, 3 * H * h2p), bf)
        b1 = (self._padded["sa_b1"] if "sa_b1" in self._padded else P["sa_b1"])[l].reshape(M, 3 * H * h2p)
        ops.gemm(h1, self.W("sa_w1")[l].reshape(M, 3 * H * h2p, C), bias=b1, act=ACT_TANH, out_bf16=hid)
        # stage 2: block-diagonal [hs x h2] per (q|k|v, head): grouped GEMM, groups (3H within the row, modality)
        qkv = nb("qkv", l, (M, BT, 3 * C), bf)
        if self._bd is not None:
            ops.gemm(self._hid_groups(hid, BT, self._bd[0]), self._bd[2][l], out_bf16=self._qkv_groups(qkv, BT, self._bd[0]))
        else:
            ops.gemm(self._hid_groups(hid, BT), self.W("sa_w2")[l].reshape(M, 3 * H, hs, -1),
                     out_bf16=self._qkv_groups(qkv, BT))
        o = nb("o", l, (M, BT, C), bf)
        lse = nb("lse", l, (1, M, B, H, T), f32)
        if dyn:
            ops.kv_append(qkv.view(M, B, 3 * C)[..., C:], cache.self_kv[l], cache.len_dev)
        elif cache is not None:    # append this call's keys / values (a strided device copy: cache bookkeeping, not math)
            cache.self_kv[l][:, :, t0:t0 + T].copy_(qkv.view(M, B, T, 3 * C)[..., C:])
        if decode:
            kc = cache.self_kv[l]
            ops.attention_decode(qkv.view(M, B, 3 * C)[..., 0:C], [kc[..., 0:C]], [kc[..., C:2 * C]], o.view(M, B, C),
                                 length=1 if dyn else t0 + 1, H=H, hs=hs, **({"len_dev": cache.len_dev} if dyn else {}))
        else:
            ops.attention_fwd(qkv[..., 0:C], [qkv[..., C:2 * C]], [qkv[..., 2 * C:3 * C]], o, lse, B=B, T=T, H=H, hs=hs,
                              **({"dropout_p": pd, "dropout_seed": S_(self.SITE_SA_ATTN, l), "dropout_step": self._drop_step} if pd > 0.0 else {}))
        # projection MLP + residual   model.py:82-91, 224
        p1 = nb("p1", l, (M, BT, C // 2), bf)
        ops.gemm(o, self.W("sa_pw1")[l], bias=P["sa_pb1"][l], act=ACT_TANH, out_bf16=p1)
        x1 = nb("x1", l, (M, BT, C), f32)
        ops.gemm(p1, self.W("sa_pw2")[l], bias=P["sa_pb2"][l], residual=x, out_f32=x1,
                 **({"dropout_p": pd, "dropout_seed": S_(self.SITE_SA_PROJ, l), "dropout_step": self._drop_step} if pd > 0.0 else {}))
        # feed-forward + residual   model.py:167-172, 225
        hh = nb("h2", l, (M, BT, C), bf)
        st2 = nb("st2", l, (2, M * BT), f32)
        ops.layernorm_fwd(x1, P["ln2_w"][l], P["ln2_b"][l], hh, st2[0], st2[1])
        ffh = nb("ffh", l, (M, BT, 4 * C), bf)
        ops.gemm(hh, self.W("ff_w1")[l], bias=P["ff_b1"][l], act=ACT_RELU, out_bf16=ffh)
        x2 = nb("x2", l, (M, BT, C), f32)
        a_bf = nb("a_bf", l, (M, BT, C), bf) if d.Xc else None
        x3 = None
        if d.Xc and d.Xc < M and pd == 0.0:
            # modalities without cross-attention leave the block unchanged (model.py:229-242 only rewrites the enabled
            # ones): their feed-forward output is written straight into the block's output buffer instead of being copied
            # there afterwards (4 of c4's 8 modalities: 0.5 GB read + written per layer). One launch per run of
            # neighbouring modalities of the same kind. (With dropout the mask is keyed by the group index of ONE launch
            # over all modalities - the copy form below is kept.)
            x3 = nb("x3", l, (M, BT, C), f32)
            flags = [i in d.xc for i in range(M)]
            a = 0
            while a < M:
                e = a + 1
                while e < M and flags[e] == flags[a]:
                    e += 1
                ops.gemm(ffh[a:e], self.W("ff_w2")[l][a:e], bias=P["ff_b2"][l][a:e], residual=x1[a:e],
                         out_f32=(x2 if flags[a] else x3)[a:e], out_bf16=a_bf[a:e])
                a = e
        else:
            ops.gemm(ffh, self.W("ff_w2")[l], bias=P["ff_b2"][l], residual=x1, out_f32=x2, out_bf16=a_bf,
                     **({"dropout_p": pd, "dropout_seed": S_(self.SITE_FFN, l), "dropout_step": self._drop_step} if pd > 0.0 else {}))
        Sl.update(h1=h1, st1=st1, hid=hid, qkv=qkv, o=o, lse=lse, p1=p1, x1=x1, h2=hh, st2=st2, ffh=ffh, x2=x2,
                  a_bf=a_bf)
        x = x2
        if d.Xc:
            x = self._cross_forward(l, Sl, x2, a_bf, B, T, nb, pd, cache, t0, x3)
        return x, Sl

    def decode_step(self, idx, cache):
        """One new position per sequence against the K/V cache (generate's growing phase; idx int64 [M, B, 1]). On a CUDA
        device in the bf16 mode the step is replayed as a CUDA graph: the number of cached positions lives in device memory
        (cache.len_dev) and is read there by the position-embedding, cache-append and decode-attention kernels, so one
        captured graph serves every position. The first decode step of a cache runs eagerly (it allocates the one-row
        buffers), the second one is captured, later ones are replays. TRADEAID_DECODE_GRAPH=0 keeps plain launches."""
        use_graph = (self.device.type == "cuda" and self.precision == "bf16" and hasattr(self.ops, "kv_append")
                     and os.environ.get("TRADEAID_DECODE_GRAPH", "1") != "0")
        if cache.length == 0 or idx.shape[2] != 1:
            raise ValueError("decode_step takes one new position per sequence after the prompt has filled the cache")
        if not use_graph and not os.environ.get("TRADEAID_DECODE_DYN"):
            return self.forward(idx, None, save=False, train=False, cache=cache)[0]
        if cache.length + 1 > cache.capacity:
            raise ValueError(f"cache holds {cache.length} of {cache.capacity} positions")
        cache.dyn = True
        try:
            if not use_graph or cache._dyn_steps == 0:
                cache._dyn_steps += 1
                return self.forward(idx, None, save=False, train=False, cache=cache)[0]
            if cache._graph is None:
                if self.shadows_stale:
                    self.refresh_shadows()
                cache._static_idx = idx.clone()
                length = cache.length
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    cache._static_logits = self.forward(cache._static_idx, None, save=False, train=False, cache=cache)[0]
                cache.length = length        # capture does not run the step: undo the host-side count
                cache._graph = g
            cache._static_idx.copy_(idx)
            cache._graph.replay()
            cache.length += 1
            cache._dyn_steps += 1
            return cache._static_logits
        finally:
            cache.dyn = False

    def _hid_groups(self, hid, BT, gsz=1):
        """[M, BT, 3H*h2p] -> [M, 3H/gsz, BT, gsz*h2]: column slices of gsz neighbouring (q|k|v, head) blocks."""
        d = self.d
        M, H, h2, h2p = d.M, d.n_head, d.h2, d.h2p
        if gsz > 1:
            return hid.as_strided((M, 3 * H // gsz, BT, gsz * h2), (BT * 3 * H * h2, gsz * h2, 3 * H * h2, 1),
                                  hid.storage_offset())
        return hid.as_strided((M, 3 * H, BT, h2), (BT * 3 * H * h2p, h2p, 3 * H * h2p, 1))

    def _qkv_groups(self, qkv, BT, gsz=1):
        d = self.d
        M, H, hs, C = d.M, d.n_head, d.hs, d.n_embd
        return qkv.as_strided((M, 3 * H // gsz, BT, gsz * hs), (BT * 3 * C, gsz * hs, 3 * C, 1), qkv.storage_offset())

    def _sub(self, t, idxs):
        """Views of t[idxs] (first dim) as a list of (view, positions-in-idxs) with arithmetic-progression strides."""
        out, pos = [], 0
        for start, step, n in _runs(idxs):
            out.append((t[start:start + step * (n - 1) + 1:step], list(range(pos, pos + n))))
            pos += n
        return out

    def _cross_forward(self, l, Sl, x2, a_bf, B, T, nb, pd=0.0, cache=None, t0=0, x3=None):
        """Selective cross-attention of layer l (model.py:136-159, 229-242). Every enabled modality queries the SAME
        snapshot x2 of all other modalities, so all K/V projections are issued before any residual update."""
        d, ops, P = self.d, self.ops, self.P
        M, C, H, hs, Xc = d.M, d.n_embd, d.n_head, d.hs, d.Xc
        BT = B * T
        bf, f32 = self.act_dtype, torch.float32
        K = M - 1
        hc = nb("hc", l, (Xc, BT, C), bf)
        stc = nb("stc", l, (2, Xc * BT), f32)
        for xv, pos in self._sub(x2, d.xc):
            # LayerNorm wants a dense [G, rows, C] block: neighbouring enabled modalities go out as one grouped launch,
            # a strided run (every other modality, ...) as one launch per modality
            dense = xv.is_contiguous()
            for k0 in ([0] if dense else range(len(pos))):
                k1 = len(pos) if dense else k0 + 1
                p0, p1 = pos[k0], pos[k1 - 1] + 1
                ops.layernorm_fwd(xv[k0:k1], P["lnc_w"][l][p0:p1], P["lnc_b"][l][p0:p1], hc[p0:p1],
                                  stc[0][p0 * BT:p1 * BT], stc[1][p0 * BT:p1 * BT])
        qc = nb("qc", l, (Xc, BT, C), bf)
        ops.gemm(hc, self.W("ca_wq")[l].reshape(Xc, C, C), out_bf16=qc)
        kvs, kv_runs = [None] * M, []
        Wkv = self.W("ca_wkv")[l]  # [sum_j nq_j, H, 2hs, C]
        for j0, j1, nq in self._kv_runs():
            # sources j0..j1-1 have the same number of queriers: one grouped launch (more tiles per launch = fewer
            # partly filled waves of the persistent grid)
            kv = nb(f"kv{j0}", l, (j1 - j0, BT, nq * 2 * C), bf)
            ops.gemm(a_bf[j0:j1], Wkv[d.qoff[j0]:d.qoff[j1]].reshape(j1 - j0, nq * 2 * C, C), out_bf16=kv)
            kv_runs.append(kv)
            for j in range(j0, j1):
                kvs[j] = kv[j - j0]
        decode = cache is not None and t0 > 0
        dyn = decode and cache.dyn
        if dyn:
            for r, kv in enumerate(kv_runs):
                ops.kv_append(kv.view(kv.shape[0], B, kv.shape[-1]), cache.cross_kv[l][r], cache.len_dev)
        elif cache is not None:
            for r, kv in enumerate(kv_runs):
                cache.cross_kv[l][r][:, :, t0:t0 + T].copy_(kv.view(kv.shape[0], B, T, kv.shape[-1]))
        oc = nb("oc", l, (Xc, BT, C), bf)
        ocs = nb("ocs", l, (Xc, K, 1, BT, C), bf) if K > 1 else None
        lsec = nb("lsec", l, (Xc, K, 1, B, H, T), f32)
        for xi, i in enumerate(d.xc):
            if decode:     # the new row of querier i against the cached K/V of every other modality
                ks, vs = self._kv_cache_views(cache.cross_kv[l], i)
                ops.attention_decode(qc[xi:xi + 1].view(1, B, C), ks, vs, oc[xi:xi + 1].view(1, B, C),
                                     length=1 if dyn else t0 + 1, H=H, hs=hs, head_stride_kv=2 * hs,
                                     **({"len_dev": cache.len_dev} if dyn else {}))
                continue
            ks, vs = self._kv_views(kvs, i)
            ops.attention_fwd(qc[xi:xi + 1], ks, vs, oc[xi:xi + 1], lsec[xi], B=B, T=T, H=H, hs=hs, head_stride_kv=2 * hs,
                              o_src=ocs[xi] if ocs is not None else None,
                              **({"dropout_p": pd, "dropout_seed": self._seed(self.SITE_CA_ATTN, l, xi), "dropout_step": self._drop_step} if pd > 0.0 else {}))
        pc1 = nb("pc1", l, (Xc, BT, C // 2), bf)
        ops.gemm(oc, self.W("ca_pw1")[l], bias=P["ca_pb1"][l], act=ACT_TANH, out_bf16=pc1)
        prefilled = x3 is not None       # the rows of the modalities without cross-attention were written by the feed-forward
        if x3 is None:
            x3 = nb("x3", l, (M, BT, C), f32)
        for (xv, pos), (ov, _) in zip(self._sub(x2, d.xc), self._sub(x3, d.xc)):
            p0, p1 = pos[0], pos[-1] + 1
            ops.gemm(pc1[p0:p1], self.W("ca_pw2")[l][p0:p1], bias=P["ca_pb2"][l][p0:p1], residual=xv, out_f32=ov,
                     **({"dropout_p": pd, "dropout_seed": self._seed(self.SITE_CA_PROJ, l, p0), "dropout_step": self._drop_step} if pd > 0.0 else {}))
        for i in range(M):
            if i not in d.xc and not prefilled:
                x3[i].copy_(x2[i])
        Sl.update(hc=hc, stc=stc, qc=qc, kvs=kvs, oc=oc, ocs=ocs, lsec=lsec, pc1=pc1, x3=x3)
        return x3

    def _kv_runs(self):
        """Maximal runs [j0, j1) of consecutive source modalities with the same non-zero number of queriers -> (j0, j1, nq):
        their K/V projections (and the backward of those) have one shape and go out as one grouped launch."""
        d, out, j = self.d, [], 0
        while j < d.M:
            nq = len(d.queriers[j])
            k = j + 1
            while k < d.M and len(d.queriers[k]) == nq:
                k += 1
            if nq > 0:
                out.append((j, k, nq))
            j = k
        return out

    def _kv_cache_views(self, run_caches, i):
        """K and V views [1, B, capacity, *] of querier i inside the cached K/V of each of its sources."""
        d = self.d
        C, hs = d.n_embd, d.hs
        ks, vs = [], []
        runs = self._kv_runs()
        for j in d.others(i):
            r = next(k for k, (j0, j1, _) in enumerate(runs) if j0 <= j < j1)
            p = d.queriers[j].index(i)
            blk = run_caches[r][j - runs[r][0]:j - runs[r][0] + 1, :, :, p * 2 * C:(p + 1) * 2 * C]
            ks.append(blk)
            vs.append(blk[..., hs:])
        return ks, vs

    def _kv_views(self, kvs, i, grads=False):
        """K and V column views (head 0, dim 0) of querier i inside each source's stacked K/V buffer."""
        d = self.d
        C, hs = d.n_embd, d.hs
        ks, vs = [], []
        for j in d.others(i):
            p = d.queriers[j].index(i)
            blk = kvs[j][:, p * 2 * C:(p + 1) * 2 * C].unsqueeze(0)
            ks.append(blk)
            vs.append(blk[..., hs:])
        return ks, vs

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, loss_weights=None, stage_hook=None):
        """Gradients of sum_i w_i * loss_i (w_i = 1: main.py:646) into self.grad; needs forward(..., save=True).
        `stage_hook(stage)` is called when every gradient of a stage has been enqueued: "heads" (output heads + final
        LayerNorm), then each layer index from n_layer-1 down to 0, then "embed" — the data-parallel trainer starts the
        all-reduce of a gradient bucket there while the kernels of the earlier layers are still to run. ("part", l) is
        reported once the first part of layer l's backward is enqueued (its cross-attention, or its feed-forward when
        there is none): the window in which the previous bucket's collective runs beside the kernels."""
        S = self._saved
        if S is None or "dlogits" not in S:
            raise RuntimeError("backward() needs a preceding forward(idx, targets, save=True)")
        d, ops, G = self.d, self.ops, self.G
        M, C, H, hs, L = d.M, d.n_embd, d.n_head, d.hs, d.n_layer
        h2p = d.h2p
        B, T = S["B"], S["T"]
        BT = B * T
        bf, f32 = self.act_dtype, torch.float32
        pd = S.get("pd", 0.0)
        restore_step = None
        if pd > 0.0 and S.get("drop_step", self._drop_step_host) != self._drop_step_host:
            # a later training-mode forward advanced the device counter: regenerate THIS forward's masks
            restore_step = self._drop_step_host
            self._drop_step.fill_(S["drop_step"])
        # upstream gradient of a dropout site = stream gradient o keep / (1 - p), same mask as the forward epilogue
        dxm = self.buf("dxm", (M, BT, C), bf) if pd > 0.0 else None

        def masked(t, site, sub=0):
            if pd == 0.0:
                return t
            out = dxm[:t.shape[0]]
            ops.dropout_bf16(t, out, pd, self._seed(site, l, sub), self._drop_step)
            return out

        cs_ws = self.buf("cs_ws", (max(ops.colsum_workspace(BT, 4 * C, M), ops.colsum_workspace(BT, max(d.vocab_sizes), M),
                                       ops.colsum_workspace(BT, 3 * H * h2p, M)),), f32)
        # the workspace depends on the launch shape: grouped launches (all modalities) and the per-modality launches
        # of the cross-attention LayerNorm share one buffer
        ln_ws = self.buf("ln_ws", (max([ops.layernorm_bwd_workspace(M, BT, C)] +
                                       [ops.layernorm_bwd_workspace(k, BT, C) for k in range(1, max(d.Xc, 1) + 1)]),), f32)
        dx = self.buf("dx", (M, BT, C), f32)
        dxb = self.buf("dxb", (M, BT, C), bf)
        # ---- output heads
        dhf = self.buf("dhf", (M, BT, C), bf)
        hg = self._heads_grouped if (loss_weights is None and S.get("dl_all") is not None) else None
        if hg is not None:     # every modality in one grouped launch per op
            dl_all, hd_all = S["dl_all"], S["hd_all"]
            ops.colsum(dl_all, hg["G"]["b2"], cs_ws)
            ops.gemm(dl_all, hd_all, a_major=1, b_major=1, out_f32=hg["G"]["w2"])
            dhd_all = self.buf("dhd_all", hd_all.shape, bf)
            ops.gemm(dl_all, hg["S"]["w2"], b_major=1, aux=hd_all, aux_mode=AUX_TANH_BWD, out_bf16=dhd_all)
            ops.gemm(dhd_all, S["hf"], a_major=1, b_major=1, out_f32=hg["G"]["w1"])
            ops.colsum(dhd_all, hg["G"]["b1"], cs_ws)
            ops.gemm(dhd_all, hg["S"]["w1"], b_major=1, out_bf16=dhf)
        for i, V in enumerate(d.vocab_sizes if hg is None else []):
            V2 = V // 2
            dl = S["dlogits"][i]
            w = 1.0 if loss_weights is None else float(loss_weights[i])
            ops.colsum(dl.unsqueeze(0), G[f"hd_b2_{i}"].unsqueeze(0), cs_ws)
            if w != 1.0:
                G[f"hd_b2_{i}"].mul_(w)
            if V2 == 0:
                dhf[i].zero_()
                continue
            hd = S["hd"][i]
            ops.gemm(dl, hd, a_major=1, b_major=1, out_f32=G[f"hd_w2_{i}"], alpha=w)
            dhd = self.buf(f"dhd{i}", (BT, r8(V2)), bf)[:, :V2]
            ops.gemm(dl, self.W(f"hd_w2_{i}"), b_major=1, aux=hd, aux_mode=AUX_TANH_BWD, out_bf16=dhd, alpha=w)
            ops.gemm(dhd, S["hf"][i], a_major=1, b_major=1, out_f32=G[f"hd_w1_{i}"])
            ops.colsum(dhd.unsqueeze(0), G[f"hd_b1_{i}"].unsqueeze(0), cs_ws)
            ops.gemm(dhd, self.W(f"hd_w1_{i}"), b_major=1, out_bf16=dhf[i])
        ops.layernorm_bwd(dhf, S["x_final"], self.P["lnf_w"], S["stf"][0], S["stf"][1], None, dx, dxb, G["lnf_w"],
                          G["lnf_b"], ln_ws)
        if stage_hook is not None:
            stage_hook("heads")
        # ---- layers, last to first
        def nbr(name, l_, shape, dtype):
            return self.buf("rc." + name, shape, dtype)

        for l in reversed(range(L)):
            Sl = S["layers"][l]
            if S.get("recompute"):      # refill the shared activation buffers with this block's forward (same dropout masks:
                _, Sl = self._layer_forward(l, Sl["x0"], B, T, nbr, pd)      # they are functions of (seed, step, coordinates))
            if d.Xc:
                self._cross_backward(l, Sl, dx, dxb, B, T, cs_ws, ln_ws, pd, dxm, stage_hook)
            # feed-forward
            dy = masked(dxb, self.SITE_FFN)
            ops.gemm(dy, Sl["ffh"], a_major=1, b_major=1, out_f32=G["ff_w2"][l])
            ops.colsum(dy, G["ff_b2"][l], cs_ws)
            dffh = self.buf("dffh", (M, BT, 4 * C), bf)
            ops.gemm(dy, self.W("ff_w2")[l], b_major=1, aux=Sl["ffh"], aux_mode=AUX_RELU_BWD, out_bf16=dffh)
            ops.gemm(dffh, Sl["h2"], a_major=1, b_major=1, out_f32=G["ff_w1"][l])
            ops.colsum(dffh, G["ff_b1"][l], cs_ws)
            dh = self.buf("dh", (M, BT, C), bf)
            ops.gemm(dffh, self.W("ff_w1")[l], b_major=1, out_bf16=dh)
            ops.layernorm_bwd(dh, Sl["x1"], self.P["ln2_w"][l], Sl["st2"][0], Sl["st2"][1], dx, dx, dxb, G["ln2_w"][l],
                              G["ln2_b"][l], ln_ws)
            if stage_hook is not None and not d.Xc:
                stage_hook(("part", l))
            # self-attention projection MLP
            dy = masked(dxb, self.SITE_SA_PROJ)
            ops.gemm(dy, Sl["p1"], a_major=1, b_major=1, out_f32=G["sa_pw2"][l])
            ops.colsum(dy, G["sa_pb2"][l], cs_ws)
            dp1 = self.buf("dp1", (M, BT, C // 2), bf)
            ops.gemm(dy, self.W("sa_pw2")[l], b_major=1, aux=Sl["p1"], aux_mode=AUX_TANH_BWD, out_bf16=dp1)
            ops.gemm(dp1, Sl["o"], a_major=1, b_major=1, out_f32=G["sa_pw1"][l])
            ops.colsum(dp1, G["sa_pb1"][l], cs_ws)
            do = self.buf("do", (M, BT, C), bf)
            ops.gemm(dp1, self.W("sa_pw1")[l], b_major=1, out_bf16=do)
            # attention
            qkv = Sl["qkv"]
            dqkv = self.buf("dqkv", (M, BT, 3 * C), bf)
            ops.attention_bwd(qkv[..., 0:C], [qkv[..., C:2 * C]], [qkv[..., 2 * C:3 * C]], Sl["o"], do, Sl["lse"],
                              dqkv[..., 0:C], [dqkv[..., C:2 * C]], [dqkv[..., 2 * C:3 * C]], B=B, T=T, H=H, hs=hs,
                              **({"dropout_p": pd, "dropout_seed": self._seed(self.SITE_SA_ATTN, l), "dropout_step": self._drop_step} if pd > 0.0 else {}))
            # q/k/v MLP stage 2 (block diagonal) and stage 1
            hidg, dqkvg = self._hid_groups(Sl["hid"], BT), self._qkv_groups(dqkv, BT)
            ops.gemm(dqkvg, hidg, a_major=1, b_major=1, out_f32=G["sa_w2"][l].reshape(M, 3 * H, hs, d.h2))
            dhid = self.buf("dhid", (M, BT, 3 * H * h2p), bf)
            if self._bd is not None:
                gsz = self._bd[0]
                ops.gemm(self._qkv_groups(dqkv, BT, gsz), self._bd[2][l], b_major=1, aux=self._hid_groups(Sl["hid"], BT, gsz),
                         aux_mode=AUX_TANH_BWD, out_bf16=self._hid_groups(dhid, BT, gsz))
            else:
                ops.gemm(dqkvg, self.W("sa_w2")[l].reshape(M, 3 * H, hs, -1), b_major=1, aux=hidg, aux_mode=AUX_TANH_BWD,
                         out_bf16=self._hid_groups(dhid, BT))
            if h2p == d.h2:
                ops.gemm(dhid, Sl["h1"], a_major=1, b_major=1, out_f32=G["sa_w1"][l].reshape(M, 3 * H * h2p, C))
                ops.colsum(dhid, G["sa_b1"][l].reshape(M, 3 * H * h2p), cs_ws)
            else:
                gw = self.buf("gw1_pad", (M, 3 * H * h2p, C), f32)
                gb = self.buf("gb1_pad", (M, 3 * H * h2p), f32)
                ops.gemm(dhid, Sl["h1"], a_major=1, b_major=1, out_f32=gw)
                ops.colsum(dhid, gb, cs_ws)
                ops.unpad(gw.view(M * 3 * H, h2p, C), G["sa_w1"][l].reshape(M * 3 * H, d.h2, C))
                ops.unpad(gb.view(M * 3 * H, 1, h2p), G["sa_b1"][l].reshape(M * 3 * H, 1, d.h2))
            ops.gemm(dhid, self.W("sa_w1")[l].reshape(M, 3 * H * h2p, C), b_major=1, out_bf16=dh)
            ops.layernorm_bwd(dh, Sl["x0"], self.P["ln1_w"][l], Sl["st1"][0], Sl["st1"][1], dx, dx, dxb, G["ln1_w"][l],
                              G["ln1_b"][l], ln_ws)
            if stage_hook is not None:
                stage_hook(l)
        ops.embed_bwd(S["idx"].view(M, BT), dx, self._voff, max(d.vocab_sizes), G["tok"], G["pos"][:T], B, T)
        if T < d.block_size:
            G["pos"][T:].zero_()
        if restore_step is not None:
            self._drop_step.fill_(restore_step)
        if stage_hook is not None:
            stage_hook("embed")

    def _cross_backward(self, l, Sl, dx, dxb, B, T, cs_ws, ln_ws, pd=0.0, dxm=None, stage_hook=None):
        d, ops, G = self.d, self.ops, self.G
        M, C, H, hs, Xc = d.M, d.n_embd, d.n_head, d.hs, d.Xc
        BT = B * T
        bf, f32 = self.act_dtype, torch.float32
        K = M - 1
        # projection MLP of every enabled modality (its upstream gradient is the stream gradient of that modality)
        dpc1 = self.buf("dpc1", (Xc, BT, C // 2), bf)
        for dv, pos in self._sub(dxb, d.xc):
            p0, p1 = pos[0], pos[-1] + 1
            if pd > 0.0:   # same mask as the forward epilogue of this launch (seed keyed by layer and first position)
                dvm = dxm[:dv.shape[0]]
                ops.dropout_bf16(dv, dvm, pd, self._seed(self.SITE_CA_PROJ, l, p0), self._drop_step)
                dv = dvm
            ops.gemm(dv, Sl["pc1"][p0:p1], a_major=1, b_major=1, out_f32=G["ca_pw2"][l][p0:p1])
            ops.colsum(dv, G["ca_pb2"][l][p0:p1], cs_ws)
            ops.gemm(dv, self.W("ca_pw2")[l][p0:p1], b_major=1, aux=Sl["pc1"][p0:p1], aux_mode=AUX_TANH_BWD,
                     out_bf16=dpc1[p0:p1])
        ops.gemm(dpc1, Sl["oc"], a_major=1, b_major=1, out_f32=G["ca_pw1"][l])
        ops.colsum(dpc1, G["ca_pb1"][l], cs_ws)
        doc = self.buf("doc", (Xc, BT, C), bf)
        ops.gemm(dpc1, self.W("ca_pw1")[l], b_major=1, out_bf16=doc)
        dqc = self.buf("dqc", (Xc, BT, C), bf)
        dkvs, dkv_runs = [None] * M, []
        for j0, j1, nq in self._kv_runs():
            t = self.buf(f"dkv{j0}", (j1 - j0, BT, nq * 2 * C), bf)
            dkv_runs.append(t)
            for j in range(j0, j1):
                dkvs[j] = t[j - j0]
        for xi, i in enumerate(d.xc):
            ks, vs = self._kv_views(Sl["kvs"], i)
            dks, dvs = self._kv_views(dkvs, i)
            ops.attention_bwd(Sl["qc"][xi:xi + 1], ks, vs, Sl["oc"][xi:xi + 1], doc[xi:xi + 1], Sl["lsec"][xi],
                              dqc[xi:xi + 1], dks, dvs, B=B, T=T, H=H, hs=hs, head_stride_kv=2 * hs,
                              o_src=Sl["ocs"][xi] if Sl["ocs"] is not None else None,
                              **({"dropout_p": pd, "dropout_seed": self._seed(self.SITE_CA_ATTN, l, xi), "dropout_step": self._drop_step} if pd > 0.0 else {}))
        if stage_hook is not None and self._dp_window == "attn":
            stage_hook(("part", l))      # the first few milliseconds of this layer's backward are enqueued
        # query projection and its LayerNorm (input: the snapshot x2 of the querying modality)
        ops.gemm(dqc, Sl["hc"], a_major=1, b_major=1, out_f32=G["ca_wq"][l].reshape(Xc, C, C))
        dhc = self.buf("dhc", (Xc, BT, C), bf)
        ops.gemm(dqc, self.W("ca_wq")[l].reshape(Xc, C, C), b_major=1, out_bf16=dhc)
        for xv, pos in self._sub(Sl["x2"], d.xc):
            dense = xv.is_contiguous()
            for k0 in ([0] if dense else range(len(pos))):
                k1 = len(pos) if dense else k0 + 1
                p0, p1 = pos[k0], pos[k1 - 1] + 1
                i0, i1 = d.xc[p0], d.xc[p1 - 1] + 1
                ops.layernorm_bwd(dhc[p0:p1], Sl["x2"][i0:i1], self.P["lnc_w"][l][p0:p1], Sl["stc"][0][p0 * BT:p1 * BT],
                                  Sl["stc"][1][p0 * BT:p1 * BT], dx[i0:i1], dx[i0:i1], dxb[i0:i1], G["lnc_w"][l][p0:p1],
                                  G["lnc_b"][l][p0:p1], ln_ws)
        # key/value projections: gradients land on the (un-normalised) stream of every SOURCE modality
        Wkv, Gkv = self.W("ca_wkv")[l], G["ca_wkv"][l]
        for (j0, j1, nq), dkv in zip(self._kv_runs(), dkv_runs):
            sl, n = slice(d.qoff[j0], d.qoff[j1]), j1 - j0
            ops.gemm(dkv, Sl["a_bf"][j0:j1], a_major=1, b_major=1, out_f32=Gkv[sl].reshape(n, nq * 2 * C, C))
            ops.gemm(dkv, Wkv[sl].reshape(n, nq * 2 * C, C), b_major=1, residual=dx[j0:j1], out_f32=dx[j0:j1],
                     out_bf16=dxb[j0:j1])
        if stage_hook is not None and self._dp_window != "attn":
            stage_hook(("part", l))

    # ------------------------------------------------------------------------------------------ bookkeeping
    def zero_grad(self):
        self.grad.zero_()

    def activation_bytes(self):
        return sum(t.numel() * t.element_size() for t in self._bufs.values())

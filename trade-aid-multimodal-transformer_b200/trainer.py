"""Engine-owned training step (the fast tier): get_batch gather -> forward -> backward -> [gradient all-reduce] ->
fused AdamW, optionally replayed as CUDA graphs. Semantics of one step are those of the reference loop body
(/root/reference/main.py:641-650): loss = sum of per-modality cross-entropies, zero_grad, backward, AdamW step with
torch defaults.

Data parallelism (BASELINE.json: 1/2/4/8 B200): one process per GPU, full replica each, per-rank batch shard, one
exchange per iteration — an all-reduce (sum) of the flat fp32 gradient buffer over NCCL, with the 1/R mean folded into
the AdamW kernel's grad_scale. The exchange is BUCKETED in the order backward produces gradients (output heads and the
last layers first, the embedding tables last): as soon as the kernels of a bucket's layers are enqueued, the all-reduce
of that bucket's gradient ranges is issued asynchronously (one coalesced NCCL call on the process group's own stream,
which waits for exactly the work enqueued so far), so it runs beside the backward of the earlier layers. In the window
in which that collective runs (the first part of the next layer's backward: a few milliseconds, the collective of a
one-layer bucket takes ~2) the persistent kernels leave `sm_margin` SMs free for its CTAs (ta_set_sm_margin) — a
one-CTA-per-SM grid would otherwise run its displaced CTAs as a second wave; the rest of the layer runs full grids
(a margin over the whole backward costs the GEMMs 11 % of their SMs all the time: measured +18 ms per step on c4). `overlap=False` keeps the plain form
(backward, then the whole-buffer all-reduce in `grad_chunks` pieces).
"""
from __future__ import annotations

import os

import torch


class Trainer:
    def __init__(self, engine, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, process_group=None, world_size=1,
                 grad_chunks=8, overlap=None, grad_buckets=None, sm_margin=None, grad_dtype=None):
        self.eng = engine
        self.ops = engine.ops
        self.lr, self.betas, self.eps, self.wd = float(lr), betas, float(eps), float(weight_decay)
        n = engine.layout.n_active
        dev = engine.device
        self.n = n
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.scalars = torch.zeros(4, dtype=torch.float32, device=dev)
        self.pg = process_group
        self.world = int(world_size)
        if self.world > 1:   # every rank draws its own dropout masks (its own batch shard, model.py:69 semantics per sample)
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                engine.dropout_base_seed += 0x632BE59BD9B4E019 * (dist.get_rank(process_group) + 1)
        self.grad_chunks = max(1, int(grad_chunks))
        if overlap is None:
            overlap = os.environ.get("TRADEAID_DP_OVERLAP", "1") != "0"
        self.overlap = bool(overlap) and self.world > 1
        L = engine.d.n_layer
        if grad_buckets is None:
            grad_buckets = int(os.environ.get("TRADEAID_DP_BUCKETS", "12"))
        self.grad_buckets = max(1, min(int(grad_buckets), L))
        if sm_margin is None:
            sm_margin = int(os.environ.get("TRADEAID_DP_SM_MARGIN", "16"))
        self.sm_margin = int(sm_margin) if (self.overlap and hasattr(self.ops, "set_sm_margin")) else 0
        # "bf16": every bucket is cast to bf16 (ta_cast_bf16, on the compute stream, right before its collective is issued),
        # all-reduced in bf16 - half the NVLink bytes - and consumed by the AdamW kernel as bf16 (no cast back). "fp32" keeps
        # the exchange exact: the data-parallel PARITY mode (DP == single process to fp32 summation order) uses it.
        if grad_dtype is None:
            grad_dtype = os.environ.get("TRADEAID_DP_GRAD_DTYPE", "fp32")
        if grad_dtype not in ("fp32", "bf16"):
            raise ValueError("grad_dtype must be 'fp32' or 'bf16'")
        self.grad_dtype = grad_dtype if self.world > 1 else "fp32"
        self.grad16 = torch.zeros(n, dtype=torch.bfloat16, device=dev) if self.grad_dtype == "bf16" else None
        self._buckets = self._make_buckets() if self.overlap else []
        self._works = []
        self.skip_exchange = False       # measurement only (bench.py): run the data-parallel step without its collectives
        self._graphs = None
        self.launches_per_step = None

    # ------------------------------------------------------------------------------------------ buckets
    def _make_buckets(self):
        """[(closing stage, [gradient views])]: bucket b is complete when backward reports `closing stage`. Layers are
        split into grad_buckets runs of consecutive layers, last layers first; the head gradients ride with the first
        bucket, the embedding tables with the last one (which closes at 'embed')."""
        lay, g = self.eng.layout, self.eng.grad
        L, nb = self.eng.d.n_layer, self.grad_buckets
        bounds = [round(L * k / nb) for k in range(nb + 1)]          # ascending layer boundaries
        runs = [list(range(bounds[k], bounds[k + 1])) for k in range(nb) if bounds[k + 1] > bounds[k]]
        runs.reverse()                                               # backward order: last layers first
        buckets = []
        for b, run in enumerate(runs):
            ranges = lay.grad_ranges("layers", run)
            if b == 0:
                ranges = lay.grad_ranges("heads") + ranges
            last = b == len(runs) - 1
            if last:
                ranges = ranges + lay.grad_ranges("embed")
            # 8-element alignment of every range start keeps the bf16 copies 16-byte aligned (cast / AdamW vector accesses)
            buckets.append(("embed" if last else run[0], [g[a:e] for a, e in ranges], ranges))
        return buckets

    def bucket_bytes(self):
        per = 2 if self.grad_dtype == "bf16" else 4
        return [sum(t.numel() * per for t in b[1]) for b in self._buckets]

    def _exchange_views(self, views, ranges):
        """What goes on the wire for a bucket: the fp32 gradient ranges themselves, or their bf16 copies (cast here)."""
        if self.grad16 is None:
            return views
        out = []
        for t, (a, e) in zip(views, ranges):
            dst = self.grad16[a:e]
            if a % 8 == 0:
                self.ops.cast_bf16(t, dst)
            else:                                  # unaligned start: the strided conversion kernel has no alignment needs
                self.ops.convert(t.view(1, -1), dst.view(1, -1))
            out.append(dst)
        return out

    def _issue_bucket(self, views, ranges=None):
        """One coalesced asynchronous all-reduce(sum) of a bucket's gradient ranges. The process group's stream waits
        for the work enqueued on the current stream so far; the current stream is not blocked."""
        import torch.distributed as dist
        if ranges is not None:
            views = self._exchange_views(views, ranges)
        if self.skip_exchange:
            return
        cmgr = getattr(dist, "_coalescing_manager", None)
        if cmgr is None:       # no coalescing support in this torch: one asynchronous all-reduce per range
            for t in views:
                self._works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
            return
        with cmgr(group=self.pg, async_ops=True) as cm:
            for t in views:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
        self._works.append(cm)

    def _wait_buckets(self):
        for w in self._works:
            w.wait()
        self._works = []

    # ------------------------------------------------------------------------------------------ pieces
    def forward_backward(self, idx, tgt, stage_hook=None):
        logits, losses = self.eng.forward(idx, tgt, save=True, train=True)
        self.eng.backward(stage_hook=stage_hook)
        return losses

    def allreduce_grads(self):
        if self.world <= 1:
            return
        import torch.distributed as dist
        if self.skip_exchange:
            return
        g = self.eng.grad[:self.n]
        if self.grad16 is not None:
            self.ops.cast_bf16(g, self.grad16)
            g = self.grad16
        per = (self.n + self.grad_chunks - 1) // self.grad_chunks
        per = (per + 63) // 64 * 64
        works = []
        for a in range(0, self.n, per):
            works.append(dist.all_reduce(g[a:min(self.n, a + per)], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        for w in works:
            w.wait()

    def optimizer_step(self):
        eng, ops = self.eng, self.ops
        ops.adamw_tick(self.step_dev, self.scalars, self.lr, self.betas[0], self.betas[1], self.wd)
        ops.adamw_step(eng.master[:self.n], self.grad16 if self.grad16 is not None else eng.grad[:self.n], self.exp_avg,
                       self.exp_avg_sq, eng.shadow[:self.n],
                       self.scalars, self.betas[0], self.betas[1], self.eps, 1.0 / self.world)
        eng.refresh_shadows(flat_cast=False)

    def _set_margin(self, sms):
        if self.sm_margin:
            self.ops.set_sm_margin(sms)

    # ------------------------------------------------------------------------------------------ eager step
    def step(self, idx, tgt):
        """One training step on device-resident int64 [M, B, T] batches. Returns the per-modality losses [M] (device)."""
        if not self.overlap:
            losses = self.forward_backward(idx, tgt)
            self.allreduce_grads()
            self.optimizer_step()
            return losses
        closing = {stage: (views, ranges) for stage, views, ranges in self._buckets}

        def hook(stage):
            if isinstance(stage, tuple):             # ("part", l): the previous bucket's collective is over by now
                self._set_margin(0)
                return
            b = closing.get(stage)
            if b is not None:
                self._issue_bucket(*b)
                self._set_margin(self.sm_margin)     # a collective now runs beside the next kernels

        try:
            losses = self.forward_backward(idx, tgt, stage_hook=hook)
        finally:
            self._set_margin(0)
        self._wait_buckets()
        self.optimizer_step()
        return losses

    # ------------------------------------------------------------------------------------------ CUDA graphs
    def capture(self, idx_static, tgt_static):
        """Capture the step into CUDA graphs over STATIC batch buffers (the feeder writes them with
        TokenStore.gather(ix, T, out=(idx_static, tgt_static)) before each replay). Single GPU: one graph holds forward,
        backward and the optimizer. Data parallel: the optimizer is its own graph after the exchange; with overlap the
        forward+backward program is cut into one graph per gradient bucket, and the bucket's all-reduce is issued
        between the replays."""
        eng = self.eng
        if eng.shadows_stale:
            eng.refresh_shadows()
        # warm-up on a side stream: allocates every buffer and sets kernel attributes; it only writes activations
        # and the gradient buffer, so parameters and optimizer state are untouched
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.forward_backward(idx_static, tgt_static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.ops.reset_launch_count()
        segments = []          # [(graph, gradient views to all-reduce after it | None)]
        if not self.overlap:
            g_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fb):
                self._losses_static = self.forward_backward(idx_static, tgt_static)
                if self.world <= 1:
                    self.optimizer_step()
            segments.append((g_fb, None))
        else:
            closing = {stage: (views, ranges) for stage, views, ranges in self._buckets}
            pool = torch.cuda.graph_pool_handle()
            cap = torch.cuda.Stream()
            cap.wait_stream(torch.cuda.current_stream())
            state = {"g": None}

            def begin():
                state["g"] = torch.cuda.CUDAGraph()
                state["g"].capture_begin(pool=pool)

            def hook(stage):
                if isinstance(stage, tuple):         # ("part", l): full grids again for the rest of the layer
                    self._set_margin(0)
                    return
                b = closing.get(stage)
                if b is None:
                    return
                if self.grad16 is not None:          # the bf16 copy of the bucket is part of the captured segment
                    b = (self._exchange_views(*b), None)
                state["g"].capture_end()
                segments.append((state["g"], b))
                state["g"] = None
                self._set_margin(self.sm_margin)
                if stage != "embed":
                    begin()

            try:
                with torch.cuda.stream(cap):
                    begin()
                    self._losses_static = self.forward_backward(idx_static, tgt_static, stage_hook=hook)
                    assert state["g"] is None, "the last bucket must close at the end of backward"
            finally:
                self._set_margin(0)
            torch.cuda.current_stream().wait_stream(cap)
        g_opt = None
        if self.world > 1:
            g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_opt):
                self.optimizer_step()
        self.launches_per_step = self.ops.launch_count()
        self._graphs = (segments, g_opt)
        return self._losses_static

    def replay(self):
        segments, g_opt = self._graphs
        for g, b in segments:
            g.replay()
            if b is not None:
                self._issue_bucket(*b)
        if g_opt is not None:
            if self.overlap:
                self._wait_buckets()
            else:
                self.allreduce_grads()
            g_opt.replay()
        return self._losses_static

    # ------------------------------------------------------------------------------------------ state
    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_dev, "lr": self.lr}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"]); self.step_dev.copy_(sd["step"])

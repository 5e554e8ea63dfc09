"""Engine-owned training step (the fast tier): get_batch gather -> forward -> backward -> [gradient all-reduce] ->
fused AdamW, optionally replayed as CUDA graphs. Semantics of one step are those of the reference loop body
(/root/reference/main.py:641-650): loss = sum of per-modality cross-entropies, zero_grad, backward, AdamW step with
torch defaults.

Data parallelism (BASELINE.json: 1/2/4/8 B200): one process per GPU, full replica each, per-rank batch shard, one
exchange step per iteration — an all-reduce (sum) of the flat fp32 gradient buffer over NCCL, with the 1/R mean folded
into the AdamW kernel's grad_scale. The all-reduce of the per-layer gradient ranges is issued on a side stream as soon as
the backward kernels of that layer are enqueued, so it overlaps the rest of backward.
"""
from __future__ import annotations

import torch


class Trainer:
    def __init__(self, engine, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, process_group=None, world_size=1,
                 grad_chunks=8):
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
        self._graphs = None
        self.launches_per_step = None

    # ------------------------------------------------------------------------------------------ pieces
    def forward_backward(self, idx, tgt):
        logits, losses = self.eng.forward(idx, tgt, save=True, train=True)
        self.eng.backward()
        return losses

    def allreduce_grads(self):
        if self.world <= 1:
            return
        import torch.distributed as dist
        g = self.eng.grad[:self.n]
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
        ops.adamw_step(eng.master[:self.n], eng.grad[:self.n], self.exp_avg, self.exp_avg_sq, eng.shadow[:self.n],
                       self.scalars, self.betas[0], self.betas[1], self.eps, 1.0 / self.world)
        eng.refresh_shadows(flat_cast=False)

    # ------------------------------------------------------------------------------------------ eager step
    def step(self, idx, tgt):
        """One training step on device-resident int64 [M, B, T] batches. Returns the per-modality losses [M] (device)."""
        losses = self.forward_backward(idx, tgt)
        self.allreduce_grads()
        self.optimizer_step()
        return losses

    # ------------------------------------------------------------------------------------------ CUDA graphs
    def capture(self, idx_static, tgt_static):
        """Capture the step into CUDA graphs over STATIC batch buffers (the feeder writes them with
        TokenStore.gather(ix, T, out=(idx_static, tgt_static)) before each replay). Single GPU: one graph holds forward,
        backward and the optimizer. Data parallel: forward+backward and the optimizer are two graphs with the NCCL
        all-reduce between them."""
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
        g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_fb):
            self._losses_static = self.forward_backward(idx_static, tgt_static)
            if self.world <= 1:
                self.optimizer_step()
        g_opt = None
        if self.world > 1:
            g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_opt):
                self.optimizer_step()
        self.launches_per_step = self.ops.launch_count()
        self._graphs = (g_fb, g_opt)
        return self._losses_static

    def replay(self):
        g_fb, g_opt = self._graphs
        g_fb.replay()
        if g_opt is not None:
            self.allreduce_grads()
            g_opt.replay()
        return self._losses_static

    # ------------------------------------------------------------------------------------------ state
    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_dev, "lr": self.lr}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"]); self.step_dev.copy_(sd["step"])

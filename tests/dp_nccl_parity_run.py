"""Data-parallel PARITY mode on real GPUs over NCCL (test infrastructure): under torchrun with world 1 or 2,
three training steps of the engine-owned step on a fixed GLOBAL batch (every rank draws the same indices and gathers its
B/world rows: training_utils.batch_shard) with the bucketed, overlapped fp32 gradient all-reduce, then estimate_loss dealt over
the ranks (training_utils.eval_process_group). Rank 0 saves the parameters, the per-step losses of its shard rows and the
evaluation output; tests/test_dp_nccl_gpu.py compares world 2 with world 1.
usage: [torchrun ...] python tests/dp_nccl_parity_run.py OUT.pt [graph]"""
import contextlib
import io
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tradeaid_b200 import synthetic as syn, training_utils as tu  # noqa: E402
from tradeaid_b200.engine import Engine  # noqa: E402
from tradeaid_b200.params import Dims  # noqa: E402
from tradeaid_b200.trainer import Trainer  # noqa: E402

out_path = sys.argv[1]
use_graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
pg = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD

V, cross, T, C, H, L, B = [256, 100, 64, 300], [True, False, False, True], 64, 256, 8, 3, 8
cfg = dict(kinds=["price", "volume", "time", "indicator"], cross=cross, vocab=V, rows=20_000)
eng = Engine(Dims(V, cross, T, C, H, L), dev)
eng.layout.init_(eng.master, generator=torch.Generator(device=dev).manual_seed(0))
start = eng.master.clone()
tr = Trainer(eng, lr=1e-3, process_group=pg, world_size=world, overlap=True, grad_buckets=3, grad_dtype="fp32")
streams = syn.make_streams(cfg, "cpu", cfg["rows"])
n_train = 18_000


class _Model:                                  # what estimate_loss needs of training_utils.m
    def eval(self): pass
    def train(self): pass

    def __call__(self, xb, yb):
        logits, losses = eng.forward(torch.stack(xb), torch.stack(yb), save=False)
        return logits, [losses[i] for i in range(len(xb))]


tu.configure(device=str(dev), batch_size=B, block_size=T, eval_iters=6)
tu.all_train_sets = [s[:n_train].tolist() for s in streams]
tu.all_val_sets = [s[n_train:] for s in streams]
tu.all_vocabularies = [[float(v) for v in range(n)] for n in V]
tu.all_modality_params = syn.modality_params(cfg)
tu.all_file_info = [[n_train]] * 4
tu.file_lengths = [n_train]
tu.num_modalities = 4
tu.is_percents = False
tu.m = _Model()
tu.eval_process_group = pg
tu.batch_shard = (rank, world) if world > 1 else None
torch.manual_seed(5)
step_losses = []
static = None
for _ in range(3):
    xb, yb = tu.get_batch("train", 1)
    idx, tgt = torch.stack(xb), torch.stack(yb)
    if use_graph:
        if static is None:
            static = (idx.clone(), tgt.clone())
            tr.capture(*static)
        static[0].copy_(idx); static[1].copy_(tgt)
        losses = tr.replay()
    else:
        losses = tr.step(idx, tgt)
    step_losses.append(losses.clone())
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    ev = tu.estimate_loss(3, 3)
torch.cuda.synchronize()
if world > 1:
    gathered = [torch.zeros_like(eng.master) for _ in range(world)]
    dist.all_gather(gathered, eng.master)
    assert all(torch.equal(gathered[0], t) for t in gathered), "replicas diverged"
    # the loss of the global batch is the mean of the shard losses (equal shard sizes)
    sl = torch.stack(step_losses)
    dist.all_reduce(sl, op=dist.ReduceOp.SUM)
    step_losses = list(sl / world)
if rank == 0:
    torch.save({"master": eng.master.cpu(), "start": start.cpu(), "step_losses": torch.stack(step_losses).cpu(), "eval": ev,
                "eval_text": buf.getvalue(), "world": world}, out_path)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

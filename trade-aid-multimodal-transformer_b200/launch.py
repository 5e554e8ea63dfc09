"""Data-parallel launcher: the reference's training program over the engine-owned step, one process per GPU.

    python -m tradeaid_b200.launch --reference /path/to/trade-AId-multimodal-transformer            (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
           -m tradeaid_b200.launch --reference /path/to/trade-AId-multimodal-transformer           (8 GPUs, NCCL)

What runs, and where it comes from:
  * everything the reference's main.py does BEFORE its loop (/root/reference/main.py:1-596: config.yaml / input_schemas.yaml,
    data loading and preparation, the module globals of training_utils, model construction or checkpoint load, the header of
    the training log) is the reference's own code, executed unchanged from `--reference` with `model` and `training_utils`
    resolved to this package (the INTEGRATION.md shims) — main.py is a script, so its text up to the loop is executed as is;
  * the loop itself (main.py:598-668) is restated here over `Trainer`: same evaluation cadence (`iter % eval_interval == 0 or
    iter == max_iters - 1` -> estimate_loss), same console lines and log-file lines, same NaN handling and early stopping,
    same checkpoint cadence (`save_model`), `get_batch('train', 1)` for every step; the step is forward + hand-written
    backward + bucketed NCCL gradient all-reduce overlapped with backward + fused AdamW (torch.optim.AdamW defaults, as
    main.py:464) instead of `sum(losses).backward(); optimizer.step()`.

Data parallelism: `--mode parity` (default) keeps the GLOBAL batch of config.yaml: every rank draws the same
batch_size indices from identically seeded host generators and gathers its own B/world rows (`training_utils.batch_shard`), the
gradients are averaged, so the run is the single-process run up to fp32 summation order. `--mode throughput` gives every rank
its own batch of batch_size windows (rank-seeded): world x the samples per step. Evaluation is dealt to the ranks
(`training_utils.eval_process_group`); rank 0 prints, writes the log file and saves checkpoints.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import random
import sys
from datetime import datetime

import torch

LOOP_MARKER = "best_val_loss = float('inf')"       # first statement of main.py's loop section (main.py:594)


def parse(argv=None):
    ap = argparse.ArgumentParser(prog="python -m tradeaid_b200.launch", description=__doc__.split("\n\n")[0])
    ap.add_argument("--reference", required=True, help="directory of the reference program (main.py, config.yaml, ...)")
    ap.add_argument("--mode", choices=["parity", "throughput"], default="parity")
    ap.add_argument("--seed", type=int, default=1337, help="seed of torch / random on every rank (main.py itself is unseeded)")
    ap.add_argument("--backend", default=None, help="torch.distributed backend (default: nccl on cuda, gloo on cpu)")
    ap.add_argument("--graph", action="store_true", help="replay the step as CUDA graphs (static batch buffers)")
    ap.add_argument("--precision", default=None, choices=["bf16", "bf16x3"])
    ap.add_argument("--recompute", action="store_true", help="activation recompute (TRADEAID_RECOMPUTE=1): one stream copy per block "
                    "is kept, backward reruns each block's forward")
    ap.add_argument("--buckets", type=int, default=None, help="gradient buckets of the overlapped exchange")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--grad-dtype", default=None, choices=["fp32", "bf16"], help="what the gradient buckets travel as; default: "
                    "fp32 in parity mode (the run equals the single-process run), bf16 in throughput mode")
    return ap.parse_args(argv)


def run_preamble(ref_dir, rank):
    """Execute main.py up to its training loop, unchanged, with `model` / `training_utils` resolved to this package.
    Returns its namespace. Ranks other than 0 run it silently and without a log file."""
    from . import model as our_model, training_utils as our_tu
    ref_dir = os.path.abspath(ref_dir)
    path = os.path.join(ref_dir, "main.py")
    with open(path) as f:
        src = f.read()
    cut = src.find(LOOP_MARKER)
    if cut < 0:
        raise RuntimeError(f"{path}: cannot find the start of the training loop ({LOOP_MARKER!r})")
    os.chdir(ref_dir)                      # the reference resolves config.yaml and its data paths against the working directory
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    sys.modules["model"] = our_model
    sys.modules["training_utils"] = our_tu
    if rank != 0:
        import compatibility_layer as cl      # the reference's; only rank 0 owns the log file
        orig = cl.get_system_configuration

        def quiet_cfg(*a, **k):
            c = dict(orig(*a, **k))
            c["output_file_name"] = ""
            return c
        cl.get_system_configuration = quiet_cfg
    ns = {"__name__": "__main__", "__file__": path}
    out = contextlib.nullcontext() if rank == 0 else contextlib.redirect_stdout(io.StringIO())
    with out:
        exec(compile(src[:cut], path, "exec"), ns)
    return ns


def train_loop(ns, trainer, rank, world, use_graph=False):
    """main.py:594-668 over Trainer.step. `ns` is the namespace main.py's preamble left behind."""
    from . import training_utils as tu
    m = ns["m"]
    max_iters, eval_interval = ns["max_iters"], ns["eval_interval"]
    output_file_name, output_file_path = ns["output_file_name"], ns["output_file_path"]
    save_model, model_file_name = ns["save_model"], ns["model_file_name"]
    say = print if rank == 0 else (lambda *a, **k: None)
    best_val_loss = float('inf')
    patience = 1000
    no_improvement_count = 0
    static = None
    last = None

    def save(final):
        if rank != 0:
            return
        model_dir = os.path.dirname(model_file_name)
        if model_dir and not os.path.exists(model_dir):
            os.makedirs(model_dir, exist_ok=True)
        current_time = datetime.now().strftime("%H:%M:%S")
        if final:
            print(f'Final Save: Model checkpoint | {current_time}')
        torch.save(m.state_dict(), model_file_name)
        if final:
            print(f"Final Save: {round(os.path.getsize(model_file_name)/1024**2,2)} MB complete")
        else:
            print()
            print(f'Saved: Model checkpoint ({round(os.path.getsize(model_file_name)/1024**2,2)} MB) | {current_time}')
            print()

    for it in range(max_iters):
        if it % 100 == 0:
            say(f'Training: Iteration {it}/{max_iters}')
        if it % eval_interval == 0 or it == max_iters - 1:
            losses = tu.estimate_loss(it, max_iters)
            current_time = datetime.now().strftime("%H:%M:%S")
            if not torch.isnan(torch.tensor([losses['train'], losses['val']])).any():
                say(f"\nLOSS METRICS: Step {it}/{max_iters} | Train: {losses['train']:.4f} | Val: {losses['val']:.4f} | Time: {current_time}")
                say("-" * 80)
                if output_file_name != '' and rank == 0:
                    with open(output_file_path, 'a', encoding='utf-8') as f:
                        progress_pct = (it / max_iters) * 100
                        f.write(f"\nSTEP {it:,}/{max_iters:,} ({progress_pct:.1f}% Complete) | Training Loss: {losses['train']:.6f} | "
                                f"Validation Loss: {losses['val']:.6f} | {current_time}\n\n")
            else:
                say(f"Warning: Step {it} losses are NaN, skipping save | {current_time}")
            if not torch.isnan(torch.tensor(losses['val'])).any():
                if losses['val'] < best_val_loss:
                    best_val_loss = losses['val']
                    no_improvement_count = 0
                else:
                    no_improvement_count += 1
                if no_improvement_count >= patience:          # the losses are all-reduced: every rank takes the same branch
                    say(f"Training: Early stopping (no improvement for {patience} evaluations)")
                    break
        if save_model == 1 and (it % eval_interval == 0 or it == max_iters - 1):
            save(final=False)
        xb_list, yb_list = tu.get_batch('train', 1)
        idx, tgt = torch.stack(xb_list), torch.stack(yb_list)
        if use_graph:
            if static is None:
                static = (idx.clone(), tgt.clone())
                trainer.capture(*static)
            static[0].copy_(idx)
            static[1].copy_(tgt)
            last = trainer.replay()
        else:
            last = trainer.step(idx, tgt)
    say("\nTRAINING COMPLETED SUCCESSFULLY")
    if save_model == 1:
        save(final=True)
    return last


def main(argv=None):
    args = parse(argv)
    import torch.distributed as dist
    from . import model as our_model, training_utils as tu
    from .trainer import Trainer
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.precision:
        our_model.set_precision(args.precision)
    if args.recompute:
        os.environ["TRADEAID_RECOMPUTE"] = "1"      # read by Engine.__init__ when main.py's preamble builds the model
    if torch.cuda.is_available():
        torch.cuda.set_device(local)       # config.yaml's `device: cuda` then means this rank's GPU
    torch.manual_seed(args.seed)
    random.seed(args.seed)
    ns = run_preamble(args.reference, rank)
    m = ns["m"]
    eng = m._require_engine()
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = args.backend or ("nccl" if eng.device.type == "cuda" else "gloo")
        if not dist.is_initialized():
            dist.init_process_group(backend, **({"device_id": eng.device} if backend == "nccl" else {}))
        pg = dist.group.WORLD
        dist.broadcast(eng.master, src=0, group=pg)      # one set of initial weights, whatever each rank drew or loaded
        eng.shadows_stale = True
        tu.eval_process_group = pg
        if args.mode == "parity":
            tu.batch_shard = (rank, world)
    kw = {}
    if args.buckets is not None:
        kw["grad_buckets"] = args.buckets
    kw["grad_dtype"] = args.grad_dtype or ("fp32" if args.mode == "parity" else "bf16")
    trainer = Trainer(eng, lr=ns["learning_rate"], process_group=pg, world_size=world, overlap=not args.no_overlap, **kw)
    # the same generator positions on every rank from here on in parity mode; own streams per rank in throughput mode
    seed = args.seed + 1 + (rank if args.mode == "throughput" else 0)
    torch.manual_seed(seed)
    random.seed(seed)
    try:
        train_loop(ns, trainer, rank, world, use_graph=args.graph)
    finally:
        tu.batch_shard = None
        tu.eval_process_group = None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — training throughput of the trade-AId multimodal-transformer hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c4] [--impl ours|reference]

A step = one pass of the hot path over one batch of synthetic input: get_batch (host torch.randint -> index upload ->
device window gather) -> forward -> backward -> [gradient all-reduce] -> fused AdamW. One JSON line on stdout (rank 0).
Metric (BASELINE.json): training tokens/s (tokens = modalities x batch x block_size per step, whole job).
`--impl reference` times the reference's own CPU implementation on the host cores: the UNMODIFIED reference from oracle/_ref
(model.MultimodalTransformer + training_utils.get_batch + torch.optim.AdamW, main.py:641-650; oracle/make_ref.py copies it in
the build container and it travels with the snapshot), falling back to the oracle port (oracle/restatement.py) when that copy
is absent. `--impl reference-eager` is the extra comparator of BASELINE.md section 3: the same unmodified reference on the
B200 in torch eager (`--ref-mode fp32|tf32|autocast`). `--tier compat` measures the drop-in the way main.py drives it
(m(xb, yb); sum(losses).backward(); stock torch.optim.AdamW) instead of the engine-owned Trainer step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-eager"])
    ap.add_argument("--ref-mode", default="fp32", choices=["fp32", "tf32", "autocast"], help="reference-eager: matmul precision")
    ap.add_argument("--tier", default="engine", choices=["engine", "compat"], help="compat: the reference's own loop body "
                    "(nn.Module forward, autograd bridge, torch.optim.AdamW) over the kernels")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--config", default=os.environ.get("TRADEAID_BENCH_CONFIG", "c4"))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--rows", type=int, default=0, help="synthetic rows override")
    ap.add_argument("--dropout", type=float, default=0.0, help="dropout probability of the training step (default 0: the parity configuration)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--all-cross", action="store_true", help="cross-attention on EVERY modality (SURVEY appendix E: 'also report "
                    "all-T'); c4 then needs --batch 32, or --recompute, to fit 180 GB")
    ap.add_argument("--recompute", action="store_true", help="activation recompute (Engine(recompute=True)): one stream copy per block "
                    "is kept, backward reruns each block's forward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=0, help="batch of the bounded CPU sample")
    ap.add_argument("--no-overlap", action="store_true", help="DP: whole-buffer all-reduce after backward instead of "
                    "bucketed all-reduces overlapped with backward")
    ap.add_argument("--grad-dtype", default="bf16", choices=["fp32", "bf16"], help="DP: what the gradient buckets travel as "
                    "(bf16: half the NVLink bytes, consumed by AdamW as bf16; fp32: the exact exchange of the parity mode)")
    ap.add_argument("--buckets", type=int, default=12, help="DP: gradient buckets (runs of consecutive layers)")
    ap.add_argument("--sm-margin", type=int, default=16, help="DP overlap: SMs the persistent kernels leave to NCCL "
                    "(also the cap on NCCL's CTAs)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d.get("hbm_gbs", 6650.0), tc_burst=d.get("bf16_tflops", 1590.0),
                    tc_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback")


def traffic_of_dominant_launch():
    """DRAM bytes (read + write) of ONE launch of the dominant kernel, from the committed `ncu --set full` capture
    (profiles/r02_roofline_traffic.json - the capture of the latest round - names the launch and its algorithmic bytes); None if
    no such file exists."""
    p = next((q for q in (os.path.join(ROOT, "profiles", f"r{r:02d}_roofline_traffic.json") for r in (2, 1)) if os.path.exists(q)), None)
    if p is None:
        return None
    with open(p) as f:
        d = json.load(f)
    return {"bytes_per_launch": d["dram_bytes_read"] + d["dram_bytes_write"], "algorithmic_bytes": d["algorithmic_bytes"],
            "launch": d["launch"], "tensor_pipe_active_pct": d.get("tensor_pipe_active_pct"), "source": d["source"]}


# ------------------------------------------------------------------------------------------------ CPU side
def cpu_oracle_run(cfg, batch, steps, warmup, lr=3e-4):
    """The reference algorithm on the host cores (oracle port): get_batch + forward + backward + AdamW per step.
    Returns (tokens_per_s, seconds_per_step, threads)."""
    from oracle import restatement as R
    from tradeaid_b200 import synthetic as syn
    torch.set_num_threads(os.cpu_count() or 1)
    T, V = cfg["block_size"], cfg["vocab"]
    M = len(V)
    ocfg = R.Cfg(T, cfg["n_embd"], cfg["n_head"], cfg["n_layer"], V, cfg["cross"])
    sd = R.random_state_dict(ocfg, seed=0)
    rows = min(cfg["rows"], 200_000)
    data = syn.make_streams(cfg, "cpu", rows)
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v2 = {k: torch.zeros_like(v) for k, v in sd.items()}
    torch.manual_seed(1)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        ix = R.batch_starting_indices(rows, T, batch, "train", [rows], False)
        xb, yb = R.gather_batch(data, ix, T)
        _, _, grads = R.forward_backward(sd, ocfg, xb, yb)
        R.adamw_step(sd, grads, m, v2, s + 1, lr)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return M * batch * T / sec, sec, torch.get_num_threads()


CPU_SAMPLE_BATCH = {"c1": 4, "c2": 16, "c3": 4, "c4": 1, "c5": 1}     # bounded CPU sample: batch of the timed reference step


def cpu_reference_run(cfg, batch, steps, warmup, lr=3e-4):
    """The UNMODIFIED reference (oracle/_ref) on the host cores through its own loop body (main.py:641-650). Returns
    (tokens_per_s, seconds_per_step, threads) or None when oracle/_ref is absent."""
    from oracle import ref_runner
    if not ref_runner.available():
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    run = ref_runner.ReferenceRun(cfg, batch, "cpu", min(cfg["rows"], 200_000), lr=lr)
    sec, _ = run.timed(steps, warmup)
    return len(cfg["vocab"]) * batch * cfg["block_size"] / sec, sec, torch.get_num_threads()


def cpu_baseline_of(cfg, config_name, batch, steps, warmup):
    """cpu_baseline object: oracle/_ref ("reference") when it is there, else the oracle port ("port")."""
    r = cpu_reference_run(cfg, batch, steps, warmup)
    if r is not None:
        tps, sec, threads = r
        return tps, sec, {"value": tps, "unit": "tokens/s", "cores": threads, "kind": "reference",
                          "sample": f"unmodified reference (oracle/_ref: model.MultimodalTransformer + training_utils.get_batch + "
                                    f"torch.optim.AdamW, main.py:641-650) on the host CPU, fp32, full model, batch {batch} of "
                                    f"{cfg['batch']}, 200 k synthetic rows, {steps} timed steps ({sec:.2f} s/step)"}
    tps, sec, threads = cpu_oracle_run(cfg, batch, steps, warmup)
    return tps, sec, {"value": tps, "unit": "tokens/s", "cores": threads, "kind": "port",
                      "sample": f"oracle port (oracle/restatement.py) of model.py fwd+bwd + AdamW + get_batch on CPU (oracle/_ref "
                                f"absent), full model, batch {batch} of {cfg['batch']}, {steps} timed steps ({sec:.2f} s/step)"}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_batch or max(1, min(cfg["batch"], CPU_SAMPLE_BATCH.get(args.config, 1)))
    # bounded: a c4 step of the unmodified reference is ~45 s on 8 host cores (plus ~60 s to build the 1.89 G-parameter model)
    heavy = args.config in ("c4", "c5")
    steps, warmup = max(1, min(args.steps, 2 if heavy else 5)), min(args.warmup, 1)
    tps, sec, cb = cpu_baseline_of(cfg, args.config, B, steps, warmup)
    line = {
        "impl": "reference", "metric": "training tokens/s", "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "timed_steps": steps, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg, B, 1), "cpu_baseline": cb,
        "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_reference_eager(args, cfg):
    """BASELINE.md section 3, "GPU comparator": the unmodified reference with device='cuda' in torch eager — fp32, TF32 matmuls,
    or torch.autocast(bf16). Not the driver's reference arm (that one is the CPU path); an extra line for profiles/."""
    from oracle import ref_runner
    if not ref_runner.available():
        print(json.dumps({"impl": "reference-eager", "unavailable": "oracle/_ref is missing"}), flush=True)
        return
    torch.cuda.set_device(0)
    B = args.batch or cfg["batch"]
    sampler = ClockSampler(0)
    run = ref_runner.ReferenceRun(cfg, B, "cuda", min(args.rows or cfg["rows"], 200_000), mode=args.ref_mode)
    for _ in range(max(1, args.warmup)):
        run.step()
    torch.cuda.synchronize()
    sampler.start()
    sec, loss = run.timed(args.steps, 0)
    clocks = sampler.stop()
    M, T = len(cfg["vocab"]), cfg["block_size"]
    tps = M * B * T / sec
    from tradeaid_b200 import synthetic as syn
    fl = syn.train_flops_per_step(cfg, B)
    print(json.dumps({"impl": "reference-eager", "ref_mode": args.ref_mode, "metric": "training tokens/s", "value": tps,
                      "unit": "tokens/s", "n_gpus": 1, "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": sec * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": {"fp32": "f32", "tf32": "tf32", "autocast": "bf16 (torch.autocast)"}[args.ref_mode],
                      "data": "synthetic", "config": workload_config(args, cfg, B, 1), "clocks": clocks,
                      "achieved_tflops": fl / sec / 1e12, "loss_sum_last_step": loss,
                      "hbm_bytes_allocated": torch.cuda.max_memory_allocated(),
                      "note": "unmodified reference (oracle/_ref) through its own loop body, torch eager on the B200"}), flush=True)


def workload_config(args, cfg, batch, world):
    return {"workload": f"{args.config}: {cfg['note']}", "modalities": len(cfg["vocab"]), "cross_attention": cfg["cross"],
            "vocab": cfg["vocab"], "rows": args.rows or cfg["rows"], "n_embd": cfg["n_embd"], "n_head": cfg["n_head"],
            "n_layer": cfg["n_layer"], "block_size": cfg["block_size"], "batch_per_gpu": batch,
            "global_batch": batch * world, "parallelism": f"dp{world}", "dropout": float(getattr(args, "dropout", 0.0)),
            "l2": "working set per step (activations + weights) exceeds the 126 MB L2 many times over"
                  if args.config in ("c3", "c4", "c5") else "inputs smaller than L2; a 256 MB buffer is written between steps"}


# ------------------------------------------------------------------------------------------------ instrumentation
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


class TimedOps:
    """Proxy of the kernel namespace that brackets every call with CUDA events on the launching stream."""

    def __init__(self, ops, blockdiag=None):
        self._ops = ops
        self.records = []   # (name, start_event, end_event, flops)
        self.detail = {}
        # (storage pointer of the packed block-diagonal stage-2 weights, blocks per group): those launches multiply exact
        # zeros off the diagonal - only 1/gsz of the launched FLOPs are algorithmic (model.py:36-50)
        self._bd = blockdiag
        self.launched_flops = 0.0

    def __getattr__(self, name):
        fn = getattr(self._ops, name)
        if not callable(fn) or name.endswith("workspace") or name in ("launch_count", "reset_launch_count", "raw"):
            return fn

        def wrapped(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flops = 0.0
            if name == "gemm":
                A, B = a[0], a[1]
                am, bm = kw.get("a_major", 0), kw.get("b_major", 0)
                K = A.shape[-1] if am == 0 else A.shape[-2]
                Mm = A.shape[-2] if am == 0 else A.shape[-1]
                N = B.shape[-2] if bm == 0 else B.shape[-1]
                G = 1
                for s in A.shape[:-2]:
                    G *= s
                flops = 2.0 * Mm * N * K * G
                self.launched_flops += flops
                if self._bd is not None and B.untyped_storage().data_ptr() == self._bd[0]:
                    flops /= self._bd[1]
                name_key = f"gemm M={Mm} N={N} K={K} G={G} maj={am}{bm} act={kw.get('act', 0)} aux={kw.get('aux_mode', 0)} " \
                           f"res={int(kw.get('residual') is not None)} f32={int(kw.get('out_f32') is not None)} " \
                           f"bf16={int(kw.get('out_bf16') is not None)}"
                self.detail.setdefault(name_key, []).append((e0, e1, flops))
            e0.record()
            out = fn(*a, **kw)
            e1.record()
            self.records.append((name, e0, e1, flops))
            return out

        return wrapped

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, e0, e1, fl in self.records:
            a = agg.setdefault(name, [0.0, 0, 0.0])
            a[0] += e0.elapsed_time(e1)
            a[1] += 1
            a[2] += fl
        return agg


# ------------------------------------------------------------------------------------------------ main arm
def run_ours(args, cfg):
    import torch.distributed as dist
    from tradeaid_b200 import lib, synthetic as syn
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.feeder import TokenStore, generate_batch_starting_indices
    from tradeaid_b200.params import Dims
    from tradeaid_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        opts = None
        if not args.no_overlap and args.sm_margin > 0:
            # the collective runs beside backward: bound its CTAs by the SMs the persistent kernels leave free
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = args.sm_margin
            opts.config.min_ctas = min(args.sm_margin, 8)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
        pg = dist.group.WORLD
    pk = peaks()
    B = args.batch or cfg["batch"]
    T, V = cfg["block_size"], cfg["vocab"]
    M = len(V)
    rows = args.rows or cfg["rows"]

    # ---- model, data, trainer
    dims = Dims(V, cfg["cross"], T, cfg["n_embd"], cfg["n_head"], cfg["n_layer"], dropout=args.dropout)
    eng = Engine(dims, dev, precision=args.precision, recompute=args.recompute)
    eng.layout.init_(eng.master, generator=torch.Generator(device=dev).manual_seed(0))   # same weights on every rank
    tr = Trainer(eng, lr=3e-4, process_group=pg, world_size=world, overlap=not args.no_overlap,
                 grad_buckets=args.buckets, sm_margin=args.sm_margin, grad_dtype=args.grad_dtype)
    streams = syn.make_streams(cfg, dev, rows)
    n_train = int(rows * 0.9)
    store = TokenStore([s[:n_train] for s in streams], dev)
    del streams
    torch.manual_seed(1 + rank)     # each rank samples its own batch shard
    idx = torch.zeros(M, B, T, dtype=torch.int64, device=dev)
    tgt = torch.zeros_like(idx)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.config in ("c1", "c2") else None

    def feed():
        ix = generate_batch_starting_indices(n_train, T, B, "train", [n_train], False)
        store.gather(ix, T, out=(idx, tgt))

    feed()
    use_graph = not args.no_graph
    if use_graph:
        tr.capture(idx, tgt)
        launches_per_step = tr.launches_per_step + 1      # + the gather kernel
        step = tr.replay
    else:
        lib.reset_launch_count()
        tr.step(idx, tgt)
        launches_per_step = lib.launch_count() + 1
        step = lambda: tr.step(idx, tgt)  # noqa: E731

    def one_step():
        if flush is not None:
            flush.zero_()
        feed()
        return step()

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses = one_step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    # what the exchange costs: the same step with its collectives left out (measurement only), max over ranks
    exposed_ms = None
    if world > 1:
        tr.skip_exchange = True
        for _ in range(2):
            one_step()
        torch.cuda.synchronize()
        dist.barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record()
        for _ in range(max(3, args.steps // 2)):
            one_step()
        n1.record()
        torch.cuda.synchronize()
        t = torch.tensor([n0.elapsed_time(n1) / max(3, args.steps // 2)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exposed_ms = ms_per_step - float(t.item())
        tr.skip_exchange = False
    tokens_per_step = M * B * T * world
    value = tokens_per_step / (ms_per_step * 1e-3)
    loss_now = float(losses.sum().item())

    # ---- end to end through the public step API with HOST buffers: pinned xb/yb -> H2D, step, losses -> D2H
    e2e = None
    if not args.no_e2e:
        host_store = store.store.cpu()
        xh = torch.empty(M, B, T, dtype=torch.int64).pin_memory()
        yh = torch.empty(M, B, T, dtype=torch.int64).pin_memory()
        lh = torch.empty(M, dtype=torch.float32).pin_memory()
        ar = torch.arange(T)
        n_e2e = max(3, min(args.steps, 10))

        def e2e_step():
            ix = generate_batch_starting_indices(n_train, T, B, "train", [n_train], False)
            pos = ix[:, None] + ar
            xh.copy_(host_store[:, pos]); yh.copy_(host_store[:, pos + 1])
            idx.copy_(xh, non_blocking=True); tgt.copy_(yh, non_blocking=True)
            ls = step()
            lh.copy_(ls, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(lh.sum())

        e2e_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": tokens_per_step * n_e2e / dt, "unit": "tokens/s", "h2d_bytes_per_step": 2 * M * B * T * 8,
               "d2h_bytes_per_step": 4 * M, "steps": n_e2e,
               "api": "host index draw + host window gather into pinned int64 [M,B,T] x2 -> cudaMemcpyAsync -> "
                      "Trainer step (fwd+bwd+AdamW) -> per-modality losses copied back and read on the host every step"}

    # ---- roofline of the dominant kernel (the tcgen05 GEMM), from one instrumented eager step
    base_ops = eng.ops
    timed = TimedOps(base_ops, (eng._bd[2].untyped_storage().data_ptr(), eng._bd[0]) if eng._bd is not None else None)
    eng.ops = tr.ops = timed
    tr.step(idx, tgt)          # un-timed settle
    timed.records.clear()
    tr.step(idx, tgt)
    agg = timed.summary()
    eng.ops = tr.ops = base_ops
    if rank == 0 and os.environ.get("TRADEAID_BENCH_DETAIL"):
        rows = []
        for k, evs in timed.detail.items():
            evs = evs[len(evs) // 2:]       # second (timed) step only
            ms = sum(a.elapsed_time(b) for a, b, _ in evs)
            fl = sum(f for _, _, f in evs)
            rows.append((ms, len(evs), fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0, k))
        with open(os.environ["TRADEAID_BENCH_DETAIL"], "w") as f:
            for ms, n, tf, k in sorted(rows, reverse=True):
                f.write(f"{ms:9.3f} ms  {n:4d} launches  {tf:8.1f} TF/s  {k}\n")
    tot_ms = sum(v[0] for v in agg.values())
    gm = agg.get("gemm", [0.0, 0, 0.0])
    gemm_tflops = gm[2] / (gm[0] * 1e-3) / 1e12 if gm[0] > 0 else 0.0
    breakdown = {k: {"ms": round(v[0], 3), "launches": v[1], "share": round(v[0] / tot_ms, 4)} for k, v in
                 sorted(agg.items(), key=lambda kv: -kv[1][0])}
    flops_step = syn.train_flops_per_step(cfg, B)
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05.mma + TMA, all nn.Linear fwd/bwd)",
                "achieved": gemm_tflops, "peak": pk["tc_sustained"], "unit": "TFLOP/s",
                "frac": gemm_tflops / pk["tc_sustained"], "traffic": (traffic_of_dominant_launch() or {}).get("bytes_per_launch"),
                "traffic_detail": traffic_of_dominant_launch(),
                "peak_source": f"{pk['source']} (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)",
                "launches_per_step": gm[1], "gemm_ms_per_step": round(gm[0], 3), "gemm_share_of_step": round(gm[0] / tot_ms, 4),
                "algorithmic_flops_per_step_gemm": gm[2], "launched_flops_per_step_gemm": timed.launched_flops / 2,
                "achieved_note": "algorithmic FLOPs (the exact zeros of the block-diagonal stage-2 packing are not counted) / "
                                 "CUDA-event time of the GEMM launches of one eager step",
                "step_mfu": {"algorithmic_train_tflop_per_step": flops_step / 1e12,
                             "achieved_tflops_per_gpu": flops_step / (ms_per_step * 1e-3) / 1e12,
                             "frac_of_measured_sustained": flops_step / (ms_per_step * 1e-3) / 1e12 / pk["tc_sustained"],
                             "frac_of_nominal_2250": flops_step / (ms_per_step * 1e-3) / 1e12 / 2250.0}}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = args.cpu_batch or CPU_SAMPLE_BATCH.get(args.config, 1)
        heavy = args.config in ("c4", "c5")        # ~45 s per step of the unmodified reference on 8 cores: one timed step
        _, _, cpu_baseline = cpu_baseline_of(cfg, args.config, cb, 1 if heavy else 2, 0 if heavy else 1)
    if rank == 0:
        line = {"metric": "training tokens/s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (split bf16 operands, fp32 activations)",
                "data": "synthetic", "config": workload_config(args, cfg, B, world), "precision": args.precision, "tier": "engine", "activation_recompute": bool(args.recompute),
                "exchange": None if world == 1 else (
                    {"kind": f"bucketed {tr.grad_dtype} all-reduce overlapped with backward", "buckets": len(tr._buckets),
                     "bucket_bytes": tr.bucket_bytes(), "bytes_per_step": sum(tr.bucket_bytes()), "sm_margin": tr.sm_margin,
                     "exposed_ms_per_step": exposed_ms,
                     "exposed_note": "ms_per_step minus the same step replayed without its collectives (max over ranks)"}
                    if tr.overlap else
                    {"kind": f"whole-buffer {tr.grad_dtype} all-reduce after backward", "chunks": tr.grad_chunks,
                     "exposed_ms_per_step": exposed_ms}),
                "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu_baseline, "breakdown_eager_step": breakdown, "cuda_graph": use_graph,
                "loss_sum_last_step": loss_now, "positions_per_s": value / M,
                "hbm_bytes_allocated": torch.cuda.max_memory_allocated(dev)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ compat tier
def run_compat(args, cfg):
    """`--tier compat`: the step the way the reference's main.py drives it (main.py:641-650) over the drop-in nn.Module —
    `logits, losses = m(xb_list, yb_list)`, `sum(losses).backward()` through the autograd bridge into the engine's backward
    program, stock `torch.optim.AdamW(m.parameters(), lr)`; no CUDA graph, the operand copies of the weights are re-cast
    after every optimizer step. One GPU (the reference has no data parallelism of its own)."""
    from tradeaid_b200 import lib, model as tmodel, synthetic as syn
    from tradeaid_b200.feeder import TokenStore, generate_batch_starting_indices
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    T, V = cfg["block_size"], cfg["vocab"]
    M = len(V)
    rows = args.rows or cfg["rows"]
    tmodel.configure(T, cfg["n_embd"], cfg["n_head"], cfg["n_layer"], dropout=args.dropout)
    tmodel.set_precision(args.precision)
    torch.manual_seed(0)
    m = tmodel.MultimodalTransformer(M, list(V), syn.modality_params(cfg)).to(dev)
    opt = torch.optim.AdamW(m.parameters(), lr=3e-4)
    streams = syn.make_streams(cfg, dev, rows)
    n_train = int(rows * 0.9)
    store = TokenStore([s[:n_train] for s in streams], dev)
    del streams
    torch.manual_seed(1)
    B = args.batch or cfg["batch"]

    def step(B):
        ix = generate_batch_starting_indices(n_train, T, B, "train", [n_train], False)
        xb, yb = store.gather(ix, T)
        logits_list, losses_list = m([xb[i] for i in range(M)], [yb[i] for i in range(M)])
        total_loss = sum(losses_list)
        opt.zero_grad(set_to_none=True)
        total_loss.backward()
        opt.step()
        return total_loss

    while True:
        try:
            for _ in range(max(args.warmup, 3)):
                step(B)
            break
        except torch.OutOfMemoryError:
            opt.zero_grad(set_to_none=True)
            torch.cuda.empty_cache()
            if B == 1:
                raise
            B //= 2                 # torch.optim.AdamW's foreach temporaries on top of the engine's buffers
    torch.cuda.synchronize()
    lib.reset_launch_count()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        loss = step(B)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    ms_per_step = e0.elapsed_time(e1) / args.steps
    value = M * B * T / (ms_per_step * 1e-3)
    pk = peaks()
    fl = syn.train_flops_per_step(cfg, B)
    print(json.dumps({"metric": "training tokens/s", "value": value, "unit": "tokens/s", "n_gpus": 1, "steps": args.steps,
                      "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3", "data": "synthetic",
                      "config": workload_config(args, cfg, B, 1), "tier": "compat", "precision": args.precision,
                      "api": "tradeaid_b200.model.MultimodalTransformer: m(xb, yb) -> sum(losses).backward() -> torch.optim.AdamW.step() "
                             "(main.py:641-650), device-side get_batch gather",
                      "wall_ms_per_step": wall / args.steps * 1e3, "gpu_launches": lib.launch_count(), "clocks": clocks,
                      "step_mfu": {"achieved_tflops_per_gpu": fl / (ms_per_step * 1e-3) / 1e12,
                                   "frac_of_measured_sustained": fl / (ms_per_step * 1e-3) / 1e12 / pk["tc_sustained"]},
                      "loss_sum_last_step": float(loss.detach()), "hbm_bytes_allocated": torch.cuda.max_memory_allocated(dev)}),
          flush=True)


# ------------------------------------------------------------------------------------------------ inference (c5)
def run_inference(args, cfg):
    """`--config c5` (BASELINE.json: long-context inference, forward/generate on 1xB200): tokens/s of the full forward over
    [M, B, T] windows, and of generate() in its growing phase (prompt T/2, K/V cache, one new position per step).
    A parity-test configuration, not the headline line; replicas only for --gpus > 1 (no exchange in inference)."""
    from tradeaid_b200 import lib, synthetic as syn
    from tradeaid_b200.engine import Engine, KVCache
    from tradeaid_b200.params import Dims
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch or cfg["batch"]
    T, V = cfg["block_size"], cfg["vocab"]
    M = len(V)
    dims = Dims(V, cfg["cross"], T, cfg["n_embd"], cfg["n_head"], cfg["n_layer"])
    eng = Engine(dims, dev)
    eng.layout.init_(eng.master, generator=torch.Generator(device=dev).manual_seed(0))
    rows = min(args.rows or cfg["rows"], 1_000_000)
    streams = syn.make_streams(cfg, dev, rows)
    g = torch.Generator().manual_seed(1 + rank)
    ix = torch.randint(0, rows - T - 1, (B,), generator=g).to(dev)
    idx = torch.stack([s[ix[:, None] + torch.arange(T, device=dev)] for s in streams]).contiguous()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    for _ in range(max(args.warmup, 3)):
        eng.forward(idx, None, save=False)
    torch.cuda.synchronize()
    lib.reset_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        logits, _ = eng.forward(idx, None, save=False)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.launch_count()
    fwd_ms = e0.elapsed_time(e1) / args.steps
    fwd_tps = M * B * T / (fwd_ms * 1e-3)
    flops = syn.train_flops_per_step(cfg, B) / 3
    # generate, growing phase: prompt of T/2 positions, then one position per step against the K/V cache
    n0, n_dec = T // 2, min(64, T // 2)
    cache = KVCache(eng, B, T)
    p0, p1 = ev(), ev()
    p0.record()
    eng.forward(idx[:, :, :n0].contiguous(), None, save=False, cache=cache)
    p1.record()
    new = idx[:, :, n0:n0 + 1].contiguous()
    eng.decode_step(new, cache)                               # eager: allocates the one-row buffers
    eng.decode_step(idx[:, :, n0 + 1:n0 + 2].contiguous(), cache)     # captures the step as a CUDA graph and replays it
    steps_in = [idx[:, :, n0 + k:n0 + k + 1].contiguous() for k in range(2, n_dec)]
    d0, d1 = ev(), ev()
    d0.record()
    for new in steps_in:
        eng.decode_step(new, cache)                           # replays (device-side length)
    d1.record()
    torch.cuda.synchronize()
    dec_ms = d0.elapsed_time(d1) / len(steps_in)
    # the same n_dec positions the way the reference generates them: a full forward over the window per new token
    r0, r1 = ev(), ev()
    r0.record()
    for k in range(2):
        eng.forward(idx[:, :, :n0 + n_dec // 2].contiguous(), None, save=False)
    r1.record()
    torch.cuda.synchronize()
    recompute_ms = r0.elapsed_time(r1) / 2
    if rank == 0:
        pk = peaks()
        line = {"metric": "inference tokens/s (forward)", "value": fwd_tps * world, "unit": "tokens/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": fwd_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args, cfg, B, world), "gpu_launches": launches, "clocks": clocks,
                "forward": {"ms": fwd_ms, "algorithmic_tflop": flops / 1e12, "tflops": flops / (fwd_ms * 1e-3) / 1e12,
                            "frac_of_measured_sustained": flops / (fwd_ms * 1e-3) / 1e12 / pk["tc_sustained"]},
                "generate": {"prompt": n0, "prefill_ms": p0.elapsed_time(p1), "decode_ms_per_position": dec_ms,
                             "decode_tokens_per_s": M * B / (dec_ms * 1e-3), "decode_as_cuda_graph": cache._graph is not None,
                             "kv_cache_bytes": cache.bytes(),
                             "kv_cache_GBs_read": cache.bytes() * (n0 + n_dec / 2) / T / (dec_ms * 1e-3) / 1e9,
                             "recompute_ms_per_position": recompute_ms,
                             "speedup_vs_recompute": recompute_ms / dec_ms},
                "hbm_bytes_allocated": torch.cuda.max_memory_allocated(dev)}
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ data preparation (row f-4)
def run_dataprep(args):
    """`--config prep` (not the headline line): the one-shot data preparation of c4's raw columns —
    8 modalities x 50 M float64 rows (--rows overrides) — through tradeaid_b200.data_utils. A "step" prepares every
    modality once: percent changes -> rounding -> 10+10+1 groups for the odd modalities, 2 whole digits / 1 decimal for
    the even ones, then vocabulary + int32 tokens into the [M, N] token store. `value`: rows/s with the raw columns
    resident in HBM (CUDA events); `e2e`: the same through prepare_dataset from pinned HOST columns (H2D inside the timed
    region, vocabularies read back); `roofline`: the radix-select histogram passes (the largest share of the step) against
    the measured copy bandwidth; `cpu_baseline`: the oracle's pure-Python port of the reference chain on a bounded
    sample. Each column (400 MB) exceeds the 126 MB L2. Replicas only for --gpus > 1 (every rank prepares its own data)."""
    import torch.distributed as dist
    from tradeaid_b200 import lib, data_utils as du
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    M, N = 8, args.rows or 50_000_000
    params = [["synthetic.csv", i + 1, None, True, None, 2, 10, None, False, f"pct{i}", 0.1, 2.2] if i % 2 else
              ["synthetic.csv", i + 1, None, False, 2, 1, None, None, False, f"ranged{i}", None, None] for i in range(M)]
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    cols = []
    for i in range(M):
        steps = torch.randn(N, dtype=torch.float64, device=dev, generator=g) * (2.0 / N ** 0.5)
        cols.append((100.0 * (i + 1) * torch.exp(torch.cumsum(steps, 0))).mul(100).round().div(100))
        del steps
    tokens = torch.empty(M, N, dtype=torch.int32, device=dev)

    def step():
        return [du.prepare_modality(cols[i], params[i], tokens_out=tokens[i])[1] for i in range(M)]

    W = max(args.warmup, 3)
    for _ in range(W):
        vocabs = step()
    lib.reset_launch_count()
    step()
    launches_per_step = lib.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = M * N * world / (ms_per_step * 1e-3)

    # roofline of the dominant kernel: the 8 histogram passes of one radix select (4 order statistics = 2 percentiles)
    st = torch.empty(lib.PREP_STATUS_WORDS, dtype=torch.int64, device=dev)
    lib.prep_status_reset(st)
    ranks = torch.tensor([N // 1000, N // 1000 + 1, N - N // 1000 - 2, N - N // 1000 - 1], dtype=torch.int64, device=dev)
    osv = torch.empty(4, dtype=torch.float64, device=dev)
    ws = torch.empty(lib.prep_select_workspace_bytes(), dtype=torch.uint8, device=dev)
    n_sel = 5
    for _ in range(2):
        lib.prep_order_statistics(cols[1], ranks, osv, ws, st)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(n_sel):
        lib.prep_order_statistics(cols[1], ranks, osv, ws, st)
    s1.record()
    torch.cuda.synchronize()
    sel_ms = s0.elapsed_time(s1) / n_sel
    sel_gbs = N * 8 * 8 / (sel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "prep_select_hist_kernel x8 + prep_select_pick_kernel x8 (one radix select)",
                "achieved": sel_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": sel_gbs / pk["hbm"], "traffic": None,
                "algorithmic_bytes": N * 8 * 8, "ms": sel_ms, "peak_source": f"{pk['source']} (MEASURED_PEAKS.json hbm_gbs)",
                "note": "8 B/row per pass x 8 passes; ncu of one pass (profiles/r01_prof_dataprep_v3_ncu_full.txt): DRAM read = "
                        "algorithmic (64.0 MB per 8 M rows); instruction-bound, DESIGN.md section 7"}

    S = min(N, 400_000)
    sample = [cols[i][:S].cpu().tolist() for i in (0, 1)]       # bounded sample for the CPU baseline

    # end to end from pinned host columns through the drop-in API
    e2e = None
    if not args.no_e2e:
        host_cols = [c.cpu().pin_memory() for c in cols]
        del cols

        def e2e_step():
            dcols = [h.to(dev, non_blocking=True) for h in host_cols]
            tr, va, vo = du.prepare_dataset(dcols, params, [N], validation_size=0.1)
            return sum(len(v) for v in vo)

        e2e_step()
        torch.cuda.synchronize()
        n_e2e = max(2, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": M * N * world * n_e2e / dt, "unit": "rows/s", "h2d_bytes_per_step": M * N * 8,
               "d2h_bytes_per_step": 8 * sum(len(v) for v in vocabs), "steps": n_e2e,
               "api": "data_utils.prepare_dataset(pinned host float64 columns -> cudaMemcpyAsync -> K9 kernels -> int32 "
                      "[M, N] token store, vocabularies read back as python lists)"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import dataprep as oracle
        t0 = time.perf_counter()
        for i in (0, 1):
            oracle.prepare_modality(sample[i], params[i])
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": 2 * S / dt, "unit": "rows/s", "cores": 1, "kind": "port",
                        "sample": f"oracle port (oracle/dataprep.py, pure Python like the reference) of both chains on the first "
                                  f"{S} rows of two modalities ({dt:.1f} s)"}
    if rank == 0:
        line = {"metric": "data preparation rows/s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
                "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"prep: raw columns of c4 ({M} modalities x {N} float64 rows), odd modalities percent(2) -> "
                                       "10 groups per sign (0.1 % band, exponent 2.2), even modalities 2 whole digits / 1 decimal, "
                                       "then vocabulary -> int32 token store", "modalities": M, "rows": N,
                           "l2": "every column (8 B/row) is larger than the 126 MB L2", "parallelism": f"replicas x{world}"},
                "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu_baseline, "vocabulary_sizes": [len(v) for v in vocabs],
                "hbm_bytes_allocated": torch.cuda.max_memory_allocated(dev)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    from tradeaid_b200 import synthetic as syn
    if args.config == "prep":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "--config prep is an extra row; its CPU baseline is in the "
                                                                   "cpu_baseline key of the default arm"}), flush=True)
            return
        return run_dataprep(args)
    cfg = dict(syn.CONFIGS[args.config])
    if args.all_cross:
        cfg["cross"] = [True] * len(cfg["vocab"])
        cfg["note"] += " (cross-attention on every modality)"
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.impl == "reference-eager":
        run_reference_eager(args, cfg)
    elif args.config == "c5":
        run_inference(args, cfg)
    elif args.tier == "compat":
        run_compat(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()

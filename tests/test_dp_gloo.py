"""Data-parallel host logic on CPU: world_size 2 over gloo with the emulated kernel namespace. Two ranks, each with half
of the batch, must produce the parameters a single process produces on the whole batch (the loss is a mean over rows, so
grad_global = mean of the per-rank grads — SURVEY.md 8e)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _setup_paths():
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)


def _build(world, pg, overlap=False, grad_dtype="fp32"):
    _setup_paths()
    import emu_ops
    from oracle import restatement as R
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.params import Dims
    from tradeaid_b200.trainer import Trainer
    V, cross, T, C, H, L = [24, 10, 20], [True, False, True], 8, 32, 2, 3
    sd = R.random_state_dict(R.Cfg(T, C, H, L, V, cross), seed=4)
    eng = Engine(Dims(V, cross, T, C, H, L), "cpu", ops=emu_ops, act_dtype=torch.float32)
    nv = eng.layout.named_views(eng.master)
    for k, v in sd.items():
        nv[k].copy_(v)
    tr = Trainer(eng, lr=1e-2, process_group=pg, world_size=world, grad_chunks=3, overlap=overlap, grad_buckets=2,
                 grad_dtype=grad_dtype)
    g = torch.Generator().manual_seed(9)
    idx = torch.stack([torch.randint(0, v, (4, T), generator=g) for v in V])
    tgt = torch.stack([torch.randint(0, v, (4, T), generator=g) for v in V])
    return eng, tr, idx, tgt


def _worker(rank, world, port, out_path, overlap, grad_dtype="fp32"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng, tr, idx, tgt = _build(world, dist.group.WORLD, overlap, grad_dtype)
    assert tr.grad_dtype == grad_dtype and (tr.grad16 is not None) == (grad_dtype == "bf16")
    assert tr.overlap == overlap and (len(tr._buckets) == 2 if overlap else not tr._buckets)
    half = idx.shape[1] // world
    sl = slice(rank * half, (rank + 1) * half)
    for _ in range(3):
        tr.step(idx[:, sl].contiguous(), tgt[:, sl].contiguous())
    if rank == 0:
        torch.save(eng.master.clone(), out_path)
    gathered = [torch.zeros_like(eng.master) for _ in range(world)]
    dist.all_gather(gathered, eng.master)
    assert all(torch.equal(gathered[0], t) for t in gathered), "replicas diverged"
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("overlap", [False, True], ids=["whole-buffer", "bucketed-overlap"])
def test_two_rank_dp_equals_single_process(tmp_path, overlap):
    out = str(tmp_path / "dp_master.pt")
    port = 29500 + (os.getpid() + 7 * overlap) % 2000
    mp.spawn(_worker, args=(2, port, out, overlap), nprocs=2, join=True)
    dp_master = torch.load(out)
    eng, tr, idx, tgt = _build(1, None)
    for _ in range(3):
        tr.step(idx, tgt)
    assert (dp_master - eng.master).abs().max().item() <= 2e-5   # fp32 summation order (mean of halves vs whole) through 3 Adam steps of lr 1e-2


@pytest.mark.parametrize("overlap", [False, True], ids=["whole-buffer", "bucketed-overlap"])
def test_two_rank_dp_with_bf16_gradient_exchange(tmp_path, overlap):
    """grad_dtype="bf16": the buckets travel as bf16 copies and AdamW consumes them as bf16 - the replicas stay identical and the
    parameters follow the single-process run within the bf16 rounding of the gradients (Adam normalises the step: lr-sized)."""
    out = str(tmp_path / "dp_master_bf16.pt")
    port = 31500 + (os.getpid() + 7 * overlap) % 2000
    mp.spawn(_worker, args=(2, port, out, overlap, "bf16"), nprocs=2, join=True)
    dp_master = torch.load(out)
    eng, tr, idx, tgt = _build(1, None)
    start = eng.master.clone()
    for _ in range(3):
        tr.step(idx, tgt)
    moved = (eng.master - start).norm().item()
    assert (dp_master - eng.master).norm().item() <= 0.05 * moved


def test_two_rank_dp_with_activation_recompute(tmp_path, monkeypatch):
    """TRADEAID_RECOMPUTE=1 under data parallelism: backward() reruns each block's forward before its backward, the stage hooks
    (bucket all-reduces overlapped with backward) fire at the same points - the two-rank run still follows the single-process
    saving run to fp32 summation order."""
    out = str(tmp_path / "dp_master_rc.pt")
    port = 33500 + os.getpid() % 2000
    monkeypatch.setenv("TRADEAID_RECOMPUTE", "1")
    mp.spawn(_worker, args=(2, port, out, True), nprocs=2, join=True)
    monkeypatch.delenv("TRADEAID_RECOMPUTE")
    dp_master = torch.load(out)
    eng, tr, idx, tgt = _build(1, None)
    assert not eng.recompute
    for _ in range(3):
        tr.step(idx, tgt)
    assert (dp_master - eng.master).abs().max().item() <= 2e-5


def test_gradient_buckets_partition_the_active_parameters():
    """Every element of every active parameter kind is in exactly one bucket (trainer.py: _make_buckets)."""
    _setup_paths()
    import emu_ops
    from tradeaid_b200.engine import Engine
    from tradeaid_b200.params import Dims
    from tradeaid_b200.trainer import Trainer
    for cross, L, nb in (([True, False, True], 5, 3), ([False, False], 2, 6), ([True, True, True, False], 12, 6)):
        V = [24, 10, 20, 8][:len(cross)]
        eng = Engine(Dims(V, cross, 8, 32, 2, L), "cpu", ops=emu_ops, act_dtype=torch.float32)
        tr = Trainer(eng, lr=1e-3, world_size=2, overlap=True, grad_buckets=nb)
        assert [b[0] for b in tr._buckets][-1] == "embed" and len(tr._buckets) == min(nb, L)
        cover = torch.zeros(eng.layout.total, dtype=torch.int32)
        base = eng.grad.data_ptr()
        for _, views, _ranges in tr._buckets:
            for v in views:
                a = (v.data_ptr() - base) // 4
                cover[a:a + v.numel()] += 1
        want = torch.zeros_like(cover)
        for kind, (off, shape) in eng.layout.kinds.items():
            n = 1
            for s_ in shape:
                n *= s_
            if off < eng.layout.n_active:
                want[off:off + n] = 1
        assert torch.equal(cover, want)


# ------------------------------------------------------------------------------------------------ data-parallel evaluation
class _OracleModel:
    """Stand-in for the nn.Module in training_utils.m: the oracle forward on CPU (estimate_loss only calls it, eval(), train())."""

    def __init__(self):
        from oracle import restatement as R
        self.R = R
        self.cfg = R.Cfg(8, 32, 2, 2, [24, 10, 20], [True, False, True])
        self.sd = R.random_state_dict(self.cfg, seed=5)
        self.training = True

    def eval(self):
        self.training = False

    def train(self):
        self.training = True

    def __call__(self, xb, yb):
        return self.R.forward(self.sd, self.cfg, xb, yb)


def _eval_setup(pg):
    _setup_paths()
    import emu_ops
    from tradeaid_b200 import training_utils as tu
    gen = torch.Generator().manual_seed(31)
    V = [24, 10, 20]
    full = [torch.randint(0, v, (300,), generator=gen) for v in V]
    tu.configure(device="cpu", batch_size=4, block_size=8, eval_iters=5)
    tu._ops_override = emu_ops
    tu._stores.clear()
    tu.all_train_sets = [d[:-60].tolist() for d in full]
    tu.all_val_sets = [d[-60:].clone() for d in full]
    tu.all_vocabularies = [[0.5 * k - 4.0 for k in range(V[0])], [float(k - 2) for k in range(V[1])], [f"s{k}" for k in range(V[2])]]
    tu.all_modality_params = [["s", i, None, i == 1, None, None, None, None, False, f"m{i}"] for i in range(3)]
    tu.file_lengths = [300]
    tu.num_modalities = 3
    tu.is_percents = False
    tu.m = _OracleModel()
    tu.eval_process_group = pg
    torch.manual_seed(77)            # every rank: the same host index stream (parity mode)
    return tu


def _eval_worker(rank, world, port, out_dir):
    import contextlib
    import io
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tu = _eval_setup(dist.group.WORLD)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = tu.estimate_loss(3, 10)
    after = int(torch.randint(0, 2 ** 31 - 1, (1,)))          # position of the host RNG stream after the evaluation
    torch.save({"out": out, "printed": buf.getvalue(), "rng_after": after, "training": tu.m.training},
               os.path.join(out_dir, f"eval_{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_evaluation_equals_single_process(tmp_path, capsys):
    """estimate_loss with eval_process_group: iterations dealt round-robin, one all-reduce of the totals — same directional
    counts, same mean loss (float64 sum order), same host RNG position as the single-process evaluation; rank 0 alone prints."""
    port = 31500 + os.getpid() % 2000
    mp.spawn(_eval_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = [torch.load(str(tmp_path / f"eval_{r}.pt")) for r in range(2)]
    tu = _eval_setup(None)
    try:
        want = tu.estimate_loss(3, 10)
        want_after = int(torch.randint(0, 2 ** 31 - 1, (1,)))
        printed = capsys.readouterr().out
    finally:
        tu._ops_override = None
        tu._stores.clear()
        tu.m = None
        tu.eval_process_group = None
    strip = lambda text: [ln for ln in text.splitlines() if not ln.startswith("Evaluation:")]     # (the header has a clock)
    assert strip(got[0]["printed"]) == strip(printed) and "DIRECTIONAL METRICS - Val Set" in printed
    assert got[1]["printed"] == ""
    for r in range(2):
        assert got[r]["rng_after"] == want_after and got[r]["training"]
        for split in ("train", "val"):
            assert abs(got[r]["out"][split] - want[split]) <= 1e-12 * abs(want[split])
    assert got[0]["out"] == got[1]["out"]


# ------------------------------------------------------------------------------------------------ the pieces together
class _EngineModel:
    """training_utils.m for the engine-owned tier: forward through the engine (evaluation never needs backward)."""

    def __init__(self, eng):
        self.eng, self.training = eng, True

    def eval(self):
        self.training = False

    def train(self):
        self.training = True

    def __call__(self, xb, yb):
        logits, losses = self.eng.forward(torch.stack(xb), torch.stack(yb), save=False)
        return logits, [losses[i] for i in range(len(xb))]


def _parity_run(world, pg, rank):
    """Three training steps + one evaluation in data-parallel PARITY mode: get_batch with batch_shard (every rank draws the
    global batch's indices, gathers its rows), Trainer with the gradient exchange, estimate_loss over eval_process_group."""
    eng, tr, _, _ = _build(world, pg)
    tu = _eval_setup(pg if world > 1 else None)
    tu.configure(device="cpu", batch_size=4, block_size=8, eval_iters=4)
    tu.m = _EngineModel(eng)
    tu.batch_shard = (rank, world) if world > 1 else None
    torch.manual_seed(5)
    try:
        for _ in range(3):
            xb, yb = tu.get_batch("train", 0)
            tr.step(torch.stack(xb), torch.stack(yb))
        # batch_shard AND eval_process_group both stay set (the natural parity set-up): estimate_loss itself draws whole
        # evaluation batches and deals the iterations to the ranks (it used to evaluate 1/world of each batch)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            out = tu.estimate_loss()
    finally:
        tu.batch_shard = None
        tu.eval_process_group = None
        tu._ops_override = None
        tu._stores.clear()
        tu.m = None
    return eng.master.clone(), out


def _parity_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    master, out = _parity_run(world, dist.group.WORLD, rank)
    torch.save({"master": master, "out": out}, os.path.join(out_dir, f"parity_{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_parity_mode_training_and_evaluation_equal_single_process(tmp_path):
    """SURVEY.md 8e end to end on CPU: fixed global batch, sharded get_batch, gradient exchange, sharded evaluation - the
    parameters after three steps and the evaluation losses equal the single-process run on the same index stream."""
    port = 33500 + os.getpid() % 2000
    mp.spawn(_parity_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = [torch.load(str(tmp_path / f"parity_{r}.pt")) for r in range(2)]
    master, out = _parity_run(1, None, 0)
    assert torch.equal(got[0]["master"], got[1]["master"])
    assert (got[0]["master"] - master).abs().max().item() <= 2e-5
    for split in ("train", "val"):
        assert got[0]["out"][split] == got[1]["out"][split]
        assert abs(got[0]["out"][split] - out[split]) <= 1e-5 * abs(out[split])

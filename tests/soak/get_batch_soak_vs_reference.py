"""Offline differential soak (BUILD CONTAINER ONLY: imports /root/reference): training_utils.get_batch of the reference against
the product's (resident token store + gather + the [2]-quirk augmentation, over the CPU emulation of the kernels) through the
module globals main.py sets - random modalities, file splits, is_percents, augmentation sizes at index 2 (None / True / 2 / 3),
and a random sequence of (split, is_training) calls. Every call: identical xb / yb (int64), identical position of torch's CPU
generator and identical state of Python's `random`; at the end the reference's cumulatively modified training lists equal the
product's resident store.    usage: cd /tmp && python /root/repo/tests/soak/get_batch_soak_vs_reference.py [trials] [seed]"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, "/root/reference")
os.chdir("/tmp")
import config_utils  # noqa: E402
import emu_ops  # noqa: E402
import training_utils as ref  # noqa: E402  (the reference's)
from tradeaid_b200 import training_utils as ours  # noqa: E402

ours._ops_override = emu_ops
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
calls_total = 0
for trial in range(trials):
    M = rng.choice([1, 2, 3])
    T, B = rng.choice([2, 4, 8, 16]), rng.choice([1, 3, 8])
    lens = [rng.randint(40, 200) for _ in range(rng.choice([1, 1, 2, 4]))]
    N = sum(lens)
    n_val = rng.randint(T + 3, max(T + 4, N // 3))
    V = [rng.choice([3, 5, 20, 57, 300]) for _ in range(M)]
    g = torch.Generator().manual_seed(rng.randint(0, 10 ** 6))
    full = [torch.randint(0, v, (N,), generator=g) for v in V]
    params = [["f", i + 1, rng.choice([None, None, True, 2, 3]), False, None, None, None, None, False, f"m{i}"] for i in range(M)]
    is_pct = rng.random() < 0.4
    vocabs = [[float(k) for k in range(v)] for v in V]
    config_utils._config_cache = dict(device="cpu", batch_size=B, block_size=T, eval_iters=2, n_embd=32, n_head=4, n_layer=1,
                                      dropout=0.0, fixed_values=[], output_file_name="", project_file_path="/tmp/")
    ours.configure(device="cpu", batch_size=B, block_size=T, eval_iters=2)
    ours._stores.clear()
    for mod, copy in ((ref, True), (ours, True)):
        mod.all_train_sets = [d[:-n_val].tolist() for d in full]
        mod.all_val_sets = [d[-n_val:].clone() for d in full]
        mod.all_vocabularies = vocabs
        mod.all_modality_params = params
        mod.file_lengths = lens
        mod.num_modalities = M
        mod.is_percents = is_pct
    seq = [(rng.choice(["train", "val"]), rng.choice([0, 1, 1])) for _ in range(rng.randint(1, 5))]
    seed_t, seed_p = rng.randint(0, 10 ** 6), rng.randint(0, 10 ** 6)
    outs = []
    for mod in (ref, ours):
        torch.manual_seed(seed_t); random.seed(seed_p)
        rec = []
        for split, tr in seq:
            try:
                xb, yb = mod.get_batch(split, tr)
                rec.append(("ok", [t.clone() for t in xb], [t.clone() for t in yb], int(torch.randint(0, 2 ** 31 - 1, (1,))),
                            random.getstate()))
            except (ValueError, TypeError, RuntimeError) as e:
                rec.append(("err", type(e).__name__, str(e)))
                break
        outs.append(rec)
    assert len(outs[0]) == len(outs[1]), (trial, seq, [r[0] for r in outs[0]], [r[0] for r in outs[1]])
    for a, b in zip(*outs):
        assert a[0] == b[0], (trial, seq, a[:3] if a[0] == "err" else a[0], b[:3] if b[0] == "err" else b[0])
        if a[0] == "err":
            assert a[1:] == b[1:], (trial, a, b)
            continue
        for s, t in zip(a[1] + a[2], b[1] + b[2]):
            assert s.dtype == t.dtype == torch.int64 and torch.equal(s, t), (trial, seq)
        assert a[3] == b[3] and a[4] == b[4], (trial, "generator position")
        calls_total += 1
    if all(r[0] == "ok" for r in outs[0]):
        store = ours._stores["train"][1].store
        for i in range(M):
            assert store[i].tolist() == ref.all_train_sets[i], (trial, "resident store vs the reference's modified lists", i)
print(f"get_batch soak: {trials} random set-ups, {calls_total} calls identical to the reference (batches, torch + random generator state, "
      f"cumulative augmentation)")

"""Offline differential soak (BUILD CONTAINER ONLY: imports /root/reference): generate_batch_starting_indices of the product
feeder, of the oracle and of the reference itself on 3000 random configurations (1-8 files, train/val, is_percents, block and
batch sizes, data sizes down to the error cases) - identical indices, identical position of the torch CPU generator
afterwards, identical exception types and messages. Round 1: 2824 ok + 176 error cases, all identical.
usage: cd /tmp && python /root/repo/tests/soak/index_soak_vs_reference.py"""
import os
import sys, random, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); sys.path.insert(0, "/root/reference")
import config_utils
config_utils._config_cache = dict(device="cpu", batch_size=4, block_size=8, eval_iters=2, n_embd=32, n_head=4, n_layer=1, dropout=0.0,
                                  fixed_values=[], output_file_name="", project_file_path="/tmp/")
import training_utils as ref
from tradeaid_b200.feeder import generate_batch_starting_indices as ours
from oracle import restatement as R
rng = random.Random(5)
n_ok = n_err = 0
for trial in range(3000):
    nf = rng.choice([1, 1, 2, 3, 5, 8])
    lens = [rng.randint(5, 400) for _ in range(nf)]
    total = sum(lens)
    split = rng.choice(["train", "val"])
    frac = rng.uniform(0.05, 0.95)
    data_size = max(2, int(total * (frac if split == "train" else 1 - frac)))
    T = rng.choice([2, 4, 8, 16, 64])
    B = rng.choice([1, 3, 16])
    pct = rng.random() < 0.5
    seed = rng.randint(0, 10**6)
    res = []
    for fn in (ref.generate_batch_starting_indices, ours, lambda *a: R.batch_starting_indices(*a)):
        torch.manual_seed(seed)
        try:
            out = fn(data_size, T, B, split, lens, pct)
            out = torch.as_tensor(out).tolist() if not isinstance(out, list) else [int(v) for v in out]
            res.append(("ok", out, int(torch.randint(0, 2**31 - 1, (1,)))))
        except (ValueError, TypeError, RuntimeError) as e:
            res.append(("err", type(e).__name__, str(e)))
    assert res[0] == res[1] == res[2], (trial, data_size, T, B, split, lens, pct, res)
    n_ok += res[0][0] == "ok"; n_err += res[0][0] == "err"
print("index soak:", n_ok, "ok", n_err, "errors, all identical to the reference (values, RNG position, messages)")

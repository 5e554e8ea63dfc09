"""Offline differential soak (BUILD CONTAINER ONLY: imports /root/reference): calculate_evaluation_metrics of the reference
against the product's (tradeaid_b200.training_utils over the CPU emulation of ta_directional_metrics) on random batches -
numeric / percentage / non-numeric vocabularies, duplicate and signed-zero vocabulary values, ties in the logits, batch
sizes 1..40, sequence lengths 1..6. Wins, losses and processed flags must be identical, the certainty sums equal to
float32-softmax round-off.      usage: cd /tmp && python /root/repo/tests/soak/metrics_soak_vs_reference.py [trials] [seed]"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, "/root/reference")
import config_utils  # noqa: E402
config_utils._config_cache = dict(device="cpu", batch_size=4, block_size=8, eval_iters=2, n_embd=32, n_head=4, n_layer=1,
                                  dropout=0.0, fixed_values=[], output_file_name="", project_file_path="/tmp/")
import training_utils as ref  # noqa: E402  (the reference's)
import emu_ops  # noqa: E402
from tradeaid_b200 import training_utils as ours  # noqa: E402

ours._ops_override = emu_ops
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
g = torch.Generator().manual_seed(rng.randint(0, 10 ** 6))
worst = 0.0
for trial in range(trials):
    M = rng.choice([1, 2, 3])
    B, T = rng.choice([1, 2, 5, 16, 40]), rng.choice([1, 2, 3, 6])
    vocabs, params, logits, xb, yb = [], [], [], [], []
    for i in range(M):
        V = rng.choice([2, 3, 7, 33, 120])
        kind = rng.choice(["float", "float", "pct", "ints", "strings", "dups"])
        if kind == "strings":
            vocab = [f"s{k}" for k in range(V)]
        elif kind == "ints":
            vocab = sorted(rng.sample(range(-200, 200), V))
        elif kind == "dups":
            vocab = sorted(rng.choice([-1.5, -0.0, 0.0, 0.5, 2.0]) for _ in range(V))
        else:
            vocab = sorted({round(rng.uniform(-5, 5), 2) for _ in range(V * 3)})[:V]
            while len(vocab) < V:
                vocab.append(vocab[-1] + 0.01)
        vocabs.append(vocab)
        params.append(["f", i + 1, None, kind == "pct" or rng.random() < 0.2, None, None, None, None, False, f"m{i}"])
        lg = torch.randn(B, T, V, generator=g)
        if rng.random() < 0.3:
            lg = (lg * 2).round() / 2          # exact ties between logits: arg-max must pick the lowest index, like torch
        logits.append(lg)
        xb.append(torch.randint(0, V, (B, T), generator=g))
        yb.append(torch.randint(0, V, (B, T), generator=g))
    want = ref.calculate_evaluation_metrics(logits, xb, yb, M, vocabs, params, [])
    got = ours.calculate_evaluation_metrics(logits, xb, yb, M, vocabs, params, [])
    assert list(want[0]) == list(got[0]) and list(want[1]) == list(got[1]) and list(want[3]) == list(got[3]), (trial, want, got)
    for a, b in zip(want[2], got[2]):
        worst = max(worst, abs(a - b))
        assert abs(a - b) <= 2e-5 * max(1.0, abs(a)), (trial, a, b)
print(f"metrics soak: {trials} random batches identical to the reference (wins, losses, processed); certainty within {worst:.2e}")

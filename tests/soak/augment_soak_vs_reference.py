"""Offline differential soak (BUILD CONTAINER ONLY: imports /root/reference): data_utils.add_rand_to_data_points of the
reference against TokenStore.augment_modality (host logic: bulk pre-draw of Mersenne-Twister words, re-positioning of
Python's `random` generator) over the CPU emulation of ta_augment_tokens, on random token rows, vocabulary sizes and
rand_size 1..3, several cumulative calls in a row: identical tokens AND identical state of the `random` module afterwards.
usage: cd /tmp && python /root/repo/tests/soak/augment_soak_vs_reference.py [trials] [seed]"""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, "/root/reference")
import data_utils as ref  # noqa: E402  (the reference's)
import emu_ops  # noqa: E402
from tradeaid_b200.feeder import TokenStore  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
for trial in range(trials):
    n = rng.choice([1, 2, 17, 300, 5000])
    V = rng.choice([2, 3, 5, 8, 57, 1024])
    s = rng.choice([1, 2, 3])
    g = torch.Generator().manual_seed(rng.randint(0, 10 ** 6))
    rows = [torch.randint(0, V, (n,), generator=g), torch.randint(0, V, (n,), generator=g)]
    store = TokenStore([r.clone() for r in rows], "cpu", ops=emu_ops)
    lists = [r.tolist() for r in rows]
    seed = rng.randint(0, 10 ** 9)
    random.seed(seed)
    for call in range(3):                          # the quirk applies the walk cumulatively, call after call
        for r in range(2):
            lists[r] = ref.add_rand_to_data_points(lists[r], s, V)
    want_state = random.getstate()
    random.seed(seed)
    for call in range(3):
        for r in range(2):
            store.augment_modality(r, s, V)
    assert random.getstate() == want_state, (trial, n, V, s)
    for r in range(2):
        assert store.store[r].tolist() == lists[r], (trial, n, V, s, r)
print(f"augmentation soak: {trials} random stores x 3 cumulative calls x 2 modalities identical to the reference (tokens and random state)")

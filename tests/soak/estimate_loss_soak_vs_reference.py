"""Offline differential soak (BUILD CONTAINER ONLY: imports /root/reference): training_utils.estimate_loss of the reference
against the product's on random set-ups - the SAME reference nn.Module as `m` on both sides (CPU), the product's feeder and
metrics over the CPU emulation of the kernels. Returned dict (train / val mean loss), every printed line except the clock in
the 'Evaluation:' header, the lines appended to the output file, the model's train/eval mode afterwards and the position of
torch's generator must be identical.
usage: cd /tmp && python /root/repo/tests/soak/estimate_loss_soak_vs_reference.py [trials] [seed]"""
import contextlib
import io
import os
import random
import re
import shutil
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, "/root/reference")
os.chdir("/tmp")
import config_utils  # noqa: E402
import emu_ops  # noqa: E402
import training_utils as ref  # noqa: E402  (the reference's)
from tradeaid_b200 import training_utils as ours  # noqa: E402

ours._ops_override = emu_ops
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
work = tempfile.mkdtemp(prefix="eval_soak_")
os.makedirs(os.path.join(work, "output"))
for trial in range(trials):
    M = rng.choice([1, 2, 3])
    T, B, C, H, L = rng.choice([4, 8]), rng.choice([1, 3, 6]), 32, rng.choice([2, 4]), 1
    iters = rng.choice([1, 2, 3])
    V = [rng.choice([3, 7, 30, 57]) for _ in range(M)]
    cross = [rng.random() < 0.5 for _ in range(M)]
    pct = [rng.random() < 0.4 for _ in range(M)]
    names = [rng.choice([f"mod{i}", "", None]) for i in range(M)]
    N, n_val = 160, 50
    lens = rng.choice([[160], [100, 60]])
    g = torch.Generator().manual_seed(rng.randint(0, 10 ** 6))
    full = [torch.randint(0, v, (N,), generator=g) for v in V]
    vocabs = []
    for i, v in enumerate(V):
        kind = rng.choice(["float", "float", "string", "int"])
        vocabs.append([f"s{k}" for k in range(v)] if kind == "string" else
                      ([k - v // 2 for k in range(v)] if kind == "int" else [round(0.37 * k - 2.0, 2) for k in range(v)]))
    params = [["f", i + 1, None, pct[i], None, None, None, None, cross[i], names[i]] for i in range(M)]
    fname = rng.choice(["", f"log_{trial}.txt"])
    base = dict(device="cpu", batch_size=B, block_size=T, eval_iters=iters, n_embd=C, n_head=H, n_layer=L, dropout=0.0,
                fixed_values=[], project_file_path=work + "/")
    sys.modules.pop("model", None)
    config_utils._config_cache = dict(base, output_file_name=fname)
    import model as ref_model
    torch.manual_seed(trial)
    net = ref_model.MultimodalTransformer(M, V, params)
    results = []
    for mod, tag in ((ref, "ref"), (ours, "ours")):
        out_name = f"{tag}_{fname}" if fname else ""
        config_utils._config_cache = dict(base, output_file_name=out_name)
        ours.configure(**dict(base, output_file_name=out_name))
        ours._stores.clear()
        mod.all_train_sets = [d[:-n_val].tolist() for d in full]
        mod.all_val_sets = [d[-n_val:].clone() for d in full]
        mod.all_vocabularies = vocabs
        mod.all_modality_params = params
        mod.all_file_info = []
        mod.file_lengths = lens
        mod.num_modalities = M
        mod.is_percents = any(pct)
        mod.m = net
        net.train()
        torch.manual_seed(1000 + trial)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            out = mod.estimate_loss(trial, 99) if trial % 2 else mod.estimate_loss()
        text = re.sub(r"\| \d\d:\d\d:\d\d", "| <clock>", buf.getvalue())
        log = open(os.path.join(work, "output", out_name)).read() if out_name else None
        results.append((out, text, log, net.training, int(torch.randint(0, 2 ** 31 - 1, (1,)))))
    a, b = results
    for split in ("train", "val"):
        assert abs(a[0][split] - b[0][split]) <= 1e-6 * abs(a[0][split]), (trial, a[0], b[0])
    assert a[1] == b[1], (trial, a[1], b[1])
    assert a[2] == b[2], (trial, a[2], b[2])
    assert a[3] == b[3] is True and a[4] == b[4], (trial, a[3:], b[3:])
shutil.rmtree(work)
print(f"estimate_loss soak: {trials} random set-ups identical to the reference (losses, console lines, log-file lines, mode, generator)")

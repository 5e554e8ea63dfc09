"""Offline differential soak (BUILD CONTAINER ONLY: imports /root/reference): the drop-in nn.Module surface of
tradeaid_b200.model.MultimodalTransformer against model.MultimodalTransformer on random configurations - state_dict keys in
order (tril buffers included), shapes and dtypes; named_parameters order (what torch.optim.AdamW and optimizer checkpoints
index by); parameter count; strict load_state_dict of a reference checkpoint and the round trip back into the reference;
torch.save / torch.load(weights_only=True) of the state dict (main.py:470, 635).
usage: cd /tmp && python /root/repo/tests/soak/module_soak_vs_reference.py [trials] [seed]"""
import io
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, "/root/reference")
os.chdir("/tmp")
import config_utils  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
for trial in range(trials):
    M = rng.choice([1, 2, 3, 4, 6])
    H = rng.choice([1, 2, 3, 4, 8])
    hs = rng.choice([8, 16, 32, 64])
    if (H * hs) % 16:
        H += 1
    C, L, T = H * hs, rng.choice([1, 2, 4]), rng.choice([2, 8, 32])
    V = [rng.choice([2, 3, 5, 57, 300]) for _ in range(M)]
    params = [["f", i + 1, None, False, None, None, None, None, rng.choice([True, False, None, 1, 0]), f"m{i}", None, None]
              for i in range(M)]
    config_utils._config_cache = dict(device="cpu", batch_size=2, block_size=T, eval_iters=2, n_embd=C, n_head=H, n_layer=L,
                                      dropout=0.0, fixed_values=[], output_file_name="", project_file_path="/tmp/")
    for name in ("model", "tradeaid_b200.model"):
        sys.modules.pop(name, None)
    import model as ref_model
    from tradeaid_b200 import model as our_model
    torch.manual_seed(trial)
    r = ref_model.MultimodalTransformer(M, V, params)
    o = our_model.MultimodalTransformer(M, V, params)
    rsd, osd = r.state_dict(), o.state_dict()
    assert list(rsd.keys()) == list(osd.keys()), (trial, set(rsd) ^ set(osd))
    for k in rsd:
        assert rsd[k].shape == osd[k].shape and rsd[k].dtype == osd[k].dtype, (trial, k, rsd[k].shape, osd[k].shape)
    assert [n for n, _ in r.named_parameters()] == [n for n, _ in o.named_parameters()], trial
    assert sum(p.numel() for p in r.parameters()) == sum(p.numel() for p in o.parameters()), trial
    buf = io.BytesIO()
    torch.save(rsd, buf)
    buf.seek(0)
    o.load_state_dict(torch.load(buf, weights_only=True), strict=True)          # a reference checkpoint into the engine's module
    for k, v in o.state_dict().items():
        assert torch.equal(v, rsd[k]), (trial, k)
    buf = io.BytesIO()
    torch.save(o.state_dict(), buf)
    buf.seek(0)
    r2 = ref_model.MultimodalTransformer(M, V, params)
    r2.load_state_dict(torch.load(buf, weights_only=True), strict=True)         # and the engine's checkpoint into the reference
    for k, v in r2.state_dict().items():
        assert torch.equal(v, rsd[k]), (trial, k)
print(f"module soak: {trials} random configurations - state dict keys / order / shapes, parameter order and count, strict checkpoint "
      f"round trips in both directions identical to the reference")

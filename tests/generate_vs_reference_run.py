"""`generate` token by token against the reference's own model.py on the same device (test infrastructure).
The reference model (oracle/_ref/model.py:404-446, unmodified) and the drop-in are loaded with the same reference-made
weights (tests/golden/model_*.pt), seeded identically before every call, and asked for the same continuations — growing
phase (K/V cache on our side) and past block_size (sliding window on both). torch.multinomial draws from the same CUDA
generator stream on both sides, so the tokens are identical unless a probability difference moves a sample across a
bin edge: in the bf16x3 mode that needs a ~1e-5 coincidence, in the bf16 mode ~1e-3 per sample.
Prints one JSON line: per precision the fraction of identical tokens and the max probability difference of the first step."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "oracle", "_ref")
dev = sys.argv[2] if len(sys.argv) > 2 else "cuda"
emu = len(sys.argv) > 3 and sys.argv[3] == "emu"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
import config_utils  # noqa: E402  (the reference's)

out = {"device": dev, "cases": []}
for name, prompt, new in (("m3", 3, 9), ("hs32", 2, 5), ("demo", 2, 6)):
    g = torch.load(os.path.join(ROOT, "tests", "golden", f"model_{name}.pt"), weights_only=False)
    c = g["cfg"]
    config_utils._config_cache = dict(device=dev, batch_size=g["idx"][0].shape[0], block_size=c["T"], n_embd=c["C"], n_head=c["H"],
                                      n_layer=c["L"], dropout=0.0, eval_iters=1, fixed_values=[], output_file_name="",
                                      project_file_path="./")
    for mod in ("model", "tradeaid_b200.model"):
        sys.modules.pop(mod, None)
    import model as refmodel  # noqa: E402
    M = len(c["V"])
    params = [["p", 1, None, False, None, None, None, None, bool(c["cross"][i]), f"m{i}", None, None] for i in range(M)]
    ref = refmodel.MultimodalTransformer(M, list(c["V"]), params)
    ref.load_state_dict(g["state_dict"], strict=False)
    ref = ref.to(dev).eval()
    from tradeaid_b200 import model as ours_model  # noqa: E402
    if emu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import emu_ops
        ours_model._ops_override = emu_ops
    idx = [t[:, :prompt].to(dev) for t in g["idx"]]
    for precision in ("bf16", "bf16x3"):
        ours_model.set_precision(precision)
        ours = ours_model.MultimodalTransformer(M, list(c["V"]), params)
        ours.load_state_dict(g["state_dict"], strict=False)
        ours = ours.to(dev).eval()
        same = total = 0
        with torch.no_grad():
            lr, _ = ref(idx)
            lo, _ = ours(idx)
        dp = max(float((torch.softmax(a[:, -1], -1) - torch.softmax(b[:, -1], -1)).abs().max()) for a, b in zip(lr, lo))
        for mg in range(M):
            torch.manual_seed(100 + mg)
            a = ref.generate([t.clone() for t in idx], max_new_tokens=new, modality_to_generate=mg)
            torch.manual_seed(100 + mg)
            b = ours.generate([t.clone() for t in idx], max_new_tokens=new, modality_to_generate=mg)
            assert [x.shape for x in a] == [x.shape for x in b], (name, mg)
            for i in range(M):
                same += int((a[i] == b[i]).sum())
                total += a[i].numel()
        out["cases"].append({"model": name, "precision": precision, "tokens": total, "identical": same,
                             "first_step_max_prob_diff": dp})
print(json.dumps(out))

"""Offline probe (BUILD CONTAINER ONLY: imports /root/reference): autograd semantics of the drop-in module against the reference
module - gradient accumulation over two backward passes, zero_grad(set_to_none=False), a weighted sum of the per-modality
losses, a single modality's loss. Relative gradient differences must be float32 round-off (round 1: 2e-7 each, after the
accumulation case - 0.72 before - was fixed in model._Bridge.backward).
usage: cd /tmp && python /root/repo/tests/soak/autograd_semantics_vs_reference.py"""
import os
import sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, "/root/reference")
import config_utils
config_utils._config_cache = dict(device="cpu", batch_size=2, block_size=4, eval_iters=2, n_embd=32, n_head=4, n_layer=1, dropout=0.0,
                                  fixed_values=[], output_file_name="", project_file_path="/tmp/")
import emu_ops, model as ref_model
from tradeaid_b200 import model as our_model
our_model._ops_override = emu_ops
params = [["f", i + 1, None, False, None, None, None, None, True, f"m{i}", None, None] for i in range(2)]
torch.manual_seed(0)
r = ref_model.MultimodalTransformer(2, [20, 7], params)
o = our_model.MultimodalTransformer(2, [20, 7], params); o.load_state_dict(r.state_dict()); o = o.to("cpu")
g = torch.Generator().manual_seed(1)
batches = [([torch.randint(0, v, (2, 4), generator=g) for v in (20, 7)], [torch.randint(0, v, (2, 4), generator=g) for v in (20, 7)]) for _ in range(2)]
for name, mod in (("ref", r), ("ours", o)):
    mod.zero_grad(set_to_none=True)
    for xb, yb in batches:                      # two backward passes, no zero_grad in between: torch accumulates
        _, ls = mod(xb, yb); sum(ls).backward()
k = "blocks.0.ffwd_layers.0.net.0.weight"
gr, go = dict(r.named_parameters())[k].grad, dict(o.named_parameters())[k].grad
print("accumulation: rel diff", float((gr - go).norm() / gr.norm()))
# zero_grad(set_to_none=False) then one backward
for name, mod in (("ref", r), ("ours", o)):
    mod.zero_grad(set_to_none=False)
    _, ls = mod(*batches[0]); sum(ls).backward()
gr, go = dict(r.named_parameters())[k].grad, dict(o.named_parameters())[k].grad
print("zero_grad(set_to_none=False) + backward: rel diff", float((gr - go).norm() / gr.norm()))
# weighted losses: (2*l0 + 0.5*l1).backward()
for name, mod in (("ref", r), ("ours", o)):
    mod.zero_grad(set_to_none=True)
    _, ls = mod(*batches[1]); (2 * ls[0] + 0.5 * ls[1]).backward()
gr, go = dict(r.named_parameters())[k].grad, dict(o.named_parameters())[k].grad
print("weighted sum of losses: rel diff", float((gr - go).norm() / gr.norm()))
# only one modality's loss
for name, mod in (("ref", r), ("ours", o)):
    mod.zero_grad(set_to_none=True)
    _, ls = mod(*batches[1]); ls[1].backward()
gr, go = dict(r.named_parameters())[k].grad, dict(o.named_parameters())[k].grad
print("single loss: rel diff", float((gr - go).norm() / (gr.norm() + 1e-12)), float(gr.norm()), float(go.norm()))

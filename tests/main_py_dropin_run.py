"""Whole-program drop-in check on a real device (test infrastructure): the reference's own `main.py`
(/root/reference/main.py:598-653 is the loop) on its demo — BASELINE.json configs[0] — twice in a scratch copy of oracle/_ref:

  (1) unmodified, on `--device` (torch eager, fp32);
  (2) with the two shim modules of INTEGRATION.md ahead on sys.path (model.py -> tradeaid_b200.model, training_utils.py ->
      tradeaid_b200.training_utils): every forward / backward / get_batch / evaluation metric through libtradeaid_b200.so.

Both start from the same seeded checkpoint (create_new_model: 0), generators re-seeded after model construction, dropout 0
(torch's dropout stream cannot be reproduced outside torch), save_model: 1 so that main.py writes checkpoints. Compared:
console text with numbers masked (identical), every printed / logged loss (abs tolerance), directional-metric counts
(tolerance in samples), checkpoint keys / shapes (identical) and values (normwise). Prints one JSON line; exit code 0 = pass.

usage: python tests/main_py_dropin_run.py --ref oracle/_ref --device cuda [--precision bf16|bf16x3] [--loss-tol 1e-2]"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ap = argparse.ArgumentParser()
ap.add_argument("--ref", default=os.path.join(ROOT, "oracle", "_ref"))
ap.add_argument("--device", default="cuda")
ap.add_argument("--precision", default="bf16")
ap.add_argument("--loss-tol", type=float, default=1e-2)
ap.add_argument("--count-tol", type=int, default=8)
ap.add_argument("--weight-tol", type=float, default=2e-2)
ap.add_argument("--emu", action="store_true", help="CPU emulation of the kernel namespace (build container, no GPU)")
ap.add_argument("--launcher", type=int, default=0, help="N > 0: the second run is `python -m tradeaid_b200.launch` on N ranks "
                "(data-parallel parity mode over Trainer) instead of main.py with the shims")
ap.add_argument("--graph", action="store_true", help="launcher: CUDA-graph replay of the step")
args = ap.parse_args()

work = tempfile.mkdtemp(prefix="mainpy_dropin_")
run = os.path.join(work, "ref_run")
shims = os.path.join(work, "shims")
shutil.copytree(args.ref, run)
cfg = open(os.path.join(run, "config.yaml")).read()
for a, b in (("dropout: 0.1", "dropout: 0.0"), ("create_new_model: 1", "create_new_model: 0"), ("save_model: 0", "save_model: 1"),
             ("device: cpu", f"device: {args.device}")):
    assert a in cfg, a
    cfg = cfg.replace(a, b)
open(os.path.join(run, "config.yaml"), "w").write(cfg)
MODEL_ARGS = ('2, [57, 3], [["p", 13, True, False, 2, 1, None, None, True, "Demo Prices", None, None], '
              '["p", 13, True, True, None, 2, 3, None, False, "Demo Changes (%)", 0.1, None]]')
DIMS = dict(batch_size=4, block_size=4, n_embd=32, n_head=4, n_layer=2)
os.makedirs(os.path.join(run, "output"), exist_ok=True)
os.makedirs(os.path.join(run, "examples", "output"), exist_ok=True)
os.makedirs(shims)
open(os.path.join(shims, "model.py"), "w").write("from tradeaid_b200.model import MultimodalTransformer  # noqa: F401\n")
open(os.path.join(shims, "training_utils.py"), "w").write(
    "from tradeaid_b200.training_utils import *  # noqa: F401,F403\n"
    "import tradeaid_b200.training_utils as _impl, sys\n"
    "sys.modules[__name__] = _impl\n")

CHECKPOINT = f"""
import os, sys, torch
os.chdir({run!r}); sys.path.insert(0, {run!r})
import config_utils
config_utils._config_cache = dict(device="cpu", eval_iters=40, dropout=0.0, fixed_values=[], output_file_name="",
                                  project_file_path="./examples/", **{DIMS!r})
import model
torch.manual_seed(42)
torch.save(model.MultimodalTransformer({MODEL_ARGS}).state_dict(), "output/demo_model.pth")
"""
RESEED = """
_orig = {mod}.MultimodalTransformer.__init__
def _init(self, *a, **k):
    _orig(self, *a, **k)
    torch.manual_seed(123); random.seed(123)          # the same generator positions after construction on both sides
{mod}.MultimodalTransformer.__init__ = _init
sys.argv = ["main.py"]
exec(compile(open("main.py").read(), "main.py", "exec"))
"""
HEAD = f"import os, random, sys, torch\nos.chdir({run!r}); sys.path.insert(0, {run!r})\n"
RUN_REF = HEAD + "import model\n" + RESEED.format(mod="model")
OURS_HEAD = HEAD + f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {shims!r})\n"
if args.emu:
    OURS_HEAD += (f"sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\nimport emu_ops\n"
                  "from tradeaid_b200 import model as _m, training_utils as _t\n_m._ops_override = emu_ops; _t._ops_override = emu_ops\n")
else:
    OURS_HEAD += "from tradeaid_b200 import model as _m, training_utils as _t, lib as _lib\n"
OURS_HEAD += f"_m.set_precision({args.precision!r})\n"
RUN_OURS = OURS_HEAD + RESEED.format(mod="_m") + (
    "" if args.emu else "print('TA_LAUNCHES', _lib.launch_count(), file=sys.stderr)\n")

if args.launcher:
    # the launcher seeds torch / random with seed + 1 after the preamble: 122 -> the 123 of the reference run above
    LAUNCH = (f"import os, sys\nsys.path.insert(0, {ROOT!r})\n"
              + (f"sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\nimport emu_ops\n"
                 "from tradeaid_b200 import model as _m, training_utils as _t\n_m._ops_override = emu_ops; _t._ops_override = emu_ops\n"
                 if args.emu else "")
              + "from tradeaid_b200 import launch\n"
              f"launch.main(['--reference', {run!r}, '--seed', '122', '--precision', {args.precision!r}"
              + (", '--graph'" if args.graph else "") + "])\n"
              + ("" if args.emu else "from tradeaid_b200 import lib as _lib\nprint('TA_LAUNCHES', _lib.launch_count(), file=sys.stderr)\n"))
    launch_file = os.path.join(work, "launch_run.py")
    open(launch_file, "w").write(LAUNCH)

ckpt = os.path.join(run, "output", "demo_model.pth")      # model_file_name is relative to the working directory
init = os.path.join(work, "init.pth")
saved = {}


def go(code, tag):
    if tag != "checkpoint":
        shutil.copyfile(init, ckpt)
    if tag == "launcher" and args.launcher > 1:
        port = 29600 + os.getpid() % 300
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.launcher}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), code]
    elif tag == "launcher":
        cmd = [sys.executable, code]
    else:
        cmd = [sys.executable, "-c", code]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    if r.returncode != 0:
        print(json.dumps({"ok": False, "stage": tag, "stdout": r.stdout[-2000:], "stderr": r.stderr[-3000:]}))
        sys.exit(1)
    if tag == "checkpoint":
        shutil.copyfile(ckpt, init)
        return None, None, r.stderr
    saved[tag] = os.path.join(work, f"saved_{tag}.pth")
    shutil.copyfile(ckpt, saved[tag])
    log = os.path.join(run, "examples", "output", "demo_training_log.txt")
    text = open(log).read() if os.path.exists(log) else ""
    if os.path.exists(log):
        os.remove(log)
    return r.stdout, text, r.stderr


NUM = re.compile(r"-?\d+(?:,\d{3})*(?:\.\d+)?")


def mask_clock(text):
    out = []
    for ln in text.splitlines():
        if ln.startswith("NCCL version "):        # torch's c10d prints it on stdout when the first communicator is made
            continue
        ln = re.sub(r"\d\d:\d\d:\d\d", "<clock>", re.sub(r"\d{4}-\d\d-\d\d", "<date>", ln))
        if ln.startswith(("Saved: Model checkpoint (", "Final Save: ")):
            ln = re.sub(r"\d+\.\d+ MB", "<size> MB", ln)     # one flat storage: fewer zip records, same keys and values
        out.append(ln)
    return out


def compare(a, b):
    """(lines whose text with numbers masked differs, max loss difference, max directional-count difference)"""
    a, b = mask_clock(a), mask_clock(b)
    if len(a) != len(b):
        import difflib
        d = [ln for ln in difflib.unified_diff([NUM.sub("#", x) for x in a], [NUM.sub("#", y) for y in b], lineterm="", n=0)
             if not ln.startswith(("---", "+++", "@@"))]
        return [f"{len(a)} lines vs {len(b)}"] + d[:10], None, None
    bad, dl, dc = [], 0.0, 0
    for x, y in zip(a, b):
        if x == y:
            continue
        if NUM.sub("#", x) != NUM.sub("#", y):
            bad += ["-" + x, "+" + y]
            continue
        nx = [float(v.replace(",", "")) for v in NUM.findall(x)]
        ny = [float(v.replace(",", "")) for v in NUM.findall(y)]
        if "DIRECTIONAL" in x or x.lstrip().startswith("- "):
            dc = max([dc] + [abs(p - q) for p, q in zip(nx, ny)])       # counts and the percentage derived from them
        else:
            dl = max([dl] + [abs(p - q) for p, q in zip(nx, ny)])
    return bad, dl, dc


go(CHECKPOINT, "checkpoint")
out_ref, log_ref, _ = go(RUN_REF, "reference")
if args.launcher:
    out_ours, log_ours, err_ours = go(launch_file, "launcher")
    saved["shimmed"] = saved["launcher"]
    # torchrun prefixes nothing to stdout here, but other ranks print nothing at all: only rank 0's console is compared
else:
    out_ours, log_ours, err_ours = go(RUN_OURS, "shimmed")
res = {"device": args.device, "precision": args.precision, "emulated": bool(args.emu), "launcher_ranks": args.launcher,
       "console_lines": len(out_ref.splitlines()), "log_lines": len(log_ref.splitlines())}
ok = "TRAINING COMPLETED SUCCESSFULLY" in out_ref and "TRAINING COMPLETED SUCCESSFULLY" in out_ours
ok = ok and "Model: Loaded successfully" in out_ref and "Model: Loaded successfully" in out_ours
bad_c, dl_c, dc_c = compare(out_ref, out_ours)
bad_l, dl_l, dc_l = compare(log_ref, log_ours)
res.update(console_text_diffs=len(bad_c), log_text_diffs=len(bad_l), max_loss_diff=max(dl_c or 0.0, dl_l or 0.0),
           max_count_diff=max(dc_c or 0, dc_l or 0), diffs=(bad_c + bad_l)[:20])
m = re.search(r"TA_LAUNCHES (\d+)", err_ours or "")
if args.launcher:      # the launcher never creates the "Optimizer" through main.py's branch text differently: same preamble, same lines
    pass
res["kernel_launches"] = int(m.group(1)) if m else None
import torch  # noqa: E402

a = torch.load(saved["reference"], weights_only=True, map_location="cpu")
b = torch.load(saved["shimmed"], weights_only=True, map_location="cpu")
res["checkpoint_keys_equal"] = list(a) == list(b) and all(a[k].shape == b[k].shape for k in a)
if res["checkpoint_keys_equal"]:
    num = sum(float((a[k].double() - b[k].double()).pow(2).sum()) for k in a)
    den = sum(float(a[k].double().pow(2).sum()) for k in a)
    res["checkpoint_normwise_diff"] = (num / den) ** 0.5
    a0 = torch.load(init, weights_only=True, map_location="cpu")
    moved = sum(float((a[k].double() - a0[k].double()).pow(2).sum()) for k in a)
    res["training_moved_weights_normwise"] = (moved / den) ** 0.5          # the run did train: the bound above is not vacuous
ok = ok and not bad_c and not bad_l and res["max_loss_diff"] <= args.loss_tol and res["max_count_diff"] <= args.count_tol
ok = ok and res["checkpoint_keys_equal"] and res.get("checkpoint_normwise_diff", 1.0) <= args.weight_tol
ok = ok and (args.emu or (res["kernel_launches"] or 0) > 1000)
res["ok"] = bool(ok)
losses = [ln for ln in mask_clock(out_ours) if "LOSS METRICS" in ln]
res["loss_lines_ours"] = losses
res["loss_lines_reference"] = [ln for ln in mask_clock(out_ref) if "LOSS METRICS" in ln]
print(json.dumps(res))
shutil.rmtree(work, ignore_errors=True)
sys.exit(0 if ok else 1)

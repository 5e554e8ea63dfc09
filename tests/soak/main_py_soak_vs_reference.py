"""Offline whole-program soak (BUILD CONTAINER ONLY: copies and runs /root/reference): the reference's own `main.py` on its
demo (config.yaml = the demo: 2 modalities of demo_stock.csv, 50 steps, evaluation at steps 0 / 25 / 49, the [2]-quirk
augmentation active because has_header is true) twice in a writable copy -
  (1) unmodified;
  (2) with the two shim modules of INTEGRATION.md section 1 ahead on sys.path (model.py -> tradeaid_b200.model,
      training_utils.py -> tradeaid_b200.training_utils), the engine's host program running over the CPU emulation of the
      kernels (there is no GPU in the build container) -
  (3) as (2) plus a data_utils shim that routes range_numeric_data / bin_numeric_data / numerical_representation /
      convert_to_percent_changes to tradeaid_b200.data_utils (lists in - lists out) -
and compares every console line (clocks masked; exact) and the training log file (its six-decimal losses to 2e-6). Both runs start from the same seeded checkpoint
(create_new_model: 0 - the engine initialises its flat buffer with its own draws) with the generators
re-seeded after model construction, and dropout 0.0 (torch's dropout stream cannot be reproduced outside torch).
Round 1: 109 console lines and 78 log-file lines, 0 differences - after this soak exposed evaluation running one optimizer step behind (model.py
_sync_shadows; fixed, regression test in tests/test_dropin_cpu.py).
usage: python tests/soak/main_py_soak_vs_reference.py [demo|folder]
("folder": three csv files in a folder, four modalities incl. a string one, file-based validation split, block_size 8)"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
work = tempfile.mkdtemp(prefix="mainpy_soak_")
run = os.path.join(work, "ref_run")
shims = os.path.join(work, "shims")
shutil.copytree(REF, run)
for d, _, files in os.walk(run):
    os.chmod(d, 0o755)
    for f in files:
        os.chmod(os.path.join(d, f), 0o644)
cfg = open(os.path.join(run, "config.yaml")).read()
cfg = cfg.replace("dropout: 0.1", "dropout: 0.0").replace("create_new_model: 1", "create_new_model: 0")
open(os.path.join(run, "config.yaml"), "w").write(cfg)
SCENARIO = sys.argv[1] if len(sys.argv) > 1 else "demo"
MODEL_ARGS = ('2, [57, 3], [["p", 13, True, False, 2, 1, None, None, True, "Demo Prices", None, None], '
              '["p", 13, True, True, None, 2, 3, None, False, "Demo Changes (%)", 0.1, None]]')
DIMS = dict(batch_size=4, block_size=4, n_embd=32, n_head=4, n_layer=2)
VOCAB_BUILDS = 2
if SCENARIO == "folder":
    # a harder set-up: a FOLDER of three csv files (multi-file index generation, per-file percent changes), four modalities -
    # ranged prices, binned percent changes, a STRING modality (time of day: non-numeric vocabulary, no directional metrics)
    # and ranged volumes - three of them with cross-attention, the last file as the validation set, block_size 8
    import random as _r
    rng = _r.Random(3)
    src = open(os.path.join(run, "examples", "demo_data", "demo_stock.csv")).read().splitlines()
    folder = os.path.join(run, "examples", "demo_data", "multi")
    os.makedirs(folder)
    for k, n_rows in enumerate((70, 100, 55)):
        rows = [src[0]]
        price = 60.0 + 25 * k
        for j in range(n_rows):
            f = src[1 + (j + 13 * k) % (len(src) - 1)].split(",")
            price = max(5.0, price * (1 + rng.gauss(0, 0.01)))
            f[12] = f"{price:.2f}"
            f[13] = str(rng.randint(900, 2500) * 1000)
            rows.append(",".join(f))
        open(os.path.join(folder, f"part_{k}.csv"), "w").write("\n".join(rows) + "\n")
    open(os.path.join(run, "input_schemas.yaml"), "w").write("""modalities:
  - modality_name: "Prices"
    path: "./examples/demo_data/multi"
    column_number: 13
    has_header: true
    processing_steps:
      - function: range_numeric_data
        args: {num_whole_digits: 2, decimal_places: 1}
        enabled: true
    cross_attention: true
    randomness_size: null
  - modality_name: "Changes (%)"
    path: "./examples/demo_data/multi"
    column_number: 13
    has_header: true
    processing_steps:
      - function: convert_to_percent_changes
        args: {decimal_places: 2}
        enabled: true
      - function: bin_numeric_data
        args: {num_bins: 4, outlier_percentile: 1, exponent: 2.0}
        enabled: true
    cross_attention: false
    randomness_size: null
  - modality_name: "Time of day"
    path: "./examples/demo_data/multi"
    column_number: 9
    has_header: true
    processing_steps: []
    cross_attention: true
    randomness_size: null
  - modality_name: "Volume"
    path: "./examples/demo_data/multi"
    column_number: 14
    has_header: true
    processing_steps:
      - function: range_numeric_data
        args: {num_whole_digits: 1, decimal_places: 1}
        enabled: true
    cross_attention: true
    randomness_size: null
""")
    cfg = open(os.path.join(run, "config.yaml")).read()
    for a, b in (("num_validation_files: 0", "num_validation_files: 1"), ("block_size: 4", "block_size: 8"), ("save_model: 0", "save_model: 1"),
                 ("max_iters: 50", "max_iters: 30"), ("eval_interval: 25", "eval_interval: 10")):
        assert a in cfg, a
        cfg = cfg.replace(a, b)
    open(os.path.join(run, "config.yaml"), "w").write(cfg)
    MODEL_ARGS = None          # vocabulary sizes are read from a first, unmodified run (below)
    DIMS = dict(batch_size=4, block_size=8, n_embd=32, n_head=4, n_layer=2)
    VOCAB_BUILDS = 3           # the string modality goes to the reference's function (NotImplementedError in the product)
os.makedirs(os.path.join(run, "output"))
os.makedirs(os.path.join(run, "examples", "output"), exist_ok=True)
os.makedirs(shims)
open(os.path.join(shims, "model.py"), "w").write("from tradeaid_b200.model import MultimodalTransformer  # noqa: F401\n")
open(os.path.join(shims, "training_utils.py"), "w").write(
    "from tradeaid_b200.training_utils import *  # noqa: F401,F403\n"
    "import tradeaid_b200.training_utils as _impl, sys\n"
    "sys.modules[__name__] = _impl\n")



def checkpoint_code(model_args):
    return f"""
import os, sys, torch
os.chdir({run!r}); sys.path.insert(0, {run!r})
import config_utils
config_utils._config_cache = dict(device="cpu", eval_iters=40, dropout=0.0, fixed_values=[], output_file_name="",
                                  project_file_path="./examples/", **{DIMS!r})
import model
torch.manual_seed(42)
torch.save(model.MultimodalTransformer({model_args}).state_dict(), "output/demo_model.pth")
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
RUN_REF = f"import os, random, sys, torch\nos.chdir({run!r}); sys.path.insert(0, {run!r})\nimport model\n" + RESEED.format(mod="model")
RUN_OURS = (f"import os, random, sys, torch\nos.chdir({run!r}); sys.path.insert(0, {run!r})\n"
            f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r}); sys.path.insert(0, {shims!r})\n"
            "import emu_ops\nfrom tradeaid_b200 import model as _m, training_utils as _t\n"
            "_m._ops_override = emu_ops; _t._ops_override = emu_ops\n" + RESEED.format(mod="_m"))


saved = {}


def go(code, tag):
    ckpt, init = os.path.join(run, "output", "demo_model.pth"), os.path.join(work, "init.pth")
    if tag == "checkpoint":
        pass
    elif os.path.exists(init):
        shutil.copyfile(init, ckpt)          # every run starts from the same checkpoint (save_model: 1 overwrites it)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1200)
    if tag == "checkpoint":
        shutil.copyfile(ckpt, init)
    elif os.path.exists(ckpt):
        saved[tag] = os.path.join(work, f"saved_{len(saved)}.pth")
        shutil.copyfile(ckpt, saved[tag])
    assert r.returncode == 0, (tag, r.stdout[-1500:], r.stderr[-1500:])
    log = os.path.join(run, "examples", "output", "demo_training_log.txt")
    text = open(log).read() if os.path.exists(log) else ""
    if os.path.exists(log):
        os.remove(log)
    if tag == "shimmed + data_utils":     # the shimmed functions really ran: one vocabulary build per modality
        assert f"K9_VOCAB_BUILDS {VOCAB_BUILDS}" in r.stderr, r.stderr[-500:]
    return r.stdout, text


def mask(text):
    # clocks and dates; and the size of the checkpoint FILE: the engine's parameters are views of one flat storage and its
    # per-head tril buffers one tensor, so torch.save writes fewer zip records (0.71 vs 0.85 MB here) - same keys, shapes and
    # values, loadable in both directions (compared below and in module_soak_vs_reference.py)
    out = []
    for ln in text.splitlines():
        ln = re.sub(r"\d\d:\d\d:\d\d", "<clock>", re.sub(r"\d{4}-\d\d-\d\d", "<date>", ln))
        if ln.startswith(("Saved: Model checkpoint (", "Final Save: ")):
            ln = re.sub(r"\d+\.\d+ MB", "<size> MB", ln)
        out.append(ln)
    return out


# (3) additionally the data-preparation functions of row f-4 (form (a): lists in, lists out) through
# a data_utils shim: everything of the reference's data_utils except the four processing functions
shims3 = os.path.join(work, "shims3")
shutil.copytree(shims, shims3)
open(os.path.join(shims3, "data_utils.py"), "w").write(
    "import importlib.util as _u\n"
    f"_spec = _u.spec_from_file_location('_reference_data_utils', {os.path.join(run, 'data_utils.py')!r})\n"
    "_ref = _u.module_from_spec(_spec); _spec.loader.exec_module(_ref)\n"
    "globals().update({k: getattr(_ref, k) for k in dir(_ref) if not k.startswith('__')})\n"
    "from tradeaid_b200.data_utils import (range_numeric_data, bin_numeric_data, convert_to_percent_changes)  # noqa: E402,F401\n"
    "from tradeaid_b200 import data_utils as _ours\n"
    "def numerical_representation(data_points):\n"
    "    try:\n"
    "        return _ours.numerical_representation(data_points)\n"
    "    except NotImplementedError:          # string modalities stay with the reference's function (INTEGRATION.md 1a)\n"
    "        return _ref.numerical_representation(data_points)\n")
RUN_OURS3 = (RUN_OURS.replace(repr(shims), repr(shims3)).replace('exec(compile(open("main.py").read(), "main.py", "exec"))',
                              'exec(compile(open("main.py").read(), "main.py", "exec"))\n'
                              'print("K9_VOCAB_BUILDS", emu_prep.stats["vocab_inserts"], file=sys.stderr)')
             .replace("_m._ops_override = emu_ops; _t._ops_override = emu_ops\n",
                      "_m._ops_override = emu_ops; _t._ops_override = emu_ops\n"
                      "import emu_prep\nfrom tradeaid_b200 import data_utils as _d\n_d._ops_override = emu_prep\n"))

if MODEL_ARGS is None:
    probe = cfg_text = open(os.path.join(run, "config.yaml")).read()
    open(os.path.join(run, "config.yaml"), "w").write(cfg_text.replace("create_new_model: 0", "create_new_model: 1")
                                                      .replace("max_iters: 30", "max_iters: 1"))
    out_probe, _ = go(RUN_REF, "probe")
    open(os.path.join(run, "config.yaml"), "w").write(cfg_text)
    sizes = [int(v) for v in re.search(r"Vocabulary sizes: \[([^\]]*)\]", out_probe).group(1).split(",")]
    flags = [True, False, True, True]
    MODEL_ARGS = f"{len(sizes)}, {sizes}, [['p', 1, True, False, None, None, None, None, {flags}[i], 'm', None, None] for i in range({len(sizes)})]"
    print("folder scenario: vocabulary sizes", sizes)
go(checkpoint_code(MODEL_ARGS), "checkpoint")
out_ref, log_ref = go(RUN_REF, "reference")
out_ours, log_ours = go(RUN_OURS, "shimmed")
out_ours3, log_ours3 = go(RUN_OURS3, "shimmed + data_utils")
assert "Model: Loaded successfully" in out_ref and "Model: Loaded successfully" in out_ours
assert "TRAINING COMPLETED SUCCESSFULLY" in out_ref and "TRAINING COMPLETED SUCCESSFULLY" in out_ours


def differing(a, b, tol):
    """Lines that differ in text - or, with tol, in any number by more than tol (the log file prints six decimals: the
    float32 round-off between torch's kernels and the emulation shows in the last one)."""
    if len(a) != len(b):
        return [f"{len(a)} lines vs {len(b)}"]
    out = []
    num = re.compile(r"-?\d+\.\d+")
    for x, y in zip(a, b):
        if x == y:
            continue
        nx, ny = num.findall(x), num.findall(y)
        if tol and num.sub("#", x) == num.sub("#", y) and all(abs(float(p) - float(q)) <= tol for p, q in zip(nx, ny)):
            continue
        out += ["-" + x, "+" + y]
    return out


d_out = differing(mask(out_ref), mask(out_ours), 0)
d_log = differing(mask(log_ref), mask(log_ours), 2e-6)
print("\n".join(ln for ln in mask(out_ours) if "LOSS METRICS" in ln))
print(f"main.py soak: {len(mask(out_ref))} console lines, {len(d_out)} differing; {len(mask(log_ref))} log-file lines, {len(d_log)} differing")
print("\n".join((d_out + d_log)[:40]))
d_out3 = differing(mask(out_ref), mask(out_ours3), 0)
d_log3 = differing(mask(log_ref), mask(log_ours3), 2e-6)
print(f"with the data_utils shim as well: {len(d_out3)} console lines and {len(d_log3)} log-file lines differing")
print("\n".join((d_out3 + d_log3)[:40]))
if SCENARIO == "folder":                 # save_model: 1 - the checkpoints main.py wrote at the end of the runs
    import torch
    a = torch.load(saved["reference"], weights_only=True)
    for tag in ("shimmed", "shimmed + data_utils"):
        b = torch.load(saved[tag], weights_only=True)
        assert list(a.keys()) == list(b.keys())
        init_sd = torch.load(os.path.join(work, "init.pth"), weights_only=True)
        worst = max(float((a[k] - b[k]).norm() / ((a[k] - init_sd[k]).norm() + 1e-12)) for k in a
                    if not k.endswith(".tril") and float((a[k] - init_sd[k]).abs().max()) > 0)
        print(f"saved checkpoint ({tag}): same keys in the same order; parameter updates within {worst:.1e} normwise of the reference's")
        assert worst <= 5e-2
shutil.rmtree(work)
assert not d_out and not d_log and not d_out3 and not d_log3

"""bench.py's JSON-line contract on a machine without a GPU: the reference arm (`--impl reference`: the unmodified reference
from oracle/_ref on the host cores, or the oracle port where that copy is absent) prints ONE line with the keys the driver reads; `--config prep --impl reference`
answers with the `unavailable` form. (The CUDA arms of bench.py run on the GPU box only.)"""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_bench(*flags):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, env=env,
                       cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--config", "c1", "--steps", "2", "--warmup", "1")
    assert d["impl"] == "reference" and d["metric"] == "training tokens/s" and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and abs(d["value"] - d["config"]["modalities"] * d["config"]["global_batch"]
                                  * d["config"]["block_size"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "model.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("c1") and "model" not in d["config"]
    assert d["vs_baseline"] is None and d["data"] == "synthetic"


def test_prep_reference_arm_is_declared_unavailable():
    d = run_bench("--config", "prep", "--impl", "reference")
    assert d["impl"] == "reference" and "unavailable" in d

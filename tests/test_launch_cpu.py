"""Row x-1 (the data-parallel launcher) on CPU: `python -m tradeaid_b200.launch` on two gloo ranks in parity mode, the engine's
host program over the CPU emulation of the kernels, against the UNMODIFIED reference main.py on its demo — every console line
and log-file line identical, checkpoints equal (tests/main_py_dropin_run.py). Needs oracle/_ref (made by build() in the build
container; absent on a bare checkout: skipped)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "main.py")), reason="oracle/_ref is missing: python oracle/make_ref.py")
def test_two_rank_launcher_prints_what_the_reference_main_py_prints():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "main_py_dropin_run.py"), "--ref", REF, "--device", "cpu", "--emu",
                        "--launcher", "2", "--loss-tol", "2e-5", "--count-tol", "0", "--weight-tol", "1e-4"],
                       capture_output=True, text=True, timeout=1500, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert line, (r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(line[-1])
    assert res["ok"] and res["console_text_diffs"] == 0 and res["log_text_diffs"] == 0, res
    assert res["console_lines"] >= 100 and res["log_lines"] >= 70
    assert res["training_moved_weights_normwise"] > 1e-2 and res["checkpoint_normwise_diff"] <= 1e-5

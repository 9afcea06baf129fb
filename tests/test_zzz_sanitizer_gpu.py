"""compute-sanitizer memcheck over tools/sanitize_small.py on the GPU box (second half): tiny train runs
(RESIDENT and forced-STREAMING merge loops, default and wide-probe kernels) and encodes (default kernels with and
without dropout, every experimental variant), each result also compared with the oracle inside the script.
Subprocess with a hard timeout.  PASS = the tool reports no error and every result is identical to the oracle.
A report, a crash or a timeout is an XFAIL with the tool's summary (never run on hardware before this round's end:
the sanitizer runs are evidence to collect, not yet a gate); the full log goes to gpurun_out/.  Runs last (zzz)."""
import os
import shutil
import subprocess
import sys

import pytest

from _bind import ROOT

pytestmark = pytest.mark.gpu


def test_memcheck_of_a_small_train_and_encode_workload(product):
    tool = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(tool):
        pytest.skip("compute-sanitizer is not installed")
    env = {k: v for k, v in os.environ.items() if not k.startswith(("YTTM_", "YT_EMU_"))}
    try:
        r = subprocess.run([tool, "--tool", "memcheck", "--error-exitcode", "9", sys.executable,
                            os.path.join(ROOT, "tools", "sanitize_small.py")], cwd=ROOT, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=240)
    except subprocess.TimeoutExpired:
        pytest.xfail("memcheck run did not finish in 240 s")
    text = r.stdout.decode(errors="replace")
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "sanitizer_memcheck.log"), "w") as fh:
            fh.write(text)
    except OSError:
        pass
    summary = [ln for ln in text.splitlines() if "ERROR SUMMARY" in ln or "sanitize_small:" in ln or "Error" in ln][-6:]
    if r.returncode != 0 or "ERROR SUMMARY: 0 errors" not in text or "checks identical to the oracle" not in text:
        pytest.xfail("memcheck: rc %d; %s" % (r.returncode, " | ".join(summary) or text[-600:]))

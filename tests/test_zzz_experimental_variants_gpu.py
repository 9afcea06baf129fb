"""The env-gated EXPERIMENTAL encode kernels (YTTM_ENC_FIND_CACHED / _FIND_VEC / _BUCKETED / _ZLIN / _DEDUP; off by default, parity
proven only under the CPU SIMT emulator: tests/test_simt_emul_cpu.py, tools/fuzz_emul.py) on real hardware: tools/ab_encode.py
in a SUBPROCESS with a hard timeout, on a 200 k-sentence cut of the bench workload.  The default kernels are what the
other GPU tests pin against the oracle; this file only asks whether every variant returns the default kernels' ids.
A mismatch, a crash or a hang of an experimental kernel is reported as XFAIL with the reason (they are not the
product path), never as a silent pass; the per-stage timings land in gpurun_out/ for the next A/B.  Runs last (zzz)."""
import json
import os
import subprocess
import sys

import pytest

from _bind import ROOT

pytestmark = pytest.mark.gpu


def test_experimental_encode_variants_return_the_default_ids(product, tmp_path):
    out_json = tmp_path / "ab.json"
    env = {k: v for k, v in os.environ.items() if not k.startswith("YTTM_ENC_")}
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ab_encode.py"), "200000", "3", str(out_json)], cwd=ROOT,
                           env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=420)
    except subprocess.TimeoutExpired:
        pytest.xfail("experimental encode variants: tools/ab_encode.py did not finish in 420 s on this box")
    if r.returncode != 0 or not out_json.exists():
        pytest.xfail("experimental encode variants: tools/ab_encode.py failed: " + r.stderr.decode(errors="replace")[-1500:])
    res = json.loads(out_json.read_text())
    assert res["default"]["ids_equal_default"] and res["default"]["n_ids"] > 0
    try:  # evidence for the next round's A/B (scratch directory; ignored if not writable)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "ab_encode_200k.json"), "w") as fh:
            json.dump(res, fh, indent=1)
    except OSError:
        pass
    bad = [k for k, v in res.items() if not v.get("ids_equal_default", True)]
    if bad:
        pytest.xfail("experimental encode variants differ from the default ids on hardware: %s" % bad)
    assert set(res) >= {"default", "find_cached", "bucketed", "both", "both+zlin", "dedup", "dedup+find_cached", "find_vec", "dedup+find_vec"}


def test_experimental_train_and_encode_variants_match_the_oracle(product):
    """tools/sanitize_small.py without a sanitizer: tiny trainings with the default AND the wide-probe merge loop
    (RESIDENT and forced STREAMING) and encodes with every experimental variant, each compared with the ORACLE
    inside the script.  Subprocess + hard timeout; a difference, a crash or a hang is an XFAIL with the reason."""
    env = {k: v for k, v in os.environ.items() if not k.startswith(("YTTM_ENC_", "YTTM_LOOP_", "YTTM_FORCE_", "YTTM_STREAM_", "YTTM_STAGES"))}
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_small.py")], cwd=ROOT, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    except subprocess.TimeoutExpired:
        pytest.xfail("experimental variants vs the oracle: tools/sanitize_small.py did not finish in 300 s")
    text = r.stdout.decode(errors="replace")
    if r.returncode != 0 or "checks identical to the oracle" not in text:
        pytest.xfail("experimental variants vs the oracle: " + text[-1500:])

#!/bin/bash
# Runs the reference's OWN Python test files (tests/unit_tests/test_python_api.py, test_cli.py, test_manual.py, read
# from /root/reference, not copied) against this repo's Python package and CLI, with the kernels under the CPU SIMT
# emulator (tests/emul/simt - test harness only; there is no GPU in the build container).  A shim package named
# `youtokentome` and a `yttm` launcher are created in a scratch directory.  Slow: the tests train vocab 16000 on a
# 1 MB corpus several times (minutes per training when emulated).   usage: tools/run_reference_pytests_emulated.sh [pytest args]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${REF:-/root/reference}
W=$(mktemp -d /tmp/yttm_refpy.XXXXXX)
mkdir -p "$W/shim/youtokentome" "$W/bin" "$W/work"
cat > "$W/shim/youtokentome/__init__.py" <<PY
import sys
sys.path.insert(0, "$ROOT"); sys.path.insert(0, "$ROOT/tests")
from youtokentome_b200 import _lib as _l
from _emu import emu_lib
_l._lib = emu_lib()          # the emulated kernels stand in for libyttm_b200.so in THIS scratch shim only
from youtokentome_b200 import *  # noqa
from youtokentome_b200 import BPE, OutputType  # noqa
PY
cat > "$W/bin/yttm" <<PY
#!$(command -v python3)
import sys
sys.path.insert(0, "$W/shim")
import youtokentome  # noqa: F401
from youtokentome_b200.yttm_cli import main
main()
PY
chmod +x "$W/bin/yttm"
cd "$W/work"
PATH="$W/bin:$PATH" PYTHONPATH="$W/shim:$REF/tests/unit_tests" YT_EMU_SMS=${YT_EMU_SMS:-12} \
  python3 -m pytest "$REF/tests/unit_tests/test_python_api.py" "$REF/tests/unit_tests/test_cli.py" \
  "$REF/tests/unit_tests/test_manual.py" -q -p no:cacheprovider --rootdir="$W/work" "$@"

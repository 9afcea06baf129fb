"""Committed golden vectors (tests/golden/*.json, produced by tests/golden/make_golden.py from the
unmodified reference built with -DDETERMINISTIC_QUEUE).  CPU: the oracle reproduces them.  GPU:
the CUDA product reproduces them through the C ABI.  Needs neither /root/reference nor oracle/_ref."""
import base64
import glob
import json
import os

import pytest

from _bind import read_model, tmp_model_path

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.json")))


def _load(path):
    with open(path) as f:
        g = json.load(f)
    g["train"] = base64.b64decode(g["train_b64"])
    g["sentences"] = [base64.b64decode(s) for s in g["sentences_b64"]]
    g["want_model"] = ({int(k): int(v) for k, v in g["model"]["char2id"]},
                       [tuple(r) for r in g["model"]["rules"]], tuple(g["model"]["special_line"]))
    return g


def test_fixtures_present():
    assert len(FIXTURES) >= 10


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-5] for p in FIXTURES])
def test_oracle_reproduces_golden(oracle, path):
    g = _load(path)
    m = tmp_model_path("orc")
    oracle.train(g["train"], m, g["vocab_size"], g["coverage"], **g["special"])
    assert read_model(m) == g["want_model"]
    enc = oracle.encoder(m)
    for item in g["ids"]:
        assert enc.encode(g["sentences"], **item["flags"]) == item["ids"]


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-5] for p in FIXTURES])
def test_gpu_reproduces_golden(product, path):
    from _gpu import GpuEncoder, gpu_train
    g = _load(path)
    m = gpu_train(g["train"], g["vocab_size"], g["coverage"], **g["special"])
    assert read_model(m) == g["want_model"]
    enc = GpuEncoder(m)
    for item in g["ids"]:
        assert enc.encode(g["sentences"], **item["flags"]) == item["ids"]

"""Generates tests/golden/*.json from the UNMODIFIED reference (oracle/_ref, -DDETERMINISTIC_QUEUE
build; only possible where /root/reference exists).  Each fixture holds a training text, its
parameters, the reference's model (char2id, rules, special ids) and the reference's ids for test
sentences under several flag sets.  The GPU box has no /root/reference: the fixtures travel."""
import base64
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _cases  # noqa: E402
from _bind import Reference, build_checkers, read_model, tmp_model_path  # noqa: E402
from youtokentome_b200 import synth  # noqa: E402

FLAGS = [dict(), dict(bos=True, eos=True), dict(reverse=True, eos=True)]


def b64(b):
    return base64.b64encode(b).decode()


def main():
    build_checkers()
    ref = Reference("det")
    cases = {}
    for name, (tr, te, vocab) in synth.GOLDEN_TEXTS.items():
        cases["manual_" + name] = (tr.encode(), vocab, 1.0, (0, 1, 2, 3), [te.encode()] + _cases.EDGE_SENTENCES)
    cases["manual_baba"] = (b"baba baaab", 9, 1.0, (0, 1, 2, 3), [b"d d", b"baba baaab"])
    cases["readme_small"] = (synth.readme_corpus(n_lines=400), 500, 1.0, (0, 1, 2, 3),
                             [synth.stress_text(k, 120, train=False) for k in range(12)])
    cases["special_ids"] = (synth.readme_corpus(n_lines=300), 300, 0.999, (7, 0, 3, 299),
                            [synth.stress_text(k, 80, train=False) for k in range(6)])
    cases["dirty_zipf_cov98"] = (_cases.dirty_zipf_text(60_000), 700, 0.98, (0, 1, 2, 3),
                                 _cases.zipf_sentences(40, 80) + _cases.EDGE_SENTENCES)
    for seed in (3, 17, 41):
        text, vocab, cov, sents = _cases.stress_case(seed)
        cases["stress_%d" % seed] = (text, vocab, cov, (0, 1, 2, 3), sents)
    for name, (text, vocab, cov, sp, sents) in cases.items():
        m = tmp_model_path(name)
        pad, unk, bos, eos = sp
        ref.train(text, m, vocab, cov, n_threads=4, pad=pad, unk=unk, bos=bos, eos=eos)
        c2i, rules, special = read_model(m)
        enc = ref.encoder(m, n_threads=2)
        out = {"train_b64": b64(text), "vocab_size": vocab, "coverage": cov,
               "special": {"pad": pad, "unk": unk, "bos": bos, "eos": eos},
               "model": {"char2id": sorted(c2i.items()), "rules": rules, "special_line": list(special)},
               "sentences_b64": [b64(s) for s in sents],
               "ids": [{"flags": f, "ids": enc.encode(sents, **f)} for f in FLAGS]}
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f, separators=(",", ":"))
        print(name, len(rules), "rules")


if __name__ == "__main__":
    main()

"""Generates tests/golden/turns_reference.json by running the REFERENCE's own ``append_speaker_turns``.

/root/reference/stac-st/inference.py cannot be imported here (it imports speechbrain, librosa, hyperpyyaml, ... at
module level), so the function definition is cut out of the file with `ast` and executed, unmodified, in a namespace that
supplies the three module-level names it uses: ``hparams`` (yaml keys turn = 7, xt = 8,
transformer_multitask.yaml:138-149), ``DOWNSAMPLING`` (inference.py:48) and the two result lists.  The input is a torch
tensor of posteriors (the function calls .argmax / .cpu().numpy() on it).

    python tests/golden/make_turns_golden.py          (in the container that has /root/reference)
"""
import ast
import json
from pathlib import Path
from types import SimpleNamespace

import torch

REF = Path("/root/reference/stac-st/inference.py")
OUT = Path(__file__).resolve().parent / "turns_reference.json"


def reference_function():
    src = REF.read_text()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "append_speaker_turns")
    ns = {"hparams": {"turn": 7, "xt": 8}, "DOWNSAMPLING": 25, "turn_rttm": [], "xt_rttm": []}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(REF), "exec"), ns)
    return ns


def make_case(seed, b, t2, vocab, p_spike):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(9, vocab, (b, t2), generator=g)
    r = torch.rand(b, t2, generator=g)
    ids[r < p_spike] = 7
    ids[(r >= p_spike) & (r < 2 * p_spike)] = 8
    ids[r > 0.7] = 0                                        # blanks
    utt = [f"spk{seed}-rec{i}-{1000 * i + 37:07d}-{1000 * i + 900:07d}" for i in range(b)]
    return ids, utt


def main():
    cases = []
    for seed, b, t2, vocab, p in [(1, 3, 40, 50, 0.1), (2, 8, 251, 5000, 0.02), (3, 1, 7, 20, 0.3), (4, 5, 300, 5000, 0.0)]:
        ids, utt = make_case(seed, b, t2, vocab, p)
        p_ctc = torch.full((b, t2, vocab), -20.0)
        p_ctc.scatter_(2, ids[..., None], -0.1)             # arg-max = ids
        ns = reference_function()
        ns["append_speaker_turns"](SimpleNamespace(id=utt), p_ctc)
        cases.append({"seed": seed, "vocab": vocab, "ids": ids.tolist(), "utt": utt,
                      "turn_rttm": ns["turn_rttm"], "xt_rttm": ns["xt_rttm"]})
    OUT.write_text(json.dumps(cases))
    print(OUT, [(len(c["turn_rttm"]), len(c["xt_rttm"])) for c in cases])


if __name__ == "__main__":
    main()

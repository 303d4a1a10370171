#!/usr/bin/env python
"""Regenerates tests/golden/ref_outputs/ by running the UNMODIFIED reference binary
(oracle/_ref/selection, built by oracle/build_ref.sh from /root/reference/src/selection.cpp)
on synthetic inputs.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import golden_cases as G  # noqa: E402
from cuda_selection_criteria_b200 import sketch_io  # noqa: E402


def main():
    ref = os.path.join(ROOT, "oracle", "_ref", "selection")
    subprocess.run([os.path.join(ROOT, "oracle", "build_ref.sh")], check=True)
    out_dir = os.path.join(HERE, "ref_outputs")
    os.makedirs(out_dir, exist_ok=True)
    cases = []
    for case in G.CASES:
        data = G.build_inputs(case)
        with tempfile.TemporaryDirectory() as td:
            crit = case["criterion"]
            names = data["names"]
            for i, nm in enumerate(names):
                base = os.path.join(td, nm)
                sketch_io.write_hll(base + ".hll", data["regs"][i], data["p"], level=1)
                if crit == "smh_a":
                    sketch_io.write_smh(base + ".smh" + str(data["aux"].shape[1]), data["aux"][i], level=1)
                elif crit in ("hll_a", "hll_an"):
                    pa = case["aux_bytes"].bit_length() - 1
                    sketch_io.write_hll(base + ".hll_" + str(pa), data["aux"][i], pa, level=1)
                else:
                    sketch_io.write_smh(base + ".smh1", np.array([42], np.uint64), level=1)
            with open(os.path.join(td, "list.txt"), "w") as f:
                f.write("\n".join(names) + "\n")
            flag_c = "smh_a" if crit == "cb" else crit
            cmd = [ref, "-l", "list.txt", "-t", "8", "-h", str(case["tau"]), "-a", str(case["aux_bytes"]), "-c", flag_c]
            out = subprocess.run(cmd, cwd=td, capture_output=True, text=True, check=True).stdout
        with open(os.path.join(out_dir, case["id"] + ".txt"), "w") as f:
            f.write(out)
        rec = dict(case)
        rec["input_sha256"] = G.digest(data)
        rec["lines"] = out.count("\n")
        rec["cmd"] = " ".join(cmd[1:])
        cases.append(rec)
        print(case["id"], rec["lines"], "lines")
    with open(os.path.join(out_dir, "cases.json"), "w") as f:
        json.dump(cases, f, indent=1)


if __name__ == "__main__":
    main()

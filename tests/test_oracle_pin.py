"""Pins the CPU oracle (oracle/oracle.cpp) to the reference: its golden output and the unmodified binary."""
import os
import subprocess

import numpy as np
import pytest

import oracle_api as O
from cuda_selection_criteria_b200 import sketch_io, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "influenza")


def _load(criterion, aux_bytes):
    files = sketch_io.load_file_list(os.path.join(GOLD, "test_influeza_filelist.txt"))
    p, regs, stored, kind, alen, aux = sketch_io.load_sketches(files, criterion, aux_bytes, base=GOLD)
    return files, p, regs, stored, aux


def test_fixture_headers_and_cardinalities():
    files, p, regs, stored, _ = _load("smh_a", 512)
    assert p == 14 and regs.shape == (10, 16384)
    assert np.all(stored == -1.0)                      # fixtures store (0,2,2,1),14,-1.0
    cards = [O.cardinality(regs[i], 14) for i in range(10)]
    assert 12892 < min(cards) < 12893 and 13323 < max(cards) < 13324   # SURVEY §8c probe


@pytest.mark.parametrize("criterion,aux_bytes", [("smh_a", 512), ("smh_a", 32), ("hll_a", 256), ("hll_an", 256)])
def test_oracle_reproduces_results_txt(criterion, aux_bytes):
    """results.txt:1-7 = `selection -l test_influeza_filelist.txt -h 0.9 -a 512 -c smh_a`; the other three
    flag sets give the same 7 lines (SURVEY §4 [probe])."""
    files, p, regs, stored, aux = _load(criterion, aux_bytes)
    res = O.select(regs, p, criterion, np.float32(0.9), aux=aux, stored=stored)
    want = open(os.path.join(GOLD, "results.txt")).read().splitlines()
    assert O.format_lines(files, res) == want


def test_band_params_match_survey_probes():
    assert O.band_params(128, 0.9) == (16, 8)
    assert O.band_params(64, 0.9) == (8, 8)
    assert O.band_params(4, 0.9) == (2, 2)
    for tau, want in [(0.70, (32, 4)), (0.75, (32, 4)), (0.80, (32, 4)), (0.85, (16, 8)), (0.95, (8, 16))]:
        assert O.band_params(128, tau) == want
    # fallthrough difference between the two drivers (selection.cpp:258-267 vs selection_cuda.cpp:119-128)
    assert O.band_params(1, 0.5, True) == (1, 1)
    assert O.band_params(4, 0.01, True) == (4, 1) and O.band_params(4, 0.01, False) == (1, 1)


REF_OUT = os.path.join(os.path.dirname(__file__), "golden", "ref_outputs")


def _golden_cases():
    import json
    with open(os.path.join(REF_OUT, "cases.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", _golden_cases() if os.path.exists(os.path.join(REF_OUT, "cases.json")) else [],
                         ids=lambda c: c["id"])
def test_oracle_matches_reference_binary_goldens(case):
    """Outputs of the UNMODIFIED reference binary on synthetic inputs (tests/golden/make_golden.py).
    The inputs are regenerated here and their digest checked, so the comparison is on identical bytes."""
    import golden_cases as G
    data = G.build_inputs(case)
    assert G.digest(data) == case["input_sha256"], "synthetic generator drifted: regenerate goldens"
    res = O.select(data["regs"], data["p"], case["criterion"], np.float32(case["tau"]), aux=data["aux"],
                   z=1.96, order_n=1)
    got = O.format_lines(data["names"], res)
    want = open(os.path.join(REF_OUT, case["id"] + ".txt")).read().splitlines()
    assert got == want


@pytest.mark.skipif(O.ref_binary() is None, reason="oracle/_ref/selection not built (needs /root/reference)")
def test_live_reference_binary_on_influenza(tmp_path):
    out = subprocess.run([O.ref_binary(), "-l", "test_influeza_filelist.txt", "-t", "4", "-h", "0.9", "-a", "512",
                          "-c", "smh_a"], cwd=GOLD, capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(GOLD, "results.txt")).read()

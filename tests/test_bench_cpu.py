"""CPU-only: bench.py's host logic — the reference arm on a small workload (the unmodified binary over files written in
the reference's layout), the identical `config` of both arms, and the two parity diffs on constructed lists."""
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest

import oracle_api as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_arm_runs_the_whole_configuration():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "1500", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    cfg = bench.Cfg(1500, 1002, "smh_a", 0.9, 1024)
    assert line["impl"] == "reference" and line["config"] == cfg.config()        # the same dict our arm prints
    assert line["steps"] == 2 and line["warmup"] == 1 and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["cores"] == os.cpu_count() and "whole workload" in cb["sample"] and cb["passes"] == 2
    if O.ref_binary():
        assert cb["kind"] == "reference"
        # the pair list the arm's passes printed is the oracle's for the same inputs
        regs, aux = bench.host_inputs(cfg)
        ora = O.select(regs, 14, "smh_a", np.float32(0.9), aux=aux, threads=4)
        assert cb["lines"] == len(ora["i"])
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_reference_arm_budget_keeps_full_passes(monkeypatch, capsys):
    """When warmup + steps passes do not fit the budget the arm still runs FULL passes, fewer of them, and says so."""
    a = types.SimpleNamespace(gpus=1, steps=20, warmup=5, ref_budget_s=0.0)
    cfg = bench.Cfg(400, 7, "smh_a", 0.9, 1024)
    bench.run_reference_arm(a, cfg, 0, sys.stdout)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["steps"] == 1 and line["warmup"] == 1 and line["steps_requested"] == 20 and "note" in line
    assert line["config"] == cfg.config() and "whole workload" in line["cpu_baseline"]["sample"]


def test_line_diff_reports_missing_extra_and_printed_error():
    ref = ["a b 0.950000", "a c 0.910000", "b c 0.990000"]
    got = ["a b 0.950001", "b c 0.990000", "c d 0.930000"]
    d = bench.diff_reference_lines(got, ref, 0.9)
    assert d["pairs_missing"] == 1 and d["pairs_extra"] == 1 and not d["lines_identical_in_order"]
    assert abs(d["max_abs_diff_printed_jaccard"] - 1e-6) < 1e-12
    assert not bench.parity_ok(d)
    same = bench.diff_reference_lines(ref, ref, 0.9)
    assert same["lines_identical_in_order"] and bench.parity_ok(same)


def test_oracle_diff_excludes_only_near_tau_pairs():
    ora = {"i": np.array([0, 0, 2], np.int32), "k": np.array([1, 3, 3], np.int32),
           "jaccard": np.array([0.95, 0.9000000001, 0.97]), "stage": [6, 5, 4, 3],
           "cards_sorted": np.array([10.5, 11.5, 12.5, 13.5]), "near_i": np.array([0], np.int32), "near_k": np.array([3], np.int32)}
    res = types.SimpleNamespace(i=np.array([0, 2], np.int32), k=np.array([1, 3], np.int32), jaccard=np.array([0.95, 0.97 * (1 + 5e-7)]),
                                near_i=np.array([0], np.int32), near_k=np.array([3], np.int32), near_jaccard=np.array([0.8999999999]),
                                stats={"pairs_cb": 5, "pairs_aux": 4, "pairs_out": 2}, cards_sorted=np.array([10.1, 11.9, 12.0, 13.99]))
    d = bench.diff_oracle(res, ora, 0.9)
    assert d["pairs_missing"] == 0 and d["pairs_extra"] == 0          # (0,3) sits within 1e-6 of tau: listed, not counted
    assert d["near_tau_count"] == 1 and d["near_tau"][0][:2] == [0, 3] and d["near_tau_equal_oracle"]
    assert 4e-7 < d["max_rel_jaccard"] < 6e-7 and bench.parity_ok(d)
    res.i = np.array([0], np.int32); res.k = np.array([1], np.int32); res.jaccard = np.array([0.95])
    d = bench.diff_oracle(res, ora, 0.9)
    assert d["pairs_missing"] == 1 and not bench.parity_ok(d)         # (2,3) is not near tau: a real miss


def test_other_configs_are_the_baseline_configs():
    labels = [l for l, _ in bench.other_configs()]
    assert labels[0].startswith("C2") and sum(l.startswith("C3") for l in labels) == 6
    assert sum(l.startswith("C5") for l in labels) == 4
    c5 = dict(bench.other_configs())["C5 hll_a p_aux=10"]
    assert (c5.n, c5.criterion, c5.aux_bytes, c5.tau) == (50_000, "hll_a", 1024, 0.9)

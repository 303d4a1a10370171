"""-m gpu: the C++ host CLIs and the link-level launcher shims."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import golden_cases as G
import oracle_api as O
import cuda_selection_criteria_b200 as S
from cuda_selection_criteria_b200 import sketch_io, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden", "influenza")
BIN = os.path.join(ROOT, "cuda_selection_criteria_b200", "bin")
SMH = "_Z17launch_kernel_smhPKhPKmPKdPK4int2idiiiiP6ResultPii"
CBSMH = "_Z19launch_kernel_CBsmhPKhPKmPKdPK4int2idiiiiP6ResultPii"


@pytest.mark.parametrize("flags", [["-h", "0.9", "-a", "512", "-c", "smh_a"], ["-h", "0.9", "-a", "32", "-c", "smh_a"],
                                   ["-h", "0.9", "-a", "256", "-c", "hll_a", "-t", "2"],
                                   ["-h", "0.9", "-a", "256", "-c", "hll_an"]])
def test_cli_selection_reproduces_results_txt(gpu, flags):
    out = subprocess.run([os.path.join(BIN, "selection"), "-l", "test_influeza_filelist.txt"] + flags, cwd=GOLD,
                         capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(GOLD, "results.txt")).read()


def test_cli_error_behaviour(gpu):
    sel = os.path.join(BIN, "selection")
    r = subprocess.run([sel, "-l", "test_influeza_filelist.txt"], cwd=GOLD, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "Option -c invalid. The accepted criteria are hll_a, hll_an and smh_a.\n"
    r = subprocess.run([sel, "-l", "missing.txt", "-c", "smh_a"], cwd=GOLD, capture_output=True, text=True)
    assert r.returncode == 255 and r.stderr == "No valid input file provided\n"
    r = subprocess.run([sel, "-c", "smh_a"], cwd=GOLD, capture_output=True, text=True)
    assert r.returncode == 255 and r.stderr == "No input file provided\n"
    r = subprocess.run([sel, "-x"], capture_output=True, text=True)
    assert r.stdout == "Usage: -l -t -a -h -c\n"
    # unreadable sketch: uncaught std::runtime_error -> abort, like the reference (selection.cpp:16)
    r = subprocess.run([sel, "-l", "test_influeza_filelist.txt", "-c", "smh_a", "-a", "64"], cwd=GOLD,
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Could not open file at" in r.stderr


@pytest.mark.parametrize("case_id", ["smh_n400_t090", "hlla_n300_p8_t080", "hllan_n400_p10_t090"])
def test_cli_on_synthetic_files_matches_reference_golden(gpu, tmp_path, case_id):
    case = [c for c in json.load(open(os.path.join(HERE, "golden", "ref_outputs", "cases.json"))) if c["id"] == case_id][0]
    data = G.build_inputs(case)
    crit = case["criterion"]
    for i, nm in enumerate(data["names"]):
        base = str(tmp_path / nm)
        sketch_io.write_hll(base + ".hll", data["regs"][i], 14, level=1)
        if crit == "smh_a":
            sketch_io.write_smh(base + ".smh" + str(data["aux"].shape[1]), data["aux"][i], level=1)
        else:
            pa = case["aux_bytes"].bit_length() - 1
            sketch_io.write_hll(base + ".hll_" + str(pa), data["aux"][i], pa, level=1)
    (tmp_path / "list.txt").write_text("\n".join(data["names"]) + "\n")
    out = subprocess.run([os.path.join(BIN, "selection"), "-l", "list.txt", "-t", "4", "-h", str(case["tau"]), "-a",
                          str(case["aux_bytes"]), "-c", crit], cwd=tmp_path, capture_output=True, text=True,
                         check=True).stdout
    assert out == open(os.path.join(HERE, "golden", "ref_outputs", case_id + ".txt")).read()


def test_cli_selection_cuda_driver(gpu):
    """selection_cuda flags (-l -b -a -h, -c ignored): same pair set as the CPU driver on the fixtures;
    similarities printed as floats (src/selection_cuda.cpp:184-186)."""
    out = subprocess.run([os.path.join(BIN, "selection_cuda"), "-l", "test_influeza_filelist.txt", "-b", "128", "-a",
                          "512", "-h", "0.9", "-c", "hll_a"], cwd=GOLD, capture_output=True, text=True, check=True).stdout
    want = open(os.path.join(GOLD, "results.txt")).read().splitlines()
    got = out.splitlines()
    assert [ln.split()[:2] for ln in got] == [ln.split()[:2] for ln in want]
    for g, w in zip(got, want):
        assert abs(float(g.split()[2]) - float(w.split()[2])) < 2e-6


def _launch(lib, name, *args):
    fn = getattr(lib, name)
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                   C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    fn(*args)


@pytest.mark.parametrize("sym", [SMH, CBSMH])
@pytest.mark.parametrize("tau", [0.9, 0.5])
def test_shims_match_reference_launchers(gpu, sym, tau):
    """Our launch_kernel_smh / launch_kernel_CBsmh against the reference's own compiled kernels
    (oracle/_ref/libref_kernels.so, built unmodified from src/selection_kernels.cu) on the same
    device buffers: identical Result sets, bit-identical float similarities."""
    import torch
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_kernels.so")
    if not os.path.exists(ref_path):
        pytest.skip("oracle/_ref/libref_kernels.so not built (needs /root/reference at build time)")
    ref = C.CDLL(ref_path)
    ours = C.CDLL(S.lib_path())
    n, m_aux = 500, 128
    plan = synth.make_plan(n, 77)
    regs = synth.hll(plan, 14)
    aux = synth.smh(plan, m_aux)
    ora = O.select(regs, 14, "cb", np.float32(2.0))                     # cardinalities + order only
    order = ora["order"]
    dev = torch.device("cuda", gpu)
    d_main = torch.from_numpy(regs[order].copy()).to(dev)
    d_aux = torch.from_numpy(aux[order].view(np.int64).copy()).to(dev)
    d_cards = torch.from_numpy(ora["cards_sorted"].copy()).to(dev)
    ii, kk = np.triu_indices(n, 1)
    pairs = np.stack([ii, kk], 1).astype(np.int32)
    d_pairs = torch.from_numpy(pairs).to(dev)
    total = pairs.shape[0]
    nb, nr = O.band_params(m_aux, tau, False)
    res = {}
    for tag, lib in (("ref", ref), ("ours", ours)):
        d_out = torch.zeros(total * 3, dtype=torch.int32, device=dev)
        d_cnt = torch.full((1,), -5, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        _launch(lib, sym, d_main.data_ptr(), d_aux.data_ptr(), d_cards.data_ptr(), d_pairs.data_ptr(), total,
                float(np.float32(tau)), m_aux, 1 << 14, nr, nb, d_out.data_ptr(), d_cnt.data_ptr(), 128)
        torch.cuda.synchronize()
        cnt = int(d_cnt.item())
        out = d_out[: cnt * 3].cpu().numpy().reshape(cnt, 3)
        key = out[:, 0].astype(np.int64) << 32 | out[:, 1]
        o = np.argsort(key)
        res[tag] = (key[o], out[o, 2].copy().view(np.float32))
    assert res["ref"][0].size > 0
    assert np.array_equal(res["ref"][0], res["ours"][0])
    assert np.array_equal(res["ref"][1], res["ours"][1])


def test_cli_time_smh_cuda(gpu):
    """Three `list;phase;tau;seconds` lines like experiments/src/time_smh_cuda.cpp:228-230,279-299."""
    r = subprocess.run([os.path.join(BIN, "time_smh_cuda"), "-l", "test_influeza_filelist.txt", "-h", "0.9", "-m", "64",
                        "-b", "128"], cwd=GOLD, capture_output=True, text=True, check=True)
    lines = r.stdout.splitlines()
    assert [ln.split(";")[:3] for ln in lines] == [["test_influeza_filelist.txt", ph, "0.9"]
                                                   for ph in ("build_smh", "smh_a", "CB+smh_a")]
    assert all(float(ln.split(";")[3]) > 0 for ln in lines)
    # SMH rebuilt from the FASTA files (M = 64) gives the 7 pairs of results.txt in both modes
    assert "smh_a 7, CB+smh_a 7" in r.stderr

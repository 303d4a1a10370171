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
    regs = synth.hll(plan, 14).copy()
    aux = synth.smh(plan, m_aux).copy()
    # four near-saturated sketches with equal auxiliary buckets: their union estimate exceeds 2^32/30, the branch where
    # criteria_sketch_cuda.cuh:61-63 negates an unsigned constant (the estimate goes negative, the pair is dropped)
    rng = np.random.default_rng(3)
    regs[10:14] = rng.integers(15, 19, size=(4, 1 << 14), dtype=np.uint8)
    aux[10:14] = aux[10]
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


def _device_ms(stderr):
    """(device ms, wall ms) of the two compare phases from the harness's `selb200: device ms: ...` lines."""
    out = []
    for ln in stderr.splitlines():
        if ln.startswith("selb200: device ms:"):
            dev, wall = ln.split("|")
            d = [float(x.split()[-1]) for x in dev.split(":")[2].split(",")]
            w = [float(x) for x in wall.split(":")[1].split(",")]
            out.append((d, w))
    return out


def test_cli_time_smh_cuda(gpu):
    """Three `list;phase;tau;seconds` lines like experiments/src/time_smh_cuda.cpp:228-230,279-299, and — unlike the
    reference, whose timers bracket an unsynchronised launch — seconds that cover the device work: each compare
    phase's printed time is at least the CUDA-event time of its run and at most a host overhead above it."""
    r = subprocess.run([os.path.join(BIN, "time_smh_cuda"), "-l", "test_influeza_filelist.txt", "-h", "0.9", "-m", "64",
                        "-b", "128"], cwd=GOLD, capture_output=True, text=True, check=True)
    lines = r.stdout.splitlines()
    assert [ln.split(";")[:3] for ln in lines] == [["test_influeza_filelist.txt", ph, "0.9"]
                                                   for ph in ("build_smh", "smh_a", "CB+smh_a")]
    assert all(float(ln.split(";")[3]) > 0 for ln in lines)
    # SMH rebuilt from the FASTA files (M = 64) gives the 7 pairs of results.txt in both modes
    assert "smh_a 7, CB+smh_a 7" in r.stderr
    (dev, wall), = _device_ms(r.stderr)
    for ph in range(2):
        printed_ms = float(lines[1 + ph].split(";")[3]) * 1e3
        assert dev[ph] > 0 and abs(printed_ms - wall[ph]) < 0.01 * max(printed_ms, 1e-3) + 1e-3
        assert dev[ph] <= printed_ms <= dev[ph] + 50.0


def test_cli_time_smh_cpu_flag_set(gpu):
    """bin/time_smh: the flags and line suffixes of the reference's CPU harness (experiments/src/time_smh.cpp:139,
    197,257,292: -t, -R repetitions, ';m:M', ';r:R_b:B') over the same GPU path."""
    r = subprocess.run([os.path.join(BIN, "time_smh"), "-l", "test_influeza_filelist.txt", "-t", "4", "-h", "0.9", "-m", "64",
                        "-R", "2"], cwd=GOLD, capture_output=True, text=True, check=True)
    lines = r.stdout.splitlines()
    assert [ln.split(";")[1] for ln in lines] == ["build_smh", "smh_a", "CB+smh_a", "smh_a", "CB+smh_a"]
    assert lines[0].split(";")[4] == "m:64"
    nb, nr = O.band_params(64, np.float32(0.9), False)
    assert all(ln.split(";")[4] == f"r:{nr}_b:{nb}" for ln in lines[1:])
    assert all(float(ln.split(";")[3]) > 0 for ln in lines)
    assert len(_device_ms(r.stderr)) == 2 and "smh_a 7, CB+smh_a 7" in r.stderr


def test_time_experiment_script(gpu, tmp_path):
    """tools/run_time_experiment.sh (the sweep of the reference's run_time_experiment.sh:17-42) over both harnesses:
    the CSV the reference's analysis expects, one row per (impl, phase, repetition), every time a positive number."""
    log = tmp_path / "times.csv"
    env = dict(os.environ, LISTA="test_influeza_filelist.txt", MH_SIZE_ARR="64", REPS="2", LOG=str(log))
    subprocess.run(["bash", os.path.join(ROOT, "tools", "run_time_experiment.sh")], cwd=GOLD, env=env, check=True,
                   capture_output=True, text=True)
    rows = log.read_text().splitlines()
    assert rows[0] == "impl,threads,mh_size,rep,criterio,tiempo"
    body = [r.split(",") for r in rows[1:]]
    assert sorted((b[0], b[4]) for b in body) == sorted((i, c) for i in ("cpu", "gpu") for c in ("build_smh", "smh_a", "CB+smh_a")
                                                        for _ in range(2))
    assert all(b[2] == "64" and float(b[5]) > 0 for b in body)
    assert {b[1] for b in body if b[0] == "cpu"} == {"8"} and {b[1] for b in body if b[0] == "gpu"} == {"256"}


def _write_synthetic_files(tmp_path, n, seed, m_smh):
    plan = synth.make_plan(n, seed)
    regs, aux = synth.hll(plan, 14), synth.smh(plan, m_smh)
    names = [f"s{i:04d}.fna.gz" for i in range(n)]
    for i, nm in enumerate(names):
        sketch_io.write_hll(str(tmp_path / nm) + ".hll", regs[i], 14, level=1)
        sketch_io.write_smh(str(tmp_path / nm) + f".smh{m_smh}", aux[i], level=1)
    (tmp_path / "list.txt").write_text("\n".join(names) + "\n")
    return regs, aux, names


@pytest.mark.parametrize("tau", ["0.9", "0.75"])
def test_comparison_experiment_script_joins_cpu_and_gpu_output(gpu, tmp_path, tau):
    """tools/run_comparison_experiment.sh = the join / diff of the reference's run_comparison_experiment.sh:36-52
    (key nameA_nameB, |sim_cpu - sim_gpu| with EPS 1e-6).  CPU side: the UNMODIFIED reference binary when it was
    built (oracle/_ref/selection), else this repo's bin/selection; GPU side: bin/selection_cuda.  Every pair of the
    CPU side must be joined, none may exist on one side only, and no similarity may differ by more than the
    float the CUDA driver prints (src/selection_cuda.cpp:184-186: 6 significant digits)."""
    regs, aux, names = _write_synthetic_files(tmp_path, 500, 31, 128)
    out = tmp_path / "cmp.csv"
    cpu = O.ref_binary() or os.path.join(BIN, "selection")
    env = dict(os.environ, LISTA="list.txt", THRESHOLD=tau, MH_SIZE_ARR="1024", THREADS="4", CPU_BINARY=cpu, OUT=str(out))
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "run_comparison_experiment.sh")], cwd=tmp_path, env=env,
                       check=True, capture_output=True, text=True)
    assert "only on the CPU side: 0, only on the GPU side: 0" in r.stderr
    rows = out.read_text().splitlines()
    assert rows[0] == "cfg,card1,card2,sim_cpu,sim_gpu,diff"
    ora = O.select(regs, 14, "smh_a", np.float32(float(tau)), aux=aux, threads=8)
    assert len(rows) - 1 == len(ora["i"]) > 100
    want = {tuple(ln.split()[:2]) for ln in O.format_lines(names, ora)}
    assert {tuple(r_.split(",")[1:3]) for r_ in rows[1:]} == want
    assert all(r_.split(",")[0] == "t4_b128_m1024_r1" for r_ in rows[1:])
    assert max(float(r_.split(",")[5]) for r_ in rows[1:]) <= 1.0000001e-6

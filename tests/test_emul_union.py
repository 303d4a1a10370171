"""CPU-only: the bit-plane union kernels compiled as host code from the .inl sources of the GPU build and run through
the warp emulator tests/emul/cuda_emul.h — k_planes_from_bytes + k_pair_hist_planes with subset counting
(k_pair_hist_planes<EpiSubsets<..>>, the default form of the union pass) and with one-hot counting
(SELB200_UNION=planes) against the byte-wise definition of the union histogram
(sketch/include/sketch/hll.h:1191-1206), for precisions 9..16, narrow and wide value ranges, empty and saturated
sketches.  The emulator checks arithmetic, index math and the
producer/consumer walk; the hardware protocol (TMA, mbarriers) is covered by the -m gpu tests."""
import os
import shutil
import subprocess

import pytest


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_union_kernels_on_the_warp_emulator(tmp_path):
    exe = tmp_path / "emul_union"
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "emul", "emul_union.cpp"), "-o", str(exe), "-ldl"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "all identical" in r.stdout
    assert r.stdout.count(" ok") >= 30 and "FAIL" not in r.stdout          # 15 cases x (planes, subsets)
    assert r.stdout.count(" subsets ") == 15
    # subset counting with the per-step group limit (per-eighth maxima), every case with 2^14 registers or more
    assert r.stdout.count(" subtops ") == 10


_MUTATIONS = [
    # (name, text in kernels/union_planes.inl, replacement, what the run must show)
    ("wrong wait parity",
     "mbar_wait(bar0 + 8 * st, (n_done / PL_STAGES) & 1u);",
     "mbar_wait(bar0 + 8 * st, ((n_done / PL_STAGES) + 1) & 1u);",
     "wait names the wrong phase parity"),
    ("ring over-issued by one stage",
     "for (int k = 0; k < PL_STAGES - 1 && prod.valid; ++k) {",
     "for (int k = 0; k < PL_STAGES && prod.valid; ++k) {",
     "previous phase was never waited for"),
    ("no __syncwarp before a stage is refilled",
     "        __syncwarp();                          // every lane has finished reading the stage about to be refilled\n",
     "\n",
     "FAIL"),
]


@pytest.mark.parametrize("name,old,new,expect", _MUTATIONS, ids=[m[0] for m in _MUTATIONS])
def test_mbarrier_model_catches_protocol_mutations(tmp_path, name, old, new, expect):
    """The emulator's TMA / mbarrier model (destination poisoned at issue, data delivered at the wait, parity and
    expect_tx checked) must reject a plane kernel whose ring protocol is broken — otherwise the CPU check of the
    union kernels would say nothing about their staging."""
    tree = tmp_path / "repo"
    shutil.copytree(os.path.join(ROOT, "tests", "emul"), tree / "tests" / "emul")
    shutil.copytree(os.path.join(ROOT, "cuda_selection_criteria_b200", "csrc", "kernels"),
                    tree / "cuda_selection_criteria_b200" / "csrc" / "kernels")
    src = tree / "cuda_selection_criteria_b200" / "csrc" / "kernels" / "union_planes.inl"
    text = src.read_text()
    assert text.count(old) == 1, name
    src.write_text(text.replace(old, new))
    exe = tmp_path / "emul_union_mut"
    subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-Wno-unknown-pragmas",
                    str(tree / "tests" / "emul" / "emul_union.cpp"), "-o", str(exe), "-ldl"], check=True)
    # the missing-barrier mutation shows through thread timing (a lane still reading when lane 0 poisons the stage):
    # seen in every run so far, but give it three runs before calling the model blind
    for attempt in range(3):
        r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, EMUL_UNION_FIRST_FAIL="1"))
        if r.returncode != 0 and expect in r.stdout + r.stderr:
            return
    assert r.returncode != 0
    assert expect in r.stdout + r.stderr, (r.stdout[-1500:], r.stderr[-500:])

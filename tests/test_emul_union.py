"""CPU-only: the bit-plane union kernels compiled as host code from the .inl sources of the GPU build and run through
the warp emulator tests/emul/cuda_emul.h — k_planes_from_bytes + k_pair_hist_planes (the default form of the union
pass) and k_split_build + k_pair_hist_split (SELB200_UNION=split) against the byte-wise definition of the union
histogram (sketch/include/sketch/hll.h:1191-1206), for precisions 9..16, narrow and wide value ranges, equal and
mixed bases, empty, saturated and overflowing high lists.  The emulator checks arithmetic, index math and the
producer/consumer walk; the hardware protocol (TMA, mbarriers) is covered by the -m gpu tests."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_union_kernels_on_the_warp_emulator(tmp_path):
    exe = tmp_path / "emul_union"
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "emul", "emul_union.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "all identical" in r.stdout
    assert r.stdout.count(" ok") >= 28 and "FAIL" not in r.stdout

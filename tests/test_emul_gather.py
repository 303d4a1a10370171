"""CPU-only: the peer-memory gather kernels of the multi-GPU path (csrc/kernels/gather.inl) compiled as host code and
run on the warp emulator (tests/emul/emul_gather.cpp, self-checking): runs back to back with ranks pushing in a
shuffled order (merged lists = multiset union of the ranks' lists, contiguous blocks, epoch parity, release),
an overflowed pass that must not be pushed, a landing zone that is too small, and the timeout of a buffer that was
never released.  The NVLink / CUDA-IPC side is covered by tests/test_gpu_gather.py on a GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gather_kernels_on_the_emulator(tmp_path):
    exe = tmp_path / "emul_gather"
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "emul", "emul_gather.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "all checks passed" in r.stdout and r.stdout.count(": ok") == 4

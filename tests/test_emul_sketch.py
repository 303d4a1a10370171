"""CPU-only: the sketch builder's kernel (csrc/kernels/sketch_kernels.inl — canonical 31-mers, WangHash, HLL registers,
SuperMinHash buckets) compiled as host code and run on the warp emulator (tests/emul/emul_sketch.cpp) over the
influenza FASTA files, held against the sketch files the REFERENCE's own build_sketch wrote next to them
(tests/golden/influenza: .hll, .smh4, .smh64, .hll_8) — byte-identical registers and buckets, without a GPU."""
import os
import struct
import subprocess

import numpy as np
import pytest

from cuda_selection_criteria_b200 import build_sketch as B, sketch_io
from cuda_selection_criteria_b200.selection import AUX_HLL, AUX_NONE, AUX_SMH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "influenza")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("emuls") / "emul_sketch"
    subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-Wno-unknown-pragmas",
                    os.path.join(ROOT, "tests", "emul", "emul_sketch.cpp"), "-o", str(out)], check=True)
    return str(out)


@pytest.fixture(scope="module")
def genomes():
    files = sketch_io.load_file_list(os.path.join(GOLD, "test_influeza_filelist.txt"))
    paths = [os.path.join(GOLD, f) for f in files]
    return paths, [B.read_fasta_clean(p) for p in paths]


def run(exe, tmp_path, seqs, p, aux_kind, aux_len):
    n = len(seqs)
    offsets = np.zeros(n + 1, np.int64)
    offsets[1:] = np.cumsum([len(s) for s in seqs])
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<4i", n, p, aux_kind, aux_len))
        f.write(offsets.tobytes())
        f.write(b"".join(seqs))
    r = subprocess.run([exe, str(inp), str(outp)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(outp, "rb").read()
    m = 1 << p
    hll = np.frombuffer(raw, np.uint8, n * m, 0).reshape(n, m)
    off = n * m
    if aux_kind == AUX_HLL:
        return hll, np.frombuffer(raw, np.uint8, n << aux_len, off).reshape(n, 1 << aux_len)
    if aux_kind == AUX_SMH:
        return hll, np.frombuffer(raw, np.uint64, n * aux_len, off).reshape(n, aux_len)
    return hll, None


@pytest.mark.parametrize("aux_kind,aux_len,suffix", [
    (AUX_SMH, 64, ".smh64"),       # build_sketch -a 512 -c smh_a
    (AUX_SMH, 4, ".smh4"),         # -a 32
    (AUX_HLL, 8, ".hll_8"),        # -a 256 -c hll_a
])
def test_builder_kernel_reproduces_the_reference_sketch_files(exe, genomes, tmp_path, aux_kind, aux_len, suffix):
    paths, seqs = genomes
    assert len(seqs) >= 5 and all(len(s) > 1000 for s in seqs)
    hll, aux = run(exe, tmp_path, seqs, 14, aux_kind, aux_len)
    for g, path in enumerate(paths):
        p, estim, jestim, value, regs = sketch_io.read_hll(path + ".hll")
        assert p == 14 and np.array_equal(hll[g], regs), path
        if aux_kind == AUX_SMH:
            assert np.array_equal(aux[g], sketch_io.read_smh(path + suffix)), path
        else:
            pa, _, _, _, ra = sketch_io.read_hll(path + suffix)
            assert pa == aux_len and np.array_equal(aux[g], ra), path


def test_builder_kernel_edge_sequences(exe, tmp_path):
    """Empty genome, a genome shorter than one k-mer, non-ACGT characters and record breaks restarting the rolling
    k-mer, lower case, a genome longer than one tile (512 x 32 positions): against a plain Python restatement of
    hll_t::add over canonical 31-mers (build_sketch.cpp:26-39,61-92; hll.h:886-894; hash.h:44-53)."""
    rng = np.random.default_rng(5)
    long_seq = bytes(rng.choice(list(b"ACGT"), 40000).tolist())
    seqs = [b"", b"ACGT" * 7, b"ACGTTGCATGCATGCAAGGTCCATGCATGGACTGACTGNACGTAGCTAGCTAGCTAGGATCGATCGATTTAGCGCGCATATAGCRYACGT" * 3,
            long_seq, long_seq[:20000].lower() + b"N" + long_seq[20000:]]
    hll, _ = run(exe, tmp_path, seqs, 14, AUX_NONE, 0)

    def wang(key):
        M = (1 << 64) - 1
        key = (~key + (key << 21)) & M
        key ^= key >> 24
        key = (key + (key << 3) + (key << 8)) & M
        key ^= key >> 14
        key = (key + (key << 2) + (key << 4)) & M
        key ^= key >> 28
        return (key + (key << 31)) & M

    code = {65: 0, 67: 1, 71: 2, 84: 3, 97: 0, 99: 1, 103: 2, 116: 3}
    for g, s in enumerate(seqs):
        want = np.zeros(1 << 14, np.uint8)
        fwd = rev = run_len = 0
        mask = (1 << 62) - 1
        for ch in s:
            c = code.get(ch)
            if c is None:
                fwd = rev = run_len = 0
                continue
            fwd = ((fwd << 2) | c) & mask
            rev = (rev >> 2) | ((3 - c) << 60)
            run_len = min(run_len + 1, 31)
            if run_len < 31:
                continue
            h = wang(min(fwd, rev))
            idx = h >> 50
            low = ((h << 1) | 1) << 13 & ((1 << 64) - 1)
            rank = 64 - low.bit_length() + 1
            want[idx] = max(want[idx], rank)
        assert np.array_equal(hll[g], want), g

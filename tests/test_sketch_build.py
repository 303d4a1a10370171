"""build_sketch: oracle pinned to the reference's shipped sketches (CPU), and the
CUDA builder against the oracle, the fixtures and the reference binary (-m gpu)."""
import ctypes as C
import glob
import gzip
import os
import subprocess

import numpy as np
import pytest

import oracle_api as O
import cuda_selection_criteria_b200 as S
from cuda_selection_criteria_b200 import build_sketch as B, sketch_io
from cuda_selection_criteria_b200.selection import AUX_HLL, AUX_NONE, AUX_SMH

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FIX = os.path.join(HERE, "golden", "influenza", "datasets", "test_influenzaA")
FASTAS = sorted(glob.glob(os.path.join(FIX, "*.fna.gz")))


def ora_hll(seq: bytes, p: int) -> np.ndarray:
    arr = np.frombuffer(seq, np.uint8) if seq else np.zeros(1, np.uint8)
    regs = np.zeros(1 << p, np.uint8)
    O.lib().oracle_sketch_hll(C.c_void_p(arr.ctypes.data), C.c_uint64(len(seq)), p, C.c_void_p(regs.ctypes.data))
    return regs


def ora_smh(seq: bytes, m_arg: int) -> np.ndarray:
    arr = np.frombuffer(seq, np.uint8) if seq else np.zeros(1, np.uint8)
    m = O.lib().oracle_smh_size(m_arg)
    h = np.zeros(m, np.uint64)
    O.lib().oracle_sketch_smh(C.c_void_p(arr.ctypes.data), C.c_uint64(len(seq)), m_arg, C.c_void_p(h.ctypes.data))
    return h


def random_genome(rng, n_records=3, length=4000):
    """FASTA text with lowercase stretches, Ns, IUPAC codes, a record shorter than k, wrapped lines."""
    recs = []
    for r in range(n_records):
        L = int(rng.integers(10, length))
        s = rng.choice(list(b"ACGT"), L).astype(np.uint8)
        for _ in range(int(rng.integers(0, 4))):
            pos = int(rng.integers(0, L)); s[pos:pos + int(rng.integers(1, 40))] = ord("N")
        for _ in range(2):
            pos = int(rng.integers(0, L)); s[pos:pos + 1] = rng.choice(list(b"RYKMU"))
        txt = bytes(s)
        if r % 2:
            txt = txt.lower()
        recs.append(txt)
    recs.append(b"ACGTACGTAC")       # shorter than k = 31
    fasta = b""
    for i, t in enumerate(recs):
        fasta += b">rec%d some description\n" % i
        fasta += b"\n".join(t[j:j + 70] for j in range(0, len(t), 70)) + b"\n"
    return fasta, b"N".join(recs)


# ---------------------------------------------------------------------------------- CPU: oracle pin
def test_oracle_sketches_equal_reference_fixtures():
    assert len(FASTAS) == 10
    for f in FASTAS:
        seq = B.read_fasta_clean(f)
        assert np.array_equal(ora_hll(seq, 14), sketch_io.read_hll(f + ".hll")[4])
        assert np.array_equal(ora_hll(seq, 8), sketch_io.read_hll(f + ".hll_8")[4])
        assert np.array_equal(ora_smh(seq, 4), sketch_io.read_smh(f + ".smh4"))
        assert np.array_equal(ora_smh(seq, 64), sketch_io.read_smh(f + ".smh64"))


def test_fasta_reader_joins_lines_and_separates_records(tmp_path):
    rng = np.random.default_rng(1)
    fasta, clean = random_genome(rng)
    (tmp_path / "a.fna").write_bytes(fasta)
    with gzip.open(tmp_path / "b.fna.gz", "wb") as f:
        f.write(fasta)
    assert B.read_fasta_clean(str(tmp_path / "a.fna")) == clean
    assert B.read_fasta_clean(str(tmp_path / "b.fna.gz")) == clean
    assert S.lib().selb200_smh_size(4) == 4 and S.lib().selb200_smh_size(5) == 8 and S.lib().selb200_smh_size(128) == 128


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "build_sketch")),
                    reason="oracle/_ref/build_sketch not built (needs /root/reference)")
def test_oracle_equals_live_reference_build_sketch(tmp_path):
    rng = np.random.default_rng(7)
    names = []
    cleans = []
    for i in range(4):
        fasta, clean = random_genome(rng, n_records=4, length=20000)
        nm = f"g{i}.fna.gz"
        with gzip.open(tmp_path / nm, "wb") as f:
            f.write(fasta)
        names.append(nm); cleans.append(clean)
    (tmp_path / "list.txt").write_text("\n".join(names) + "\n")
    ref = os.path.join(ROOT, "oracle", "_ref", "build_sketch")
    subprocess.run([ref, "-l", "list.txt", "-t", "2", "-a", "1024", "-c", "smh_a"], cwd=tmp_path, check=True)
    subprocess.run([ref, "-l", "list.txt", "-t", "2", "-a", "1024", "-c", "hll_a"], cwd=tmp_path, check=True)
    for nm, clean in zip(names, cleans):
        assert np.array_equal(ora_hll(clean, 14), sketch_io.read_hll(str(tmp_path / nm) + ".hll")[4])
        assert np.array_equal(ora_hll(clean, 10), sketch_io.read_hll(str(tmp_path / nm) + ".hll_10")[4])
        assert np.array_equal(ora_smh(clean, 128), sketch_io.read_smh(str(tmp_path / nm) + ".smh128"))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "build_sketch")),
                    reason="oracle/_ref/build_sketch not built (needs /root/reference)")
def test_parse_error_rule_matches_live_reference(tmp_path):
    """A non-IUPAC character stops the reference at that record (SeqAn ParseError -> break)."""
    rng = np.random.default_rng(11)
    r = [bytes(rng.choice(list(b"ACGT"), 3000).astype(np.uint8)) for _ in range(4)]
    fasta = b">a\n" + r[0] + b"\n>b x\n" + r[1][:1500] + b"\n" + r[1][1500:] + b"\n>c\n" + r[2][:100] + b"*" + r[2][100:] + b"\n>d\n" + r[3] + b"\n"
    with gzip.open(tmp_path / "g.fna.gz", "wb") as f:
        f.write(fasta)
    (tmp_path / "list.txt").write_text("g.fna.gz\n")
    clean = B.read_fasta_clean(str(tmp_path / "g.fna.gz"))
    assert clean == r[0] + b"N" + r[1]
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "build_sketch"), "-l", "list.txt", "-a", "512", "-c", "smh_a"],
                   cwd=tmp_path, check=True, capture_output=True)
    assert np.array_equal(ora_hll(clean, 14), sketch_io.read_hll(str(tmp_path / "g.fna.gz") + ".hll")[4])
    assert np.array_equal(ora_smh(clean, 64), sketch_io.read_smh(str(tmp_path / "g.fna.gz") + ".smh64"))


# ---------------------------------------------------------------------------------- GPU: the builder
@pytest.mark.gpu
def test_gpu_builder_regenerates_the_shipped_sketches(gpu, tmp_path):
    """build_sketch -a 512 -c smh_a / -a 32 -c smh_a / -a 256 -c hll_a over the 10 influenza genomes:
    every file equals the reference's fixture after gunzip."""
    lst = os.path.join(HERE, "golden", "influenza", "test_influeza_filelist.txt")
    base = os.path.join(HERE, "golden", "influenza")
    for aux_bytes, crit, suffix in ((512, "smh_a", ".smh64"), (32, "smh_a", ".smh4"), (256, "hll_a", ".hll_8")):
        out = tmp_path / f"{crit}{aux_bytes}"
        n = B.build_filelist(lst, aux_bytes=aux_bytes, criterion=crit, device=gpu, base=base, out_base=str(out))
        assert n == 10
        for f in sketch_io.load_file_list(lst):
            for sfx in (".hll", suffix):
                assert gzip.open(os.path.join(out, f) + sfx).read() == gzip.open(os.path.join(base, f) + sfx).read()


@pytest.mark.gpu
@pytest.mark.parametrize("aux_kind,aux_len", [(AUX_SMH, 128), (AUX_SMH, 5), (AUX_SMH, 512), (AUX_HLL, 10), (AUX_HLL, 4),
                                              (AUX_NONE, 0)])
def test_gpu_builder_equals_oracle_on_messy_fasta(gpu, aux_kind, aux_len):
    rng = np.random.default_rng(aux_len + 3)
    seqs = [random_genome(rng, n_records=5, length=60000)[1] for _ in range(6)]
    seqs.append(b"")                         # empty genome
    seqs.append(b"ACGT" * 5)                 # shorter than one k-mer
    seqs.append(b"A" * 200)                  # poly-A: canonical k-mer 0 -> RNG seed 1337 (wy.h:108)
    hll, aux = B.sketch_sequences(seqs, 14, aux_kind, aux_len, gpu)
    for i, s in enumerate(seqs):
        assert np.array_equal(hll[i], ora_hll(s, 14)), i
        if aux_kind == AUX_SMH:
            assert np.array_equal(aux[i], ora_smh(s, aux_len)), i
        elif aux_kind == AUX_HLL:
            assert np.array_equal(aux[i], ora_hll(s, aux_len)), i


@pytest.mark.gpu
def test_gpu_built_sketches_feed_the_selection(gpu, tmp_path):
    """End to end: sketches built on the GPU from the FASTA files select the pairs of results.txt."""
    base = os.path.join(HERE, "golden", "influenza")
    lst = os.path.join(base, "test_influeza_filelist.txt")
    B.build_filelist(lst, aux_bytes=512, criterion="smh_a", device=gpu, base=base, out_base=str(tmp_path))
    lines = S.run_filelist(lst, tau=0.9, aux_bytes=512, criterion="smh_a", device=gpu, base=str(tmp_path))
    assert lines == open(os.path.join(base, "results.txt")).read().splitlines()


@pytest.mark.gpu
@pytest.mark.parametrize("flags,suffix", [(["-a", "512", "-c", "smh_a"], ".smh64"), (["-a", "256", "-c", "hll_an", "-t", "3"], ".hll_8")])
def test_cpp_build_sketch_cli_regenerates_fixtures(gpu, tmp_path, flags, suffix):
    import shutil
    names = []
    for f in FASTAS:
        shutil.copy(f, tmp_path / os.path.basename(f))
        names.append(os.path.basename(f))
    (tmp_path / "list.txt").write_text("\n".join(names) + "\n")
    exe = os.path.join(ROOT, "cuda_selection_criteria_b200", "bin", "build_sketch")
    r = subprocess.run([exe, "-l", "list.txt"] + flags, cwd=tmp_path, capture_output=True, text=True, check=True)
    assert r.stdout == ""
    for nm in names:
        for sfx in (".hll", suffix):
            assert gzip.open(tmp_path / (nm + sfx)).read() == gzip.open(os.path.join(FIX, nm) + sfx).read()
    # invalid -c: primary sketches are still written, then the reference's message (build_sketch.cpp:290-292)
    r = subprocess.run([exe, "-l", "list.txt", "-c", "nope"], cwd=tmp_path, capture_output=True, text=True, check=True)
    assert r.stdout == "Option -c invalid. The accepted criteria are hll_a, hll_an and smh_a.\n"

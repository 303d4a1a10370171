"""CPU-only tests: the C-ABI library loads and exports what include/*.h declares, host logic, I/O,
and the multi-rank plumbing over gloo (world_size 2).  No compute call is made here."""
import ctypes as C
import gzip
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle_api as O
import cuda_selection_criteria_b200 as S
from cuda_selection_criteria_b200 import _lib, sketch_io, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "selb200.h")).read()
    declared = set(re.findall(r"\b(selb200_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = C.CDLL(S.lib_path())
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/selb200.h but not exported"
    assert declared == {n for n, _, _ in _lib.SYMBOLS}
    # link-level shims (include/selb200_shims.h) with the reference's mangled names
    for name in ("_Z17launch_kernel_smhPKhPKmPKdPK4int2idiiiiP6ResultPii",
                 "_Z19launch_kernel_CBsmhPKhPKmPKdPK4int2idiiiiP6ResultPii"):
        assert hasattr(L, name)
    assert S.lib().selb200_abi_version() == 1


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors vs the real C structs: sizes and a few offsets from a gcc-compiled probe."""
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "selb200.h"\nint main(void){'
                   'printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(selb200_params), sizeof(selb200_stats),'
                   'offsetof(selb200_params, sort_output), offsetof(selb200_stats, n_bands),'
                   'offsetof(selb200_stats, ms_bounds), offsetof(selb200_stats, ms_total));return 0;}')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.Params), C.sizeof(_lib.Stats), _lib.Params.sort_output.offset, _lib.Stats.n_bands.offset,
            _lib.Stats.ms_bounds.offset, _lib.Stats.ms_total.offset]
    assert got == want


def test_no_cpu_fallback():
    """Without a device every compute entry point fails loudly (this container has no GPU)."""
    if S.lib().selb200_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(S.SelB200Error) as e:
        S.Selection(0)
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "cuda_selection_criteria_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".inl", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "oracle_api" not in txt and "oracle/" not in txt, f


def test_band_params_equal_oracle_everywhere():
    for m in (1, 2, 3, 4, 8, 12, 32, 64, 100, 128, 256, 512):
        for tau in np.linspace(0.01, 0.999, 41):
            for v in (True, False):
                assert S.band_params(m, float(tau), v) == O.band_params(m, float(tau), v)


def test_sort_order_equals_oracle_with_ties():
    rng = np.random.default_rng(3)
    cards = rng.integers(0, 50, 5000).astype(np.float64)       # many ties: std::sort's unstable order matters
    a = np.empty(cards.size, np.int32)
    b = np.empty(cards.size, np.int32)
    assert S.lib().selb200_sort_order(cards.size, cards.ctypes.data, a.ctypes.data) == 0
    O.lib().oracle_sort_order(cards.size, cards.ctypes.data, b.ctypes.data)
    assert np.array_equal(a, b)
    assert np.all(np.diff(cards[a]) >= 0)


def test_sketch_io_roundtrip_and_reference_fixture_bytes(tmp_path):
    gold = os.path.join(ROOT, "tests", "golden", "influenza", "datasets", "test_influenzaA")
    name = sorted(f for f in os.listdir(gold) if f.endswith(".hll"))[0]
    p, estim, jestim, value, regs = sketch_io.read_hll(os.path.join(gold, name))
    assert (p, estim, jestim, value) == (14, 2, 2, -1.0)
    sketch_io.write_hll(str(tmp_path / "x.hll"), regs, 14)
    assert gzip.open(tmp_path / "x.hll").read() == gzip.open(os.path.join(gold, name)).read()
    smh = sketch_io.read_smh(os.path.join(gold, name[:-4] + ".smh64"))
    assert smh.size == 64
    sketch_io.write_smh(str(tmp_path / "x.smh64"), smh)
    assert gzip.open(tmp_path / "x.smh64").read() == gzip.open(os.path.join(gold, name[:-4] + ".smh64")).read()
    assert sketch_io.aux_suffix("smh_a", 512) == ".smh64" and sketch_io.aux_suffix("hll_a", 256) == ".hll_8"
    (tmp_path / "l.txt").write_text("  a.fna.gz \r\n\n\tb.fna.gz\n")
    assert sketch_io.load_file_list(str(tmp_path / "l.txt")) == ["a.fna.gz", "b.fna.gz"]
    with pytest.raises(FileNotFoundError):
        sketch_io.read_hll(str(tmp_path / "missing.hll"))


def test_synth_is_deterministic_and_follows_the_model():
    plan = synth.make_plan(200, 1001)
    a = synth.hll(plan, 14)
    b = synth.hll(plan, 14)
    assert np.array_equal(a, b) and a.max() <= 51
    # the sketched cardinality tracks the planned set size |core| + |private|
    want = plan.n_core[plan.cluster] + plan.n_priv
    got = np.array([O.cardinality(a[i], 14) for i in range(200)])
    assert np.median(np.abs(got - want) / want) < 0.02
    # a prefix plan reproduces the prefix of the full data (bench's CPU sample relies on it)
    assert np.array_equal(synth.hll(plan.head(50), 14), a[:50])
    s = synth.smh(plan, 128)
    assert np.array_equal(synth.smh(plan.head(50), 128), s[:50])
    # strangers share (almost) no buckets; cluster mates share many
    assert (s[0] == s[-1]).mean() < 0.05
    mates = [(i, i + 1) for i in range(199) if plan.cluster[i] == plan.cluster[i + 1]]
    assert np.mean([(s[i] == s[j]).mean() for i, j in mates]) > 0.3


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from cuda_selection_criteria_b200 import dist as sdist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
# sketch broadcast: every rank ends with rank 0's bytes
regs = torch.full((8, 16), rank + 1, dtype=torch.uint8)
aux = torch.full((8, 4), rank + 7, dtype=torch.int64)
sdist.broadcast_sketches(regs, aux, src=0)
assert int(regs.sum()) == 8 * 16 and int(aux.sum()) == 8 * 4 * 7
# per-rank host slices -> full matrices on every rank (all-gather in place), ragged last slice
n = 11
full = torch.arange(n * 16, dtype=torch.int64).reshape(n, 16)
g0, rows, per = sdist.slice_rows(n, rank, world)
assert sum(sdist.slice_rows(n, r, world)[1] for r in range(world)) == n
for chunks in (1, 2, 4):
    sh = sdist.ShardedSketches(n, 16, 4, torch.int64, "cpu", rank, world, chunks=chunks)
    r_all, a_all = sh.assemble(full[g0:g0 + rows].to(torch.uint8), full[g0:g0 + rows, :4].contiguous())
    f = sh.row_to_file
    assert sorted(f[f >= 0].tolist()) == list(range(n))            # every file lands in exactly one row
    real = torch.from_numpy(f >= 0)
    src = torch.from_numpy(np.where(f >= 0, f, 0))
    assert torch.equal(r_all[real], full.to(torch.uint8)[src][real]) and torch.equal(a_all[real], full[:, :4][src][real])
    assert int(r_all[~real].sum()) == 0                            # padding rows are empty sketches
# shard ranges tile the list exactly
T = 1001
tiles = sorted(t for r in range(world) for t in sdist.shard_tiles(T, r, world))
assert tiles == list(range(T))
# variable-length gather (unequal lengths), merged in (i,k) order
full = np.array([(i << 32) | k for i in range(40) for k in range(i + 1, 40, 7)], dtype=np.int64)
cut = full.size // 3
mine = full[:cut] if rank == 1 else full[cut:]
keys = torch.from_numpy(mine.copy()); jac = keys.to(torch.float64) * 0.5
k, j = sdist.gather_lists(keys, jac, dst=0)
if rank == 0:
    assert np.array_equal(k.numpy(), np.sort(full)) and np.array_equal(j.numpy(), np.sort(full) * 0.5)
    i32, k32 = sdist.split_keys(k)
    assert (i32 < k32).all()
else:
    assert k is None
# one rank empty, then all ranks empty
e = torch.zeros(0, dtype=torch.int64)
k, j = sdist.gather_lists(keys if rank == 0 else e, jac if rank == 0 else e.to(torch.float64), dst=0)
if rank == 0:
    assert k.numel() == mine.size
k, j = sdist.gather_lists(e, e.to(torch.float64), dst=0)
assert (k is None) or k.numel() == 0
dist.destroy_process_group()
print("OK", rank)
"""


def test_multi_rank_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29513", str(script), ROOT],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") == 2

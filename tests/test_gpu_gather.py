"""-m gpu: the peer-memory gather of the pair lists (include/selb200.h, selb200_gather_*).

Two ranks on ONE device are enough to exercise the protocol: the root's landing zone is reached through
the same pointer (one process, two contexts) or through a CUDA IPC mapping (two processes); on an
NVSwitch box the only difference is that the stores travel over NVLink.  The merged list on the
root must equal the unsharded run bit for bit (union of shards == 1-GPU list)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cuda_selection_criteria_b200 as S
from cuda_selection_criteria_b200 import synth
from cuda_selection_criteria_b200.selection import AUX_HLL, AUX_SMH

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _same(a, b):
    return (np.array_equal(a.i, b.i) and np.array_equal(a.k, b.k) and np.array_equal(a.jaccard, b.jaccard)
            and np.array_equal(np.sort(a.near_i), np.sort(b.near_i)))


@pytest.mark.parametrize("criterion,world", [("smh_a", 2), ("smh_a", 3), ("hll_a", 2), ("cb", 2)])
def test_gather_contexts_in_one_process(gpu, criterion, world):
    n = 700 if criterion == "cb" else 2500
    plan = synth.make_plan(n, 21)
    regs = synth.hll(plan, 14)
    aux, kind = None, 0
    if criterion == "smh_a":
        aux, kind = synth.smh(plan, 128), AUX_SMH
    elif criterion == "hll_a":
        aux, kind = synth.hll(plan, 10, synth.TAG_AUX_HLL), AUX_HLL
    tau = np.float32(0.85)
    ctxs = [S.Selection(gpu) for _ in range(world)]
    try:
        for c in ctxs:
            c.load(regs, aux, kind)
        whole = ctxs[0].run(tau=tau, criterion=criterion)
        assert whole.i.size > 0
        handle = ctxs[0].gather_create(1 << 20)
        for r, c in enumerate(ctxs):
            c.gather_attach(r, world, handle)
        # five runs: both buffer parities, and the "root merged run e-2" hand-back, get used
        for it in range(5):
            t = tau if it != 3 else np.float32(0.8)          # a different result size in between
            ref = whole if it != 3 else ctxs[0].run(tau=t, criterion=criterion)
            parts = [ctxs[r].run(tau=t, criterion=criterion, gather=True) for r in range(world - 1, 0, -1)]
            root = ctxs[0].run(tau=t, criterion=criterion, gather=True)
            assert all(p.i.size == 0 for p in parts)          # their pairs live on the root now
            assert _same(root, ref), it
            assert sum(p.stats["pairs_out"] for p in parts) + root.stats["pairs_out"] == ref.i.size
    finally:
        for c in ctxs:
            c.close()


def test_gather_needs_attach_and_matching_shard(gpu):
    plan = synth.make_plan(300, 2)
    regs, aux = synth.hll(plan, 14), synth.smh(plan, 128)
    with S.Selection(gpu) as sel:
        sel.load(regs, aux, AUX_SMH)
        with pytest.raises(RuntimeError):
            sel.run(gather=True)
        h = sel.gather_create(1024)
        sel.gather_attach(0, 1, h)
        one = sel.run(tau=np.float32(0.9), criterion="smh_a", gather=True)      # world of one: a plain run
        ref = sel.run(tau=np.float32(0.9), criterion="smh_a")
        assert _same(one, ref)
        with pytest.raises(ValueError):
            sel.gather_attach(0, 1, b"short")
        # a landing zone that is too small fails loudly
        h = sel.gather_create(1)
        sel.gather_attach(0, 1, h)
        if ref.i.size > 1:
            with pytest.raises(S.SelB200Error) as e:
                sel.run(tau=np.float32(0.9), criterion="smh_a", gather=True)
            assert "landing zone" in str(e.value)


_IPC_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
import cuda_selection_criteria_b200 as S
from cuda_selection_criteria_b200 import dist as sdist, synth
from cuda_selection_criteria_b200.selection import AUX_SMH
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")                 # two ranks on ONE GPU: NCCL refuses that, the gather does not need it
plan = synth.make_plan(3000, 1002)
regs, aux = synth.hll(plan, 14), synth.smh(plan, 128)
sel = S.Selection(0)
sel.load(regs, aux, AUX_SMH)
whole = sel.run(tau=np.float32(0.9), criterion="smh_a")
sdist.setup_gather(sel, 1 << 20)
for it in range(4):
    res = sel.run(tau=np.float32(0.9), criterion="smh_a", gather=True)
    if rank == 0:
        assert np.array_equal(res.i, whole.i) and np.array_equal(res.k, whole.k)
        assert np.array_equal(res.jaccard, whole.jaccard)
    else:
        assert res.i.size == 0 and res.stats["pairs_out"] > 0
dist.barrier()
sel.close()
dist.destroy_process_group()
print("OK", rank, whole.i.size)
"""


def test_gather_two_processes_cuda_ipc(gpu, tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_IPC_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script), ROOT],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == 2


def test_gather_after_multi_range_pass(gpu):
    """CB-only with more than 8 Mi band pairs per rank: the pass runs as several tile ranges, so the push into
    the root's landing zone is queued after the pass is known to be final (not the optimistic form)."""
    plan = synth.make_plan(9000, 5)
    regs = synth.hll(plan, 14, device=gpu)
    import torch
    torch.cuda.synchronize()
    tau = np.float32(0.5)
    ctxs = [S.Selection(gpu) for _ in range(2)]
    try:
        for c in ctxs:
            c.load(regs)
        whole = ctxs[0].run(tau=tau, criterion="cb")
        assert whole.stats["batches"] >= 2 and whole.stats["pairs_aux"] > (8 << 20)
        handle = ctxs[0].gather_create(1 << 21)
        for r, c in enumerate(ctxs):
            c.gather_attach(r, 2, handle)
        for _ in range(2):
            part = ctxs[1].run(tau=tau, criterion="cb", gather=True)
            root = ctxs[0].run(tau=tau, criterion="cb", gather=True)
            assert part.stats["batches"] >= 2 or root.stats["batches"] >= 2
            assert np.array_equal(root.i, whole.i) and np.array_equal(root.k, whole.k)
            assert np.array_equal(root.jaccard, whole.jaccard)
    finally:
        for c in ctxs:
            c.close()

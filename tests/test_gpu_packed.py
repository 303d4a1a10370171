"""-m gpu: packed transport of the register rows end to end (include/selb200.h "Packed transport").

selb200_load_host packs on the host, copies half the bytes and unpacks on the device; the result must be the load of
the same bytes from a device matrix (no packing involved) — cardinalities, order and the pair list of a run — also when
rows sit at the exception capacity, have to travel raw inside a piece, or force a whole piece to travel raw.  The
per-rank route (ShardedSketches.assemble with on_piece_packed -> selb200_load_device_rows_packed) is run with a world of
one: same pieces, same entry points, the all-gather is the only step left out."""
import numpy as np
import pytest
import torch

import cuda_selection_criteria_b200 as S
from cuda_selection_criteria_b200 import dist as sdist, synth
from cuda_selection_criteria_b200.selection import AUX_SMH

pytestmark = pytest.mark.gpu


def _odd_rows(regs, seed):
    """rows that stress the format: at the exception capacity, raw (33 far registers / no band at all), a piece with
    more raw rows than slots, empty and saturated registers"""
    rng = np.random.default_rng(seed)
    n, m = regs.shape
    far = int(regs.max()) + 16
    r = regs.copy()
    r[3, rng.choice(m, 32, replace=False)] = min(far, 51)
    r[5, rng.choice(m, 33, replace=False)] = min(far, 51)
    r[6, :] = rng.integers(0, 52, size=m)
    r[9, ::97] = 0
    r[n - 1, :] = 0                                         # an empty sketch
    if n > 2100:
        r[2050:2057, :] = rng.integers(0, 52, size=(7, m))  # second piece of 1024: 7 raw rows > 4 slots
    return r


@pytest.mark.parametrize("n,odd", [(2500, False), (3300, True)])
def test_packed_host_load_equals_device_load(gpu, n, odd):
    plan = synth.make_plan(n, 77)
    regs = synth.hll(plan, 14)
    aux = synth.smh(plan, 128)
    if odd:
        regs = _odd_rows(regs, 1)
    with S.Selection(gpu) as a, S.Selection(gpu) as b:
        a.load(regs, aux, AUX_SMH)                                       # host rows: packed transport
        info = a.load_info()
        assert info["rows_packed"] + info["rows_raw"] == n
        if odd:                                                          # the piece with seven raw rows went raw as a whole
            assert info["rows_raw"] >= 1024 and info["rows_packed"] >= 1024
        else:                                                            # pageable source: no piece goes raw to feed the link
            assert info["rows_raw"] == 0 and info["h2d_register_bytes"] < 0.53 * regs.size
        regs_d = torch.from_numpy(regs).to(f"cuda:{gpu}")
        aux_d = synth.smh(plan, 128, device=gpu)
        b.load(regs_d, aux_d, AUX_SMH)                                   # device matrix: no transport at all
        assert b.load_info() == {"h2d_register_bytes": 0, "rows_packed": 0, "rows_raw": 0}
        ca, oa = a.order()
        cb, ob = b.order()
        assert np.array_equal(ca, cb) and np.array_equal(oa, ob)
        ra = a.run(tau=np.float32(0.85), criterion="smh_a")
        rb = b.run(tau=np.float32(0.85), criterion="smh_a")
        assert ra.i.size > 100
        assert np.array_equal(ra.i, rb.i) and np.array_equal(ra.k, rb.k) and np.array_equal(ra.jaccard, rb.jaccard)


def test_packed_load_still_validates_registers(gpu):
    plan = synth.make_plan(600, 8)
    regs = synth.hll(plan, 14).copy()
    regs[17, 1234] = 52                                                   # 64 - p + 2: not a register of a p = 14 sketch
    with S.Selection(gpu) as sel:
        with pytest.raises(S.SelB200Error) as e:
            sel.load(regs)
        assert "register value 52" in str(e.value)


@pytest.mark.parametrize("chunks", [1, 3])
def test_sharded_assemble_packed_world_of_one(gpu, chunks, monkeypatch):
    monkeypatch.setattr(sdist.ShardedSketches, "SUB_ROWS_BYTES", 300 << 14)       # several small pieces per piece, a short last one
    n = 2100
    plan = synth.make_plan(n, 12)
    regs = _odd_rows(synth.hll(plan, 14), 2)
    regs[2050:2057] = regs[100:107]                                      # keep every piece within its four raw slots
    aux = synth.smh(plan, 128)
    dev = torch.device("cuda", gpu)
    regs_h = torch.from_numpy(regs).pin_memory()
    aux_h = torch.from_numpy(aux.view(np.int64)).pin_memory()
    # one explicit stream for the torch copies and the library's kernels: load_device_rows(_packed) declares rows complete
    # in the order of the CONTEXT's stream, so the context must run on the stream the copies are queued behind
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        sh = sdist.ShardedSketches(n, regs.shape[1], aux.shape[1], torch.int64, dev, 0, 1, chunks=chunks)
        with S.Selection(gpu, stream=st.cuda_stream) as a, S.Selection(gpu) as b:
            a.load_device_begin(sh.regs, sh.aux, AUX_SMH)
            calls = []
            sh.assemble(regs_h, aux_h, on_piece=a.load_device_rows,
                        on_piece_packed=lambda g0, cnt, pieces, pr: (calls.append((g0, cnt)), a.load_device_rows_packed(g0, cnt, pieces, pr)))
            a.load_end()
            assert len(calls) == chunks and sum(c for _, c in calls) == sh.n_dev
            # the device matrix holds the host rows again (padding rows: all zero)
            got = sh.regs.cpu().numpy()
            keep = sh.row_to_file >= 0
            assert np.array_equal(got[keep], regs[sh.row_to_file[keep]]) and not got[~keep].any()
            b.load(regs, aux, AUX_SMH)
            ra = a.run(tau=np.float32(0.85), criterion="smh_a")
            rb = b.run(tau=np.float32(0.85), criterion="smh_a")
            # padding rows shift sorted positions: compare in file indices
            fa = np.sort(np.stack([sh.row_to_file[a.order()[1][ra.i]], sh.row_to_file[a.order()[1][ra.k]]], 1), axis=1)
            fb = np.sort(np.stack([b.order()[1][rb.i], b.order()[1][rb.k]], 1), axis=1)
            ka = np.lexsort((fa[:, 1], fa[:, 0]))
            kb = np.lexsort((fb[:, 1], fb[:, 0]))
            assert rb.i.size > 100 and np.array_equal(fa[ka], fb[kb]) and np.array_equal(ra.jaccard[ka], rb.jaccard[kb])

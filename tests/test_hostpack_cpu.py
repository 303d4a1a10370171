"""Packed transport of register rows, host side (csrc/hostpack.h through the C-ABI: selb200_nib4_piece_bytes /
selb200_nib4_pack_piece).  The piece is decoded here in numpy straight from the documented layout and must give back
the bytes that went in: natural sketches, rows at the exception capacity, rows that have to stay raw, more raw rows
than a piece has slots for.  No GPU involved (the device side is tests/test_gpu_packed.py)."""
import ctypes as C

import numpy as np
import pytest

from cuda_selection_criteria_b200 import _lib, synth

EXC_CAP, RAW_CAP = 32, 4


def up256(x):
    return (x + 255) & ~255


def layout(rows, m):
    off_exc = up256(rows * 4)
    off_rawidx = off_exc + up256(rows * EXC_CAP * 4)
    off_nib = off_rawidx + 256
    off_raw = off_nib + up256(rows * (m // 2))
    return off_exc, off_rawidx, off_nib, off_raw, off_raw + RAW_CAP * m


def pack(regs, threads=2):
    L = _lib.lib()
    rows, m = regs.shape
    p = m.bit_length() - 1
    nbytes = int(L.selb200_nib4_piece_bytes(rows, p))
    assert nbytes == layout(rows, m)[4]
    piece = np.full(nbytes, 0xAB, np.uint8)
    n_raw = int(L.selb200_nib4_pack_piece(regs.ctypes.data, rows, p, piece.ctypes.data, threads))
    return piece, n_raw


def unpack(piece, rows, m):
    off_exc, off_rawidx, off_nib, off_raw, _ = layout(rows, m)
    hdr = piece[:rows * 4].reshape(rows, 4)
    base, raw = hdr[:, 0].astype(np.int64), hdr[:, 1]
    n_exc = hdr[:, 2].astype(np.int64) | (hdr[:, 3].astype(np.int64) << 8)
    nib = piece[off_nib:off_nib + rows * (m // 2)].reshape(rows, m // 2)
    out = np.empty((rows, m), np.int64)
    out[:, 0::2] = nib & 15
    out[:, 1::2] = nib >> 4
    out += base[:, None]
    exc = piece[off_exc:off_exc + rows * EXC_CAP * 4].view(np.uint32).reshape(rows, EXC_CAP)
    for g in range(rows):
        for e in range(n_exc[g]):
            out[g, exc[g, e] >> 8] = exc[g, e] & 0xff
    raw_idx = piece[off_rawidx:off_rawidx + 256].view(np.int32)
    for r in range(RAW_CAP):
        if raw_idx[r] >= 0:
            out[raw_idx[r]] = piece[off_raw + r * m:off_raw + (r + 1) * m]
    assert np.all(raw_idx[RAW_CAP:] == -1)
    return out.astype(np.uint8), raw, n_exc, raw_idx[:RAW_CAP]


@pytest.mark.parametrize("p,rows", [(14, 300), (10, 77), (9, 1)])
def test_natural_sketches_round_trip(p, rows):
    plan = synth.make_plan(rows, 40 + p)
    regs = synth.hll(plan, 14)[:, :1 << p].copy()
    piece, n_raw = pack(regs)
    assert n_raw == 0
    back, raw, n_exc, raw_idx = unpack(piece, rows, 1 << p)
    assert np.array_equal(back, regs)
    assert not raw.any() and np.all(raw_idx == -1)
    assert n_exc.max() <= EXC_CAP and (p < 14 or n_exc.sum() > 0)       # real sketches do have a few far registers
    # half the bytes at the precision the path runs at (the fixed 128 B of exception slots weigh more on short rows):
    # what crosses the bus is [0, off_raw)
    if p == 14:
        assert layout(rows, 1 << p)[3] < 0.52 * regs.size


def test_exception_capacity_raw_rows_and_overflow():
    rng = np.random.default_rng(5)
    rows, m = 64, 1 << 12
    regs = rng.integers(9, 20, size=(rows, m), dtype=np.uint8)           # band of 11 values: nothing far away
    regs[:, 7] = 9
    regs[3, rng.choice(m, EXC_CAP, replace=False)] = 40                   # exactly at the capacity: still packed
    regs[5, rng.choice(m, EXC_CAP + 1, replace=False)] = 41               # one more: raw
    regs[6, :] = rng.integers(0, 52, size=m)                              # not a sketch of anything: raw
    regs[9, 0] = 0                                                        # base 0, everything else 9+: 15+ above -> raw
    regs[11, 100] = 255                                                   # a byte no HLL holds travels unchanged (validated on the device)
    piece, n_raw = pack(regs)
    assert n_raw == 3
    back, raw, n_exc, raw_idx = unpack(piece, rows, m)
    assert np.array_equal(back, regs)
    assert list(np.nonzero(raw)[0]) == [5, 6, 9] and sorted(raw_idx[raw_idx >= 0]) == [5, 6, 9]
    assert n_exc[3] == EXC_CAP and n_exc[11] == 1
    # more raw rows than slots: the count says so and the caller must not use the piece
    regs[20:26, :] = rng.integers(0, 52, size=(6, m))
    piece, n_raw = pack(regs)
    assert n_raw == 9 > RAW_CAP


def test_thread_counts_agree():
    plan = synth.make_plan(200, 3)
    regs = synth.hll(plan, 14)
    a, _ = pack(regs, threads=1)
    b, _ = pack(regs, threads=5)
    used = layout(200, 1 << 14)[3]
    assert np.array_equal(a[:used], b[:used])


def test_bad_arguments():
    L = _lib.lib()
    assert L.selb200_nib4_piece_bytes(10, 8) < 0 and L.selb200_nib4_piece_bytes(-1, 14) < 0
    assert L.selb200_nib4_pack_piece(None, 3, 14, None, 1) < 0

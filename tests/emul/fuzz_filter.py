"""Offline differential fuzz of the emulated filter chains (tests/emul/emul_filter*.cpp: the CB + smh_a / hll_a / hll_an
kernels compiled as host code) against the oracle's per-pair decisions: random n, tau, sketch sizes, shards, grids,
empty sketches, cardinality ties, outlier registers.  usage: python tests/emul/fuzz_filter.py SEED COUNT
The union kernels have the same in tests/emul/emul_union.cpp:  emul_union LIB fuzz SEED COUNT."""
import os, sys, struct, subprocess, tempfile, pathlib, random
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import oracle_api as O
from cuda_selection_criteria_b200 import synth
import test_emul_filter as T


tmp = pathlib.Path(tempfile.mkdtemp())
exe = str(tmp / 'emul_filter'); exe_h = str(tmp / 'emul_filter_hll')
for src, out in (('emul_filter.cpp', exe), ('emul_filter_hll.cpp', exe_h)):
    subprocess.run(['g++', '-O2', '-std=c++20', '-pthread', '-ffp-contract=off', '-Wno-unknown-pragmas',
                    os.path.join(ROOT, 'tests', 'emul', src), '-o', out], check=True)
rnd = random.Random(int(sys.argv[1]))
fails = 0
for it in range(int(sys.argv[2])):
    n = rnd.randint(2, 420); seed = rnd.randint(1, 10**6)
    tau32 = np.float32(rnd.choice([0.5, 0.7, 0.8, 0.9, 0.95, 0.99, rnd.uniform(0.3, 0.999)]))
    plan = synth.make_plan(n, seed)
    regs = synth.hll(plan, 14)
    cards = np.array([O.cardinality(regs[g], 14) for g in range(n)])
    if rnd.random() < 0.5: cards[::rnd.randint(2, 50)] = 0.0
    if rnd.random() < 0.2: cards[:] = np.round(cards, -3)        # ties in e
    order = np.argsort(cards, kind='stable'); e = cards[order].astype(np.uint64)
    shards = rnd.randint(1, 4); grid = rnd.randint(1, 6)
    d = tmp / f'c{it}'; d.mkdir()
    try:
        if rnd.random() < 0.5:
            m_aux = rnd.choice([1, 2, 4, 8, 16, 32, 64, 128, 256])
            smh = synth.smh(plan, m_aux)
            aux_sorted = np.ascontiguousarray(smh[order])
            nb, nr = O.band_params(m_aux, tau32)
            lo, hi, p_cb, tiles, cand, pairs = T.run_emulated(exe, d, e, aux_sorted, tau32, nr, nb, shards, grid)
            olo, ohi, op_cb, opairs = T.oracle_decisions(e, aux_sorted, tau32, nr, nb)
            ok = p_cb == op_cb and pairs == opairs
            what = f'smh n={n} seed={seed} tau={tau32} m={m_aux} {nb}x{nr} shards={shards} grid={grid}'
        else:
            p_aux = rnd.choice([4, 5, 6, 7, 8, 9, 10]); an = rnd.random() < 0.5
            form = 'planes' if p_aux >= 6 and rnd.random() < 0.7 else 'bytes'
            aux = synth.hll(plan, p_aux, synth.TAG_AUX_HLL).copy()
            if rnd.random() < 0.4: aux[::rnd.randint(2, 9), rnd.randint(0, (1 << p_aux) - 1)] = 64 - p_aux + 1
            aux_sorted = np.ascontiguousarray(aux[order])
            zs = np.float32(1.96) * np.float32(O.lib().oracle_sigma(p_aux))
            inp, outp = d / 'in.bin', d / 'out.bin'
            with open(inp, 'wb') as f:
                f.write(struct.pack('<5i', n, p_aux, int(an), 1, int(np.count_nonzero(e == 0))))
                f.write(struct.pack('<d', float(tau32))); f.write(struct.pack('<f', float(zs)))
                f.write(e.tobytes()); f.write(aux_sorted.tobytes())
            r = subprocess.run([exe_h, str(inp), str(outp), form, str(shards), str(grid)], capture_output=True, text=True, timeout=900)
            assert r.returncode == 0, r.stdout + r.stderr
            raw = open(outp, 'rb').read()
            p_cb, npairs = struct.unpack_from('<2q', raw, 0)
            pairs = [tuple(x) for x in np.frombuffer(raw, np.uint32, 2 * npairs, 16).reshape(-1, 2).tolist()]
            op_cb, opairs = T.oracle_hll_decisions(e, aux_sorted, p_aux, tau32, an, 1)
            ok = p_cb == op_cb and pairs == opairs
            what = f'hll n={n} seed={seed} tau={tau32} p_aux={p_aux} an={an} {form} shards={shards} grid={grid}'
    except Exception as ex:
        ok = False; what = f'EXC {ex!r} n={n} seed={seed}'
    if not ok:
        fails += 1
        print('FAIL', what, flush=True)
    elif it % 10 == 0:
        print('ok', it, what, flush=True)
print('done, failures:', fails)

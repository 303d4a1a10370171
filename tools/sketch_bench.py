#!/usr/bin/env python
"""Sketch-builder measurement: the CUDA builder vs the reference's build_sketch.

Random genomes (default 64 x 5 Mbp, one record each, 70-column FASTA, gzip level 1) are written to a
temp dir; timed: (1) the reference binary oracle/_ref/build_sketch -t nproc, (2) our C++ CLI
bin/build_sketch -t nproc (inflate + de-line on the host, sketches on the GPU), (3) the device call
selb200_sketch_host alone on in-memory sequences (H2D + kernel + D2H).  Outputs are compared
byte for byte after gunzip.  One JSON line on stdout.
"""
import argparse, gzip, json, os, shutil, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=64)
    ap.add_argument("--mbp", type=float, default=5.0)
    ap.add_argument("--aux-bytes", type=int, default=1024)
    ap.add_argument("--criterion", default="smh_a")
    a = ap.parse_args()
    from cuda_selection_criteria_b200 import build_sketch as B
    from cuda_selection_criteria_b200.selection import AUX_HLL, AUX_SMH
    cores = os.cpu_count() or 1
    td = tempfile.mkdtemp(prefix="selb200_sk_")
    try:
        rng = np.random.default_rng(5)
        L = int(a.mbp * 1e6)
        names, seqs = [], []
        lut = np.frombuffer(b"ACGT", np.uint8)
        for g in range(a.genomes):
            s = lut[rng.integers(0, 4, L)]
            body = s.reshape(-1, 70) if L % 70 == 0 else None
            txt = bytes(s)
            fasta = b">g%d\n" % g + b"\n".join(txt[j:j + 70] for j in range(0, L, 70)) + b"\n"
            nm = f"g{g:04d}.fna.gz"
            with gzip.open(os.path.join(td, nm), "wb", compresslevel=1) as f:
                f.write(fasta)
            names.append(nm); seqs.append(txt)
        open(os.path.join(td, "list.txt"), "w").write("\n".join(names) + "\n")
        res = {"genomes": a.genomes, "mbp_per_genome": a.mbp, "criterion": a.criterion, "aux_bytes": a.aux_bytes, "cores": cores}
        flags = ["-l", "list.txt", "-t", str(cores), "-a", str(a.aux_bytes), "-c", a.criterion]
        ref = os.path.join(ROOT, "oracle", "_ref", "build_sketch")
        sfx = ".smh%d" % (a.aux_bytes // 8) if a.criterion == "smh_a" else ".hll_%d" % (a.aux_bytes.bit_length() - 1)
        keep = {}
        if os.path.exists(ref):
            t0 = time.perf_counter(); subprocess.run([ref] + flags, cwd=td, check=True); t = time.perf_counter() - t0
            res["reference_s"] = t; res["reference_mbp_per_s"] = a.genomes * a.mbp / t
            for nm in names:
                keep[nm] = (gzip.open(os.path.join(td, nm + ".hll")).read(), gzip.open(os.path.join(td, nm + sfx)).read())
                os.unlink(os.path.join(td, nm + ".hll")); os.unlink(os.path.join(td, nm + sfx))
        exe = os.path.join(ROOT, "cuda_selection_criteria_b200", "bin", "build_sketch")
        subprocess.run([exe] + flags, cwd=td, check=True)        # warm-up (CUDA context, page cache)
        t0 = time.perf_counter(); subprocess.run([exe] + flags, cwd=td, check=True); t = time.perf_counter() - t0
        res["cli_s"] = t; res["cli_mbp_per_s"] = a.genomes * a.mbp / t
        if keep:
            res["identical_to_reference"] = all(
                keep[nm] == (gzip.open(os.path.join(td, nm + ".hll")).read(), gzip.open(os.path.join(td, nm + sfx)).read())
                for nm in names)
        kind, alen = (AUX_SMH, a.aux_bytes // 8) if a.criterion == "smh_a" else (AUX_HLL, a.aux_bytes.bit_length() - 1)
        B.sketch_sequences(seqs[:2], 14, kind, alen)
        t0 = time.perf_counter(); B.sketch_sequences(seqs, 14, kind, alen); t = time.perf_counter() - t0
        res["device_call_s"] = t; res["device_call_mbp_per_s"] = a.genomes * a.mbp / t
        print(json.dumps(res))
    finally:
        shutil.rmtree(td, ignore_errors=True)


if __name__ == "__main__":
    main()

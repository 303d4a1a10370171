"""Static instruction mix of the innermost loops of a kernel, read from the built library (cuobjdump -sass): for every
backward branch whose body holds no other backward branch, the body's opcode counts.  Forward branches inside a body
(warp-uniform skips) are not followed, so the numbers are those of the all-paths-taken iteration.  Used to compare two
forms of a kernel when no GPU is at hand:  python tools/sass_loops.py k_pair_hist_planes [min_instructions]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cuda_selection_criteria_b200", "libselb200.so")


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if cur and m:
            d[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return d


def opcode(ins):
    parts = ins.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    return op.split(".")[0]


def main():
    pat = sys.argv[1]
    min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    for name, ins in functions(LIB).items():
        if pat not in name:
            continue
        short = re.sub(r"_GLOBAL__N__[0-9a-f]+_[0-9]+_[A-Za-z0-9_]+?_cu_[0-9a-f]+", "", name)
        back = []
        for addr, text in ins:
            m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) <= addr:
                back.append((int(m.group(1), 16), addr))
        inner = [(a, b) for a, b in back if not any((a2, b2) != (a, b) and a <= a2 and b2 <= b for a2, b2 in back)]
        print(f"{short}: {len(ins)} instructions, {len(inner)} innermost loops")
        for a, b in inner:
            body = [t for ad, t in ins if a <= ad <= b]
            if len(body) < min_len:
                continue
            c = collections.Counter(opcode(t) for t in body)
            alu = sum(c[k] for k in ("LOP3", "PRMT", "SEL", "ISETP", "IMNMX", "VIMNMX", "SHF", "IADD3", "IADD", "LEA"))
            print(f"  loop 0x{a:04x}..0x{b:04x}: {len(body):4d} instr  LOP3 {c['LOP3']:3d}  POPC {c['POPC']:2d}  IADD3/IADD {c['IADD3'] + c['IADD']:2d}"
                  f"  IMAD {c['IMAD']:2d}  LDS {c['LDS']:2d}  BRA {c['BRA']:2d}  ALU-pipe {alu:3d}")


if __name__ == "__main__":
    main()

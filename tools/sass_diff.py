"""Device-code identity check between a git revision and the working tree: builds the revision's library in a
temporary worktree, dumps the SASS of both libraries (cuobjdump), drops addresses and the per-build hash of the
anonymous namespace, and compares kernel by kernel.  Use after source-only refactors (moves, guards, comments) made
when no GPU is at hand:  python tools/sass_diff.py <rev>      (nvcc cross-compiles; no GPU needed)"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join("cuda_selection_criteria_b200", "libselb200.so")


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_[0-9]+_[A-Za-z0-9_]+?_cu_[0-9a-f]+", "_ANON_", m.group(1))
            d[cur] = []
        elif cur and ";" in line and not line.strip().startswith("."):
            d[cur].append(re.sub(r"\s+", " ", re.sub(r"/\*[0-9a-f]{4,}\*/", "", line).strip()))
    return d


def main():
    rev = sys.argv[1]
    with tempfile.TemporaryDirectory() as td:
        wt = os.path.join(td, "wt")
        subprocess.run(["git", "-C", ROOT, "worktree", "add", "-q", wt, rev], check=True)
        try:
            subprocess.run(["make", "-C", os.path.join(wt, "cuda_selection_criteria_b200", "csrc"), "../libselb200.so"],
                           check=True, capture_output=True)
            old = kernels(os.path.join(wt, LIB))
        finally:
            subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", wt], check=True)
    subprocess.run(["make", "-C", os.path.join(ROOT, "cuda_selection_criteria_b200", "csrc"), "../libselb200.so"],
                   check=True, capture_output=True)
    new = kernels(os.path.join(ROOT, LIB))
    same = [k for k in old if old[k] == new.get(k)]
    diff = [k for k in old if k in new and old[k] != new[k]]
    print(f"{rev}: {len(old)} kernels, working tree: {len(new)}; identical SASS: {len(same)}, different: {len(diff)}")
    for k in diff:
        print("  DIFFERENT", k, len(old[k]), "->", len(new[k]), "instructions")
    for k in new:
        if k not in old:
            print("  NEW      ", k, len(new[k]), "instructions")
    for k in old:
        if k not in new:
            print("  GONE     ", k)
    return 1 if diff else 0


if __name__ == "__main__":
    sys.exit(main())

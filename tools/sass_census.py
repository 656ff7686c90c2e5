"""SASS / ptxas census of the shipped libsedb200.so -- needs no GPU.

    python tools/sass_census.py > profiles/r02_sass_census.csv

Per kernel: how many tcgen05 (UTCHMMA = kind::f16, UTCQMMA = kind::f8f6f4), TMEM (LDTM / STTM), TMA (UTMALDG /
UTMASTG) and mbarrier (SYNCS) instructions the SASS holds, and ptxas' registers / spill bytes / static shared memory
from the build logs (`sed_crnn_b200/csrc/build/*.cu.log`, written by `sed_crnn_b200/build.py` with -Xptxas -v).
The mnemonics are the ones /opt/skills/guides/B200_PROFILING.md names as the proof of tcgen05 / TMA code.
"""
from __future__ import annotations

import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sed_crnn_b200", "libsedb200.so")
LOGS = os.path.join(ROOT, "sed_crnn_b200", "csrc", "build", "*.cu.log")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "ELECT", "FFMA2", "MUFU",
             "HMMA"]


def demangle(names: list[str]) -> dict[str, str]:
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name: str) -> str:
    """`void sedb200::(anonymous namespace)::k<1, 2>(args...)` -> `k<1, 2>`"""
    name = re.sub(r"^void ", "", name)
    name = name.replace("sedb200::(anonymous namespace)::", "").replace("sedb200::", "")
    depth, cut = 0, len(name)
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    return name[:cut]


def ptxas_table() -> dict[str, tuple[int, int, int]]:
    """mangled name -> (registers, spill store + load bytes, static smem bytes)"""
    tab = {}
    for log in glob.glob(LOGS):
        cur = None
        spill = 0
        for line in open(log):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                cur, spill = m.group(1), 0
                continue
            m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                spill = int(m.group(1)) + int(m.group(2))
            m = re.search(r"Used (\d+) registers", line)
            if m and cur:
                sm = re.search(r"(\d+) bytes smem", line)
                tab[cur] = (int(m.group(1)), spill, int(sm.group(1)) if sm else 0)
    return tab


def main() -> None:
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    pat = re.compile(r"\b(" + "|".join(MNEMONICS) + r")\b")
    counts: dict[str, collections.Counter] = collections.defaultdict(collections.Counter)
    n_instr: collections.Counter = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            n_instr[cur] += 1
            m = pat.search(line)
            if m:
                counts[cur][m.group(1)] += 1
    regs = ptxas_table()
    names = demangle(sorted(n_instr))
    w = sys.stdout.write
    w("kernel,sass_instructions,registers,spill_bytes,static_smem_bytes," + ",".join(MNEMONICS) + "\n")
    rows = sorted(n_instr, key=lambda k: (-(counts[k]["UTCHMMA"] + counts[k]["UTCQMMA"]), short(names[k])))
    for k in rows:
        r = regs.get(k, ("", "", ""))
        w(f"\"{short(names[k])}\",{n_instr[k]},{r[0]},{r[1]},{r[2]}," + ",".join(str(counts[k][m]) for m in MNEMONICS) + "\n")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    w(f"\"TOTAL ({len(rows)} kernels)\",{sum(n_instr.values())},,,," + ",".join(str(tot[m]) for m in MNEMONICS) + "\n")


if __name__ == "__main__":
    main()

"""Per-kernel SASS mnemonic counts of the built objects (cuobjdump -sass): the proof that the tcgen05 / TMA / TMEM paths
are what was compiled (UTCIMMA = tcgen05.mma kind::i8, UTCHMMA = kind::f16/tf32, UTMALDG / UTMASTG = TMA tensor load / store,
LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier).  Also ptxas resource usage per kernel.
Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sdformerflow_b200", "_lib")
MNEMONICS = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS",
             "HMMA", "IMMA", "LDGSTS", "LDG", "STG", "LDS", "STS", "RED", "ATOM"]


def main():
    print("# cuobjdump -sass mnemonic counts per kernel (objects built by sdformerflow_b200/build.py, sm_100a)\n")
    for obj in sorted(f for f in os.listdir(LIB) if f.endswith(".o")):
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(LIB, obj)], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", os.path.join(LIB, obj)], capture_output=True, text=True).stdout
        usage = {}
        name = None
        for line in res.splitlines():
            m = re.search(r"Function (\S+):", line)
            if m:
                name = m.group(1)
            elif name and "REG:" in line:
                usage[name] = line.strip()
                name = None
        kernels = collections.OrderedDict()
        cur = None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = kernels.setdefault(m.group(1), collections.Counter())
                continue
            if cur is None:
                continue
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                op = m.group(1)
                cur["_total"] += 1
                for mn in MNEMONICS:
                    if op == mn or op.startswith(mn + "."):
                        cur[mn] += 1
        print(f"## {obj}")
        for k, c in kernels.items():
            dem = subprocess.run(["cu++filt", k], capture_output=True, text=True).stdout.strip() or k
            i = dem.rfind(">(")
            dem = (dem[:i + 1] if i > 0 else dem.split("(")[0])[:150]
            cnt = " ".join(f"{mn}={c[mn]}" for mn in MNEMONICS if c[mn])
            print(f"- `{dem}`: {c['_total']} instr; {cnt}")
            if k in usage:
                print(f"    {usage[k]}")
        print()


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Per-kernel SASS inventory of libbrainseg_b200.so (cuobjdump -sass): the mnemonics that prove the Blackwell-native
path (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG, tcgen05.commit /
mbarrier traffic -> UTCBAR / SYNCS).  Writes a markdown table.

    python scripts/sass_inventory.py > profiles/r02_sass_inventory.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "brainseg_b200", "libbrainseg_b200.so")
PATTERNS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\S*\.2CTA"), ("UTMALDG", r"\bUTMALDG"),
            ("UTMALDG.MULTICAST", r"\bUTMALDG\S*MULTICAST"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"),
            ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("HMMA (legacy)", r"\bHMMA"),
            ("STG.E.ENL2.256", r"\bSTG\.E\.ENL2\.256"), ("ATOM/RED", r"\b(ATOMG|ATOMS|RED|REDG)\b")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            kernels[cur]["_instr"] = 0
            continue
        if cur is None or "/*" not in line:
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", line):
            kernels[cur]["_instr"] += 1
            for name, pat in PATTERNS:
                if re.search(pat, line):
                    kernels[cur][name] += 1
    demangled = {}
    names = list(kernels)
    try:
        dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangled = dict(zip(names, dm))
    except Exception:
        pass
    cols = [n for n, _ in PATTERNS]
    total = collections.Counter()
    print("# SASS inventory of `brainseg_b200/libbrainseg_b200.so`\n")
    print("`cuobjdump -sass` of the in-tree library, built with `-gencode arch=compute_100a,code=sm_100a`; produced by "
          "`scripts/sass_inventory.py`.  `UTCHMMA` = `tcgen05.mma.kind::f16`, `LDTM` = `tcgen05.ld`, `UTMALDG` / "
          "`UTMASTG` = TMA tensor loads / stores, `UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier operations.\n")
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for k, c in kernels.items():
        name = demangled.get(k, k)
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*\)$", "", name)
        print(f"| `{name}` | {c['_instr']} | " + " | ".join(str(c[n]) if c[n] else "" for n in cols) + " |")
        total.update(c)
    print(f"| **total ({len(kernels)} kernels)** | {total['_instr']} | " + " | ".join(str(total[n]) for n in cols) + " |")


if __name__ == "__main__":
    sys.exit(main())

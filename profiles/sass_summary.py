#!/usr/bin/env python
"""Per-kernel SASS opcode summary of the shipped library (evidence that the contractions are tcgen05 / TMEM / TMA):
    python profiles/sass_summary.py > profiles/<round>_sass_opcodes.txt
UTCHMMA = tcgen05.mma (bf16), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA load/store/reduce,
UTCBAR = tcgen05.commit, HMMA = legacy mma.sync (must be 0)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vjepa2_b200", "libvjepa2_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "HMMA", "MUFU.EX2", "SYNCS", "LDG", "STG",
       "RED", "ATOM", "FFMA2", "FMUL2", "FADD2"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            cnt[cur]["_total"] += 1
            for o in OPS:
                if m.group(1).startswith(o):
                    cnt[cur][o] += 1
    names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
    print(f"SASS opcode summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a): instructions per kernel")
    tot = collections.Counter()
    for (k, c), name in zip(cnt.items(), names):
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        print(f"{name[:70]:70s} {c['_total']:6d} | " + " ".join(f"{o}={c[o]}" for o in OPS if c[o]))
        tot.update(c)
    print(f"TOTAL over {len(cnt)} kernels: " + " ".join(f"{o}={tot[o]}" for o in OPS))


if __name__ == "__main__":
    sys.exit(main())

"""Per-kernel SASS opcode counts of the built library — the evidence that the kernels are sm_100a-native.

    python tools/sass_opcodes.py [path/to/lib.so] > profiles/r02_sass_opcodes.txt

For every kernel in the fat binary: registers are in lib/ptxas.log; here the instruction mnemonics that prove TMA
(UTMALDG / UTMASTG / UBLKCP), the 5th-generation tensor cores (UTC*MMA, UTCBAR) and tensor memory (LDTM / STTM), next to the
scan's arithmetic mix (MUFU, FFMA2 / FMUL2 / FADD2 packed fp32, LDS, LDGSTS = cp.async).  HMMA (legacy mma.sync) must
not appear.  Needs cuobjdump (CUDA toolkit), no GPU.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "actalker_b200", "lib", "libactalker_b200.so")
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "HMMA", "MUFU.EX2", "MUFU.LG2", "MUFU.RCP",
         "FFMA2", "FMUL2", "FADD2", "FFMA", "LDS", "STS", "LDGSTS", "SYNCS", "BAR.SYNC", "ATOMG", "LDG", "STG"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except OSError:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
    names = demangle(list(kernels))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, SASS opcode counts (static instructions per kernel)")
    print("# columns: " + " ".join(WATCH) + " | total")
    tot = collections.Counter()
    for k, c in kernels.items():
        tot.update(c)
        short = re.sub(r"\(.*", "", names[k]).replace("void actk::", "")
        print(f"{short[:92]:92s} " + " ".join(f"{w}={c[w]}" for w in WATCH if c[w]) + f" | {c['_total']}")
    print("# library totals: " + " ".join(f"{w}={tot[w]}" for w in WATCH))
    assert tot["HMMA"] == 0, "legacy mma.sync (HMMA) found"


if __name__ == "__main__":
    main()

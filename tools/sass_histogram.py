"""Opcode histogram of the kernels in an object file (cuobjdump -sass), for profiles/sass_*.txt: the evidence that a kernel
really issues tcgen05 (UTCHMMA / UTCBAR), tensor-memory loads / stores (LDTM / STTM), packed FP32 (FFMA2 / FMUL2) ...

    python tools/sass_histogram.py autorally_b200/lib/rollout_tc.o [kernel-name regex] > profiles/sass_rollout_tc.txt
"""
import collections
import re
import subprocess
import sys


def main():
    obj = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    print("# %s: SASS opcode histogram per kernel (cuobjdump -sass, static instruction counts)" % obj)
    for name, hist in kernels.items():
        demangled = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        if pat and not pat.search(demangled):
            continue
        total = sum(hist.values())
        base = collections.Counter()
        for op, n in hist.items():
            base[op.split(".")[0]] += n
        print("\n== %s\n   %d instructions" % (demangled[:150], total))
        key = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "F2FP",
               "HFMA2", "TEX", "LDG", "STG", "LDS", "STS", "BAR", "IMAD", "LOP3", "DFMA", "F2F"]
        print("   " + "  ".join("%s %d" % (k, base[k]) for k in key if base[k]))
        print("   top: " + ", ".join("%s %d" % (op, n) for op, n in hist.most_common(24)))


if __name__ == "__main__":
    main()

"""SASS opcode histogram of the GEMM kernels in libchromvar_b200.so (cuobjdump -sass; no GPU needed):
the evidence that the hot path is hand-written tcgen05 / TMEM / TMA code.

    python tools/sass_histogram.py > profiles/r02_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "chromvar_b200", "lib", "libchromvar_b200.so")
KERNELS = ["k_gather_contract", "k_dense_contract_fp4", "k_dense_contract_pair", "k_dense_contract<", "k_contract", "k_f4_rows_main",
           "k_dense_rows_gather", "k_counts_sums", "k_bg_sample", "k_sum_peers", "k_var_merge_peers"]
# mnemonics that prove tensor-core / TMA / cluster use on sm_100a (B200_PROFILING.md)
MARK = ["UTCOMMA", "UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS",
        "UCGABAR", "ELECT", "REDUX", "DFMA", "DMUL", "DADD", "I2F", "F2I", "IMAD", "FFMA", "LDG", "STG", "RED", "ATOM", "ATOMS",
        "LDS", "STS", "SHFL", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = name if any(k in name for k in KERNELS) else None
            if cur:
                hist[cur] = collections.Counter()
            continue
        if cur:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
            if m:
                hist[cur][m.group(1)] += 1
    print("# SASS opcode histogram of the hot kernels (sm_100a)\n")
    print("`cuobjdump -sass chromvar_b200/lib/libchromvar_b200.so`, per kernel: instruction count, the tensor-core / TMEM / TMA /")
    print("cluster mnemonics with their modifiers, and the top opcodes by base mnemonic.  UTCOMMA = tcgen05.mma kind::mxf4")
    print("(block-scaled e2m1), UTCIMMA = tcgen05.mma kind::i8, LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG = TMA")
    print("tensor load, UTCBAR = tcgen05.commit, SYNCS = mbarrier, UCGABAR = cluster barrier.\n")
    for name, h in hist.items():
        total = sum(h.values())
        short = name.split("(")[0].replace("void ", "")
        print("## `%s` -- %d instructions\n" % (short, total))
        special = [(k, v) for k, v in h.items() if any(k.startswith(p) for p in ("UTC", "LDTM", "STTM", "UTMA", "UCGABAR", "SYNCS", "ELECT"))]
        if special:
            print("| tensor / TMEM / TMA / cluster instruction | count |\n|---|---|")
            for k, v in sorted(special, key=lambda kv: -kv[1]):
                print("| `%s` | %d |" % (k, v))
            print()
        base = collections.Counter()
        for k, v in h.items():
            base[k.split(".")[0]] += v
        print("top opcodes: " + ", ".join("%s %d" % kv for kv in base.most_common(14)) + "\n")


if __name__ == "__main__":
    main()

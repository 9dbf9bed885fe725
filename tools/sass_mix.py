"""SASS instruction mix of the five kernels of a frame-step (cuobjdump, no GPU needed).

    python tools/sass_mix.py [path/to/lib.so] > profiles/rNN_vMM_sass_mix.txt

Per kernel: registers / shared memory (cuobjdump -res-usage), instruction count and the histogram of opcodes (first
mnemonic component, with the width suffix kept for memory operations).  Static counts: loops are counted once.
"""
import collections
import re
import subprocess
import sys
import os

LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "nnnoiseless_b200", "lib", "libnnnoiseless_b200.so")
KERNELS = ["hp_filter_kernelIf", "pitch_kernel", "analysis_warp_kernel", "rnn_tc_kernel", "synthesis_warp_kernelIf"]
MEM = ("LDS", "STS", "LDG", "STG", "LDSM", "LDTM", "STTM", "LDC", "LDL", "STL", "ATOMS", "RED", "ATOMG")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    blocks = re.split(r"\n\s*Function : ", sass)
    print("SASS instruction mix, %s (sm_100a, cuobjdump %s)" % (os.path.basename(LIB), "12.9"))
    for blk in blocks[1:]:
        name = blk.split("\n", 1)[0].strip()
        if not any(k in name for k in KERNELS):
            continue
        ops = collections.Counter()
        n = 0
        for line in blk.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(2)
            parts = op.split(".")
            key = parts[0]
            if key in MEM:
                w = [p for p in parts[1:] if p in ("64", "128", "U8", "U16", "S16", "32x32b", "x8", "x4", "16x256b")]
                key = ".".join([key] + w)
            ops[key] += 1
            n += 1
        print("\n== %s" % name)
        for fn, u in usage.items():
            if fn == name:
                print("   " + u)
        print("   %d instructions" % n)
        print("   " + "  ".join("%s:%d" % kv for kv in ops.most_common(40)))
        tc = {k: v for k, v in ops.items() if k.startswith(("UTC", "LDTM", "STTM", "UBLKCP", "UTMA", "SYNCS", "REDUX", "BAR", "HMMA"))}
        if tc:
            print("   blackwell / sync: " + "  ".join("%s:%d" % kv for kv in sorted(tc.items())))


if __name__ == "__main__":
    main()

"""Per-kernel count of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UTMAREDG (TMA tensor
load / store / reduce), UBLKCP (cp.async.bulk), HMMA (mma.sync), plus registers and spill bytes from the ELF.

    python tools/sass_summary.py > profiles/rNN_sass_mnemonics.txt        (CPU only: cuobjdump on the built library)
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multimodalaggressionrecognition_b200", "libmar.so")
PATTERNS = collections.OrderedDict([
    ("UTC*MMA", re.compile(r"\bUTC[A-Z]*MMA\b")), ("LDTM", re.compile(r"\bLDTM\b")), ("STTM", re.compile(r"\bSTTM\b")),
    ("UTMALDG", re.compile(r"\bUTMALDG\b")), ("UTMASTG", re.compile(r"\bUTMASTG\b")), ("UTMAREDG", re.compile(r"\bUTMAREDG\b")),
    ("UBLKCP", re.compile(r"\bUBLKCP\b")), ("HMMA", re.compile(r"\bHMMA\b")), ("SYNCS", re.compile(r"\bSYNCS\b")),
    ("RED/ATOM", re.compile(r"\b(RED|ATOMG|ATOMS|REDG)\b")),
])


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None or not re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):      # instruction lines carry their address
            continue
        counts[cur]["instructions"] += 1
        for key, pat in PATTERNS.items():
            if pat.search(line):
                counts[cur][key] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?STACK:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            regs[fn] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    names = demangle(order)
    keys = list(PATTERNS)
    print(f"# {os.path.relpath(LIB)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
    print("# " + " ".join(f"{k:>8}" for k in ["instr"] + keys + ["regs", "stack", "smem"]) + "  kernel")
    for fn in order:
        c = counts[fn]
        short = re.sub(r"\(.*", "", names.get(fn, fn).replace("(anonymous namespace)::", "")).replace("void ", "")
        r = regs.get(fn, ("-", "-", "-"))
        print("  " + " ".join(f"{c[k]:>8}" for k in ["instructions"] + keys) + " " + " ".join(f"{x:>8}" for x in r) + "  " + short[:110])


if __name__ == "__main__":
    sys.exit(main())

"""Per-kernel SASS facts of the shipped library (CPU only: cuobjdump on pointslot_b200/lib/libpslot_orb.so): instruction count,
registers, static shared memory, stack, and how often the mnemonics that characterise each kernel occur (TMA, mbarrier,
three-input min/max, byte-SWAR, integer dot products ...).

    python tools/sass_facts.py > profiles/r02e_sass_facts.txt
"""
import collections
import re
import subprocess
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "pointslot_b200" / "lib" / "libpslot_orb.so"
KEYS = ("UTMALDG", "UTMASTG", "SYNCS", "VIMNMX3", "VABSDIFF4", "IDP.4A", "IDP.2A", "PRMT", "POPC", "MATCH", "REDUX", "LDG.E.128",
        "BAR.SYNC", "ATOMS", "ATOMG", "SHFL", "VOTE")


def run(*a):
    return subprocess.run(a, capture_output=True, text=True, check=True).stdout


def demangle(n):
    d = run("c++filt", n).strip().replace("(anonymous namespace)::", "").replace("void ", "")
    return d.split("(")[0]


def main():
    cur = None
    counts = collections.defaultdict(collections.Counter)
    ninstr = collections.Counter()
    pat = re.compile(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
    for line in run("cuobjdump", "-sass", str(LIB)).split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = pat.match(line)
        if m and cur:
            ninstr[cur] += 1
            for key in KEYS:
                if m.group(1).startswith(key):
                    counts[cur][key] += 1
    res = {}
    lines = run("cuobjdump", "-res-usage", str(LIB)).split("\n")
    for i, l in enumerate(lines):
        m = re.match(r"\s*Function (\S+):", l)
        if m:
            res[m.group(1)] = lines[i + 1].strip()
    print("SASS facts per kernel of pointslot_b200/lib/libpslot_orb.so (sm_100a; cuobjdump -sass / -res-usage on the shipped library;\n"
          "tools/sass_facts.py).  UTMALDG / UTMASTG = TMA bulk-tensor loads / stores (cp.async.bulk.tensor), SYNCS = mbarrier operations,\n"
          "VIMNMX3 = three-input integer min/max, VABSDIFF4 = byte-SWAR absolute differences, IDP.4A / IDP.2A = integer dot products,\n"
          "MATCH = match.any, REDUX = warp reductions, LDG.E.128 = 16-byte global loads.\n")
    for k in sorted(ninstr, key=demangle):
        r = res.get(k, "")

        def g(key):
            m = re.search(key + r":(\d+)", r)
            return m.group(1) if m else "?"
        print(f"{demangle(k):26s} {ninstr[k]:5d} instr  regs {g('REG'):>3s}  static smem {g('SHARED'):>6s}  stack {g('STACK'):>3s}  "
              + " ".join(f"{a}={b}" for a, b in sorted(counts[k].items())))


if __name__ == "__main__":
    main()

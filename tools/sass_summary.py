"""Per-kernel SASS opcode census of libbdrc_b200.so: python tools/sass_summary.py > profiles/r02_sass_summary.txt
Counts the Blackwell tensor-core / TMEM / TMA mnemonics (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG /
UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops) next to the total instruction
count of every kernel; needs only cuobjdump (no GPU)."""
import collections
import os
import re
import subprocess
import sys

SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ocr_inference_b200", "libbdrc_b200.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOM", "SYNCS", "FFMA", "HMMA", "IMMA", "LDG", "STG",
        "LDS", "STS", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    name, counts, total = None, collections.OrderedDict(), {}
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name)
            while name in counts:
                name += "'"
            counts[name] = collections.Counter()
            total[name] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and name:
            op = m.group(1)
            total[name] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                    if k == "UTCHMMA" and ".2CTA" in op:
                        continue
                    counts[name][k] += 1
                    break
    print(f"SASS opcode census of {os.path.basename(SO)} (sm_100a), cuobjdump -sass")
    print(f"{'kernel':70s} {'instr':>7s}  " + "  ".join(f"{k}" for k in KEYS))
    agg = collections.Counter()
    for n, c in counts.items():
        agg.update(c)
        cells = "  ".join(f"{c.get(k, 0):>{len(k)}d}" for k in KEYS)
        print(f"{n[:70]:70s} {total[n]:7d}  {cells}")
    print(f"{'TOTAL':70s} {sum(total.values()):7d}  " + "  ".join(f"{agg.get(k, 0):>{len(k)}d}" for k in KEYS))


if __name__ == "__main__":
    sys.exit(main())

"""Per-kernel SASS opcode counts of the in-tree library: the evidence that the hot kernels are Blackwell-native (tcgen05 MMA =
UTCHMMA / UTCQMMA ..., TMEM loads / stores = LDTM / STTM, TMA = UTMALDG / UTMASTG) and which ones still use the legacy warp-level
HMMA path.

  python profiles/sass_opcounts.py [followyourclick_b200/libfyc_sm100a.so] > profiles/round2_sass_opcounts.md
"""
import collections
import hashlib
import re
import subprocess
import sys

OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "MUFU.EX2", "LDGSTS", "FFMA"]


def main(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    dig = hashlib.sha256(open(lib, "rb").read()).hexdigest()[:16]
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = name.replace("(anonymous namespace)::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name)
            cur = per.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                cur[o] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            cur["UTCHMMA.2CTA"] += 1
    print(f"# SASS opcode counts per kernel - `{lib}` (sha256 {dig}...), `cuobjdump -sass`\n")
    print("tcgen05 MMA = `UTCHMMA` (`.2CTA` = cta_group::2 pair mode), TMEM load / store = `LDTM` / `STTM`, TMA = `UTMALDG`, legacy warp MMA = `HMMA`.\n")
    cols = [o for o in OPS]
    print("| kernel | instr | " + " | ".join(cols) + " |\n|---|---:|" + "---:|" * len(cols))
    tot = collections.Counter()
    for k, c in sorted(per.items(), key=lambda kv: -(kv[1]["UTCHMMA"] * 1000 + kv[1]["HMMA"])):
        if not any(c[o] for o in OPS[:8]):
            continue
        print(f"| `{k[:110]}` | {c['_total']} | " + " | ".join(str(c[o]) for o in cols) + " |")
        tot.update(c)
    print(f"| **all kernels with tensor / TMA instructions** | {tot['_total']} | " + " | ".join(str(tot[o]) for o in cols) + " |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "followyourclick_b200/libfyc_sm100a.so")

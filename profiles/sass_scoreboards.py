"""Decode the scheduling control words of a kernel's SASS (stall count, write / read scoreboard, wait mask) to see WHERE a warp waits for
WHICH loads.  `cuobjdump -sass` prints every 128-bit instruction as two 64-bit words; bits 105..125 of the instruction are
[stall:4][yield:1][write-scoreboard:3][read-scoreboard:3][wait-mask:6][reuse:4] (Volta .. Blackwell).  A load names the scoreboard it
arms (W<n>); a later instruction whose wait mask has bit n set blocks until every load armed on n has landed.

  python profiles/sass_scoreboards.py followyourclick_b200/libfyc_sm100a.so 'gemm_tc_kernelILi0E' [--all]

Prints the global loads, TMEM loads, stores and every instruction that waits on a scoreboard armed by a global load.  Used for
profiles/round2_gemm_epilogue.md (the "prefetched" residual / bias registers of the GEMM epilogue were being waited for in the middle of
the group that had just issued the loads).
"""
import re
import subprocess
import sys


def load(obj, kernel_sub):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout.splitlines()
    starts = [i for i, l in enumerate(out) if "Function :" in l]
    seg = None
    for j, st in enumerate(starts):
        if kernel_sub in out[st]:
            seg = out[st:(starts[j + 1] if j + 1 < len(starts) else len(out))]
            break
    if seg is None:
        raise SystemExit(f"no function matching {kernel_sub!r}")
    ins, i = [], 0
    while i + 1 < len(seg):
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", seg[i])
        m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", seg[i + 1]) if m else None
        if m2:
            ctrl = (int(m2.group(1), 16) >> 41) & 0x7FFFFF
            ins.append(dict(addr=int(m.group(1), 16), s=m.group(2).strip(), stall=ctrl & 0xF, wr=(ctrl >> 5) & 7, rd=(ctrl >> 8) & 7,
                            wait=(ctrl >> 11) & 0x3F))
            i += 2
        else:
            i += 1
    return ins


def fmt(k, x):
    w = "-" if x["wr"] == 7 else x["wr"]
    r = "-" if x["rd"] == 7 else x["rd"]
    return f"{k:5d} {x['addr']:06x} stall{x['stall']:2d} W{w} R{r} wait={x['wait']:06b}  {x['s'][:110]}"


def main():
    obj, sub = sys.argv[1], sys.argv[2]
    ins = load(obj, sub)
    print(f"{len(ins)} instructions in {sub}")
    ldg_sb = set(x["wr"] for x in ins if "LDG" in x["s"] and x["wr"] != 7)
    mask = sum(1 << b for b in ldg_sb)
    for k, x in enumerate(ins):
        s = x["s"]
        if "--all" in sys.argv or "LDG" in s or "LDTM" in s or "STG" in s or "UTMALDG" in s or (x["wait"] & mask and not s.startswith(("LDC", "LDCU"))):
            print(fmt(k, x))


if __name__ == "__main__":
    main()

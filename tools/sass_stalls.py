"""Static issue-time estimate of a kernel's SASS: sum of the encoded stall counts (control word bits
41..44) over a range of instructions = cycles ONE warp needs for that stretch with no contention.
    python tools/sass_stalls.py model.so knot_kernel_wsI Li12E [lo hi]"""
import re
import subprocess
import sys
from collections import Counter

so, pat1, pat2 = sys.argv[1:4]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blk = [b for b in out.split("Function : ") if pat1 in b.split("\n")[0] and pat2 in b.split("\n")[0]][0]
lines = blk.split("\n")
ins = []
i = 0
while i < len(lines):
    m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s*/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ins.append((m.group(2).strip(), (hi >> 41) & 0xF))
            i += 2
            continue
    i += 1
isfp = lambda t: re.match(r"(@!?U?P\d+\s+)?(DFMA|DMUL|DADD)", t) is not None
fp = [k for k, (t, s) in enumerate(ins) if isfp(t)]
lo = int(sys.argv[4]) if len(sys.argv) > 4 else fp[0]
hi_ = int(sys.argv[5]) if len(sys.argv) > 5 else fp[-1]
seg = ins[lo:hi_ + 1]
nf = sum(1 for t, s in seg if isfp(t))
tot = sum(s for t, s in seg)
print(f"{len(ins)} instructions; range {lo}..{hi_}: {len(seg)} instructions, {nf} FP64, sum of stall counts {tot} "
      f"-> {tot / max(nf, 1):.2f} clk per FP64 instruction for one warp alone (pipe needs 2.0)")
print("stall histogram:", sorted(Counter(s for t, s in seg).items()))
# windows of 100 instructions
for a in range(lo, hi_ + 1, 100):
    w = ins[a:a + 100]
    print(f"  [{a:5d}) stalls {sum(s for t, s in w):4d} fp64 {sum(1 for t, s in w if isfp(t)):3d}")

"""profiles/traffic_*.json and profiles/fp64_*.json from one `ncu --set full` capture (raw + source CSV pages):
DRAM bytes per launch, executed FP64 instructions per knot (SASS, from the source page) and the FP64-pipe
utilisation counter -- the figures bench.py quotes in `roofline.traffic` and `fp64`.
    python tools/ncu_to_json.py gpurun_out/ws_r02_b B T profiles/ncu_ws_r02_b.txt"""
import csv
import json
import sys

base, B, T, src_name = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(base + "_raw.csv")))
hdr, units, d = rows[0], rows[1], rows[2]


def val(k):
    v, u = float(d[hdr.index(k)].replace(",", "")), units[hdr.index(k)]
    return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
rows = list(csv.reader(open(base + "_src.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h, data = rows[hi], rows[hi + 1:]
isrc, iex, ith = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
fp = {"DFMA": 0, "DMUL": 0, "DADD": 0}
tot_thread = 0
for r in data:
    if len(r) <= ith:
        continue
    t = [x for x in r[isrc].split() if not x.startswith("@")]
    if not t:
        continue
    op = t[0].split(".")[0]
    tot_thread += int(r[ith] or 0)
    if op in fp:
        fp[op] += int(r[ith] or 0)
knots = B * T
out_t = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": d[hdr.index("Kernel Name")],
         "source": f"{src_name} (ncu --set full --clock-control none, one launch of the shipped kernel; outputs still dirty in the "
                   "126 MB L2 at kernel end are not counted by dram__bytes_write)"}
out_f = {"sass_fp64_thread_instructions_per_knot": sum(fp.values()) / knots, "mix_per_knot": {k: v / knots for k, v in fp.items()},
         "all_thread_instructions_per_knot": tot_thread / knots,
         "ncu_pipe_fp64_pct": float(d[hdr.index("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")]),
         "ncu_issue_active_pct": float(d[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
         "ncu_duration_us": float(d[hdr.index("gpu__time_duration.sum")]),
         "source": f"{src_name}: executed thread-instructions of DFMA/DMUL/DADD (source page) / (B*T = {knots} knots); "
                   "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active of the same capture"}
print(json.dumps({"traffic": out_t, "fp64": out_f}, indent=1))

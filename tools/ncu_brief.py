"""Key numbers of the first kernel in an .ncu-rep: python tools/ncu_brief.py file.ncu-rep [extra-substring ...]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
hdr, row = r[0], r[2] if len(r) > 2 else r[1]
want = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct",
        "per_issue_active.ratio", "sm__warps_active.avg.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op", "dram__bytes_read.sum ",
        "dram__bytes_write.sum ", "smsp__inst_executed_pipe", "launch__registers", "sm__inst_executed_pipe_"] + sys.argv[2:]
for h, v in zip(hdr, row):
    if any(w.strip() in h for w in want):
        try:
            if float(v) == 0: continue
        except ValueError:
            pass
        print(h, "=", v)

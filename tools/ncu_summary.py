"""Developer tool: compact per-launch table from an `ncu -i X.ncu-rep --page raw --csv` dump (one row per kernel launch)."""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "ctas"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("lts__t_bytes.sum", "l2_bytes"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma_pct"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma_cyc_pct"),
        ("sm__inst_executed_pipe_uniform.sum", "uniform_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, body = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
have = [(h, n) for h, n in COLS if h in idx]
tens = [h for h in hdr if "tensor" in h and "pct_of_peak_sustained_elapsed" in h][:2]
print(",".join(["kernel"] + [n + ("[" + units[idx[h]] + "]" if units[idx[h]] else "") for h, n in have] + tens))
for r in body:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("smot::", "")
    print(",".join([name] + [r[idx[h]].replace(",", "") for h, _ in have] + [r[idx[h]] for h in tens]))

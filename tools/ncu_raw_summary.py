#!/usr/bin/env python3
"""Headline metrics of one kernel from `ncu -i rep --page raw --csv`:  tools/ncu_raw_summary.py raw.csv [out.txt]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_active.min", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed", "smsp__sass_inst_executed_op_shared_ld.sum",
        "smsp__sass_inst_executed_op_shared_st.sum"]
out = [f"Kernel Name [] = {d.get('Kernel Name', ('', ''))[1]}"]
out += [f"{w} [{d[w][0]}] = {d[w][1]}" for w in want if w in d]
st = [(h, float(v)) for h, u, v in zip(hdr, units, vals)
      if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
out.append("# warp stall cycles per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
out += [f"  {h.split('stalled_')[1].replace('_per_issue_active.ratio', ''):28s} {v:.3f}" for h, v in sorted(st, key=lambda x: -x[1])[:10]]
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}
tr = sum(float(d[k][1]) * scale[d[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
out.append(f"traffic bytes/launch {tr}")
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
    if len(sys.argv) > 3:
        json.dump({"4096": tr, "source": f"ncu --set full, {sys.argv[2]} (dram__bytes_read.sum + dram__bytes_write.sum of one k_step launch, N=4096)"},
                  open(sys.argv[3], "w"))
print(text)

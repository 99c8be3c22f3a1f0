#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report (run where ncu is installed; no GPU needed):
    python profiles/ncu_summary.py gpurun_out/r2_step_kernel_c3.ncu-rep "title" > profiles/r2_step_kernel_c3_ncu_full.md
Per captured launch: duration, DRAM bytes / throughput, occupancy, issue activity, pipe utilisation, the warp-stall
breakdown of the PC samples."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__shared_mem_per_block_static", "static smem / block"),
    ("launch__occupancy_limit_registers", "occupancy limit: registers (blocks/SM)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit: shared memory (blocks/SM)"),
    ("launch__waves_per_multiprocessor", "waves per SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy, % of 64 warps"),
    ("sm__warps_active.avg.per_cycle_active", "warps resident per active cycle"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed (avg)"),
    ("sm__cycles_active.avg", "SM cycles active (avg)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used, %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe, % active"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe, % active"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe, % active"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe, % active"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe, % active"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor instructions, % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput, % of peak"),
]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# %s\n" % title)
    print("`%s` - %d captured launch(es); `ncu --set full --clock-control none --import-source on` on a B200 (each launch "
          "replayed from flushed caches, so absolute times are cold-cache and serialised).\n" % (rep.split("/")[-1], len(data)))
    print("| metric | " + " | ".join("launch %d" % i for i in range(len(data))) + " |")
    print("|---|" + "---|" * len(data))
    print("| kernel | " + " | ".join(r[idx["Kernel Name"]][:40] for r in data) + " |")
    for key, name in KEYS:
        if key in idx:
            print("| %s | " % name + " | ".join("%s %s" % (r[idx[key]], units[idx[key]]) for r in data) + " |")
    print("\nWarp-stall breakdown (PC samples, launch 0):\n")
    st = {}
    for h in hdr:
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                st[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(data[0][idx[h]])
            except ValueError:
                pass
    tot = sum(st.values()) or 1.0
    print("| reason | share |\n|---|---|")
    for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]:
        print("| %s | %.1f %% |" % (k, 100 * v / tot))


if __name__ == "__main__":
    main()

"""Turns an `ncu --set full` report (.ncu-rep, read with the local ncu CLI) into the small JSON / text summary that is
committed under profiles/: headline counters of the captured kernel plus the hottest source lines.

    python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/r2_X_summary.json [--lines 25] [--samples N]
(--samples: network evaluations of the captured launch, from the same command run without ncu - adds the per-sample figures)
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_sectors_read_from_l1",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_throughput_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_rate_pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "l1_global_load_sectors",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "l1_global_load_requests",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_utilisation_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active": "pipe_tmem_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__block_size": "block_size",
    "launch__grid_size": "grid_size",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_blocks",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def ncu(*args):
    return subprocess.run(["ncu", *args], check=True, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    n_lines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"report": rep.split("/")[-1], "kernel": vals[hdr.index("Kernel Name")]}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            if u in UNIT_SCALE and ("bytes" in h or "duration" in h):
                x *= UNIT_SCALE[u]
            d[KEYS[h]] = x
    if "duration" in d:
        d["duration_us"] = d.pop("duration") * 1e6
    if "--samples" in sys.argv:
        n = float(sys.argv[sys.argv.index("--samples") + 1])
        d["samples"] = n
        if "l1_global_load_sectors" in d:
            d["l1_sectors_per_sample"] = d["l1_global_load_sectors"] / n
        if "l2_sectors_read_from_l1" in d:
            d["l2_to_l1_bytes_per_sample"] = d["l2_sectors_read_from_l1"] * 32.0 / n
        if "warp_instructions" in d:
            d["warp_instructions_per_32_samples"] = d["warp_instructions"] * 32.0 / n
        if "duration_us" in d:
            d["msamples_per_s"] = n / d["duration_us"]
    # hottest source lines by executed warp instructions
    src = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"))))
    cur, h2, lines = None, None, []
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) >= 2 and r[0] == "Line No":
            h2 = r
        elif h2 and len(r) == len(h2) and r[0].isdigit():
            try:
                lines.append((int(r[h2.index("Instructions Executed")]), int(r[h2.index("# Samples")]), cur, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    tot_i = sum(l[0] for l in lines) or 1
    tot_s = sum(l[1] for l in lines) or 1
    lines.sort(reverse=True)
    d["hot_lines"] = [{"file": f, "line": ln, "inst_pct": round(100.0 * i / tot_i, 2), "stall_sample_pct": round(100.0 * s / tot_s, 2), "source": t[:110]}
                      for i, s, f, ln, t in lines[:n_lines]]
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in d.items() if k != "hot_lines"}, indent=1))


if __name__ == "__main__":
    main()

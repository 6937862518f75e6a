"""Turns the raw files of scripts/gpu_r02_evidence.sh (gpurun_out/<TAG>_*) into the committed summaries under profiles/:
key counters of each captured simulation kernel + executed instructions / stall samples per source function (ncu SASS page joined
with nvdisasm line info), and the counters of the clustering kernels at N = 65,536.
  python scripts/collect_r02_profiles.py r02g"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02g"
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

SIM_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "lts__t_sectors.sum",
    "lts__t_sectors_srcunit_tex.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
    "smsp__pcsamp_warps_issue_stalled_dispatch_stall", "smsp__pcsamp_warps_issue_stalled_not_selected",
]
CLUSTER_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
]


def read_raw(path):
    rows = list(csv.reader(open(path)))
    return rows[0], rows[1], rows[2:]


def sim_summary(name, note):
    raw = os.path.join(OUT, "%s_sim_%s_raw.csv" % (TAG, name))
    if not os.path.exists(raw):
        return
    hdr, units, rows = read_raw(raw)
    d = dict(zip(hdr, rows[0]))
    u = dict(zip(hdr, units))
    kernel = d.get("Kernel Name", "?")
    lines = ["# %s" % note, "# kernel: %s, grid %s x block %s" % (kernel, d.get("Grid Size"), d.get("Block Size")),
             "# one launch under `ncu --set full --clock-control none --import-source on` (scripts/gpu_r02_evidence.sh); values per launch", ""]
    for k in SIM_KEYS:
        if k in d:
            lines.append("%-86s %s %s" % (k, d[k], u[k]))
    sass = os.path.join(OUT, "%s_sim_%s_sass.csv" % (TAG, name))
    dis = os.environ.get("UPC_DISASM_%s" % name.upper())
    mangled = os.environ.get("UPC_KERNEL_%s" % name.upper())
    if os.path.exists(sass) and dis and mangled:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_by_line.py"), sass, dis, mangled, "30"], capture_output=True, text=True).stdout
        lines += ["", "# executed warp instructions and stall samples per source function / line (scripts/ncu_by_line.py)", out]
    open(os.path.join(PROF, "%s_sim_%s_metrics.txt" % (TAG, name)), "w").write("\n".join(lines) + "\n")


def cluster_summary():
    raw = os.path.join(OUT, "%s_cluster_big_raw.csv" % TAG)
    if not os.path.exists(raw):
        return
    hdr, units, rows = read_raw(raw)
    u = dict(zip(hdr, units))
    lines = ["# clustering kernels of the N = 65,536 sweep points (8 blobs, distance pass): SE(2), SE(3), 7-DoF; adjacency pass forced, then pivot pass",
             "# `ncu --set full --clock-control none` over scripts/gpu_cluster_big.py once; times under ncu are serialised and cold",
             "# columns: " + " | ".join("%s [%s]" % (k, u.get(k, "")) for k in CLUSTER_KEYS), ""]
    for r in rows:
        d = dict(zip(hdr, r))
        lines.append("%-58s %s" % (d["Kernel Name"][:58], " | ".join(d.get(k, "?") for k in CLUSTER_KEYS)))
    txt = os.path.join(OUT, "%s_cluster_big.txt" % TAG)
    if os.path.exists(txt):
        lines += ["", "# un-profiled run of the same script (best of three after a warm-up)", open(txt).read()]
    open(os.path.join(PROF, "%s_cluster_kernels_n65536.txt" % TAG), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    sim_summary("c5", "forward_simulate_kernel<Se2Model, 1 lane, seeded> on a quarter of BASELINE config 5 (1,024 edges x 1,000 particles, 64 intervals)")
    sim_summary("c2", "forward_simulate_kernel<Se3Model, 4 lanes, seeded> on BASELINE config 2 (10,000 particles, 128^3 SDF, 64 intervals)")
    sim_summary("c3", "forward_simulate_kernel<LinkedModelSmall, 1 lane, seeded> on BASELINE config 3 (100,000 particles, 7-DoF arm, 256^3 SDF, 64 intervals)")
    cluster_summary()
    probe = os.path.join(OUT, "%s_probe.txt" % TAG)
    if os.path.exists(probe):
        open(os.path.join(PROF, "%s_probe.txt" % TAG), "w").write(open(probe).read())

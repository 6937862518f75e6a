"""Turns the raw outputs of scripts/gpu_profile_c2.sh (under gpurun_out/) into the committed summaries under profiles/."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PRO = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def last_json_line(path):
    for line in reversed(open(path).read().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("no JSON line in " + path)


# 1. bench lines
json.dump(last_json_line(os.path.join(OUT, "bench_c2_final.json")), open(os.path.join(PRO, tag + "_bench_c2.json"), "w"), indent=1)
json.dump(last_json_line(os.path.join(OUT, "bench_c2_ref_final.json")), open(os.path.join(PRO, tag + "_bench_c2_reference_arm.json"), "w"), indent=1)

# 2. launch list
raw = open(os.path.join(OUT, "launches_c2.csv")).read()
start = raw.index('"ID"')
open(os.path.join(PRO, tag + "_launches_c2.csv"), "w").write(raw[start:])
rows = list(csv.DictReader(io.StringIO(raw[start:])))
agg = {}
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    name = r["Kernel Name"].split("(")[0][:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
total = sum(a[1] for a in agg.values())
with open(os.path.join(PRO, tag + "_launches_c2_summary.md"), "w") as f:
    f.write("ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline` (device-resident steps, then the end-to-end steps whose "
            "simulation runs as four launches over step slices, then the two roofline microbenchmarks).\n\n")
    f.write("| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|\n")
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.3f | %.4f | %.1f%% |\n" % (name, cnt, ms, ms / cnt, 100.0 * ms / total))

# 3. key counters of the full capture
rep = os.path.join(OUT, "prof_sim_c2.ncu-rep")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rd[0], rd[1], rd[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum"]
got = {}
with open(os.path.join(PRO, tag + "_forward_simulate_kernel_c2_metrics.csv"), "w") as f:
    f.write("metric,value,unit\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write('"%s","%s","%s"\n' % (w, vals[i], units[i]))
            got[w] = (vals[i], units[i])


def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


rd_b = to_bytes(*got["dram__bytes_read.sum"])
wr_b = to_bytes(*got["dram__bytes_write.sum"])
json.dump({"kernel": "forward_simulate_kernel<Se3Model,4>", "workload": "c2", "dram_bytes_read": rd_b, "dram_bytes_write": wr_b, "traffic": rd_b + wr_b,
           "source": "ncu --set full --clock-control none, one launch, profiles/%s_forward_simulate_kernel_c2_metrics.csv" % tag},
          open(os.path.join(PRO, "r01_sim_kernel_traffic.json"), "w"), indent=1)

# 4. text logs
for src, dst in (("phases.log", "_sim_phase_cycles.txt"), ("breakdown3.log", "_sim_breakdown.txt")):
    p = os.path.join(OUT, src)
    if os.path.exists(p):
        open(os.path.join(PRO, tag + dst), "w").write(open(p).read())
print("profiles refreshed:", sorted(os.listdir(PRO)))

"""Turn what a GPU evidence run left in gpurun_out/ into the tracked files of profiles/ (run here, on the CPU box, after
`gpurun -- 'bash scripts/gpu_profile_r2.sh; bash scripts/gpu_traffic.sh; bash scripts/gpu_bench_all.sh <tag>; ...'`).
Usage: python profiles/refresh_evidence.py <bench tag>"""
import csv, collections, json, os, re, shutil, subprocess, sys
tag = sys.argv[1]
here = os.path.dirname(os.path.abspath(__file__)); root = os.path.dirname(here)
go = os.path.join(root, "gpurun_out")
py = sys.executable

# ncu summaries (details / phases / stalls) + stall lines + opcode mix of the config-5 capture
for rep, out in (("prof_r2_final", "r2_final"), ("prof_r2_config4", "r2_config4")):
    subprocess.run([py, os.path.join(here, "ncu_summary.py"), os.path.join(go, rep + ".ncu-rep"), os.path.join(here, out)], check=True, capture_output=True)
    if out == "r2_final":
        for script, name in (("ncu_stall_lines.py", "r2_final_stall_lines.txt"), ("ncu_opmix.py", "r2_final_opmix.txt")):
            open(os.path.join(here, name), "w").write(subprocess.run([py, os.path.join(here, script), "/tmp/_src.csv"], capture_output=True, text=True).stdout)

# launch list of the default bench command
shutil.copy(os.path.join(go, "r2_final_launches.csv"), os.path.join(here, "r2_final_launches.csv"))
rows = [r for r in csv.reader(open(os.path.join(here, "r2_final_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(re.sub(r"\(.*", "", r[4]).strip(), [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1000.0
tot = sum(v[1] for v in agg.values())
c0 = sum(v[1] for k, v in agg.items() if "compose_frames_kernel<0>" in k or "compose_frames_kernel<(bool)0>" in k)
c1 = sum(v[1] for k, v in agg.items() if "compose_frames_kernel<1>" in k or "compose_frames_kernel<(bool)1>" in k)
with open(os.path.join(here, "r2_final_launches_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 60, python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --no-check (profiles/r2_final_launches.csv)\n")
    f.write("per-launch times are cold-cache and serialised; the kernel-only timed region of the bench is steps x (compose<false> + compose<true>): "
            "the compose kernels are all of it, compose<false> alone %.1f per cent\n" % (100 * c0 / max(c0 + c1, 1e-9)))
    for k, v in agg.items():
        f.write("%-46s launches %4d  total %10.1f us  share %5.1f%%  avg %8.1f us\n" % (k[:46], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))

# DRAM traffic of the full-size config-5 launch
m = {}
for r in csv.reader(open(os.path.join(go, "r2_traffic.csv"))):
    if len(r) > 10 and r[0].isdigit():
        m[r[-3]] = float(r[-1].replace(",", ""))
t = json.load(open(os.path.join(here, "traffic.json")))
t.update({"dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_write": m["dram__bytes_write.sum"],
          "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
          "duration_ms_under_ncu": m["gpu__time_duration.sum"] / 1e6, "inst_executed": m["smsp__inst_executed.sum"],
          "thread_inst_per_inst": m["smsp__thread_inst_executed_per_inst_executed.ratio"],
          "sm_throughput_pct": m["sm__throughput.avg.pct_of_peak_sustained_elapsed"]})
json.dump(t, open(os.path.join(here, "traffic.json"), "w"), indent=1)

# bench lines
with open(os.path.join(here, "r2_final_bench.jsonl"), "w") as out:
    for wl in ("default", "config4", "config2", "config1", "config5spec", "latency", "reference"):
        p = os.path.join(go, "bench_%s_%s.log" % (tag, wl))
        if not os.path.exists(p):
            continue
        lines = [l for l in open(p).read().strip().split("\n") if l.startswith("{")]
        if lines:
            json.loads(lines[-1]); out.write(lines[-1] + "\n")

# SASS features of the library as built
open(os.path.join(here, "r2_final_sass_features.txt"), "w").write(
    subprocess.run([py, os.path.join(here, "sass_features.py"), os.path.join(root, "h264-scroll-encoder_b200", "libh264scroll_b200.so")], capture_output=True, text=True).stdout)
print(open(os.path.join(here, "r2_final_launches_summary.txt")).read())
print(open(os.path.join(here, "r2_final_stalls.txt")).read())

"""Summaries of one .ncu-rep for profiles/: <out_prefix>_details.txt, _phases.txt, _stalls.txt (run here, on the CPU box).
Usage: python profiles/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r1_x"""
import csv, subprocess, sys, os
rep, out = sys.argv[1], sys.argv[2]
here = os.path.dirname(os.path.abspath(__file__))
open(out + "_details.txt", "w").write(subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
open("/tmp/_src.csv", "w").write(src)
open(out + "_phases.txt", "w").write(subprocess.run([sys.executable, os.path.join(here, "ncu_phases.py"), "/tmp/_src.csv"], capture_output=True, text=True).stdout)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
st = sorted(((float(v.replace(",", "")), h) for h, v in d.items() if h.startswith("smsp__average_warps_issue_stalled_")
             and h.endswith("_per_issue_active.ratio") and "not_issued" not in h), reverse=True)
with open(out + "_stalls.txt", "w") as f:
    f.write("warp stall cycles per issued instruction\n")
    for v, h in st[:10]:
        f.write("  %6.3f %s\n" % (v, h.split("stalled_")[1].split("_per_issue")[0]))
    for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "dram__bytes_read.sum",
              "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "smsp__thread_inst_executed_per_inst_executed.ratio"):
        if k in d:
            f.write("%s = %s %s\n" % (k, d[k], u.get(k, "")))
print(open(out + "_stalls.txt").read())

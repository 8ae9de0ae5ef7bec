"""Reads the `ncu --set full` capture of track_kernel taken by tools/gpu_round_end_r2.sh and writes profiles/ncu_traffic.json:
DRAM bytes and executed warp-instructions of ONE launch of the bench's 148-frame step, stamped with the SHA-256 of the kernel
source it was captured from. bench.py reports `roofline.traffic` / `roofline.issue` from this file only while the stamp matches
the tree (a kernel edit makes the figures stale: they are then reported as null until the capture is repeated).
usage: python tools/ncu_traffic.py gpurun_out/r02_track_f148.ncu-rep gpurun_out/r02_bench_n1.json"""
import csv, hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, bench_json = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, v = rows[0], rows[1], rows[2]
get = lambda k: (float(v[hdr.index(k)]), units[hdr.index(k)])
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
rd, ru = get("dram__bytes_read.sum")
wr, wu = get("dram__bytes_write.sum")
inst, _ = get("smsp__inst_executed.sum")
dur, du = get("gpu__time_duration.sum")
line = json.loads(open(bench_json).read().strip().splitlines()[-1])
F = line["config"]["frames_per_step_per_gpu"]
residuals_per_launch = line["residuals_per_frame"] * F
src = os.path.join(ROOT, "nalo_slam_b200", "csrc", "nalo_track.cu")
out = {
    "track_kernel_dram_bytes_per_launch": rd * scale[ru] + wr * scale[wu],
    "track_kernel_warp_inst_per_launch": inst,
    "track_kernel_warp_inst_per_32_residuals": inst / (residuals_per_launch / 32.0),
    "track_kernel_issue_active_pct_ncu": get("smsp__issue_active.avg.pct_of_peak_sustained_active")[0],
    "track_kernel_lsu_data_pipe_pct_ncu": get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")[0],
    "track_kernel_duration_under_ncu": f"{dur} {du}",
    "residuals_per_launch": residuals_per_launch,
    "track_kernel_source_sha256": hashlib.sha256(open(src, "rb").read()).hexdigest(),
    "source": f"ncu --set full --clock-control none -k regex:track_kernel on `python bench.py --steps 3 --warmup 3` ({os.path.basename(rep)}): one launch of track_kernel, {F} frames",
}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

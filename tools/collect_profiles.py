"""Turn one gpurun_out/<tag>/ visit (tools/gpu_round.sh) into the tracked evidence under profiles/.
usage: python tools/collect_profiles.py gpurun_out/<tag> [round_suffix]"""
import csv
import json
import os
import shutil
import subprocess
import sys

src = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)
KERNELS = ["front", "raster", "raster_big", "tiles"]

for a, b in [("bench.json", "bench_%s.json"), ("bench_reference.json", "bench_%s_reference.json"),
             ("bench_inview.json", "bench_%s_inview.json"), ("configs.jsonl", "configs_%s.jsonl"), ("launches.csv", "launches_%s.csv"),
             ("pipes.txt", "pipelines_%s.txt"), ("pytest_gpu.log", "pytest_gpu_%s.log"),
             ("memcheck_smoke.log", "sanitizer_%s_memcheck_smoke.log"), ("racecheck_smoke.log", "sanitizer_%s_racecheck_smoke.log"),
             ("memcheck_fused.log", "sanitizer_%s_memcheck_fused.log")]:
    if os.path.exists(os.path.join(src, a)):
        shutil.copy(os.path.join(src, a), os.path.join(dst, b % rnd))


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


traffic = {}
with open(os.path.join(dst, "ncu_%s_summary.txt" % rnd), "w") as f:
    f.write("# ncu --set full --clock-control none, one launch of each kernel of a fused pass (10 views 1280x720, xArm7 links 1-7),\n"
            "# single pipeline, eager launches, B200.  Raw metric excerpts (tools/ncu_raw.py); cold-cache, serialised replays.\n")
    for k in KERNELS:
        rep = os.path.join(src, k + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        f.write(run([sys.executable, os.path.join(ROOT, "tools", "ncu_raw.py"), rep]) + "\n")
        out = run(["ncu", "-i", rep, "--page", "raw", "--csv"])
        rows = list(csv.reader(out.splitlines()))
        h = rows[0]
        d = {h[i]: rows[2][i] for i in range(len(h))}
        units = {h[i]: rows[1][i] for i in range(len(h))}

        def to_bytes(key):
            v = float(d[key].replace(",", ""))
            u = units[key].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        traffic[k if k != "raster_big" else "raster_big"] = int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
# bench.py looks the dominant stage up by its stage name: the raster stage = k_raster + k_raster_big
traffic["raster"] = traffic.get("raster", 0) + traffic.pop("raster_big", 0)
json.dump(traffic, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
with open(os.path.join(dst, "ncu_%s_source_lines.txt" % rnd), "w") as f:
    for k in ["raster", "raster_big", "tiles", "front"]:
        rep = os.path.join(src, k + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        f.write("# ehb_k_%s: instructions and stall samples by function, then the top source lines\n" % k)
        f.write(run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), rep]))
        f.write(run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "ehb_k_" + k, "0", "25"]) + "\n")
print(open(os.path.join(dst, "traffic.json")).read())

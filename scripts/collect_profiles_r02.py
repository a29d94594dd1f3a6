#!/usr/bin/env python
"""gpurun_out/r02_* -> profiles/r02/: text summaries of the `ncu --set full` reports, the launch lists, and traffic.json
(DRAM bytes per launch of the dominant kernels, read by bench.py for roofline.traffic)."""
import csv, glob, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02")   # on the GPU box: a directory under gpurun_out/
os.makedirs(DST, exist_ok=True)
reps = sorted(glob.glob(os.path.join(SRC, "r02_prof_*.ncu-rep")))
subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "summarize_ncu_full.py"), os.path.join(DST, "ncu_full_r02.summary.txt")] + reps,
               stdout=subprocess.DEVNULL, check=False)
for f in glob.glob(os.path.join(SRC, "r02_launches_*")):
    shutil.copy(f, os.path.join(DST, os.path.basename(f).replace("r02_", "")))


def metrics(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return None
    head, units, vals = rows[0], rows[1], rows[2]
    def get(name):
        i = head.index(name)
        v = float(vals[i].replace(",", ""))
        u = units[i].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
        return v * scale
    return dict(dram_bytes_read=get("dram__bytes_read.sum"), dram_bytes_write=get("dram__bytes_write.sum"),
                duration_ms_under_ncu=get("gpu__time_duration.sum"), source="profiles/r02/ncu_full_r02.summary.txt <- " + os.path.basename(rep))


traffic = {}
for key, name in {"k_dist_tc:msmt17:3xfp16": "r02_prof_k_dist_tc_rect", "k_dist_tc_fused:msmt17:3xfp16": "r02_prof_k_dist_tc_fused_msmt17",
                  "k_dist_tc_fused:market:3xfp16": "r02_prof_k_dist_tc_fused_market", "k_rank_count:msmt17": "r02_prof_k_rank_count",
                  "k_prep_rows_warp:msmt17": "r02_prof_k_prep_rows_warp", "k_jaccard_bucket:msmt17": "r02_prof_k_jaccard_bucket_msmt17",
                  "k_blend_default:market": "r02_prof_k_blend_default", "k_cand_topk:market": "r02_prof_k_cand_topk"}.items():
    rep = os.path.join(SRC, name + ".ncu-rep")
    if os.path.exists(rep):
        m = metrics(rep)
        if m:
            traffic[key] = m
json.dump(traffic, open(os.path.join(DST, "traffic.json"), "w"), indent=1)
print(json.dumps({k: {kk: (round(vv / 1e9, 3) if "bytes" in kk else vv) for kk, vv in v.items()} for k, v in traffic.items()}, indent=1))

"""Condense gpurun_out/step_metrics.csv (ncu --metrics ... of every kdcc kernel of ONE bench step) into a table
per bench kernel family and a small JSON (measured DRAM traffic per step) that bench.py cites as roofline.traffic.

    python tools/step_metrics_summary.py gpurun_out/step_metrics.csv profiles/r01_step_metrics.txt profiles/r01_traffic.json
"""
import collections
import csv
import json
import re
import sys

src, txt, js = sys.argv[1:4]
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
launch = collections.OrderedDict()
for r in rows:
    lid = int(r[0])
    name = re.sub(r"<.*", "", r[4].split("(")[0]).replace("void ", "").replace("kdcc::", "")
    e = launch.setdefault(lid, {"name": name})
    try:
        e[r[12]] = float(r[14].replace(",", ""))
    except ValueError:
        pass
    e.setdefault("units", {})[r[12]] = r[13]

# bench family of every launch, from the fixed per-site order of kdcc.hotpath.HotPathStep.step
fam_of = {"cast_f32_to_bf16_kernel": "cast_w", "hint_loss_kernel": "hint_loss", "loss_finalize_kernel": "loss_finalize",
          "kd_loss_kernel": "kd_loss", "reduce_splits_kernel": "pw_bwd_dw", "dw_tc_wgrad2_kernel": "dw_bwd(dW)", "dw_tc_wgrad3_kernel": "dw_bwd(dW)",
          "dw_tc_wgrad2_reduce_kernel": "dw_bwd(dW)"}
order = list(launch.values())
# kdcc.hotpath.HotPathStep.step in the reference loop's order: forward of every site (conv, GEMM), the hint losses, then the
# backward in reverse site order (GEMM dW + split reduce, GEMM dX, [conv dX], weight-gradient kernel + reduce).  In the
# site-by-site order a conv that directly follows a GEMM is the dX conv.  Both are told apart by what has been seen so far.
hints_seen = bwd_gemms = 0
interleaved = False
names = [e["name"] for e in order]
first_hint = names.index("hint_loss_kernel") if "hint_loss_kernel" in names else len(names)
interleaved = names[:first_hint].count("dw_tc_conv2_kernel") == 1   # site-by-site: one forward conv before the first hint
prev = ""
gemm_seen = 0
for idx, e in enumerate(order):
    n = e["name"]
    if interleaved:
        if n == "dw_tc_conv2_kernel":
            e["fam"] = "dw_bwd(dX)" if prev == "pw_gemm_sm100_kernel" else "dw_fwd"
        elif n == "pw_gemm_sm100_kernel":
            e["fam"] = ("pw_fwd", "pw_bwd_dw", "pw_bwd_dx")[gemm_seen % 3]
            gemm_seen += 1
        else:
            e["fam"] = fam_of.get(n, n)
    else:
        fwd_phase = idx < first_hint
        if n == "dw_tc_conv2_kernel":
            e["fam"] = "dw_fwd" if fwd_phase else "dw_bwd(dX)"
        elif n == "pw_gemm_sm100_kernel":
            if fwd_phase:
                e["fam"] = "pw_fwd"
            else:
                e["fam"] = ("pw_bwd_dw", "pw_bwd_dx")[bwd_gemms % 2]
                bwd_gemms += 1
        else:
            e["fam"] = fam_of.get(n, n)
    prev = n

def to_bytes(e, key):
    v, u = e.get(key, 0.0), e.get("units", {}).get(key, "byte")
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

def to_us(e):
    v, u = e.get("gpu__time_duration.sum", 0.0), e.get("units", {}).get("gpu__time_duration.sum", "ns")
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1e-3)

fams = collections.OrderedDict()
for e in order:
    f = fams.setdefault(e["fam"], {"launches": 0, "us": 0.0, "dram_rd": 0.0, "dram_wr": 0.0, "l2": 0.0, "tensor": [], "dram_pct": []})
    f["launches"] += 1
    f["us"] += to_us(e)
    f["dram_rd"] += to_bytes(e, "dram__bytes_read.sum")
    f["dram_wr"] += to_bytes(e, "dram__bytes_write.sum")
    f["l2"] += to_bytes(e, "lts__t_bytes.sum")
    if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in e:
        f["tensor"].append((to_us(e), e["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]))
    if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in e:
        f["dram_pct"].append((to_us(e), e["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]))

def wavg(lst):
    t = sum(a for a, _ in lst)
    return sum(a * b for a, b in lst) / t if t else 0.0

total_us = sum(f["us"] for f in fams.values())
with open(txt, "w") as out:
    out.write("# source: %s  (ncu --metrics, every kdcc kernel of one bench step: batch 4, 51M plan, NCHW bf16)\n" % src)
    out.write("# per-launch times under ncu are serialised and cold-cache: compare SHARES with bench.py's CUDA-event shares\n")
    out.write("%-14s %8s %10s %7s %12s %12s %12s %10s %10s\n" % ("family", "launches", "us", "share", "dram_rd_MB", "dram_wr_MB", "L2_MB", "tensor%", "dram%"))
    for k, f in fams.items():
        out.write("%-14s %8d %10.1f %7.3f %12.1f %12.1f %12.1f %10.1f %10.1f\n" % (
            k, f["launches"], f["us"], f["us"] / total_us, f["dram_rd"] / 1e6, f["dram_wr"] / 1e6, f["l2"] / 1e6,
            wavg(f["tensor"]), wavg(f["dram_pct"])))
    out.write("%-14s %8d %10.1f\n" % ("total", sum(f["launches"] for f in fams.values()), total_us))
traffic = {}
for k, f in fams.items():
    traffic[k] = {"dram_bytes_per_step": f["dram_rd"] + f["dram_wr"], "launches": f["launches"], "ncu_us": round(f["us"], 1)}
traffic["dw_bwd"] = {"dram_bytes_per_step": traffic["dw_bwd(dX)"]["dram_bytes_per_step"] + traffic["dw_bwd(dW)"]["dram_bytes_per_step"],
                     "launches": traffic["dw_bwd(dX)"]["launches"] + traffic["dw_bwd(dW)"]["launches"],
                     "ncu_us": round(traffic["dw_bwd(dX)"]["ncu_us"] + traffic["dw_bwd(dW)"]["ncu_us"], 1)}
import subprocess
try:
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or None
except Exception:
    commit = None
json.dump({"source": src, "config": "bench.py default: batch 4, 51M plan, k9d5p20, nchw bf16; no input gradient for the first site",
           "summarised_at_commit": commit, "families": traffic}, open(js, "w"), indent=1)
print(open(txt).read())

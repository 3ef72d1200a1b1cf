"""Summarise an ncu per-launch metrics CSV of one chunk (tools/profile_step.py under
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,
      sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,
      sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv)
into per-kernel-shape rows, the per-block DRAM traffic and the time-weighted tensor-pipe activity, and refresh
profiles/roofline_traffic.json (bench.py's roofline.traffic).

  python tools/ncu_summary.py gpurun_out/r02_kernels_b1024.csv profiles/r02_ncu_kernels_summary.json [frames]
Per-launch ncu times are cold-cache and serialised: read shares and bytes, not absolute times."""
import csv
import json
import re
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
src, dst = sys.argv[1], sys.argv[2]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "%": 1.0,
        "cycle/second": 1.0, "cycle/nsecond": 1e9, "Ghz": 1e9, "Mhz": 1e6, "hz": 1.0}
rows = list(csv.DictReader([l for l in open(src) if l.startswith('"')]))
ker = OrderedDict()
for r in rows:
    k = ker.setdefault(r["ID"], {"name": r["Kernel Name"]})
    k[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)


def short(n):
    m = re.match(r"(?:void )?(?:arp::)?([\w:]+)(<[^(]*>)?", n)
    return (m.group(1) + (m.group(2) or "")) if m else n


ROLE = {"gemm_tcgen05_kernel<__half, 0, 2, 2>": "QKV (ln_1 folded)", "gemm_tcgen05_kernel<__half, 1, 2, 2>": "c_fc (ln_2 folded, QuickGELU)",
        "gemm_tcgen05_kernel<float, 0, 2, 0>": "patch embed"}
seq = []
for k in ker.values():
    s = short(k["name"])
    seq.append({"kernel": s, "us": k["gpu__time_duration.sum"], "dram_read_MB": k["dram__bytes_read.sum"] / 1e6,
                "dram_write_MB": k["dram__bytes_write.sum"] / 1e6,
                "tensor_pct": k["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"],
                "sm_ghz": k.get("sm__cycles_elapsed.avg.per_second", 0.0) / 1e9,
                "sm_throughput_pct": k.get("sm__throughput.avg.pct_of_peak_sustained_elapsed")})
# the residual GEMM template serves out_proj (small read) and c_proj (large read): split by DRAM read volume
for s in seq:
    if s["kernel"] == "gemm_tcgen05_kernel<__half, 0, 2, 3>":
        s["role"] = "c_proj (x +=, row statistics)" if s["dram_read_MB"] > 1000 else "out_proj (x +=, row statistics)"
    else:
        s["role"] = ROLE.get(s["kernel"], "")
agg = OrderedDict()
for s in seq:
    key = s["kernel"] + (" | " + s["role"] if s["role"] else "")
    a = agg.setdefault(key, {"launches": 0, "us": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0, "tensor_us": 0.0, "ghz_us": 0.0})
    a["launches"] += 1; a["us"] += s["us"]; a["dram_read_MB"] += s["dram_read_MB"]; a["dram_write_MB"] += s["dram_write_MB"]
    a["tensor_us"] += s["tensor_pct"] * s["us"]; a["ghz_us"] += s["sm_ghz"] * s["us"]
tot_us = sum(a["us"] for a in agg.values())
table = []
for key, a in agg.items():
    table.append({"kernel": key, "launches": a["launches"], "avg_us": round(a["us"] / a["launches"], 1),
                  "share_of_chunk": round(a["us"] / tot_us, 4), "tensor_pipe_pct": round(a["tensor_us"] / a["us"], 1),
                  "dram_read_MB_per_launch": round(a["dram_read_MB"] / a["launches"], 1),
                  "dram_write_MB_per_launch": round(a["dram_write_MB"] / a["launches"], 1),
                  "dram_GBps": round((a["dram_read_MB"] + a["dram_write_MB"]) / a["us"] * 1e3, 0),
                  "sm_ghz": round(a["ghz_us"] / a["us"], 3)})
gemm = [s for s in seq if s["kernel"].startswith("gemm_tcgen05")]
block = [s for s in seq if s["role"] and "patch" not in s["role"]] + [s for s in seq if s["kernel"].startswith("attention_tc")]
n_blocks = max(1, sum(1 for s in seq if s["kernel"].startswith("attention_tc")))
out = {
    "source": f"ncu per-launch metrics of one {frames}-frame chunk (python tools/profile_step.py --batch {frames} --chunks 1), "
              "--clock-control none, round-2 kernels (fp16 operands, 16-bit residual stream, LayerNorm folded, statistics from the "
              "residual GEMM epilogue)",
    "kernels_per_chunk": len(seq), "chunk_us_under_ncu": round(tot_us, 1),
    "time_weighted_tensor_pipe_pct": round(sum(s["tensor_pct"] * s["us"] for s in seq) / tot_us, 1),
    "dram_GB_per_full_block": round(sum(s["dram_read_MB"] + s["dram_write_MB"] for s in block) / 1e3 / n_blocks, 2),
    "dram_GB_per_chunk": round(sum(s["dram_read_MB"] + s["dram_write_MB"] for s in seq) / 1e3, 1),
    "by_kernel": table,
}
Path(dst).write_text(json.dumps(out, indent=1))
traffic = {"source": out["source"], "gemm_launches_per_chunk": len(gemm),
           "gemm_dram_bytes_per_launch": sum(s["dram_read_MB"] + s["dram_write_MB"] for s in gemm) * 1e6 / len(gemm),
           "gemm_avg_us_per_launch_under_ncu": sum(s["us"] for s in gemm) / len(gemm),
           "dram_GB_per_full_block": out["dram_GB_per_full_block"]}
(ROOT / "profiles" / "roofline_traffic.json").write_text(json.dumps(traffic, indent=1))
print(json.dumps(out, indent=1))

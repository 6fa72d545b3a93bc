"""Summarises an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`
launch list of one denoising step into per-kernel-family DRAM bytes per launch (bench.py `roofline.traffic`).
The summary is stamped with the digest of the CUDA sources it was captured from; bench.py reports traffic only when
that digest matches the build it runs (a stale file would silently describe other kernels).
usage: python tools/ncu_traffic.py gpurun_out/ncu_step_dram.csv profiles/r02_ncu_traffic.json"""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

src, dst = sys.argv[1], sys.argv[2]
rows = []
with open(src) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
fam = {}
for r in rows:
    name = r["Kernel Name"]
    key = next((k for k in ("tapgemm", "fmha", "gn_fused", "gn_cluster", "gn_stats", "gn_apply", "layernorm", "ln_stats", "temporal_attn", "concat",
                            "upsample2x") if k in name), None)
    if key is None:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"].lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    d = fam.setdefault(key, {"ids": set(), "read": 0.0, "write": 0.0, "us": 0.0})
    d["ids"].add(r["ID"])
    m = r["Metric Name"]
    if m == "dram__bytes_read.sum":
        d["read"] += v * scale
    elif m == "dram__bytes_write.sum":
        d["write"] += v * scale
    elif m == "gpu__time_duration.sum":
        d["us"] += v * scale
out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over one cfg2 "
                 f"denoising step ({src.split('/')[-1]}); cold-cache, serialised launches"}
from bench import source_digest  # noqa: E402
out["source_digest"] = source_digest()
fam = {{"temporal_attn": "temporal_attention"}.get(k, k): v for k, v in fam.items()}
for k, d in fam.items():
    n = len(d["ids"])
    out[k] = {"launches": n, "dram_bytes_per_launch": (d["read"] + d["write"]) / n,
              "dram_read_bytes_per_launch": d["read"] / n, "dram_write_bytes_per_launch": d["write"] / n,
              "ncu_us_per_launch": d["us"] / n}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))

"""Per-launch DRAM traffic and tensor-pipe activity of gemm_tc_kernel from an ncu metrics CSV of one profiled step
(tools/profile_step.py under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active -k regex:gemm_tc_kernel --csv`).
Usage: gemm_traffic.py metrics.csv workload out.json"""
import csv
import json
import sys

src, workload, out = sys.argv[1:4]
rows = [r for r in csv.reader(open(src, errors="replace")) if r]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ik, im, iv, iu, iid = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("ID")
per = {}
for r in rows[hdr + 1:]:
    if len(r) <= iv or "gemm_tc_kernel" not in r[ik]:
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if r[im].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if r[im].startswith("gpu__time"):
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    per.setdefault(r[iid], {})[r[im]] = v
n = len(per)
t = sum(d["gpu__time_duration.sum"] for d in per.values())
rd = sum(d["dram__bytes_read.sum"] for d in per.values())
wr = sum(d["dram__bytes_write.sum"] for d in per.values())
tp = sum(d["gpu__time_duration.sum"] * d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] for d in per.values()) / t
json.dump({"kernel": "gemm_tc_kernel", "workload": workload, "launches": n, "total_us": t, "dram_bytes_per_launch": (rd + wr) / n,
           "dram_read_bytes_total": rd, "dram_write_bytes_total": wr, "tensor_pipe_active_pct_time_weighted": tp,
           "source": f"{src} (ncu, cold L2 per launch, serialised; final round-2 build)"}, open(out, "w"), indent=1)
print(open(out).read())

#!/usr/bin/env python
"""profiles/spmv_traffic.json and profiles/spgemm_traffic.json from the ncu outputs of tools/gpu_prof.sh (gpurun_out/r02_*):
the dominant kernel's share of a bench step (launch list) and its DRAM bytes per launch (raw page of the full capture)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    return [(r[kn], float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > mv]


def raw_metric(path, names, row=0):
    rows = list(csv.reader(open(path)))
    h = rows[0]
    out = {}
    for n in names:
        if n in h:
            out[n] = (rows[1][h.index(n)], rows[2 + row][h.index(n)])
    return out, len(rows) - 2


def to_bytes(unit, val):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


L = launches(os.path.join(G, "r02_launches_spmv_bench.csv"))
names = [k for k, _ in L]
i0 = next(i for i, k in enumerate(names) if "spmv_hot2_prep" in k)
step = L[i0:i0 + 3]
total = sum(t for _, t in step)
hot = next(t for k, t in step if "spmv_run_hot2_kernel" in k)
m, _ = raw_metric(os.path.join(G, "r02_prof_spmv_raw.csv"), ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "lts__t_sectors.sum"])
dram = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
out = {"kernel": "spmv_run_hot2_kernel<float,float,PLUS,TIMES> (TMA-staged runs + 128 KB hot table)",
       "dominant_kernel_share_of_step": hot / total, "step_kernels_ns": [[k[:60], t] for k, t in step],
       "dram_bytes_per_launch": dram, "l2_sectors_per_launch": float(m["lts__t_sectors.sum"][1].replace(",", "")),
       "ncu_kernel_us": float(m["gpu__time_duration.sum"][1].replace(",", "")) * (1e-3 if m["gpu__time_duration.sum"][0] == "ns" else 1.0),
       "source": "gpurun_out/r02_launches_spmv_bench.csv, r02_prof_spmv_raw.csv (tools/gpu_prof.sh); copies under profiles/"}
json.dump(out, open(os.path.join(P, "spmv_traffic.json"), "w"), indent=1)
print(out)
try:
    tot = 0.0
    path = os.path.join(G, "r02_prof_mstream_raw.csv")
    _, nrows = raw_metric(path, [])
    per = []
    for r in range(nrows):
        mm, _ = raw_metric(path, ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"], r)
        b = to_bytes(*mm["dram__bytes_read.sum"]) + to_bytes(*mm["dram__bytes_write.sum"])
        per.append(b); tot += b
    out2 = {"kernels": "masked_stream_kernel (classes S, M, L of one C<L> = L (+.pair) L call at scale 20)", "dram_bytes_per_class": per, "dram_bytes_per_call": tot,
            "source": "gpurun_out/r02_prof_mstream_raw.csv"}
    json.dump(out2, open(os.path.join(P, "spgemm_traffic.json"), "w"), indent=1)
    print(out2)
except Exception as e:
    print("spgemm traffic:", e)

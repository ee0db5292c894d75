"""DRAM traffic of the tcgen05 GEMM launches of a train step, from an ncu CSV (feeds bench.py's roofline.traffic).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --profile-from-start off -k regex:gemm_tcgen05 --csv --log-file gpurun_out/gemm_dram.csv python tools/step_profile.py --steps 2
    python tools/gemm_traffic.py gpurun_out/gemm_dram.csv 2 > profiles/rNN_gemm_dram_traffic.json"""
import collections
import csv
import json
import sys

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
per = collections.defaultdict(dict)
for r in rows:
    per[r["ID"]][r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}
rd = wr = t = 0.0
for m in per.values():
    rd += m["dram__bytes_read.sum"][0] * scale[m["dram__bytes_read.sum"][1]]
    wr += m["dram__bytes_write.sum"][0] * scale[m["dram__bytes_write.sum"][1]]
    t += m["gpu__time_duration.sum"][0] * scale[m["gpu__time_duration.sum"][1]]
n = len(per)
print(json.dumps({"kernel": "gemm_tcgen05_kernel", "launches_captured": n, "steps_captured": steps,
                  "dram_bytes_read_per_launch": rd / n, "dram_bytes_write_per_launch": wr / n,
                  "dram_bytes_per_launch": (rd + wr) / n, "dram_bytes_per_step": (rd + wr) / steps,
                  "kernel_ms_per_step_under_ncu": t / 1e6 / steps,
                  "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                            "--profile-from-start off -k regex:gemm_tcgen05 python tools/step_profile.py --steps 2 "
                            "(every GEMM launch of two consecutive eager train steps)"}, indent=1))

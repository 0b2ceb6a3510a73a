"""Writes profiles/r02/gemm_traffic.json from an `ncu --set full` capture of tools/quick_rloop.py (development helper):
mean dram__bytes_read.sum + dram__bytes_write.sum per legendre_gemm_kernel launch, which bench.py reports as roofline.traffic
when its workload and level chunk match."""
import csv
import json
import subprocess
import sys

rep, workload, chunk = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
tot, n, per = 0.0, 0, {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if "legendre_gemm_kernel" not in name:
        continue
    unit_r, unit_w = rows[1][idx["dram__bytes_read.sum"]], rows[1][idx["dram__bytes_write.sum"]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    b = float(r[idx["dram__bytes_read.sum"]]) * scale[unit_r] + float(r[idx["dram__bytes_write.sum"]]) * scale[unit_w]
    per.setdefault("analysis" if "<1>" in name or "(bool)1" in name or "true" in name else "synthesis", []).append(b)
    tot += b
    n += 1
out = {"workload": workload, "level_chunk": chunk, "dram_bytes_per_launch": tot / max(n, 1), "launches": n,
       "per_kind_bytes": {k: sum(v) / len(v) for k, v in per.items()}, "source": rep.replace("gpurun_out/", "profiles/r02/")}
json.dump(out, open("profiles/r02/gemm_traffic.json", "w"), indent=1)
print(out)

"""Per-launch DRAM traffic of captured kernels -> JSON (bench.py reads profiles/ncu_traffic_r01.json for roofline.traffic):
python tools/ncu_traffic.py gpurun_out/bench_gemm.ncu-rep [...] > profiles/ncu_traffic_r01.json"""
import csv
import io
import json
import subprocess
import sys

MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for path in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * MULT[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * MULT[units[col["dram__bytes_write.sum"]]]
        e = out.setdefault(name, {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        e["launches"] += 1
        e["dram_read_bytes"] += rd
        e["dram_write_bytes"] += wr
for e in out.values():
    e["traffic_bytes_per_launch"] = (e["dram_read_bytes"] + e["dram_write_bytes"]) / e["launches"]
allg = [e for k, e in out.items() if k.startswith("gemm_bf16_tn_pair_kernel")]
if allg:
    n = sum(e["launches"] for e in allg)
    out["gemm_bf16_tn_pair_kernel (all captured epilogues)"] = {
        "launches": n, "traffic_bytes_per_launch": sum(e["dram_read_bytes"] + e["dram_write_bytes"] for e in allg) / n,
        "note": "8 consecutive GEMM launches = one conformer block of the bench step (ncu --set full, bench.py --steps 1 --warmup 3)"}
print(json.dumps(out, indent=1))

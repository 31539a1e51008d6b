"""profiles/gemm_traffic_<tag>.json from gpurun_out/gemm_<tag>.ncu-rep: dram__bytes_read.sum + dram__bytes_write.sum per
launch of the dominant kernel (what bench.py reports as roofline.traffic).   python tools/gemm_traffic.py <tag>"""
import csv, io, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1]
rep = ROOT / "gpurun_out" / f"gemm_{tag}.ncu-rep"
txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]


def col(r, name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)


launches = []
for r in rows[2:]:
    if "gemm_bf16_tcgen05_persistent_kernel" not in r[hdr.index("Kernel Name")]:
        continue
    launches.append({"kernel": r[hdr.index("Kernel Name")][:60], "grid": r[hdr.index("Grid Size")],
                     "dram_read_bytes": col(r, "dram__bytes_read.sum"), "dram_write_bytes": col(r, "dram__bytes_write.sum"),
                     "us": col(r, "gpu__time_duration.sum")})
out = {"source": f"ncu --set full --clock-control none, {len(launches)} consecutive launches of gemm_bf16_tcgen05_persistent_kernel inside a live "
                 f"step (gpurun_out/gemm_{tag}.ncu-rep, summarised in profiles/ncu_summary_{tag}.*); ncu flushes the caches before every "
                 "launch, so this is the cold-cache DRAM traffic",
       "dram_bytes_per_launch": sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches) / max(len(launches), 1),
       "launches": launches}
(ROOT / "profiles" / f"gemm_traffic_{tag}.json").write_text(json.dumps(out, indent=1))
print(out["dram_bytes_per_launch"], len(launches))

"""Turn the ncu outputs brought back in gpurun_out/ into small tracked summaries under profiles/.

    python tools/summarize_profile.py r1a      # reads gpurun_out/launches_r1a.csv + gpurun_out/gemm_r1a.ncu-rep

launches: per-kernel-name launch count, total / mean device time and SHARE of the captured window (ncu times are
cold-cache and serialised: compare shares, not absolutes). full capture: the metrics B200_PROFILING.md names.
"""
from __future__ import annotations

import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"

KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "gpu__dram_throughput",
        "dram__cycles_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
        "launch__occupancy_limit", "smsp__cycles_active.avg", "sm__inst_executed_pipe_tensor", "l1tex__t_sector_hit_rate",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic"]


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("splice::", "")
    return name[:110]


def launches(tag: str) -> dict:
    path = OUT / f"launches_{tag}.csv"
    text = "\n".join(l for l in path.read_text().splitlines() if l.startswith('"'))
    rows = list(csv.DictReader(io.StringIO(text)))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        a = agg[short(r["Kernel Name"])]
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    table = sorted(([k, v[0], v[1] / 1e3, v[1] / v[0] / 1e3, 100.0 * v[1] / total] for k, v in agg.items()),
                   key=lambda t: -t[2])
    ours = sum(t[2] for t in table if not t[0].startswith("at::") and "cudnn" not in t[0] and "cutlass" not in t[0]
               and "nchw" not in t[0].lower() and "Memcpy" not in t[0])
    return {"captured_launches": sum(v[0] for v in agg.values()), "total_us": total / 1e3, "splice_kernels_share_pct": 100 * ours / (total / 1e3),
            "kernels": [{"kernel": t[0], "launches": t[1], "total_us": round(t[2], 1), "mean_us": round(t[3], 2), "share_pct": round(t[4], 2)}
                        for t in table]}


def full(tag: str, stem: str) -> list:
    rep = OUT / f"{stem}_{tag}.ncu-rep"
    if not rep.exists():
        return []
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": short(r[hdr.index("Kernel Name")]), "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for i, h in enumerate(hdr):
            if any(k in h for k in KEYS):
                d[h] = f"{r[i]} {units[i]}".strip()
        out.append(d)
    return out


def main() -> None:
    tag = sys.argv[1]
    PROF.mkdir(exist_ok=True)
    res = {"tag": tag, "launch_list": launches(tag)}
    for stem in ("gemm", "attn", "conv"):
        f = full(tag, stem)
        if f:
            res[f"full_{stem}"] = f
    (PROF / f"ncu_summary_{tag}.json").write_text(json.dumps(res, indent=1))
    ll = res["launch_list"]
    lines = [f"# ncu launch list {tag}: {ll['captured_launches']} launches, {ll['total_us']:.0f} us serialised; "
             f"splice_b200 kernels = {ll['splice_kernels_share_pct']:.1f}% of device time", "",
             "| kernel | launches | total us | mean us | share % |", "|---|---:|---:|---:|---:|"]
    for k in ll["kernels"][:40]:
        lines.append(f"| `{k['kernel']}` | {k['launches']} | {k['total_us']} | {k['mean_us']} | {k['share_pct']} |")
    for stem in ("gemm", "attn", "conv"):
        for i, d in enumerate(res.get(f"full_{stem}", [])):
            lines += ["", f"## --set full: {d['kernel']} grid {d['grid']} block {d['block']} (capture {i})", ""]
            lines += [f"- {k}: {v}" for k, v in d.items() if k not in ("kernel", "grid", "block")]
    (PROF / f"ncu_summary_{tag}.md").write_text("\n".join(lines) + "\n")
    print("\n".join(lines[:30]))


if __name__ == "__main__":
    main()

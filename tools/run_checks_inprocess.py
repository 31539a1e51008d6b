"""Runs the named checks of tools/gpu_checks.py in ONE process (tools/gpu_checks.py itself isolates every check in a subprocess,
which costs a torch import each - too slow for a short gpurun call), then __graft_entry__.smoke(). Default selection: the generator
checks and one full-step check. Writes gpurun_out/checks_inprocess.json.
    python tools/run_checks_inprocess.py [check names ...]"""
import json
import sys
import time
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)

from tools import gpu_checks  # noqa: E402

names = sys.argv[1:] or ["generator_inversion_variant", "generator_native_small", "generator_concurrent_calls", "train_step_golden"]
report = []
for n in names:
    t0 = time.time()
    try:
        res = gpu_checks.CHECKS[n]()
        rows = res if isinstance(res, list) else [res]
        rec = {"name": n, "ok": all(r.get("ok", False) for r in rows), "results": rows}
    except Exception as e:  # noqa: BLE001
        rec = {"name": n, "ok": False, "error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-2000:]}
    rec["secs"] = time.time() - t0
    report.append(rec)
    print(("PASS " if rec["ok"] else "FAIL ") + n, f"{rec['secs']:.1f}s", flush=True)
    (OUT / "checks_inprocess.json").write_text(json.dumps(report, indent=1, default=str))
try:
    import __graft_entry__

    __graft_entry__.smoke()
    print("PASS smoke", flush=True)
except Exception:  # noqa: BLE001
    print("FAIL smoke", traceback.format_exc()[-800:], flush=True)

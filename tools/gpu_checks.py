"""GPU bring-up checks, each run in its own subprocess so that a faulting kernel (trap, illegal address)
cannot poison the CUDA context of the checks that follow. Writes gpurun_out/checks.json.

    python tools/gpu_checks.py                 # run everything
    python tools/gpu_checks.py --only gemm     # run checks whose name contains "gemm"
    python tools/gpu_checks.py --run NAME      # (internal) run one check in this process
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"

CHECKS = {}


def check(fn):
    CHECKS[fn.__name__] = fn
    return fn


def _rel(a, b):
    import torch

    return (a.float() - b.float()).norm().item() / max(b.float().norm().item(), 1e-30)


def _maxabs(a, b):
    return (a.float() - b.float()).abs().max().item()


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
def _gemm_case(M, N, K, bn, impl=0, seed=0, dump=None):
    import torch
    from splice_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(seed)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    C = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(A, B, out32=C, impl=impl, bn_hint=bn)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = _maxabs(C, ref)
    rel = _rel(C, ref)
    ok = bool(rel < 1e-4) and not bool(torch.isnan(C).any())
    if not ok and dump:
        OUT.mkdir(exist_ok=True)
        torch.save({"A": A.cpu(), "B": B.cpu(), "C": C.cpu(), "ref": ref.cpu()}, OUT / f"{dump}.pt")
    return {"M": M, "N": N, "K": K, "bn": bn, "impl": impl, "maxabs": err, "rel": rel, "ok": ok,
            "nan": int(torch.isnan(C).sum().item())}


@check
def gemm_simt_small():
    return [_gemm_case(128, 128, 64, 0, impl=1), _gemm_case(200, 96, 192, 0, impl=1)]


@check
def gemm_tc_single_tile_k64():
    # one tile, one k-block: isolates descriptor / swizzle / TMEM-lane mapping from pipeline logic
    return [_gemm_case(128, 128, 64, 128, dump="gemm_fail_128x128x64")]


@check
def gemm_tc_single_tile_bn64():
    return [_gemm_case(128, 64, 64, 64, dump="gemm_fail_128x64x64")]


@check
def gemm_tc_single_tile_bn256():
    return [_gemm_case(128, 256, 64, 256, dump="gemm_fail_128x256x64")]


@check
def gemm_tc_multi_k():
    # K spans several stages and wraps the ring (K/64 = 12 > STAGES)
    return [_gemm_case(128, 128, 768, 128), _gemm_case(128, 64, 768, 64), _gemm_case(128, 256, 3072, 256)]


@check
def gemm_tc_multi_tile_edges():
    out = []
    for bn in (64, 128, 256):
        out.append(_gemm_case(785, 768, 768, bn))   # ragged M
        out.append(_gemm_case(3140, 192, 768, bn))  # N not a multiple of the wide tiles
        out.append(_gemm_case(100, 2304, 192, bn))  # M smaller than one tile
    out.append(_gemm_case(3140, 2304, 768, 0))
    out.append(_gemm_case(1570, 768, 3072, 0))
    return out


@check
def gemm_tc_epilogues():
    import torch
    from splice_b200 import ops

    res = []
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 400, 384, 256
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    base = A.float() @ B.float().t()
    for impl in (1, 0):
        # bias + residual (in place) + bf16 copy + fp32 slice
        c32 = resid.clone()
        c16 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        sl = torch.zeros(M, 128, device="cuda")
        ops.gemm(A, B, out32=c32, out16=c16, bias=bias, residual=c32, slice32=sl, slice_cols=(128, 256), impl=impl)
        ref = base + bias + resid
        res.append({"case": "bias+residual", "impl": impl, "rel32": _rel(c32, ref), "rel16": _rel(c16, ref),
                    "relslice": _rel(sl, ref[:, 128:256]),
                    "ok": _rel(c32, ref) < 1e-5 and _rel(c16, ref) < 5e-3 and _rel(sl, ref[:, 128:256]) < 1e-5})
        # GELU with pre-activation save
        h = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        pre = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(A, B, out16=h, bias=bias, act=ops.ACT_GELU, aux16=pre, impl=impl)
        refpre = base + bias
        refh = torch.nn.functional.gelu(refpre)
        res.append({"case": "gelu", "impl": impl, "relh": _rel(h, refh), "relpre": _rel(pre, refpre),
                    "ok": _rel(h, refh) < 5e-3 and _rel(pre, refpre) < 5e-3})
        # GELU grad
        d = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(A, B, out16=d, act=ops.ACT_GELU_GRAD, aux16=pre, impl=impl)
        x = pre.float().requires_grad_(True)
        torch.nn.functional.gelu(x).backward(base)
        res.append({"case": "gelu_grad", "impl": impl, "rel": _rel(d, x.grad), "ok": _rel(d, x.grad) < 5e-3})
        # token remap + pos
        S, P = 2, 200
        pos = torch.randn(P + 1, N, device="cuda", generator=g)
        tok = torch.zeros(S * (P + 1), N, device="cuda")
        ops.gemm(A, B, out32=tok, bias=bias, rows_per_seq=P, pos=pos, impl=impl)
        reft = torch.zeros_like(tok)
        for s in range(S):
            reft[s * (P + 1) + 1:(s + 1) * (P + 1)] = base[s * P:(s + 1) * P] + bias + pos[1:]
        res.append({"case": "remap", "impl": impl, "rel": _rel(tok, reft), "ok": _rel(tok, reft) < 1e-5})
    return res


@check
def gemm_tc_timing():
    """TFLOP/s of the ViT-B/8 shapes at M = 4*785 (forward) and 2*785 (backward), CUDA events, L2 flushed."""
    import torch
    from splice_b200 import ops

    shapes = [(3140, 2304, 768), (3140, 768, 768), (3140, 3072, 768), (3140, 768, 3072),
              (1570, 3072, 768), (1570, 768, 3072), (1570, 768, 2304), (3136, 768, 192)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    out = []
    for (M, N, K) in shapes:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        row = {"M": M, "N": N, "K": K}
        for bn in (64, 128, 256):
            for _ in range(3):
                ops.gemm(A, B, out16=C, bn_hint=bn)
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.gemm(A, B, out16=C, bn_hint=bn)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            row[f"bn{bn}_us"] = ms * 1e3
            row[f"bn{bn}_tflops"] = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        # cuBLAS for context
        for _ in range(3):
            torch.matmul(A, B.t())
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(A, B.t())
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        row["cublas_us"] = ms * 1e3
        row["cublas_tflops"] = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        row["ok"] = True
        out.append(row)
    return out


# ------------------------------------------------------------------------------------------------
def _run_one(name: str) -> int:
    t0 = time.time()
    try:
        import torch

        assert torch.cuda.is_available(), "no CUDA device"
        res = CHECKS[name]()
        torch.cuda.synchronize()
        ok = all(r.get("ok", False) for r in res) if isinstance(res, list) else bool(res.get("ok", False))
        print(json.dumps({"name": name, "ok": ok, "secs": time.time() - t0, "results": res}))
        return 0
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"name": name, "ok": False, "secs": time.time() - t0, "error": repr(e),
                          "trace": traceback.format_exc()[-2000:]}))
        return 1


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--run")
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=180)
    ap.add_argument("--out", default="checks.json")
    args = ap.parse_args()
    if args.run:
        return _run_one(args.run)
    OUT.mkdir(exist_ok=True)
    names = [n for n in CHECKS if all(tok in n for tok in args.only.split(",") if tok)] if args.only else list(CHECKS)
    if args.only:
        names = [n for n in CHECKS if any(tok in n for tok in args.only.split(","))]
    report = []
    for n in names:
        try:
            r = subprocess.run([sys.executable, __file__, "--run", n], capture_output=True, text=True,
                               timeout=args.timeout, cwd=str(ROOT))
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            rec = json.loads(line[-1]) if line else {"name": n, "ok": False, "error": "no output"}
            rec["returncode"] = r.returncode
            rec["stderr_tail"] = r.stderr[-1500:]
            rec["stdout_tail"] = "\n".join(l for l in r.stdout.splitlines() if not l.startswith("{"))[-1500:]
        except subprocess.TimeoutExpired:
            rec = {"name": n, "ok": False, "error": f"timeout after {args.timeout}s"}
        report.append(rec)
        print(("PASS " if rec.get("ok") else "FAIL ") + n + "  " + json.dumps(rec.get("results", rec.get("error", "")))[:1500],
              flush=True)
        (OUT / args.out).write_text(json.dumps(report, indent=1))
    nfail = sum(1 for r in report if not r.get("ok"))
    print(f"{len(report) - nfail}/{len(report)} checks passed")
    return 0 if nfail == 0 else 1


if __name__ == "__main__":
    sys.exit(main())

"""Crash-tolerant single-kernel probe for the GPU box: runs every conv / elementwise case against the fp32 torch-CPU
reference, one record per case into gpurun_out/probe.jsonl.  A hung or faulting case kills only the child process;
the parent restarts after it.  Usage: python tools/gpu_probe.py [--only substring] [--timeout 30]"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "probe.jsonl")

# name, (n, c, h, w), c_out, k, stride, pad, dil, groups, act, bias, residual, force_path
CONV_CASES = [
    ("tiled_64_256_56", (2, 64, 56, 56), 256, 1, 1, 0, 1, 1, 1, True, False, 0),
    ("tiled_256_64_56_res", (2, 256, 56, 56), 64, 1, 1, 0, 1, 1, 1, True, True, 0),
    ("tiled_2048_512_7", (3, 2048, 7, 7), 512, 1, 1, 0, 1, 1, 1, True, False, 0),
    ("tiled_24_144_56_clip", (1, 24, 56, 56), 144, 1, 1, 0, 1, 1, 2, True, False, 0),
    ("tiled_16_10_5_tail", (1, 16, 5, 5), 10, 1, 1, 0, 1, 1, 0, False, False, 0),
    ("tiled_1024_2048_bn256", (2, 1024, 14, 14), 2048, 1, 1, 0, 1, 1, 0, True, False, 0),
    ("im2col_3x3_64_56", (2, 64, 56, 56), 64, 3, 1, 1, 1, 1, 1, True, False, 0),
    ("im2col_3x3_s2_128_56", (2, 128, 56, 56), 128, 3, 2, 1, 1, 1, 1, True, False, 0),
    ("im2col_1x1_s2_256_512", (2, 256, 56, 56), 512, 1, 2, 0, 1, 1, 0, True, False, 0),
    ("im2col_3x3_odd_16", (3, 16, 13, 17), 24, 3, 1, 1, 1, 1, 0, True, False, 0),
    ("im2col_3x3_dil2", (1, 32, 20, 20), 32, 3, 1, 2, 2, 1, 0, True, False, 0),
    ("im2col_3x3_512_7", (2, 512, 7, 7), 512, 3, 1, 1, 1, 1, 1, True, True, 0),
    ("im2col_3x3_p0_128", (1, 128, 34, 34), 128, 3, 1, 0, 1, 1, 0, True, False, 0),
    ("im2col_9x9_32_3", (1, 32, 40, 40), 3, 9, 1, 0, 1, 1, 0, True, False, 2),
    ("rows_9x9_32_3_p0", (1, 32, 40, 40), 3, 9, 1, 0, 1, 1, 0, True, False, 0),
    ("rows_3x3_s2_c32_p0", (2, 32, 37, 41), 64, 3, 2, 0, 1, 1, 1, True, False, 0),
    ("rows_5x5_c24_p0", (1, 24, 30, 30), 40, 5, 1, 0, 1, 1, 0, True, False, 0),
    ("rows_3x3_c16_p0_res", (3, 16, 21, 21), 64, 3, 1, 0, 1, 1, 1, True, True, 0),
    ("im2col_as_1x1", (2, 64, 28, 28), 64, 1, 1, 0, 1, 1, 0, True, False, 2),
    ("rows_7x7_s2_stem", (2, 3, 224, 224), 64, 7, 2, 3, 1, 1, 1, True, False, 0),
    ("rows_3x3_s2_mbv2", (1, 3, 224, 224), 32, 3, 2, 1, 1, 1, 2, True, False, 0),
    ("rows_9x9_tnet", (1, 3, 72, 72), 32, 9, 1, 0, 1, 1, 0, True, False, 0),
    ("rows_3x3_c8", (1, 8, 19, 23), 16, 3, 1, 1, 1, 1, 0, True, False, 0),
    ("splitk_3x3_512_7_b32", (32, 512, 7, 7), 512, 3, 1, 1, 1, 1, 1, True, True, 0),
    ("splitk_1x1_2048_512_b32", (32, 2048, 7, 7), 512, 1, 1, 0, 1, 1, 1, True, False, 0),
    ("splitk_3x3_256_14_b32", (32, 256, 14, 14), 256, 3, 1, 1, 1, 1, 1, True, False, 0),
    ("splitk_1x1_1024_256_b8", (8, 1024, 14, 14), 256, 1, 1, 0, 1, 1, 0, False, True, 0),
    ("splitk_gemm_like_b4", (4, 2048, 1, 1), 1000, 1, 1, 0, 1, 1, 0, True, False, 0),
    ("dw_3x3_32_112", (1, 32, 112, 112), 32, 3, 1, 1, 1, 32, 2, True, False, 0),
    ("dw_3x3_s2_144", (2, 144, 56, 56), 144, 3, 2, 1, 1, 144, 2, True, False, 0),
    ("dw_3x3_c20", (1, 20, 9, 9), 20, 3, 1, 1, 1, 20, 0, True, False, 0),
]


def run_cases(start: int, only: str, timeout: float) -> None:
    import numpy as np
    import torch
    import torch.nn.functional as F

    from smelter_b200.api import Context, Image, run_conv

    state = {"t0": time.time(), "idx": start, "name": "init"}

    def watchdog():
        while True:
            time.sleep(1.0)
            if time.time() - state["t0"] > timeout:
                with open(OUT, "a") as f:
                    f.write(json.dumps({"case": state["name"], "idx": state["idx"], "status": "HANG"}) + "\n")
                os._exit(3)

    threading.Thread(target=watchdog, daemon=True).start()
    ctx = Context(0)
    for idx in range(start, len(CONV_CASES)):
        name, shape, co, k, s, p, d, g, act, has_bias, has_res, force = CONV_CASES[idx]
        if only and only not in name:
            continue
        state.update(t0=time.time(), idx=idx, name=name)
        rng = np.random.default_rng(idx)
        n, c, h, w = shape
        x = rng.standard_normal(shape).astype(np.float16)
        wt = (rng.standard_normal((co, c // g, k, k)) * np.sqrt(2.0 / (c // g * k * k))).astype(np.float16)
        b = rng.standard_normal(co).astype(np.float32) if has_bias else None
        xt, wtt = torch.from_numpy(x.astype(np.float32)), torch.from_numpy(wt.astype(np.float32))
        ref = F.conv2d(xt, wtt, torch.from_numpy(b) if has_bias else None, stride=s, padding=p, dilation=d, groups=g)
        res = None
        if has_res:
            r = rng.standard_normal(tuple(ref.shape)).astype(np.float16)
            ref = ref + torch.from_numpy(r.astype(np.float32))
            res = Image.fromArray(ctx, r)
        if act == 1:
            ref = ref.relu()
        elif act == 2:
            ref = ref.clamp(0.0, 6.0)
        rec = {"case": name, "idx": idx}
        try:
            xi = Image.fromArray(ctx, x)
            y, ms = run_conv(ctx, xi, wt, b, stride=(s, s), pads=(p, p, p, p), dilation=(d, d), groups=g, act=act, clip=(0.0, 6.0),
                             residual=res, force_path=force, iters=5)
            out = y.toFloatArray()
            refn = ref.numpy()
            err = np.abs(out - refn)
            bad = ~np.isfinite(out)
            rec.update(status="ok", ms=ms, max_err=float(np.nanmax(err)), ref_max=float(np.abs(refn).max()), n_nonfinite=int(bad.sum()),
                       tflops=2.0 * refn.size * (c // g) * k * k / (ms * 1e-3) / 1e12)
            tol = 4e-3 * max(1.0, rec["ref_max"])
            if bad.any() or rec["max_err"] > tol:
                rec["status"] = "MISMATCH"
                worst = np.unravel_index(np.nanargmax(np.where(bad, np.inf, err)), err.shape)
                rec["worst_index"] = [int(v) for v in worst]
                rec["got"], rec["want"] = float(out[worst]), float(refn[worst])
                # error structure hints: per-channel / per-row / per-col max error
                rec["err_by_channel_head"] = [float(v) for v in err.max(axis=(0, 2, 3))[:16]]
                rec["err_by_row_head"] = [float(v) for v in err.max(axis=(0, 1, 3))[:16]]
                rec["err_by_col_head"] = [float(v) for v in err.max(axis=(0, 1, 2))[:16]]
                rec["frac_bad"] = float((err > tol).mean())
        except Exception as e:  # noqa: BLE001
            rec.update(status="ERROR", error=str(e)[:400])
            with open(OUT, "a") as f:
                f.write(json.dumps(rec) + "\n")
            print(json.dumps(rec), flush=True)
            os._exit(4)  # CUDA errors are sticky: restart the process for the next case
        with open(OUT, "a") as f:
            f.write(json.dumps(rec) + "\n")
        print(json.dumps(rec), flush=True)
    os._exit(0)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--start", type=int, default=-1)
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=float, default=40.0)
    args = ap.parse_args()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if args.start >= 0:
        run_cases(args.start, args.only, args.timeout)
        return 0
    if os.path.exists(OUT):
        os.remove(OUT)
    start = 0
    while start < len(CONV_CASES):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--start", str(start), "--only", args.only, "--timeout", str(args.timeout)],
                           timeout=args.timeout * (len(CONV_CASES) + 2) + 120)
        if r.returncode == 0:
            break
        last = start
        if os.path.exists(OUT):
            with open(OUT) as f:
                lines = [json.loads(l) for l in f if l.strip()]
            if lines:
                last = max(l["idx"] for l in lines)
        start = max(last, start) + 1
    ok = bad = 0
    if os.path.exists(OUT):
        with open(OUT) as f:
            for l in f:
                if l.strip():
                    if json.loads(l)["status"] == "ok":
                        ok += 1
                    else:
                        bad += 1
    print(f"probe: {ok} ok, {bad} not ok (details in gpurun_out/probe.jsonl)")
    return 0


if __name__ == "__main__":
    sys.exit(main())

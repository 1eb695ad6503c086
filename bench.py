#!/usr/bin/env python
"""Headline benchmark: ResNet-50 fp16 224x224 images/s through the drop-in inference path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one `encode` of one batch of 32 synthetic images per GPU (BASELINE.json configs[2]; weak scaling: every
rank runs its own batch of 32, the only collective is the one-time NCCL weight-arena broadcast).  The model is the
repo's seeded ResNet-50 (random weights: no network for checkpoints) with its 53 BatchNormalization nodes still in the
graph, taken through the ONNX2MPS restatement with --half (BN fold, fp16, OHWI, producer stamp) exactly as the
reference intends (README.md:54), then built and run by libsmelter_b200.so.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, CUDA events around K back-to-back encodes on the
engine's stream, max over ranks.  `e2e`: the same metric through the public API with pinned HOST buffers — every step
copies its fp16 input batch host->device and reads the fp32 logits back, all inside the timed region.
`--impl reference` times the reference arm: the oracle port of the path on the host CPU cores (the reference itself
cannot run here: Swift + Apple MPS).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PER_GPU_BATCH = 32
IMAGE = (3, 224, 224)
FLOPS_PER_IMAGE = 2 * 4_089_184_256  # 53 conv + FC, algorithmic
N_INPUT_SETS = 16                    # 16 x 9.6 MB = 154 MB of distinct inputs > 126 MB L2


def model_bytes() -> bytes:
    from smelter_b200 import modelzoo, onnx2mps

    return onnx2mps.convert_bytes(modelzoo.resnet50(seed=0, fold_bn=False).serialize(), half=True)


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


def ncu_conv_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the conv kernels of one encode, summed over the launches of a step, from
    the committed `ncu --set full` capture of this same command (profiles/r1_encode_full.csv; cold-cache replay).  None when the
    capture is absent."""
    path = os.path.join(ROOT, "profiles", "r1_encode_full.csv")
    try:
        import csv
        with open(path) as f:
            rows = list(csv.reader(f))
        head = rows[0]
        ir, iw, ik = head.index("dram_read"), head.index("dram_write"), head.index("kernel")
        return sum(float(r[ir]) + float(r[iw]) for r in rows[1:] if "conv_igemm" in r[ik] or "conv_mega" in r[ik] or "conv_pair" in r[ik])
    except (OSError, ValueError):
        return None


class stdout_to_stderr:
    """Rank 0 prints ONE JSON line on stdout; whatever libraries write to file descriptor 1 while the communicators come up
    (NCCL's "NCCL version ..." banner) is sent to stderr instead."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def mark(self) -> int:
        return len(self.rows)

    def stop(self, lo: int = 0, hi: int = None) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = self.rows[lo:hi]
        if len(rows) < 3:  # a short timed region sees few 50 ms samples: widen to everything under load after its start (the e2e loop)
            rows = self.rows[lo:] or self.rows
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(steps: int, warmup: int, budget_s: float = 150.0):
    """The oracle port (torch-CPU fp32 ONNX interpreter; onnxruntime is not installable here) on all host cores."""
    import torch

    from oracle.onnx_interp import Interpreter

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    interp = Interpreter(model_bytes())
    torch.manual_seed(1)
    x1 = torch.rand(1, *IMAGE).half().float()
    interp.run(x1)  # weight conversion + first-touch
    t0 = time.perf_counter()
    interp.run(x1)
    per_image = time.perf_counter() - t0
    total_steps = max(1, steps + warmup)
    sample = int(max(1, min(PER_GPU_BATCH, (budget_s / total_steps) / max(per_image, 1e-4))))
    x = torch.rand(sample, *IMAGE).half().float()
    for _ in range(warmup):
        interp.run(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        interp.run(x)
    dt = time.perf_counter() - t0
    return {"value": sample * steps / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps x {sample} images (of the {PER_GPU_BATCH}-image batch), fp32 torch-CPU ONNX interpreter "
                      f"(oracle/onnx_interp.py; onnxruntime unavailable), {torch.get_num_threads()} threads", "ms_per_step": dt / steps * 1e3}


def run_reference(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    base = cpu_reference(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "ResNet-50 fp16 224x224 images/sec", "value": base["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ResNet-50 224x224 batch=32 per GPU (BASELINE.json configs[2]); CPU arm runs a bounded sample per step",
                       "onnx_graph": "seeded random ResNet-50 -> ONNX2MPS --half", "note": "reference = Swift + Apple MPS, not runnable here; "
                       "this arm is the oracle port on host cores"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def run_native(args) -> int:
    import numpy as np
    import torch

    from smelter_b200 import dist as sdist
    from smelter_b200.api import Configuration, Context, Format, Image, ONNXGraph

    rank, local_rank, world = sdist.env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    quiet = stdout_to_stderr()
    quiet.__enter__()  # until the weight replicas are in place (below)
    if world > 1:
        sdist.init_process_group("nccl")
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    K, W, B = args.steps, max(args.warmup, 3), args.batch
    stream = torch.cuda.Stream(device=dev)
    ctx = Context(local_rank, stream=stream.cuda_stream)
    data = model_bytes()
    graph = ONNXGraph(data, Configuration(deferWeights=(world > 1 and rank != 0)), context=ctx)
    assert graph.modelFormat == Format.mpsFlavor
    nn = graph.metalGraph()
    bcast_ms = None
    if world > 1:  # one-time weight replica over NVLink
        uid = sdist.share_bytes(Context.ncclUniqueId() if rank == 0 else b"", 0)
        ctx.initNCCL(uid, rank, world)
        barrier()
        t0 = time.perf_counter()
        nn.broadcastWeights(0)
        bcast_ms = (time.perf_counter() - t0) * 1e3
        checksum, _ = nn.weightChecksum()
        if not sdist.all_equal(checksum, dev):
            raise SystemExit("weight replicas differ after the broadcast")
    barrier()
    quiet.__exit__()

    rng = np.random.default_rng(1 + rank)
    host_in = torch.empty((N_INPUT_SETS, B) + IMAGE, dtype=torch.float16).pin_memory()
    host_in.numpy()[...] = rng.random(host_in.shape, dtype=np.float32).astype(np.float16)
    images = [Image(ctx, B, *IMAGE) for _ in range(N_INPUT_SETS)]
    n_in = B * IMAGE[0] * IMAGE[1] * IMAGE[2]
    for i, img in enumerate(images):
        img.copyFromPointer(host_in[i].data_ptr(), n_in, stream.cuda_stream)
    ctx.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------------------------
    for i in range(W):
        nn.encode(sourceImages=[images[i % N_INPUT_SETS]])
    ctx.synchronize()
    sampler = ClockSampler(local_rank)
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    lo = sampler.mark()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(K):
            nn.encode(sourceImages=[images[i % N_INPUT_SETS]])
        e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    hi = sampler.mark()
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        elapsed_ms = sdist.max_over_ranks(elapsed_ms, dev)
    value = world * B * K / (elapsed_ms * 1e-3)
    launches = nn.numLaunches(B) * K

    # ---- end to end through the public API with host buffers --------------------------------------------------------
    copy_stream = torch.cuda.Stream(device=dev)
    dev_in = [Image(ctx, B, *IMAGE), Image(ctx, B, *IMAGE)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    host_out = torch.empty((2, B, 1000), dtype=torch.float32).pin_memory()  # double-buffered results
    out_np = [host_out[0].numpy(), host_out[1].numpy()]
    landed = [torch.cuda.Event(), torch.cuda.Event()]
    checksum = [0.0]

    def e2e_loop(n_steps: int) -> float:
        """Every step: H2D of its input batch (copy stream), encode, D2H of its logits into pinned host memory; the host reads
        step i-1's logits while step i runs (one step of queue depth, like committing the next Metal command buffer before
        waiting on the previous one), and the last step's logits before the clock stops."""
        t0 = time.perf_counter()
        dev_in[0].copyFromPointer(host_in[0].data_ptr(), n_in, copy_stream.cuda_stream)  # step 0's upload, inside the timed region
        copied[0].record(copy_stream)
        for i in range(n_steps):
            cur, nxt = i % 2, (i + 1) % 2
            if i + 1 < n_steps:  # overlap the next batch's upload with this batch's compute (double-buffered source images)
                if i > 0:
                    copy_stream.wait_event(landed[nxt])  # step i-1 read dev_in[nxt]: do not overwrite it before that encode is done
                dev_in[nxt].copyFromPointer(host_in[(i + 1) % N_INPUT_SETS].data_ptr(), n_in, copy_stream.cuda_stream)
                copied[nxt].record(copy_stream)
            stream.wait_event(copied[cur])
            res = nn.encode(sourceImages=[dev_in[cur]])
            res.toFloatArrayAsync(out_np[cur])   # fp32 logits -> pinned host buffer, enqueued behind the encode
            landed[cur].record(stream)
            if i > 0:
                landed[nxt].synchronize()         # step i-1's logits are on the host: use them
                checksum[0] += float(out_np[nxt][0, 0])
        landed[(n_steps - 1) % 2].synchronize()
        checksum[0] += float(out_np[(n_steps - 1) % 2][0, 0])
        return time.perf_counter() - t0

    e2e_loop(W)
    barrier()
    torch.cuda.synchronize()
    e2e_s = e2e_loop(K)
    torch.cuda.synchronize()
    barrier()
    if world > 1:
        e2e_s = sdist.max_over_ranks(e2e_s, dev)
    e2e_value = world * B * K / e2e_s
    clocks = sampler.stop(lo, hi)

    # ---- roofline of the dominant kernel (tcgen05 implicit-GEMM conv), measured live with CUDA events --------------------
    # Every kernel of one encode is bracketed by CUDA-event record nodes inside a captured graph (smelter_graph_profile), which
    # gives each kernel's duration in situ.  The event nodes themselves cost a few us per kernel and defeat PDL overlap, so the
    # per-kernel numbers are used for the conv kernels' SHARE of the step; the absolute duration is pinned to the event-timed
    # step above: conv_ms = ms_per_step x share.  Both the raw and the pinned figures are reported.
    prof = nn.profile([images[0]], iters=5, stream=stream.cuda_stream)
    conv = [p for p in prof if p["tensor"]]
    conv_ms_raw = sum(p["ms"] for p in conv)
    conv_flops = sum(p["flops"] for p in conv)
    total_ms_raw = sum(p["ms"] for p in prof)
    share = conv_ms_raw / total_ms_raw if total_ms_raw else 0.0
    step_ms = elapsed_ms / K
    conv_ms = step_ms * share
    pk = peaks()
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tflops_sustained"],
                "traffic": ncu_conv_traffic(), "kernel": f"conv_pair_kernel / conv_igemm_kernel <BLOCK_N,HAS_RES> ({len(conv)} conv/gemm launches per step, aggregated; "
                          f"{sum('+conv1x1(' in p['desc'] for p in conv)} of them carry a folded 1x1 projection shortcut)",
                "peak_source": f"{pk['source']} bf16 sustained (MEASURED_PEAKS.json)", "kernel_share_of_step": share,
                "launch_ms_sum": conv_ms, "launch_ms_sum_raw_with_event_nodes": conv_ms_raw, "flops_per_step": conv_flops,
                "achieved_raw_with_event_nodes": conv_flops / (conv_ms_raw * 1e-3) / 1e12 if conv_ms_raw > 0 else 0.0,
                "how": "algorithmic FLOPs (2*M*N*K, SURVEY 8d) / (event-timed step x conv share from in-graph per-kernel events)",
                "step_frac_of_conv_roofline": value / world / (pk["tflops_sustained"] * 1e12 / FLOPS_PER_IMAGE)}
    if rank == 0:
        outdir = os.path.join(ROOT, "gpurun_out")
        try:
            os.makedirs(outdir, exist_ok=True)
            with open(os.path.join(outdir, f"bench_profile_n{world}.json"), "w") as f:
                json.dump({"batch": B, "steps": prof, "plan": nn.planDump(B)}, f, indent=1)
        except OSError:
            pass

    line = {"metric": "ResNet-50 fp16 224x224 images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "ResNet-50 fp16 224x224 batch=32 per GPU (BASELINE.json configs[2]; N GPUs = configs[4] weak-scaled)",
                       "onnx_graph": "seeded random ResNet-50 (53 Conv+BN) -> ONNX2MPS --half -> libsmelter_b200", "global_batch": world * B,
                       "per_gpu_batch": B, "parallelism": f"batch-sharded dp{world}, one NCCL weight broadcast, no steady-state collective",
                       "l2": f"inputs larger than L2: {N_INPUT_SETS} distinct resident batches (154 MB) rotated, no flush",
                       "cuda_graph": True, "accumulate": "fp32 (TMEM)"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": n_in * 2, "d2h_bytes_per_step": B * 1000 * 4,
                    "ms_per_step": e2e_s / K * 1e3, "how": "pinned host fp16 batch -> Image.copyFromPointer (copy stream, double-buffered) -> encode -> "
                    "toFloatArrayAsync (fp32 logits into pinned host memory every step; the host consumes step i-1's logits while step i runs, "
                    "the last step's before the clock stops); wall clock"},
            "gpu_launches": launches, "launches_per_step": nn.numLaunches(B), "clocks": clocks, "roofline": roofline}
    if bcast_ms is not None:
        line["weight_broadcast_ms"] = bcast_ms
    if world == 1 and not args.no_cpu:
        base = cpu_reference(steps=3, warmup=1, budget_s=20.0)
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="images per GPU per step (the metric is quoted at 32)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 400), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())

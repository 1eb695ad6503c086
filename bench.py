#!/usr/bin/env python
"""Headline benchmark: ResNet-50 fp16 224x224 images/s through the drop-in inference path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one `encode` of one batch of synthetic images per GPU.  One GPU: batch 32 (BASELINE.json configs[2]).  N > 1 GPUs:
BASELINE.json configs[4] as written — GLOBAL batch 256 split 256 / N per rank (strong scaling; the 32-per-GPU weak-scaling figure,
the one-GPU batch-256 base and a bit-exact shard-parity check are measured next to it); the only collective is the one-time NCCL
weight-arena broadcast.  `--in-flight S` (default 2) keeps S encodes in flight per GPU, each on its own stream with its own
activation arena, the way Metal command buffers overlap; `one_in_flight` reports the strictly serial figure.  The model is the
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

PER_GPU_BATCH = 32      # BASELINE.json configs[2]
GLOBAL_BATCH = 256      # BASELINE.json configs[4]: split 256 / N over N GPUs
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
    """dram__bytes_read.sum + dram__bytes_write.sum of the conv kernels of one encode, summed over the launches of a step, from the
    newest committed `ncu --set full` capture of this same command (profiles/r*_encode_full.csv; cold-cache replay, so an upper
    bound on the in-situ traffic).  Returns (bytes, source file) or (None, reason)."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_encode_full.csv")))
    if not files:
        return None, "no committed ncu capture"
    path = files[-1]
    try:
        with open(path) as f:
            rows = list(csv.reader(f))
        head = rows[0]
        ir, iw, ik = head.index("dram_read"), head.index("dram_write"), head.index("kernel")
        total = sum(float(r[ir]) + float(r[iw]) for r in rows[1:] if "conv_" in r[ik])
        return total, os.path.relpath(path, ROOT) + " (committed `ncu --set full` capture of one whole-chip encode of this workload: cold-cache replay, not this run)"
    except (OSError, ValueError) as e:
        return None, f"unreadable capture: {e}"


def ncu_conv_traffic_warm():
    """The same sum from the newest committed warm-cache pass (profiles/r*_dram_insitu.csv: `ncu --cache-control none`, kernels
    serialised but caches left as the previous kernel left them -- what a kernel moves to and from DRAM inside the step)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_dram_insitu.csv")))
    if not files:
        return None, "no committed warm-cache capture"
    try:
        import csv
        total = 0.0
        with open(files[-1]) as f:
            for row in csv.reader(line for line in f if not line.startswith("# ")):
                if len(row) >= 5 and "conv_" in row[1]:
                    total += float(row[2]) + float(row[3])
        return total, os.path.relpath(files[-1], ROOT)
    except (OSError, ValueError) as e:
        return None, f"unreadable capture: {e}"


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
    world = max(1, int(os.environ.get("WORLD_SIZE", args.gpus or 1)))
    base = cpu_reference(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "ResNet-50 fp16 224x224 images/sec", "value": base["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"], "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the native arm's workload, graph and batch keys word for word (the keys that describe its GPU execution do not apply)
            "config": {"workload": ("ResNet-50 fp16 224x224 batch=32 on one GPU (BASELINE.json configs[2])" if world == 1 else
                                    f"ResNet-50 fp16 224x224 global batch={GLOBAL_BATCH} batch-sharded {GLOBAL_BATCH // world} per GPU over {world} GPU(s) "
                                    "(BASELINE.json configs[4])"),
                       "onnx_graph": "seeded random ResNet-50 (53 Conv+BN) -> ONNX2MPS --half",
                       "global_batch": PER_GPU_BATCH if world == 1 else GLOBAL_BATCH, "per_gpu_batch": PER_GPU_BATCH if world == 1 else GLOBAL_BATCH // world,
                       "note": "reference = Swift + Apple MPS, not runnable here; this arm is the fp32 oracle port of the same graph on the host "
                               "cores of rank 0, each step a bounded sample of the batch (cpu_baseline.sample)"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def _sets_for(batch: int) -> int:
    """Distinct resident input batches so that the rotation is larger than the 126 MB L2 (at least two)."""
    per = batch * IMAGE[0] * IMAGE[1] * IMAGE[2] * 2
    return max(2, min(N_INPUT_SETS, -(-160_000_000 // per)))


class Runner:
    """One graph on one context; `encodes(K, S)` times K back-to-back encodes with S of them in flight (one stream each)."""

    def __init__(self, torch, ctx, nn, dev, batch, seed, max_in_flight):
        import numpy as np
        from smelter_b200.api import Image

        self.torch, self.ctx, self.nn, self.B = torch, ctx, nn, batch
        self.sets = _sets_for(batch)
        self.n_in = batch * IMAGE[0] * IMAGE[1] * IMAGE[2]
        rng = np.random.default_rng(seed)
        self.host_in = torch.empty((self.sets, batch) + IMAGE, dtype=torch.float16).pin_memory()
        self.host_in.numpy()[...] = rng.random(self.host_in.shape, dtype=np.float32).astype(np.float16)
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(max_in_flight)]
        self.images = [Image(ctx, batch, *IMAGE) for _ in range(self.sets)]
        for i, img in enumerate(self.images):
            img.copyFromPointer(self.host_in[i].data_ptr(), self.n_in, self.streams[0].cuda_stream)
        torch.cuda.synchronize()

    def encodes(self, K: int, S: int, e0=None, e1=None):
        """K encodes round-robin over S streams.  With events: e0 is recorded before the first encode can start on any stream, e1
        after the last encode of every stream has finished (fork / join through events on stream 0)."""
        torch = self.torch
        st = self.streams[:S]
        if e0 is not None:
            e0.record(st[0])
            for s in st[1:]:
                s.wait_event(e0)
        for i in range(K):
            self.nn.encode(to=st[i % S].cuda_stream, sourceImages=[self.images[i % self.sets]])
        if e1 is not None:
            for s in st[1:]:
                ev = torch.cuda.Event()
                ev.record(s)
                st[0].wait_event(ev)
            e1.record(st[0])

    def timed(self, K: int, W: int, S: int, barrier, max_over_ranks) -> float:
        torch = self.torch
        self.encodes(max(W, S), S)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        self.encodes(K, S, e0, e1)
        torch.cuda.synchronize()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))


def other_configs(torch, ctx, stream) -> dict:
    """BASELINE.json configs[1] and configs[3] at their named shapes, timed the same way (resident input, CUDA events, graph replay)."""
    import numpy as np
    from smelter_b200 import modelzoo, onnx2mps
    from smelter_b200.api import Configuration, Image, ONNXGraph

    out = {}
    for name, model, shape, iters, gflop in (
            ("MobileNetV2 fp16 1x3x224x224 (configs[1])", modelzoo.mobilenet_v2(seed=0, fold_bn=False), (1, 3, 224, 224), 300, 0.6015),
            ("TransformerNet fp16 1x3x512x512 (configs[3])", modelzoo.transformer_net(seed=0, hw=512), (1, 3, 512, 512), 100, 80.63)):
        g = ONNXGraph(onnx2mps.convert_bytes(model.serialize(), half=True), Configuration(), context=ctx)
        nn = g.metalGraph()
        img = Image.fromArray(ctx, np.random.default_rng(1).random(shape, dtype=np.float32).astype(np.float16))
        for _ in range(5):
            nn.encode(to=stream.cuda_stream, sourceImages=[img])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            nn.encode(to=stream.cuda_stream, sourceImages=[img])
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[name] = {"ms_per_encode": ms, "images_per_s": shape[0] / ms * 1e3, "launches": nn.numLaunches(shape[0]), "tflops": gflop * shape[0] / ms}
        g.close()
    return out


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
    t_init = time.perf_counter()
    if world > 1:
        sdist.init_process_group("nccl")
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    def max_over_ranks(v: float) -> float:
        return sdist.max_over_ranks(v, dev) if world > 1 else v

    # Workload: one GPU = BASELINE.json configs[2] (batch 32); N > 1 = configs[4]: ResNet-50, GLOBAL batch 256 split 256 / N per rank
    # (strong scaling; the 32-per-GPU weak-scaling figure is measured next to it).
    K, W = args.steps, max(args.warmup, 3)
    G = args.global_batch or (PER_GPU_BATCH if world == 1 else GLOBAL_BATCH)
    if args.batch:
        G = args.batch * world
    if G % world:
        raise SystemExit(f"global batch {G} does not split over {world} ranks")
    B = G // world
    # Every layer of the step is latency-bound (a few tiles per SM between two kernel boundaries, ~4 us of hand-over per launch): the
    # throughput graph is planned for half of the SMs (Configuration.smShare = 2) and three encodes are kept in flight (four from 128
    # images per GPU), so two of them always co-run on disjoint halves of the chip.  Measured per-GPU batch 16 / 32 / 64 / 128:
    # 65.7 k / 73.3 k / 77.7 k / 78.2 k images/s against 50.6 k / 60.8 k / 74.4 k / 74.5 k with whole-chip kernels (profiles/r2_batch_share.txt).
    share = args.sm_share or 2
    S = max(1, args.in_flight or (4 if B >= 128 else 3))
    stream = torch.cuda.Stream(device=dev)
    ctx = Context(local_rank, stream=stream.cuda_stream)
    data = model_bytes()
    # Two plans of the same model: smShare = 2 (kernels sized for half of the SMs; the small-batch throughput plan) and the whole-chip
    # plan (latency of one encode with nothing else in flight, per-kernel roofline, large batches).  Each holds its own weight replica.
    graphs = {k: ONNXGraph(data, Configuration(deferWeights=(world > 1 and rank != 0), smShare=k), context=ctx) for k in sorted({1, 2, share})}
    assert all(g.modelFormat == Format.mpsFlavor for g in graphs.values())
    nns = {k: g.metalGraph() for k, g in graphs.items()}
    graph, nn, nn_full = graphs[share], nns[share], nns[1]
    comm_ms = bcast_ms = None
    if world > 1:  # one-time weight replica over NVLink
        uid = sdist.share_bytes(Context.ncclUniqueId() if rank == 0 else b"", 0)
        barrier()
        t0 = time.perf_counter()
        ctx.initNCCL(uid, rank, world)
        nn.broadcastWeights(0)          # first collective on the engine's communicator: includes its bring-up (channels, NVLS setup)
        ctx.synchronize()
        comm_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        t0 = time.perf_counter()
        nn.broadcastWeights(0)          # the 51 MB broadcast by itself
        ctx.synchronize()
        bcast_ms = (time.perf_counter() - t0) * 1e3
        for k, other in nns.items():
            if other is not nn:
                other.broadcastWeights(0)
        ctx.synchronize()
        checksum, _ = nn.weightChecksum()
        if not sdist.all_equal(checksum, dev):
            raise SystemExit("weight replicas differ after the broadcast")
    barrier()
    quiet.__exit__()

    run = Runner(torch, ctx, nn, dev, B, 1 + rank, max(S, 2))
    # ---- device-resident throughput: K encodes, S in flight (each on its own stream / activation arena) -----------------------
    sampler = ClockSampler(local_rank)
    time.sleep(0.3)
    lo = sampler.mark()
    elapsed_ms = run.timed(K, W, S, barrier, max_over_ranks)
    hi = sampler.mark()
    value = world * B * K / (elapsed_ms * 1e-3)
    launches = nn.numLaunches(B) * K
    # latency of one encode with nothing else in flight on the whole chip (what round 1 reported as the step)
    K1 = min(K, 300)
    run_full = run if nn_full is nn else Runner(torch, ctx, nn_full, dev, B, 1 + rank, 2)
    one_ms = run_full.timed(K1, W, 1, barrier, max_over_ranks) / K1

    # ---- end to end through the public API with host buffers, S batches in flight ----------------------------------------------
    copy_stream = torch.cuda.Stream(device=dev)
    n_in = run.n_in
    slots = max(4, 2 * S)   # uploads run this many steps ahead of the host's read-back
    dev_in = [Image(ctx, B, *IMAGE) for _ in range(slots)]
    copied = [torch.cuda.Event() for _ in range(slots)]
    landed = [torch.cuda.Event() for _ in range(slots)]
    host_out = torch.empty((slots, B, 1000), dtype=torch.float32).pin_memory()
    out_np = [host_out[i].numpy() for i in range(slots)]
    checksum = [0.0]

    def e2e_loop(n_steps: int) -> float:
        """Every step: H2D of its input batch (copy stream) into the slot's source image, encode on the slot's stream, D2H of its fp32
        logits into pinned host memory.  `slots` batches are in flight: before a slot is reused the host waits for that slot's
        previous logits and reads them (committing the next Metal command buffers before waiting on an earlier one); the last
        steps' logits are read before the clock stops."""
        t0 = time.perf_counter()
        for i in range(n_steps):
            s = i % slots
            if i >= slots:
                landed[s].synchronize()          # step i - slots is complete: its logits are on the host, its source image is free
                checksum[0] += float(out_np[s][0, 0])
            dev_in[s].copyFromPointer(run.host_in[i % run.sets].data_ptr(), n_in, copy_stream.cuda_stream)
            copied[s].record(copy_stream)
            st = run.streams[s % S]
            st.wait_event(copied[s])
            res = nn.encode(to=st.cuda_stream, sourceImages=[dev_in[s]])
            res.toFloatArrayAsync(out_np[s], stream=st.cuda_stream)  # fp32 logits -> pinned host buffer, enqueued behind the encode
            landed[s].record(st)
        for i in range(max(0, n_steps - slots), n_steps):
            landed[i % slots].synchronize()
            checksum[0] += float(out_np[i % slots][0, 0])
        return time.perf_counter() - t0

    e2e_loop(W + slots)
    barrier()
    torch.cuda.synchronize()
    e2e_s = e2e_loop(K)
    torch.cuda.synchronize()
    barrier()
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = world * B * K / e2e_s
    clocks = sampler.stop(lo, hi)
    # the host link alone: the same pinned -> device copies back to back with the GPU otherwise idle (what bounds e2e when the step is
    # shorter than its upload)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    h0.record(copy_stream)
    for i in range(32):
        dev_in[i % slots].copyFromPointer(run.host_in[i % run.sets].data_ptr(), n_in, copy_stream.cuda_stream)
    h1.record(copy_stream)
    torch.cuda.synchronize()
    h2d_alone_ms = h0.elapsed_time(h1) / 32

    # ---- multi-GPU: weak-scaling figure, 1-GPU batch-256 base, shard parity ----------------------------------------------------
    extra = {}
    if world > 1:
        if B == PER_GPU_BATCH:
            weak_value = value
        else:
            weak = Runner(torch, ctx, nns[2], dev, PER_GPU_BATCH, 101 + rank, 3)   # the small-batch plan: smShare 2, three in flight
            weak_value = world * PER_GPU_BATCH * K / (weak.timed(K, W, 3, barrier, max_over_ranks) * 1e-3)
            del weak
        extra["weak_scaling"] = {"per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH * world, "value": weak_value, "unit": "images/s"}
        # strong-scaling base: the whole global batch on ONE GPU (every rank times it on its own GPU; they do not share anything)
        if B == G:
            base_value = value
        else:
            k256 = max(20, min(K, 100))
            full = Runner(torch, ctx, nns[2], dev, G, 201, 4)   # the same kind of plan as the sharded run: smShare 2, four in flight
            base_value = G * k256 / (full.timed(k256, W, 4, barrier, max_over_ranks) * 1e-3)
        extra["strong_scaling"] = {"global_batch": G, "one_gpu_value": base_value, "efficiency_vs_one_gpu": value / (world * base_value),
                                   "note": "value / (N x the same global batch on one GPU, measured in this run)"}
        # shard parity: every rank holds the SAME seeded global batch, encodes its contiguous slice, rank 0 compares the gathered
        # logits with its own one-GPU encode of the whole batch.  Bit-exact with one k-reduction order per layer (SMELTER_NO_SPLITK=1:
        # the default plans may split K for some batch sizes, which only re-associates fp32 sums; that difference is reported too).
        xg = np.random.default_rng(7).random((G,) + IMAGE, dtype=np.float32).astype(np.float16)
        lo_i, hi_i = sdist.shard_range(G, rank, world)

        def logits(nn_, x):
            r = nn_.encode(sourceImages=[Image.fromArray(ctx, x)])
            return torch.from_numpy(r.toHalfArray().reshape(x.shape[0], -1).copy()).to(dev)  # fp16: all_gather moves the bits

        def gathered(nn_):
            mine = logits(nn_, xg[lo_i:hi_i])
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            return torch.cat(parts).cpu().numpy()

        g_default = gathered(nn)
        os.environ["SMELTER_NO_SPLITK"] = "1"   # read when a plan is made: fresh graph, same weights
        graph2 = ONNXGraph(data, Configuration(), context=ctx)
        nn2 = graph2.metalGraph()
        g_exact = gathered(nn2)
        if rank == 0:
            full_exact = logits(nn2, xg).cpu().numpy()
            full_default = logits(nn, xg).cpu().numpy()
            extra["shard_parity"] = bool(np.array_equal(g_exact.view(np.uint16), full_exact.view(np.uint16)))
            extra["shard_parity_how"] = {"bit_exact_with_one_k_order_per_layer": extra["shard_parity"],
                                         "default_plans_max_abs_diff": float(np.abs(g_default.astype(np.float32) - full_default.astype(np.float32)).max()),
                                         "images": G, "gathered_over": "torch.distributed all_gather of fp16 logits"}
        del os.environ["SMELTER_NO_SPLITK"]
        graph2.close()

    # ---- roofline of the dominant kernel (tcgen05 implicit-GEMM conv), measured live with CUDA events --------------------
    # Every kernel of one encode is bracketed by CUDA-event record nodes inside a captured graph (smelter_graph_profile), which
    # gives each kernel's duration in situ.  The event nodes themselves cost a few us per kernel and defeat PDL overlap, so the
    # per-kernel numbers are used for the conv kernels' SHARE of the step; the absolute duration is pinned to the event-timed
    # step above: conv_ms = ms_per_step x share.  Both the raw and the pinned figures are reported.
    prof = nn_full.profile([run.images[0]], iters=5, stream=stream.cuda_stream)
    conv = [p for p in prof if p["tensor"]]
    conv_ms_raw = sum(p["ms"] for p in conv)
    conv_flops = sum(p["flops"] for p in conv)
    total_ms_raw = sum(p["ms"] for p in prof)
    conv_share = conv_ms_raw / total_ms_raw if total_ms_raw else 0.0
    step_ms = elapsed_ms / K
    conv_ms = step_ms * conv_share
    pk = peaks()
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    traffic, traffic_src = ncu_conv_traffic()
    traffic_warm, traffic_warm_src = ncu_conv_traffic_warm()
    roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tflops_sustained"],
                "frac_of_burst_peak": achieved / pk["tflops_burst"], "burst_peak": pk["tflops_burst"],
                "traffic": traffic, "traffic_source": traffic_src, "traffic_warm_cache": traffic_warm, "traffic_warm_cache_source": traffic_warm_src,
                "kernel": f"conv_pair_kernel / conv_igemm_kernel <BLOCK_N,HAS_RES> ({len(conv)} conv/gemm launches per step, aggregated; "
                          f"{sum('+conv1x1(' in p['desc'] for p in conv)} of them carry a folded 1x1 projection shortcut)",
                "peak_source": f"{pk['source']} bf16 sustained (MEASURED_PEAKS.json); frac_of_burst_peak uses bf16_tflops", "kernel_share_of_step": conv_share,
                "launch_ms_sum": conv_ms, "launch_ms_sum_raw_with_event_nodes": conv_ms_raw, "flops_per_step": conv_flops,
                "achieved_raw_with_event_nodes": conv_flops / (conv_ms_raw * 1e-3) / 1e12 if conv_ms_raw > 0 else 0.0,
                "achieved_one_in_flight": conv_flops / (one_ms * conv_share * 1e-3) / 1e12 if one_ms > 0 else 0.0,
                "how": "algorithmic FLOPs (2*M*N*K, SURVEY 8d) / (event-timed step x conv share from in-graph per-kernel events); "
                       f"step = elapsed / steps with {S} encodes in flight",
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
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": ("ResNet-50 fp16 224x224 batch=32 on one GPU (BASELINE.json configs[2])" if world == 1 and G == PER_GPU_BATCH else
                                    f"ResNet-50 fp16 224x224 global batch={G} batch-sharded {B} per GPU over {world} GPU(s) (BASELINE.json configs[4])"),
                       "onnx_graph": "seeded random ResNet-50 (53 Conv+BN) -> ONNX2MPS --half", "engine": "libsmelter_b200 (C ABI, ctypes)", "global_batch": G,
                       "per_gpu_batch": B, "parallelism": f"batch-sharded dp{world}, one NCCL weight broadcast, no steady-state collective",
                       "in_flight": S, "sm_share": share,
                       "in_flight_note": f"{S} encodes in flight per GPU, each on its own stream with its own activation arena "
                       "(the reference's encode(to: commandBuffer) is asynchronous in the same way)" +
                       (f", kernels planned for 1/{share} of the SMs (Configuration.smShare) so that {share} encodes co-run" if share > 1 else "") +
                       "; ms_per_step = elapsed / steps; one_in_flight = one encode at a time on the whole chip",
                       "l2": f"inputs larger than L2: {run.sets} distinct resident batches ({run.sets * run.n_in * 2 / 1e6:.0f} MB) rotated, no flush",
                       "cuda_graph": True, "accumulate": "fp32 (TMEM)"},
            "one_in_flight": {"ms_per_step": one_ms, "value": world * B / (one_ms * 1e-3), "unit": "images/s", "steps": K1},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": n_in * 2, "d2h_bytes_per_step": B * 1000 * 4,
                    "ms_per_step": e2e_s / K * 1e3, "h2d_alone_ms_per_step": h2d_alone_ms, "h2d_alone_gbs": n_in * 2 / (h2d_alone_ms * 1e-3) / 1e9,
                    "link_bound_images_per_s": world * B / (h2d_alone_ms * 1e-3), "how": "pinned host fp16 batch -> Image.copyFromPointer (copy stream) -> encode -> "
                    f"toFloatArrayAsync (fp32 logits into pinned host memory every step), {slots} batches in flight: the host reads a slot's logits "
                    "before it reuses the slot, the last steps' before the clock stops; wall clock"},
            "gpu_launches": launches, "launches_per_step": nn.numLaunches(B), "clocks": clocks, "roofline": roofline}
    line.update(extra)
    if comm_ms is not None:
        line["weight_broadcast_ms"] = bcast_ms
        line["weight_broadcast"] = {"communicator_bring_up_plus_first_broadcast_ms": comm_ms, "broadcast_ms": bcast_ms, "bytes": nn.weightChecksum()[1]}
    if world == 1 and not args.no_extra:
        try:
            line["other_configs"] = other_configs(torch, ctx, stream)
        except Exception as e:  # never lose the headline line to an extra
            line["other_configs"] = {"error": str(e)[:200]}
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
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: 32 on one GPU, 256 / N on N GPUs)")
    ap.add_argument("--global-batch", type=int, default=0, help="global batch split over the ranks (default: 32 for one GPU, 256 for N > 1)")
    ap.add_argument("--in-flight", type=int, default=0, help="encodes in flight per GPU, each on its own stream (1 = strictly one at a time; "
                    "default: 3, or 4 from 128 images per GPU)")
    ap.add_argument("--sm-share", type=int, default=0, help="Configuration.smShare of the throughput graph: kernels sized for 1/k of the SMs so that "
                    "encodes in flight co-run (default 2)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[1] / configs[3] latency fields")
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

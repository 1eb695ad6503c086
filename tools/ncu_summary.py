#!/usr/bin/env python
"""Turn an ncu report (or a launch-list csv) from gpurun_out/ into the small tracked summaries under profiles/.

  python tools/ncu_summary.py full  gpurun_out/encode_r1.ncu-rep profiles/r1_encode_full.csv
  python tools/ncu_summary.py list  gpurun_out/launches_r1.csv   profiles/r1_launches.csv

`full`: one row per profiled launch with duration, grid, registers, tensor-pipe activity, DRAM bytes / throughput and
the L2 / SM throughput percentages of an `ncu --set full --clock-control none` capture.
`list`: the `--metrics gpu__time_duration.sum` pass: per launch time, plus each kernel's share of the captured window.
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "dur_us"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct_active"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed", "hmma_pct_elapsed"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
]


def to_bytes(value: str, unit: str) -> float:
    v = float(value.replace(",", "")) if value not in ("", "n/a") else 0.0
    u = unit.lower()
    for k, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
        if k in u:
            return v * m
    return v


def to_us(value: str, unit: str) -> float:
    v = float(value.replace(",", "")) if value else 0.0
    u = unit.lower()
    return v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(u, 1.0)


def short(name: str) -> str:
    name = name.replace("void ", "").replace("smelter::k::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("unnamed>::", "")
    return name.split("(")[0]


def full(src: str, dst: str) -> None:
    if src.endswith(".csv"):  # already exported on the GPU box with `ncu -i rep --page raw --csv`
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    ix = {n: i for i, n in enumerate(head)}
    cols = [(m, label) for m, label in WANT if m in ix]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["#", "kernel"] + [label for _, label in cols])
        for n, r in enumerate(body):
            out = [n, short(r[ix["Kernel Name"]])]
            for m, label in cols:
                v, u = r[ix[m]], units[ix[m]]
                if label == "dur_us":
                    out.append(f"{to_us(v, u):.2f}")
                elif label.startswith("dram_r") or label.startswith("dram_w"):
                    out.append(f"{to_bytes(v, u):.0f}")
                else:
                    out.append(v.replace(",", ""))
            w.writerow(out)
    print(f"{dst}: {len(body)} launches, metrics: {[l for _, l in cols]}")
    missing = [m for m, _ in WANT if m not in ix]
    if missing:
        print("not in report:", missing)


def insitu(src: str, dst: str) -> None:
    """`ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log -> one row per
    launch: kernels serialised but caches left warm, i.e. the DRAM bytes a kernel moves inside the step."""
    txt = open(src).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    agg = {}
    for r in rows:
        agg.setdefault((int(r["ID"]), r["Kernel Name"]), {})[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    with open(dst, "w", newline="") as f:
        f.write("# ncu --cache-control none --clock-control none: one whole-chip ResNet-50 batch-32 encode, kernels serialised, caches left warm\n")
        w = csv.writer(f)
        w.writerow(["#", "kernel", "dram_read_bytes", "dram_write_bytes", "dur_us"])
        tr = tw = tt = 0.0
        for (i, name), m in sorted(agg.items()):
            rd = to_bytes(*m["dram__bytes_read.sum"])
            wr = to_bytes(*m["dram__bytes_write.sum"])
            us = to_us(*m["gpu__time_duration.sum"])
            tr += rd; tw += wr; tt += us
            w.writerow([i, short(name), int(rd), int(wr), round(us, 2)])
        w.writerow(["total", "", int(tr), int(tw), round(tt, 2)])
    print(f"{dst}: {len(agg)} launches, DRAM read {tr / 1e6:.1f} MB, written {tw / 1e6:.1f} MB")


def launch_list(src: str, dst: str) -> None:
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    items = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        items.append((int(r["ID"]), short(r["Kernel Name"]), to_us(r["Metric Value"], r["Metric Unit"]), r.get("Grid Size", ""), r.get("Block Size", "")))
    total = sum(t for _, _, t, _, _ in items) or 1.0
    by = {}
    for _, k, t, _, _ in items:
        n, s = by.get(k, (0, 0.0))
        by[k] = (n + 1, s + t)
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["# per-kernel share of the captured window (cold-cache, serialised ncu replay: compare shares, not absolutes)"])
        w.writerow(["kernel", "launches", "sum_us", "share"])
        for k, (n, s) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{s:.1f}", f"{s / total:.4f}"])
        w.writerow([])
        w.writerow(["id", "kernel", "dur_us", "grid", "block"])
        for i, k, t, g, b in items:
            w.writerow([i, k, f"{t:.2f}", g, b])
    print(f"{dst}: {len(items)} launches, {total:.1f} us total")
    for k, (n, s) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"  {s / total:6.1%} {n:4d} x {k}")


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "insitu":
        insitu(sys.argv[2], sys.argv[3])
        sys.exit(0)
    if len(sys.argv) != 4 or sys.argv[1] not in ("full", "list"):
        sys.exit(__doc__)
    (full if sys.argv[1] == "full" else launch_list)(sys.argv[2], sys.argv[3])

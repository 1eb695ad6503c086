#!/usr/bin/env python
"""Top warp-stall instructions per profiled launch from an `ncu --set full --import-source on` report.

  python tools/ncu_hotspots.py gpurun_out/conv_src_r1.ncu-rep profiles/r1_conv_source_hotspots.txt [top]
"""
import csv
import io
import subprocess
import sys


def main(src: str, dst: str, top: int = 14) -> None:
    raw = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    launches, cur = [], None
    for row in csv.reader(io.StringIO(raw)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "head": None}
            launches.append(cur)
        elif cur is not None and row[0] == "Address":
            cur["head"] = row
        elif cur is not None and cur["head"] is not None and row[0].startswith("0x"):
            cur["rows"].append(row)
    with open(dst, "w") as f:
        f.write("# ncu --set full --import-source on: per-SASS-instruction warp-stall samples of the captured conv launches of one\n"
                f"# ResNet-50 batch-32 encode (tools/gpu_round.sh); top {top} instructions by samples per launch\n")
        for l in launches:
            ix = {n: i for i, n in enumerate(l["head"])}
            col = ix.get("# Samples", ix.get("Warp Stall Sampling (All Samples)"))
            rows = [(int(r[col] or 0), r[ix["Source"]].strip()) for r in l["rows"]]
            total = sum(n for n, _ in rows) or 1
            f.write(f"\n== {l['name'][:100]}  total samples {total}\n")
            for n, ins in sorted(rows, key=lambda t: -t[0])[:top]:
                f.write(f"  {n:5d} {n / total:6.1%}  {ins}\n")
    print(f"{dst}: {len(launches)} launches")


if __name__ == "__main__":
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 14)

"""TMA load-rate probes (B200).  See DESIGN.md section 5 for the findings."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smelter_b200 import _lib as L
from smelter_b200.api import Context, _check
ctx = Context(0)
lib = L.lib()
def run(label, mode, c, stages, iters, grid, extra, kib_per_iter):
    ms = C.c_float()
    _check(lib.smelter_tma_probe(ctx._h, mode, c, 56, 56, 32, stages, iters, grid, extra, C.byref(ms)))
    gb = grid * iters * kib_per_iter * 1024 / (ms.value * 1e-3) / 1e9
    print(f"{label:44s} stages={stages} grid={grid:3d}: {ms.value*1e3/iters:7.3f} us/iter  {gb/grid:6.1f} GB/s/SM delivered  {gb/1e3:6.2f} TB/s total", flush=True)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "issuers"):
    run("1 issuer, 16KiB box", 0, 256, 4, 400, 148, 1, 16)
    for n in (2, 4, 6):
        run(f"{n} issuer warps x 16KiB", 5, 256, 2, 400, 148, n, 16 * n)
    for slabs in (2, 4):
        run(f"1 issuer, 3-D box {slabs}x16KiB per instr", 4, 256, 2, 400, 148, slabs, 16 * slabs)
if which in ("all", "overlap"):
    # rows of 64 elements (128 B) at a pitch of 16 elements (32 B): the stem's packed-row operand
    for pitch in (16, 32, 64, 128):
        run(f"tiled noswizzle {{64,128}}, row pitch {pitch*2} B", 2, pitch, 4, 400, 148, 64 * 1000 + 128, 16)
if which in ("sweep",):
    # per-SM ingest ceiling: does the delivered rate per SM change with the number of SMs pulling?  (port limit vs chip limit)
    for grid in (8, 37, 74, 148):
        for n, st in ((4, 2), (6, 2), (7, 2)):
            run(f"{n} issuer warps x 16KiB", 5, 256, st, 400, grid, n, 16 * n)

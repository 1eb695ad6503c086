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
    print(f"{label:40s} stages={stages} grid={grid:3d}: {ms.value*1e3/iters:7.3f} us/iter  {gb/grid:6.1f} GB/s/SM delivered  {gb/1e3:6.2f} TB/s total", flush=True)
for n in (1, 2, 4, 6):
    run(f"{n} issuer LANES of one warp x 16KiB", 6, 256, 2, 400, 148, n, 16 * n)
    run(f"{n} issuer LANES of one warp x 16KiB", 6, 256, 2, 400, 1, n, 16 * n)
for n in (2, 4, 6):
    run(f"{n} issuer warps x 16KiB", 5, 256, 2, 400, 148, n, 16 * n)

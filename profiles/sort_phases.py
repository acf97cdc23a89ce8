"""Where a tile of the Onesweep pass spends its life: thread 0 of every CTA adds the device-clock time of each phase
to a global table (diagnostics build only).
    RTR_BUILD_ONLY=sort.cu RTR_NVCC_EXTRA=-DRTR_SORT_PHASE_CLOCKS python -m realtimeraytracing_b200.build --force
    python profiles/sort_phases.py"""
import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np
from realtimeraytracing_b200 import capi, synth
names = ["ticket + barrier", "tile load (TMA wait, keys to registers)", "tile histogram + publish aggregate",
         "ranking (+ interleaved look-back)", "rest of the look-back", "barrier + (warp, digit) slots",
         "reorder through shared memory", "scatter issued"]
ctx = capi.Context(0)
L = capi.load_library()
n = 10_000_000
keys = synth.random_keys_u32(n, seed=1); vals = np.arange(n, dtype=np.uint32)
d_k, d_v = ctx.dev_alloc(4 * n), ctx.dev_alloc(4 * n)
out = (C.c_ulonglong * 8)()
for rep in range(3):
    ctx.upload(d_k, keys); ctx.upload(d_v, vals)
    L.rtr_debug_sort_phases(out, 1)
    ctx.sort_pairs_u32_dev(d_k, d_v, n, 0, 32); ctx.sync()
L.rtr_debug_sort_phases(out, 0)
tiles = 4 * ((n + 6143) // 6144)
tot = sum(out)
print("per tile (4 passes x %d tiles), microseconds:" % (tiles // 4))
for nm, v in zip(names, out):
    print("  %-44s %6.2f us  %4.1f %%" % (nm, v / tiles / 1e3, 100.0 * v / tot))
print("  %-44s %6.2f us" % ("tile lifetime", tot / tiles / 1e3))

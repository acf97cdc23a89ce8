"""Times the Onesweep sort of 10 M (u32 key, u32 index) pairs and of 10 M / 1 M u32 keys (CUDA events, median of reps).
    python profiles/sort_time.py [reps]"""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from realtimeraytracing_b200 import capi, synth
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx = capi.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
for n, pairs, bits in ((10_000_000, True, 32), (10_000_000, True, 30), (10_000_000, False, 32), (1_000_000, False, 32)):
    keys = synth.random_keys_u32(n, seed=1)
    if bits == 30:
        keys = keys >> np.uint32(2)
    vals = np.arange(n, dtype=np.uint32)
    d_src, d_k, d_v = ctx.dev_alloc(4 * n), ctx.dev_alloc(4 * n), ctx.dev_alloc(4 * n)
    ctx.upload(d_src, keys)
    ms = []
    for r in range(reps + 2):
        t = torch.from_numpy(keys)  # keep alive
        ctx.upload(d_k, keys); ctx.upload(d_v, vals)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ctx.sort_pairs_u32_dev(d_k, d_v if pairs else None, n, 0, bits)
        e1.record(st)
        ctx.sync()
        if r >= 2:
            ms.append(e0.elapsed_time(e1))
    out = np.zeros(n, np.uint32); ctx.download(out, d_k)
    order = np.argsort(keys, kind="stable")
    ok = np.array_equal(out, keys[order])
    if pairs:
        ov = np.zeros(n, np.uint32); ctx.download(ov, d_v)
        ok = ok and np.array_equal(ov, order.astype(np.uint32))
    med = float(np.median(ms))
    print("n=%d %s bits=%d: %.1f us  %.1f Gkeys/s  correct=%s" % (n, "pairs" if pairs else "keys", bits, med * 1e3, n / med / 1e6, ok))
    for d in (d_src, d_k, d_v):
        ctx.dev_free(d)

#!/usr/bin/env python
"""Device time of the Onesweep sort of n (u32 key, u32 index) pairs and of n u32 keys (CUDA events inside the
library, keys re-uploaded before every sort).  RTR_NVCC_EXTRA="-DRTR_SORT_BLOCK=512" python profiles/time_sort.py"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n", type=int, default=10_000_000)
    p.add_argument("--reps", type=int, default=5)
    p.add_argument("--force-build", action="store_true")
    args = p.parse_args()
    import numpy as np
    from realtimeraytracing_b200 import build as rbuild, capi, synth
    rbuild.build(force=args.force_build)
    with capi.Context(0) as ctx:
        n = args.n
        keys = synth.random_keys_u32(n, seed=1)
        vals = np.arange(n, dtype=np.uint32)
        d_k = ctx.dev_alloc(keys.nbytes); d_v = ctx.dev_alloc(vals.nbytes)
        out = {}
        for pairs in (True, False):
            tot = {}
            for rep in range(args.reps + 1):
                ctx.upload(d_k, keys); ctx.upload(d_v, vals)
                ctx.profile_enable(rep > 0)
                ctx.sort_pairs_u32_dev(d_k, d_v if pairs else None, n)
                if rep > 0:
                    for k, v in ctx.profile_read().items():
                        tot[k] = tot.get(k, 0.0) + v[0]
                ctx.profile_enable(False)
            ms = sum(tot.values()) / args.reps
            out["pairs" if pairs else "keys"] = ms
            if pairs:
                got_k = np.zeros(n, np.uint32); got_v = np.zeros(n, np.uint32)
                ctx.download(got_k, d_k); ctx.download(got_v, d_v)
                order = np.argsort(keys, kind="stable")
                assert np.array_equal(got_k, keys[order]) and np.array_equal(got_v, order.astype(np.uint32)), "sort is wrong"
        bytes_pairs, bytes_keys = 68.0 * n, 36.0 * n
        print("RTR_NVCC_EXTRA=%r n=%d  pairs %.3f ms = %.1f Gkeys/s = %.0f GB/s algorithmic | keys %.3f ms = %.1f Gkeys/s = %.0f GB/s (kernel time only)" % (
            os.environ.get("RTR_NVCC_EXTRA", ""), n, out["pairs"], n / out["pairs"] / 1e6, bytes_pairs / out["pairs"] / 1e6,
            out["keys"], n / out["keys"] / 1e6, bytes_keys / out["keys"] / 1e6))

if __name__ == "__main__":
    main()

import sys; sys.path.insert(0,'.')
import numpy as np
from realtimeraytracing_b200 import capi, synth
n=10_000_000
tris, meshes, L = synth.triangle_soup(n)
ctx = capi.Context(0)
bvh = capi.Bvh(ctx)
for rep in range(3):
    bvh.build(tris, meshes)
t = bvh.iteration_times().astype(np.int64)
a, m = bvh.iteration_trace()
d = np.diff(t)
print("iters", a.size, "loop total us", (t[-1]-t[0])/1e3)
for i in range(len(d)):
    print(i, int(a[i]), int(m[i]), "%.1f us" % (d[i]/1e3), "%.1f ps/cluster" % (d[i]*1e3/a[i]))

"""Debug helper: records the per-layer timeline of the MLP kernels on CTA 0 (adn_set_option "trace") and prints
issue / accumulator-ready / epilogue-done times so pipeline bubbles can be read off directly.
    python profiles/trace_mlp.py [net]"""
import ctypes as C
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adanerf_b200 import Renderer
from adanerf_b200 import synthetic

net = int(sys.argv[1]) if len(sys.argv) > 1 else 1
scene = synthetic.SCENE_BARBERSHOP
sd0, sd1 = synthetic.make_weights("rand", seed=0)
r = Renderer(scene, device=0, sampling_net=sd0, shading_net=sd1)
pose = torch.tensor(scene["view_cell_center"]); rot = torch.eye(3)
for _ in range(2):
    r.render_camera(pose, rot, 800, 800, 0.2, 8)
torch.cuda.synchronize()
r.set_option("trace", net)
r.render_camera(pose, rot, 800, 800, 0.2, 8)
torch.cuda.synchronize()
buf = np.zeros(65536, dtype=np.int64)
r.lib.adn_debug_read_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
assert r.lib.adn_debug_read_trace(r.handle, buf.ctypes.data, 65536) == 0
rows = []
for region in range(7):
    base = buf[region * 8192:(region + 1) * 8192]
    n = int(base[0])
    ev = base[2:2 + 2 * n].reshape(-1, 2)
    role = 1 if region == 0 else (5 if region == 5 else (6 if region == 6 else (2 if (region - 1) % 2 == 0 else 3)))
    rows += [(int(t), role, (int(c) >> 16) & 255, (int(c) >> 8) & 255, int(c) & 255) for t, c in ev]
n = len(rows)
t0 = min(x[0] for x in rows)
rows = sorted((x[0] - t0,) + x[1:] for x in rows)
names = {0: "mma wait", 1: "mma issue", 2: "mma committed", 3: "acc seen", 4: "epi done", 5: "kb: before w_full", 6: "kb: weights ready", 7: "kb: issued", 8: "prod: wait empty", 9: "prod: empty seen, copy", 10: "help: wait full", 11: "help: local full", 12: "help: peer full"}
print("events", n)
# skip the first 3 tiles of each slot, then print ~2 tiles worth of events
start = [i for i, x in enumerate(rows) if x[1] == 1 and x[3] == 0 and x[4] == 0][6] if n > 400 else 0
for x in rows[start:start + 130]:
    print(f"{x[0]:9d}  role={x[1]} slot={x[2]} layer={x[3]:2d}  {names[x[4]]}")
# per layer statistics over the whole trace
import collections
d = collections.defaultdict(dict)
tile = collections.Counter()
for t, role, g, l, e in rows:
    if role == 1 and e == 0 and l == 0:
        tile[g] += 1
    key = (g, tile[g] if role == 1 else None)
per = collections.defaultdict(list)
last = {}
for t, role, g, l, e in rows:
    last[(role, g, l, e)] = t
    if role == 1 and e == 2 and (1, g, l, 1) in last:
        per[("issue", l)].append(t - last[(1, g, l, 1)])
    if role == 2 and e == 3 and (1, g, l, 2) in last:
        per[("commit->acc_seen", l)].append(t - last[(1, g, l, 2)])
    if role == 2 and e == 4 and (2, g, l, 3) in last:
        per[("epilogue(e0)", l)].append(t - last[(2, g, l, 3)])
    if role == 3 and e == 4 and (3, g, l, 3) in last:
        per[("epilogue(eLast)", l)].append(t - last[(3, g, l, 3)])
    if role == 1 and e == 1 and (1, g, l, 0) in last:
        per[("mma wait", l)].append(t - last[(1, g, l, 0)])
for k in sorted(per):
    v = np.array(per[k][8:])
    if len(v):
        print(f"{k[0]:20s} layer {k[1]:2d}: median {int(np.median(v)):6d}  p90 {int(np.percentile(v, 90)):6d}  n={len(v)}")

kb = [x for x in rows if x[1] == 1 and x[4] in (5, 6, 7)]
w, i = [], []
for a, b, c in zip(kb[0::3], kb[1::3], kb[2::3]):
    if a[4] == 5 and b[4] == 6 and c[4] == 7:
        w.append(b[0] - a[0]); i.append(c[0] - b[0])
w, i = np.array(w[16:]), np.array(i[16:])
print(f"layer-2 kb iterations: wait for weights median {int(np.median(w))} p90 {int(np.percentile(w,90))}; issue 4 MMAs + 2 commits median {int(np.median(i))} p90 {int(np.percentile(i,90))}  (n={len(w)})")

"""Debug helper: timeline of mlp_sh_kernel on CTA 0 (adn_set_option "trace", 1): per (layer, slot) issue / accumulator ready /
epilogue done, the issuers' wait per step, the producer's wait per stage, and the real SM clock of the launch (clock64
against globaltimer).
    python profiles/trace_sh.py"""
import ctypes as C
import collections
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adanerf_b200 import Renderer
from adanerf_b200 import synthetic

scene = synthetic.SCENE_BARBERSHOP
sd0, sd1 = synthetic.make_weights("rand", seed=0)
r = Renderer(scene, device=0, sampling_net=sd0, shading_net=sd1)
pose = torch.tensor(scene["view_cell_center"]); rot = torch.eye(3)
for _ in range(3):
    r.render_camera(pose, rot, 800, 800, 0.2, 8)
torch.cuda.synchronize()
r.set_option("trace", 1)
r.render_camera(pose, rot, 800, 800, 0.2, 8)
torch.cuda.synchronize()
buf = np.zeros(131072, dtype=np.int64)
r.lib.adn_debug_read_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
assert r.lib.adn_debug_read_trace(r.handle, buf.ctypes.data, 131072) == 0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "trace_sh_raw.npy"), buf)
c0, g0, c1, g1 = (int(x) for x in buf[7 * 8192:7 * 8192 + 4])
if g1 > g0:
    print(f"kernel (CTA 0): {c1 - c0} SM cycles in {(g1 - g0) / 1e3:.1f} us -> SM clock {(c1 - c0) / (g1 - g0) * 1e3:.0f} MHz")


def region(ri):
    base = buf[ri * 8192:(ri + 1) * 8192]
    n = int(base[0])
    e = base[2:2 + 2 * n].reshape(-1, 2)
    return [(int(t), ri, (int(c) >> 16) & 255, (int(c) >> 8) & 255, int(c) & 255) for t, c in e]


rows = []
for ri in range(5):
    rows += region(ri)
if not rows:
    raise SystemExit("no trace events (is the shading net on mlp_sh_kernel?)")
t0 = min(x[0] for x in rows)
rows = sorted((x[0] - t0,) + x[1:] for x in rows)
names = {0: "step: wait", 6: "step: go", 1: "layer: first step go", 2: "layer: last step issued", 3: "acc seen", 4: "epi done"}
who = {0: "issuer0", 1: "epi w0 slot0", 2: "epi w15 slot0", 3: "epi w0 slot1", 4: "epi w15 slot1"}
print("events", len(rows))
starts = [i for i, x in enumerate(rows) if x[1] == 0 and x[3] == 0 and x[4] == 1]
if len(starts) > 6:
    a, b = starts[5], starts[6]
    print(f"tile-pair period (issuer 0, layer 0 -> next): {rows[b][0] - rows[a][0]} cycles")
    for x in rows[a:b]:
        if x[4] in (0, 6):
            continue
        print(f"{x[0] - rows[a][0]:8d}  {who[x[1]]:14s} layer={x[3]}  {names[x[4]]}")
    per = [rows[starts[i + 1]][0] - rows[starts[i]][0] for i in range(2, len(starts) - 1)]
    print(f"tile-pair period: median {int(np.median(per))} min {min(per)} max {max(per)} (n={len(per)}); ideal tensor time 37376")
last = {}
per = collections.defaultdict(list)
for t, reg, g, l, e in rows:
    if reg == 0 and e in (1, 6) and (0, "w") in last:
        per[("issuer 0: wait per step", -1)].append(t - last[(0, "w")])
    if reg == 0 and e == 0:
        last[(0, "w")] = t
    if reg == 0 and e == 1:
        last[(0, l, 1)] = t
    if reg == 0 and e == 2 and (0, l, 1) in last:
        per[("issuer 0: layer issue time", l)].append(t - last[(0, l, 1)])
        last[(0, l, 2)] = t
    if reg in (1, 3) and e == 3:
        last[(reg, l, 3)] = t
        if reg == 1 and (0, l, 2) in last:
            per[("slot 0: last step issued -> acc seen", l)].append(t - last[(0, l, 2)])
    if reg in (1, 3) and e == 4 and (reg, l, 3) in last:
        per[("epilogue event (warp 0), slot %d" % (0 if reg == 1 else 1), l)].append(t - last[(reg, l, 3)])
for k in sorted(per):
    v = np.array(per[k][8:])
    if len(v):
        lab = "" if k[1] < 0 else f" layer {k[1]}"
        print(f"{k[0]:40s}{lab}: median {int(np.median(v)):6d}  p90 {int(np.percentile(v, 90)):6d}  n={len(v)}")
pr = region(5)
w = [b[0] - a[0] for a, b in zip(pr[0::2], pr[1::2]) if a[4] == 8 and b[4] == 9]
if w:
    w = np.array(w[16:])
    print(f"producer: wait for an empty stage median {int(np.median(w))} p90 {int(np.percentile(w, 90))} (n={len(w)})")

"""Debug helper: per-(layer, half) timeline of mlp_sh_kernel on CTA 0 (adn_set_option "trace", 1): issue / accumulator
ready / epilogue done, plus the real SM clock of the launch (clock64 against globaltimer).
    python profiles/trace_sh.py [workload: rand|shaped]"""
import ctypes as C
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adanerf_b200 import Renderer
from adanerf_b200 import synthetic

kind = sys.argv[1] if len(sys.argv) > 1 else "rand"
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
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'trace_sh_raw.npy'), buf)
c0, g0, c1, g1 = (int(x) for x in buf[7 * 8192:7 * 8192 + 4])
if g1 > g0:
    print(f"kernel (CTA 0): {c1 - c0} SM cycles in {(g1 - g0) / 1e3:.1f} us -> SM clock {(c1 - c0) / (g1 - g0) * 1e3:.0f} MHz")
rows = []
for region in range(3):
    base = buf[region * 8192:(region + 1) * 8192]
    n = int(base[0])
    ev = base[2:2 + 2 * n].reshape(-1, 2)
    rows += [(int(t), region, (int(c) >> 16) & 255, (int(c) >> 8) & 255, int(c) & 255) for t, c in ev]
if not rows:
    raise SystemExit("no trace events (is the shading net on mlp_sh_kernel?)")
t0 = min(x[0] for x in rows)
rows = sorted((x[0] - t0,) + x[1:] for x in rows)
names = {0: "step: before weight sync", 5: "step: weights ok", 6: "slot go", 1: "slot go (first step of half)", 2: "mma half issued", 3: "acc seen", 4: "epi done"}
print("events", len(rows))
# one tile pair in the middle of the trace
starts = [i for i, x in enumerate(rows) if x[1] == 0 and x[2] == 0 and x[3] == 0 and x[4] == 1]
if len(starts) > 6:
    a, b = starts[5], starts[6]
    print(f"tile-pair period (issuer, slot 0 layer 0 -> next): {rows[b][0] - rows[a][0]} cycles")
    for x in rows[a:b]:
        print(f"{x[0] - rows[a][0]:8d}  {'mma ' if x[1] == 0 else 'epi%d' % x[1]} slot={x[2]} layer={x[3] >> 1} half={x[3] & 1}  {names[x[4]]}")
    per = [rows[starts[i + 1]][0] - rows[starts[i]][0] for i in range(2, len(starts) - 1)]
    print(f"tile-pair period: median {int(np.median(per))} min {min(per)} max {max(per)} (n={len(per)}); ideal tensor time 37376")
# statistics
last = {}
per = collections.defaultdict(list)
for t, reg, g, lh, e in rows:
    last[(reg, g, lh, e)] = t
    if reg == 0 and e == 1 and (0, g, lh, 0) in last:
        per[("mma wait (sync)", lh)].append(t - last[(0, g, lh, 0)])
    if reg == 0 and e == 2 and (0, g, lh, 1) in last:
        per[("issue half", lh)].append(t - last[(0, g, lh, 1)])
    if reg in (1, 2) and e == 3 and (0, g, lh, 2) in last:
        per[("issued -> acc seen", lh)].append(t - last[(0, g, lh, 2)])
    if reg in (1, 2) and e == 4 and (reg, g, lh, 3) in last:
        per[("epilogue event slot %d" % g, lh)].append(t - last[(reg, g, lh, 3)])
for k in sorted(per):
    v = np.array(per[k][8:])
    if len(v):
        print(f"{k[0]:26s} layer {k[1] >> 1} half {k[1] & 1}: median {int(np.median(v)):6d}  p90 {int(np.percentile(v, 90)):6d}  n={len(v)}")

# producer (region 5) and helper (region 8)
def region(ri):
    base = buf[ri * 8192:(ri + 1) * 8192]
    n = int(base[0])
    e = base[2:2 + 2 * n].reshape(-1, 2)
    return [(int(t) - t0, (int(c) >> 16) & 255, (int(c) >> 8) & 255, int(c) & 255) for t, c in e]
pr = region(5)
w = [b[0] - a[0] for a, b in zip(pr[0::2], pr[1::2]) if a[3] == 8 and b[3] == 9]
if w:
    w = np.array(w[16:])
    print(f"producer: wait for an empty stage median {int(np.median(w))} p90 {int(np.percentile(w, 90))} (n={len(w)})")
hp = region(8)
d = collections.defaultdict(list)
cur = {}
for t, g, lh, e in hp:
    if e == 10:
        cur[g] = t
    elif e == 11 and g in cur:
        d[g].append(t - cur[g])
for g in sorted(d):
    v = np.array(d[g][16:])
    print(f"dependency helper, slot {g}: wait per sync point median {int(np.median(v))} p90 {int(np.percentile(v, 90))} mean {v.mean():.0f} (n={len(v)})")

# issuer: per step, time spent at the weight barrier and at the dependency barriers
iss = [x for x in rows if x[1] == 0]
wsync, dsync = [], collections.defaultdict(list)
prev = None
for t, reg, g, lh, e in iss:
    if e == 5 and prev and prev[4] == 0:
        wsync.append(t - prev[0])
    if e in (1, 6) and prev and prev[4] in (5, 1, 6, 2):
        dsync[(g, e)].append(t - prev[0])
    prev = (t, reg, g, lh, e)
ws = collections.defaultdict(list)
prev = None
for t, reg, g, lh, e in iss:
    if e in (1, 6) and prev and prev[4] == 0:
        ws[e].append(t - prev[0])
    prev = (t, reg, g, lh, e)
for e in sorted(ws):
    v = np.array(ws[e][16:])
    print(f"issuer 0, wait for all barriers of a step ({'first step of a half' if e == 1 else 'later step'}): median {int(np.median(v))} p90 {int(np.percentile(v, 90))} mean {v.mean():.0f} (n={len(v)})")
st = [b[0] - a[0] for a, b in zip([x for x in iss if x[4] == 0][:-1], [x for x in iss if x[4] == 0][1:])]
if st:
    v = np.array(st[16:])
    print(f"issuer 0, step period: median {int(np.median(v))} p90 {int(np.percentile(v, 90))} mean {v.mean():.0f}; 1024 tensor cycles per full step")
if wsync:
    v = np.array(wsync[16:])
    print(f"issuer at the weight barrier: median {int(np.median(v))} p90 {int(np.percentile(v, 90))} mean {v.mean():.0f}")
for k in sorted(dsync):
    v = np.array(dsync[k][16:])
    print(f"issuer before slot {k[0]} go ({'first step of half' if k[1] == 1 else 'later step'}): median {int(np.median(v))} p90 {int(np.percentile(v, 90))} mean {v.mean():.0f}")
hw = region(9); hp2 = region(10)
if hw and hp2:
    a = {(x[1], i): x[0] for i, x in enumerate(hw)}
    d = [p[0] - l[0] for l, p in zip(hw, hp2)]
    v = np.array(d[16:])
    print(f"weight helper: peer's share seen after the local one by median {int(np.median(v))} p90 {int(np.percentile(v, 90))} cycles")

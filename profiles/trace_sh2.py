"""Issuer micro-timeline of mlp_sh_kernel (library built with -DADN_SH_TRACE2=1): per step, cycles spent in the barrier wait,
the early probe, the fence and the MMA issue.  Reads gpurun_out/trace_sh_raw.npy written by profiles/trace_sh.py."""
import numpy as np, sys, os
buf = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "trace_sh_raw.npy"))
base = buf[0:8192]; n = int(base[0]); ev = base[2:2 + 2 * n].reshape(-1, 2)
E = [(int(t), (int(c) >> 16) & 255, (int(c) >> 8) & 255, int(c) & 255) for t, c in ev]
seg = {"wait": [], "probe": [], "fence": [], "issue": [], "loop": []}
last = {}
prev9 = None
for t, g, lh, e in E:
    if e == 0:
        if prev9 is not None: seg["loop"].append(t - prev9)
        last[0] = t
    elif e == 7 and 0 in last: seg["wait"].append(t - last[0]); last[7] = t
    elif e == 8 and 7 in last: seg["probe"].append(t - last[7]); last[8] = t
    elif e in (1, 6) and 8 in last: seg["fence"].append(t - last[8]); last[1] = t
    elif e == 9 and 1 in last: seg["issue"].append(t - last[1]); prev9 = t
for k, v in seg.items():
    if v:
        v = np.array(v[20:])
        print(f"{k:6s} median {np.median(v):7.0f}  mean {v.mean():7.0f}  p10 {np.percentile(v, 10):6.0f}  p90 {np.percentile(v, 90):6.0f}  n={len(v)}")

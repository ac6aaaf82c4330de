"""Frame-to-frame determinism stress: every workload rendered `reps` times, each frame compared bitwise with the first.
    python profiles/determinism_stress.py [reps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
import __graft_entry__ as ge
ge.build()
from adanerf_b200 import Renderer, synthetic

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
bad = 0
for name in ("800x800_thr0.2_K8_shaped", "800x800_thr0.2_K8", "800x800_pav_thr0.5_K16", "800x800_pav_thr0.3_K8", "800x800_ndc_thr0.15_K16", "800x800_dense_K128"):
    cfg = bench.WORKLOADS[name]
    W, H = cfg["W"], cfg["H"]
    r, scene, _, _ = bench.make_renderer_inputs(cfg, torch, Renderer, synthetic, 0, W, H)
    pose, rot = torch.tensor(scene["view_cell_center"], dtype=torch.float32), torch.eye(3)
    n = reps if cfg["K"] < 128 else max(2, reps // 10)
    first = r.render_camera(pose, rot, W, H, cfg["thr"], cfg["K"], want_nsamples=True)
    f_rgb, f_n = first["rgb"].clone(), first["n_samples"].clone()
    diff = 0
    for i in range(n):
        o = r.render_camera(pose, rot, W, H, cfg["thr"], cfg["K"], want_nsamples=True)
        if not (torch.equal(o["rgb"], f_rgb) and torch.equal(o["n_samples"], f_n)):
            diff += 1
            print(f"  {name}: frame {i + 1} differs in {int((o['rgb'] != f_rgb).any(1).sum())} rays")
    print(f"{name}: {n} frames, {diff} differ from the first")
    bad += diff
    r.close()
print("TOTAL differing frames:", bad)
sys.exit(1 if bad else 0)

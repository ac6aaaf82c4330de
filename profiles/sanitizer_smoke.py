"""Small mixed workload (adaptive K = 8 / 32, dense, aux outputs, fused encoder, metrics) for compute-sanitizer:
    compute-sanitizer --tool memcheck python profiles/sanitizer_smoke.py"""
import sys, torch, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as g
g.build()
from adanerf_b200 import Renderer, synthetic
scene = synthetic.SCENE_BARBERSHOP
sd0, sd1 = synthetic.make_weights("rand", seed=0)
r = Renderer(scene, device=0, sampling_net=sd0, shading_net=sd1)
pose = torch.tensor(scene["view_cell_center"]); rot = torch.eye(3)
for (rows, thr, K) in ((3, 0.2, 8), (1, 0.2, 32), (1, 0.0, 128)):
    o = r.render_camera(pose, rot, 800, 800, thr, K, row0=5, rows=rows, want_nsamples=True)
    print(rows, thr, K, bool(torch.isfinite(o["rgb"]).all()), int(o["n_samples"].sum()))
dirs = r.generate_ray_directions(800, 800, row0=0, rows=2)
a = r.render_rays(pose, rot, dirs, 0.2, 8, want_aux=True)
r.set_option("fuse_encoder", 1)
b = r.render_rays(pose, rot, dirs, 0.2, 8)
print("fused equal", torch.equal(a["rgb"], b["rgb"]))
m = r.image_metrics(a["rgb"], b["rgb"] + 0.01)
print(m)
r.close()
print("done")

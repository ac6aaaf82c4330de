"""One workload of bench.py rendered `warm` + 1 times with nothing else on the device, for ncu:
    ncu --set full --metrics lts__t_bytes.sum -s <warm * launches_per_frame> -c <launches of one chunk> python profiles/ncu_frame.py <workload> [warm]
Prints the launches per frame so the skip count can be checked."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
import __graft_entry__ as ge
ge.build()
from adanerf_b200 import Renderer, synthetic

name = sys.argv[1]
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = bench.WORKLOADS[name]
W, H = cfg["W"], cfg["H"]
r, scene, _, _ = bench.make_renderer_inputs(cfg, torch, Renderer, synthetic, 0, W, H)
pose, rot = torch.tensor(scene["view_cell_center"], dtype=torch.float32), torch.eye(3)
out = torch.empty((W * H, 3), dtype=torch.float32, device="cuda")
l0 = r.stats()["kernel_launches"]
for i in range(warm + 1):
    r.render_camera(pose, rot, W, H, cfg["thr"], cfg["K"], out=out)
    torch.cuda.synchronize()
    if i == 0:
        print("launches before the first frame:", l0, " per frame:", r.stats()["kernel_launches"] - l0)
print("samples (last chunk):", r.stats()["n_samples"], "finite:", bool(torch.isfinite(out).all()))
r.close()

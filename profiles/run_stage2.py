"""Runs stage 2 alone on a full 800x800 frame of sampling-net logits (bench weights), for ncu and timing."""
import sys
import torch
sys.path.insert(0, ".")
import __graft_entry__ as ge


def main():
    ge.build()
    from adanerf_b200 import Renderer
    from adanerf_b200 import synthetic
    scene = synthetic.SCENE_BARBERSHOP
    sd0, sd1 = synthetic.make_weights("rand", seed=0)
    r = Renderer(scene, device=0, sampling_net=sd0, shading_net=sd1)
    pose = torch.tensor(scene["view_cell_center"], dtype=torch.float32)
    dirs = r.generate_ray_directions(800, 800)
    x0, ro, rd = r.stage0(pose, torch.eye(3), dirs)
    raw0 = r.mlp0(x0)
    torch.cuda.synchronize()
    for _ in range(3):
        s2 = r.stage2(raw0, 0.2, 8)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        s2 = r.stage2(raw0, 0.2, 8)
    e1.record()
    torch.cuda.synchronize()
    cnt = s2["count"].float()
    print(f"stage2 incl. wrapper allocs: {e0.elapsed_time(e1) / 20:.4f} ms, total {s2['total']}, mean {cnt.mean():.3f}, "
          f"rays at K: {(cnt == 8).float().mean():.3f}")


if __name__ == "__main__":
    main()

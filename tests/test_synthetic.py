"""adanerf_b200.synthetic (bench / demo inputs) agrees with the oracle's seeded construction, so bench.py's GPU arm and
its CPU legs (which use the oracle) run the same networks."""
import torch

from adanerf_b200 import synthetic
from oracle import adanerf_oracle as orc


def test_rand_weights_equal_the_oracles():
    a0, a1 = synthetic.make_weights("rand", seed=0)
    b0, b1 = orc.make_weights("rand", seed=0)
    assert a0.keys() == b0.keys() and a1.keys() == b1.keys()
    assert all(torch.equal(a0[k], b0[k]) for k in a0) and all(torch.equal(a1[k], b1[k]) for k in a1)
    assert synthetic.SCENE_BARBERSHOP == orc.SCENE_BARBERSHOP


def test_shaped_recipe_matches_when_the_probe_logits_do():
    scene = orc.SCENE_BARBERSHOP
    dirs = torch.from_numpy(orc.generate_ray_directions(800, 800, scene["fov"]).reshape(-1, 3)[::157]).float()
    pose = torch.tensor(scene["view_cell_center"], dtype=torch.float32)

    def logits(sd0):
        x0, _, _ = orc.stage0_sphere_pos_dir(pose, torch.eye(3), dirs, scene)
        return orc.mlp0_forward(x0, sd0)

    a0, _ = synthetic.make_weights("shaped", seed=0, logits_fn=logits)
    b0, _ = orc.make_weights("shaped", seed=0)
    assert all(torch.equal(a0[k], b0[k]) for k in a0)

"""TEST INFRASTRUCTURE ONLY -- harness that imports the UNMODIFIED reference (thomasneff/AdaNeRF)
from /root/reference/src on CPU and drives TrainConfig.inference (src/train_data.py:278-299).

Only usable in the build container (the GPU box has no /root/reference). It is used to
  * validate oracle/adanerf_oracle.py (the travelling restatement), and
  * generate the golden fixtures under tests/golden/ (oracle/gen_golden.py).

Recipe: stub `configargparse` + `imageio`, build FeatureSets/models from a plain
namespace, skip TrainConfig.initialize (needs CUDA + dataset on disk, src/train_data.py:75).
"""
import math
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch

REF_ROOT = os.environ.get("ADANERF_REFERENCE", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")


def available():
    return os.path.isdir(REF_SRC)


def _install_stubs():
    for name in ("configargparse", "imageio"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)


def make_config(K=8, thr=0.2, pos_enc_args=("10-4", "10-4"), ndc=False):
    """Namespace with exactly the fields the two FeatureSets/models read (configs/fine_training.ini;
    ndc=True: configs/fine_training_ndc.ini -- posEncArgs [2-2, 10-4], FromClassifiedDepthAdaptiveNoDepthRange,
    rayMarchNormalization [.., None], depthTransform linear, useNDC)."""
    if ndc:
        pos_enc_args = ("2-2", "10-4")
    cfg = Namespace(
        inFeatures=["SpherePosDir", "RayMarchFromPoses"],
        outFeatures=["RawSigmoid", "RGBARayMarch"],
        posEnc=["nerf", "nerf"], posEncArgs=list(pos_enc_args),
        raySampleInput=[0, 0], multiDepthFeatures=[128, 128],
        multiDepthIgnoreValue=[1.01, 1.01], multiDepthWindowSize=[5, 5],
        activation=["relu", "nerf"], layers=[8, 8], layerWidth=[256, 256], skips=["", "auto"],
        numRaymarchSamples=[K, K],
        rayMarchSampler=["none", "FromClassifiedDepthAdaptive"],
        rayMarchSamplingStep=[1.0 / 128, 1.0 / 128],
        rayMarchNormalization=["InverseSqrtDistCentered", "InverseSqrtDistCentered"],
        rayMarchSamplingNoise=[0.0, 0.0], zNear=[0.001, 0.001], zFar=[1.0, 1.0],
        adaptiveSamplingThreshold=thr, accumulationMult="alpha",
        losses=["NeRFWeightMultiplicationLoss", "MSE"], trainWithGTDepth=False,
        deterministicSampling=True, useNDC=False, perturb=False,
        rayMarchNormalizationCenter=[], device="cpu", storeFullData=False, depthTransform="log",
        lossComponents=["One", "Zero", "NerfA"], lossComponentBlending=[-1.0, -1.0, -1.0],
        lossBlendingStart=0, lossBlendingDuration=1, lossWeights=[0.025, 1.0],
        scale=1,
    )
    if ndc:
        cfg.rayMarchSampler = ["none", "FromClassifiedDepthAdaptiveNoDepthRange"]
        cfg.rayMarchNormalization = ["InverseSqrtDistCentered", "None"]
        cfg.useNDC = True
        cfg.depthTransform = "linear"
    return cfg


def make_dataset_info(scene, w, h, ndc=False):
    """Fake DatasetInfo with the attributes FeatureSet.initialize reads (src/datasets.py:146-213)."""
    _install_stubs()
    from util.depth_transformations import LogTransform, LinearTransform
    view = Namespace(view_cell_center=list(scene["view_cell_center"]),
                     view_cell_size=list(scene["view_cell_size"]),
                     fov=scene["fov"], focal=0.5 * w / math.tan(0.5 * scene["fov"]), camera_scale=1.0)
    return Namespace(view=view, w=w, h=h, depth_max=scene["max_depth"],
                     depth_range=list(scene["depth_range"]),
                     depth_range_warped=list(scene["depth_range"]),
                     depth_transform=LinearTransform if ndc else LogTransform, use_warped_depth_range=[True, True])


class RefRenderer:
    """Builds f_in/f_out/models exactly as TrainConfig.initialize would and exposes inference()."""

    def __init__(self, scene, K=8, thr=0.2, w=800, h=800, seed=0, ndc=False):
        _install_stubs()
        torch.manual_seed(seed)
        from features import FeatureSet
        from models import ModelSelection
        from train_data import TrainConfig
        self.cfg = make_config(K=K, thr=thr, ndc=ndc)
        self.dataset_info = make_dataset_info(scene, w, h, ndc=ndc)
        f_in, f_out = FeatureSet.get_sets(self.cfg, "cpu")
        for f in list(f_in) + list(f_out):
            f.initialize(self.cfg, self.dataset_info, "cpu")
        models = [ModelSelection.getModel(self.cfg, f_in[i].n_feat, 128 if i == 0 else 4, "cpu", i)
                  for i in range(2)]
        tc = TrainConfig()
        tc.f_in, tc.f_out, tc.models, tc.config_file = f_in, f_out, models, self.cfg
        tc.device = "cpu"
        self.tc = tc

    @property
    def models(self):
        return self.tc.models

    def load_state_dicts(self, sd0, sd1):
        self.tc.models[0].load_state_dict(sd0, strict=True)
        self.tc.models[1].load_state_dict(sd1, strict=True)

    def inference(self, pose, rot, dirs):
        """pose [3], rot [3,3], dirs [n,3] float32 tensors -> (outs, dicts) of the reference."""
        from datasets import SampleDataWrapper, DatasetKeyConstants as D
        d = {D.image_pose: pose.reshape(1, 3), D.image_rotation: rot.reshape(1, 3, 3),
             D.ray_directions_samples: dirs.reshape(1, -1, 3)}
        batch = SampleDataWrapper([dict(d), dict(d)], [], False)
        with torch.no_grad():
            return self.tc.inference(batch, gradient=False, is_inference=True)

    def stages(self, pose, rot, dirs):
        """Per-stage tensors of one inference call (numpy, float32 / int64)."""
        from features import FeatureSetKeyConstants as F
        outs, dicts = self.inference(pose, rot, dirs)
        d0, d1 = dicts
        res = dict(
            x0=d0[F.input_feature_batch], raw0=d0[F.network_output],
            ray_o=d0[F.input_feature_ray_origins], ray_d=d0[F.input_feature_ray_directions],
            rgb=outs[1], weights=d1[F.nerf_weights_output], alpha=d1[F.nerf_alpha_output],
            depth_est=d1[F.nerf_estimated_depth],                 # [N,1] LogTransform.from_world(depth_map)
        )
        if F.adaptive_sample_positions in d1:
            res["asp"] = d1[F.adaptive_sample_positions]
            res["z_nan"] = d1[F.nerf_input_feature_z_vals]        # [N,K] NaN padded (world depth)
            res["x1_nan"] = d1[F.input_feature_batch]             # [N,K,90] NaN padded
            res["raw1_pad"] = d1[F.network_output]                # [N,K,4] zero padded
        else:
            res["z"] = d1[F.nerf_input_feature_z_vals]
            res["x1"] = d1[F.input_feature_batch]
            res["raw1"] = d1[F.network_output]
        return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in res.items()}


def generate_ray_directions(w, h, fov, focal):
    _install_stubs()
    from util.raygeneration import generate_ray_directions as g
    return g(w, h, fov, focal)

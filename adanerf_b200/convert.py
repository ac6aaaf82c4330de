"""`.weights` checkpoints -> export directory.

The reference saves a network as `torch.save(self.state_dict(), f"{path}{name}_{suffix}.weights")`
(src/models.py:87-90) and loads either a state_dict or a whole module (`load_weights`, src/models.py:105-112).
`src/export.py:28-93` turns a trained run into the directory the C++ viewer takes with `-mp`
({config.ini, dataset_info.txt, model0.onnx, model1.onnx}); this module does the same from two `.weights` files
without instantiating the reference's classes, so trained checkpoints run through `adn_create_from_export_dir`,
`Renderer.from_export_dir` and `adn_viewer_headless`.

    python -m adanerf_b200.convert --weights0 Net0_opt.weights --weights1 Net1_opt.weights \
        --dataset-info dataset_info.txt --threshold 0.2 --samples 8 --out export_dir
"""
import argparse
import ast
from collections import OrderedDict

import torch

from .onnx_weights import write_export_dir

SAMPLING_KEYS = ("layers.0.weight", "layers.0.bias")
SHADING_KEYS = ("pts_linears.0.weight", "views_linears.0.weight", "feature_linear.weight", "alpha_linear.weight",
                "rgb_linear.weight")


def load_weights_file(path, allow_pickle=False):
    """state_dict (name -> fp32 CPU tensor) of a `.weights` file.  The reference saves a plain state_dict
    (src/models.py:87-90), which torch loads without executing pickled code; a file that holds a pickled nn.Module
    (older checkpoints) is only read with `allow_pickle=True` -- unpickling runs arbitrary code from the file, so that is
    the caller's explicit decision (--allow-pickle on the command line), never a silent fallback."""
    try:
        obj = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:
        if not allow_pickle:
            raise ValueError(f"{path}: not a plain state_dict ({type(e).__name__}); pass allow_pickle=True / --allow-pickle "
                             "to unpickle it (only for files you trust)") from e
        obj = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(obj, (dict, OrderedDict)):
        obj = obj.state_dict()
    return OrderedDict((k, v.detach().to(torch.float32).contiguous()) for k, v in obj.items() if torch.is_tensor(v))


def check_state_dicts(sd0, sd1):
    """The architecture the hot path implements: BaseNet sampling net, NeRF shading net with a view branch."""
    for k in SAMPLING_KEYS:
        if k not in sd0:
            raise ValueError(f"sampling net: missing {k} (expected BaseNet layers.{{i}}.weight/bias, src/models.py:71-76)")
    for k in SHADING_KEYS:
        if k not in sd1:
            raise ValueError(f"shading net: missing {k} (expected NeRF with use_viewdirs, src/models.py:214-250)")
    if sd0["layers.0.weight"].shape[1] not in (90, 30) or sd1["pts_linears.0.weight"].shape[1] != 63:
        raise ValueError("posEncArgs other than [10-4, 10-4] / [2-2, 10-4] (90 or 30 / 63+27 input features) are not supported")


def read_dataset_info(path):
    """`key = value` lines of dataset_info.txt (src/train_data.py:180-195) -> scene dict."""
    vals = {}
    with open(path) as f:
        for line in f:
            if "=" in line:
                k, v = line.split("=", 1)
                vals[k.strip()] = ast.literal_eval(v.strip())
    need = ("view_cell_center", "view_cell_size", "depth_range", "fov", "max_depth")
    missing = [k for k in need if k not in vals]
    if missing:
        raise ValueError(f"{path}: missing {missing}")
    return {k: vals[k] for k in need}


def weights_to_export_dir(weights0, weights1, out_dir, scene, threshold, num_samples, allow_pickle=False):
    sd0, sd1 = load_weights_file(weights0, allow_pickle), load_weights_file(weights1, allow_pickle)
    check_state_dicts(sd0, sd1)
    write_export_dir(out_dir, scene, sd0, sd1, float(threshold), int(num_samples))
    return sd0, sd1


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--weights0", required=True, help="sampling net checkpoint (.weights)")
    ap.add_argument("--weights1", required=True, help="shading net checkpoint (.weights)")
    ap.add_argument("--dataset-info", required=True, help="dataset_info.txt of the run (scene constants)")
    ap.add_argument("--threshold", type=float, required=True, help="adaptiveSamplingThreshold")
    ap.add_argument("--samples", type=int, required=True, help="numRaymarchSamples of the shading net (K)")
    ap.add_argument("--out", required=True)
    ap.add_argument("--allow-pickle", action="store_true",
                    help="also read checkpoints that hold a pickled nn.Module (executes code from the file: trusted files only)")
    a = ap.parse_args(argv)
    weights_to_export_dir(a.weights0, a.weights1, a.out, read_dataset_info(a.dataset_info), a.threshold, a.samples, a.allow_pickle)
    print(f"wrote {a.out}")


if __name__ == "__main__":
    main()

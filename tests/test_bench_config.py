"""bench.py's workload table and the config object both arms print (CPU only, no GPU work).

BASELINE.json configs: [1] 800x800 thr 0.2 K = 8, [2] dense K = 128, [3] 1600x1600 strong scaling, [4] threshold sweep."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_workloads_cover_the_baseline_configs():
    b = _bench()
    w = b.WORKLOADS
    assert w["800x800_thr0.2_K8"] == dict(W=800, H=800, thr=0.2, K=8, weights="rand", scaling="weak")
    assert w["800x800_dense_K128"]["thr"] == 0.0 and w["800x800_dense_K128"]["K"] == 128
    assert w["1600x1600_thr0.2_K8"]["scaling"] == "strong" and w["1600x1600_thr0.2_K8"]["W"] == 1600
    for k in (8, 16):
        for t in (0.05, 0.1, 0.2, 0.3, 0.5):
            c = w[f"800x800_pav_thr{t}_K{k}"]
            assert c["weights"] == "pavillon" and c["thr"] == t and c["K"] == k
    assert os.path.exists(b.PAVILLON_NPZ)


def test_config_object_is_the_same_on_both_arms():
    """The driver compares the two arms' `config`: it depends on (workload, number of GPUs) only."""
    b = _bench()
    for name, cfg in b.WORKLOADS.items():
        for n in (1, 2, 8):
            a, c = b.workload_config(name, cfg, n), b.workload_config(name, dict(cfg), n)
            assert a == c and a["workload"] == name
            if cfg["scaling"] == "strong":
                assert a["frame"] == f"{cfg['W']}x{cfg['H']}" and a["rays_per_gpu_per_step"] == cfg["W"] * cfg["H"] // n
            else:
                assert a["frame"] == f"{cfg['W']}x{cfg['H'] * n}" and a["rays_per_gpu_per_step"] == cfg["W"] * cfg["H"]

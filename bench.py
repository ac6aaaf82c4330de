#!/usr/bin/env python
"""bench.py -- frames/s of the AdaNeRF hot path at 800x800 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the whole hot path (rays -> sampling MLP -> threshold/compaction -> posenc ->
shading MLP -> composite) over one 800x800 frame (640 000 rays) of synthetic input: random-init
sampling + shading nets (seed 0, the reference's own initialisers) and procedurally generated pinhole
rays from the view-cell centre (no datasets / checkpoints offline).

  value : frames/s with every input already resident in HBM (camera pose only; rays generated on the
          device), timed with CUDA events on the launching stream, max over ranks.
  e2e   : the same metric through the reference-facing host-buffer call adn_render_rays_host:
          ray directions start in HOST memory (H2D inside the timed region), RGB ends in host memory.
  N > 1 : weak scaling -- rank r renders rows [800 r, 800 (r+1)) of an 800 x 800N frame (fixed 640 000
          rays per GPU) followed by ONE NCCL all-gather of the RGB tiles; value = 800x800-frame
          equivalents per second over all ranks.
  --impl reference : the reference's CPU path (the oracle port of TrainConfig.inference, torch CPU,
          all host threads) on a bounded ray sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W = H = 800
WORKLOADS = {
    # BASELINE.json configs[1]: 800x800, thr 0.2, K = 8 (~8 samples/ray with random-init nets: every ray saturates at K)
    "800x800_thr0.2_K8": dict(thr=0.2, K=8, weights="rand"),
    # BASELINE.json configs[2]: 800x800, dense 128 samples/ray
    "800x800_dense_K128": dict(thr=0.0, K=128, weights="rand"),
    # ragged variant (shaped sampling net: 1..8 samples per ray)
    "800x800_thr0.2_K8_shaped": dict(thr=0.2, K=8, weights="shaped"),
}
FLOP_PER_SAMPLE_MLP1 = 1186816.0   # 2 * 593 408 MAC, unpadded
FLOP_PER_RAY_MLP0 = 898048.0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), tflops=d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0)),
                    source="MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full summary (None when absent)."""
    p = os.path.join(ROOT, "profiles", "ncu_r1_summary.json")
    try:
        table = json.load(open(p))
        key = kernel if kernel in table else next(k for k in table if k.startswith(kernel.rstrip(">")))   # trailing template args
        return int(table[key]["dram_rd"] + table[key]["dram_wr"])
    except Exception:
        return None


def stage_rooflines(stage_ms, n_rays, rays_first_chunk, m_first_chunk, thr, peaks):
    """Algorithmic work (SURVEY.md 8d / DESIGN.md 4) / measured stage time for the first chunk of the frame."""
    r, m = float(rays_first_chunk), float(m_first_chunk)
    hbm = peaks["hbm_gbs"]
    out = {}
    def add(name, ms, work, unit, peak, bound):
        if ms > 0:
            ach = work / (ms * 1e-3) / (1e9 if unit == "GB/s" else 1e12)
            out[name] = dict(bound=bound, achieved=ach, peak=peak, unit=unit, frac=ach / peak)
    add("stage0_features", stage_ms[0], r * (12 + 512 + 24), "GB/s", hbm, "hbm")          # dirs in, packed hi/lo tile + ray o/d out
    add("mlp0", stage_ms[1], r * FLOP_PER_RAY_MLP0, "TFLOP/s", peaks["tflops"], "tensor")   # algorithmic flops (x3 MMAs issued for the split)
    if thr > 0:
        add("stage2_sample", stage_ms[2], r * (512 + 8) + m * 16, "GB/s", hbm, "hbm")
    add("stage3_posenc", stage_ms[3], m * (8 + 256) + r * 24, "GB/s", hbm, "hbm")           # (ray, z) in, packed bf16 P+V blocks out
    add("mlp1", stage_ms[4], m * FLOP_PER_SAMPLE_MLP1, "TFLOP/s", peaks["tflops"], "tensor")
    add("stage5_composite", stage_ms[5], m * 20 + r * 20, "GB/s", hbm, "hbm")
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    samples=len(sm), reasons=sorted(reasons))


def cpu_reference_rate(cfg, budget_s, threads=None):
    """Oracle port of the reference CPU path on a bounded sample: chunks of 8192 rays (inferenceChunkSize,
    configs/*.ini:31) of the same frame, 1 warm-up chunk, then as many chunks as fit in ~budget_s."""
    import torch
    from oracle import adanerf_oracle as orc
    scene = orc.SCENE_BARBERSHOP
    sd0, sd1 = orc.make_weights(cfg["weights"], seed=0)
    dirs = torch.from_numpy(orc.generate_ray_directions(W, H, scene["fov"]).reshape(-1, 3)).float()
    pose = torch.tensor(scene["view_cell_center"], dtype=torch.float32)
    rot = torch.eye(3)
    chunk = 8192 if cfg["K"] <= 16 else 1024
    g = torch.Generator().manual_seed(0)
    starts = torch.randint(0, W * H - chunk, (4096,), generator=g).tolist()
    if threads is None:
        # oneMKL on many-core hosts is fastest well below the core count for these 256-wide GEMMs:
        # probe a few thread counts on a quarter chunk and keep the best (this is "all the host threads it can use")
        ncpu = os.cpu_count() or 8
        cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
        best, threads = None, cands[0]
        q = max(256, chunk // 4)
        for c in cands:
            torch.set_num_threads(c)
            orc.render_rays(pose, rot, dirs[:q], sd0, sd1, scene, cfg["thr"], cfg["K"])
            t = time.perf_counter()
            orc.render_rays(pose, rot, dirs[q:2 * q], sd0, sd1, scene, cfg["thr"], cfg["K"])
            t = time.perf_counter() - t
            if best is None or t < best:
                best, threads = t, c
    torch.set_num_threads(threads)
    orc.render_rays(pose, rot, dirs[starts[0]:starts[0] + chunk], sd0, sd1, scene, cfg["thr"], cfg["K"])  # warm-up
    rays, t0, i = 0, time.perf_counter(), 1
    while True:
        orc.render_rays(pose, rot, dirs[starts[i]:starts[i] + chunk], sd0, sd1, scene, cfg["thr"], cfg["K"])
        rays += chunk
        i += 1
        el = time.perf_counter() - t0
        if el >= budget_s or i >= len(starts):
            break
    return dict(rays_per_s=rays / el, frames_per_s=rays / el / (W * H), cores=torch.get_num_threads(),
                sample=f"{rays} rays ({rays // chunk} chunks of {chunk}, random windows of the {W}x{H} frame) in {el:.1f} s, "
                       f"torch {torch.__version__} CPU fp32")


def run_reference(args, cfg, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_reference_rate(cfg, per_step / 4)
    vals, last = [], None
    for _ in range(args.steps):
        last = cpu_reference_rate(cfg, per_step)
        vals.append(last["frames_per_s"])
    v = sum(vals) / len(vals)
    line = dict(impl="reference", metric="frames_per_sec_800x800", value=v, unit="frames/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1000.0 / v, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(workload=name, rays_per_frame=W * H, thr=cfg["thr"], K=cfg["K"], weights=cfg["weights"],
                                              note="CPU path measured on a bounded ray sample and scaled to a full frame"),
                rays_per_sec=v * W * H,
                cpu_baseline=dict(value=v, unit="frames/s", cores=last["cores"], kind="port", sample=last["sample"]),
                e2e=dict(value=v, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def run_ours(args, cfg, name):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    ge.build()
    from adanerf_b200 import Renderer
    from adanerf_b200 import synthetic   # product-side synthetic scene / weights; nothing under oracle/ on this arm
    scene = synthetic.SCENE_BARBERSHOP
    pose = torch.tensor(scene["view_cell_center"], dtype=torch.float32)
    rot = torch.eye(3)
    r = Renderer(scene, device=local)

    def probe_logits(sd0):   # W-shaped recipe: raw sampling-net outputs on every 157th ray of the 800x800 grid
        r.set_weights(0, sd0)
        x0, _, _ = r.stage0(pose, rot, r.generate_ray_directions(W, H)[::157].contiguous())
        return r.mlp0(x0)

    sd0, sd1 = synthetic.make_weights(cfg["weights"], seed=0, logits_fn=probe_logits)
    r.set_weights(0, sd0)
    r.set_weights(1, sd1)
    Hn = H * world                 # weak scaling: an 800 x 800N frame, one 800-row band per rank
    row0 = H * rank
    thr, K = cfg["thr"], cfg["K"]
    n_rays = W * H
    band = torch.empty((n_rays, 3), dtype=torch.float32, device="cuda")
    gathered = torch.empty((world * n_rays, 3), dtype=torch.float32, device="cuda") if world > 1 else None

    def step():
        r.render_camera(pose, rot, W, Hn, thr, K, row0=row0, rows=H, out=band)
        if world > 1:
            dist.all_gather_into_tensor(gathered, band)   # the one NCCL gather of RGB tiles per frame

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    l0 = r.stats()["kernel_launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    st = r.stats()
    launches = st["kernel_launches"] - l0
    m_samples = st["n_samples"]
    ms_per_step = ms / args.steps
    value = world * 1000.0 / ms_per_step          # 800x800-frame equivalents per second, all ranks

    # ---- end to end through the host-buffer entry point (H2D dirs + D2H rgb inside the timed region)
    dirs_host = np.ascontiguousarray(r.generate_ray_directions(W, Hn, row0=row0, rows=H).cpu().numpy())   # this rank's band
    rgb_host = np.empty((n_rays, 3), dtype=np.float32)   # caller-owned result buffer, reused every frame
    for _ in range(3):
        r.render_rays_host(pose, rot, dirs_host, thr, K, want_nsamples=False, out=rgb_host)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host = r.render_rays_host(pose, rot, dirs_host, thr, K, want_nsamples=False, out=rgb_host)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * args.steps / e2e_s
    finite = bool(np.isfinite(host["rgb"]).all())

    # ---- per-stage device times (CUDA events around each stage inside the context) for the roofline
    r.set_option("profile", 1)
    stage_ms = np.zeros(6)
    n_prof = 5
    for _ in range(n_prof):
        r.render_camera(pose, rot, W, Hn, thr, K, row0=row0, rows=H, out=band)
        stage_ms += np.array(r.stats()["ms_stage"])
    stage_ms /= n_prof
    r.set_option("profile", 0)

    if rank == 0:
        peaks = measured_peaks()
        chunk_rays = min(n_rays, max(8192, (8 << 20) // K))
        chunk_rays = ((chunk_rays + 127) // 128) * 128
        chunk_rays = (chunk_rays // W + (1 if chunk_rays % W else 0)) * W if chunk_rays % W else chunk_rays
        chunk_rays = min(chunk_rays, n_rays)
        prof_samples = m_samples if chunk_rays >= n_rays else m_samples  # stats hold the last chunk's M
        mlp1_flop = FLOP_PER_SAMPLE_MLP1 * (chunk_rays * K if thr == 0.0 else prof_samples)
        achieved = mlp1_flop / (stage_ms[4] * 1e-3) / 1e12 if stage_ms[4] > 0 else 0.0
        cpu = cpu_reference_rate(cfg, args.cpu_seconds) if args.cpu_seconds > 0 else dict(
            frames_per_s=None, cores=0, sample="skipped (--cpu-seconds 0)")
        line = dict(
            metric="frames_per_sec_800x800", value=value, unit="frames/s", n_gpus=world, steps=args.steps,
            warmup=max(args.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="bf16", data="synthetic",
            config=dict(workload=name, rays_per_gpu_per_step=n_rays, frame=f"{W}x{Hn}", thr=thr, K=K, weights=cfg["weights"],
                        samples_last_chunk=int(m_samples), parallelism=f"row-bands x{world} + 1 NCCL all-gather of RGB tiles",
                        l2="per-frame working set (packed features + activations I/O, >1 GB) exceeds the 126 MB L2; no explicit flush",
                        mlp0="bf16x3 split precision (fp32-class)", mlp1="bf16 operands, fp32 accumulate"),
            rays_per_sec=value * n_rays,
            e2e=dict(value=e2e_value, unit="frames/s", h2d_bytes_per_step=int(dirs_host.nbytes + 48),
                     d2h_bytes_per_step=int(n_rays * 12), api="adn_render_rays_host", finite=finite),
            gpu_launches=int(launches),
            clocks=clocks,
            stage_ms=dict(zip(["stage0_features", "mlp0", "stage2_sample", "stage3_posenc", "mlp1", "stage5_composite"],
                              [round(float(x), 4) for x in stage_ms])),
            roofline=dict(kernel="mlp_umma_kernel<1,2,2> (shading MLP, first chunk of the frame)", bound="tensor", achieved=achieved,
                          peak=peaks["tflops"], unit="TFLOP/s", frac=achieved / peaks["tflops"], traffic=ncu_traffic("mlp_umma_kernel<1, 2, 2>"),
                          traffic_unit="bytes of DRAM read+write per launch (profiles/ncu_r1_summary.json, ncu --set full)",
                          peak_source=peaks["source"]),
            roofline_stages=stage_rooflines(stage_ms, n_rays, chunk_rays if thr == 0.0 else n_rays,
                                            (chunk_rays * K) if thr == 0.0 else int(m_samples), thr, peaks),
            cpu_baseline=dict(value=cpu["frames_per_s"], unit="frames/s", cores=cpu["cores"], kind="port", sample=cpu["sample"]),
        )
        print(json.dumps(line))
    r.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="800x800_thr0.2_K8", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg, args.workload)
    else:
        run_ours(args, cfg, args.workload)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py - throughput of the HOSNeRF per-ray hot path on B200 (one JSON line on stdout).

Workload (BASELINE.json configs[1], "C2"): stage-1 background branch, 4096 rays x
(128 proposal + 128 NeRF) samples per step, PropMLP 4x256 + NeRFMLP 8x256, forward
(``LitMipNeRF360.render_rays``), fp16 tensor-core MLP with fp32 accumulation, synthetic rays and
seeded weights (hosnerf_b200.synth).  metric = ray-samples/s (rays/s reported beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this framework
    python bench.py --impl reference ...                          # the reference algorithm on the
                                                                  # host CPU (oracle port, torch fp32)
N > 1: launched by torchrun, one rank per GPU, every rank renders its own 4096-ray batch (rays
shard with no data-path collective: weak scaling); time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS = 4096
S_PROP, S_NERF = 128, 128
SAMPLES_PER_RAY = S_PROP + S_NERF
MODEL_KW = dict(num_levels=2, num_prop_samples=S_PROP, num_nerf_samples=S_NERF, nerf_netwidth=256,
                opaque_background=True)
NEAR, FAR = 0.1, 1e6
# algorithmic MLP work per sample: 2 * MACs of the reference layers
FLOP_PROP = 2 * 342272
FLOP_NERF = 2 * 851968
MLP_KERNEL = {0: "mlp_pair_kernel", 1: "mlp_tc_kernel", 2: "mlp_pair_kernel"}
WORKLOAD = "C2: stage-1 bkg render_rays, 4096 rays x (128 prop + 128 nerf) samples, PropMLP 4x256 + NeRFMLP 8x256"


def ncu_traffic():
    """DRAM bytes of the dominant kernel (both MLP launches of a step) from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "r1_mlp_pair_ncu.json")
    try:
        return json.load(open(p))["dram_bytes_per_step"]
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs: NVML (a few ms per sample) when
    nvidia_ml_py is importable, else nvidia-smi (one sample per ~100 ms)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and all(x.strip().isdigit() for x in visible.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = get(self.handle)
        flags = [bool(r & n.nvmlClocksEventReasonHwSlowdown), bool(r & n.nvmlClocksEventReasonHwThermalSlowdown),
                 bool(r & n.nvmlClocksEventReasonSwThermalSlowdown), bool(r & n.nvmlClocksEventReasonSwPowerCap)]
        return [str(mhz), str(self.max_mhz)] + ["Active" if f else "Not Active" for f in flags]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    time.sleep(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                if self.nvml is not None:
                    self.nvml = None       # fall back to nvidia-smi
                    continue
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "n_samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def cpu_reference_steps(n_rays, steps, warmup, threads):
    """Time the oracle (reference algorithm, torch CPU fp32) on `n_rays` rays of the workload."""
    from hosnerf_b200 import MipNeRF360, synth
    from oracle import mip360_ref as R
    torch.set_num_threads(threads)
    net = MipNeRF360("/nonexistent", **MODEL_KW)
    synth.fill_params_(net, 0)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    batch = synth.make_bkg_batch(n_rays, seed=1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            R.mip360_forward(sd, batch, 1.0, False, NEAR, FAR, num_levels=2, num_prop_samples=S_PROP,
                             num_nerf_samples=S_NERF)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 512
    sec = cpu_reference_steps(n, args.steps, min(args.warmup, 1), threads)
    value = n * SAMPLES_PER_RAY / sec
    line = {"impl": "reference", "metric": "ray_samples_per_s", "value": value, "unit": "ray-samples/s",
            "rays_per_s": n / sec, "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step": n, "samples_per_ray": SAMPLES_PER_RAY,
                       "note": "bounded sample of the same workload (512 of 4096 rays per step)"},
            "cpu_baseline": {"value": value, "unit": "ray-samples/s", "cores": threads, "kind": "port",
                             "sample": f"{n} rays x {SAMPLES_PER_RAY} samples per step, {args.steps} steps, torch CPU fp32"},
            "e2e": {"value": value, "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mlp-variant", type=int, default=0, help="A/B: 0 auto, 1 single-CTA MLP kernel, 2 cluster-pair kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from hosnerf_b200 import LitMipNeRF360, _lib, ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps

    ops.set_mlp_variant(args.mlp_variant)
    lit = LitMipNeRF360("/nonexistent", precision=args.precision, **MODEL_KW)
    synth.fill_params_(lit.model, 0)
    lit = lit.to(dev)
    host = synth.make_bkg_batch(N_RAYS, seed=1 + rank)       # every rank its own rays
    host = {k: v.contiguous().pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(batch):
        return lit.render_rays(batch, 0)["rgb"]

    # ---------------- device-resident throughput ----------------
    for _ in range(W):
        step(resident)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    # The MLP launches are timed with CUDA events on the launch stream inside the timed region.  When the whole level
    # loop runs behind one library call (hos_render_bkg) the library records them; otherwise ops.PROFILE does.
    one_call = lit.model.fused_render_supported()
    n_lvl = lit.model.num_levels
    mlp_ev = None
    if one_call:
        mlp_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2 * n_lvl)] for _ in range(K)]
        for row in mlp_ev:
            for e in row:
                e.record()                 # instantiates the cudaEvent_t handed to the library
    else:
        ops.PROFILE = []
    _lib.LAUNCHES = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    torch.cuda.synchronize()
    t_wall = time.perf_counter()
    for i in range(K):
        flush.zero_()                     # evict L2 between timed iterations (not timed)
        ev[i][0].record()
        if one_call:
            with torch.no_grad():
                lit.model.render_fused(resident, lit._frac(), False, lit.near, lit.far, mlp_events=mlp_ev[i])
        else:
            step(resident)
        ev[i][1].record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall
    launches = _lib.LAUNCHES
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    if one_call:
        mlp_ms = sum(row[2 * l].elapsed_time(row[2 * l + 1]) for row in mlp_ev for l in range(n_lvl))
        mlp_flops = K * N_RAYS * (S_PROP * FLOP_PROP * (n_lvl - 1) + S_NERF * FLOP_NERF)
    else:
        prof, ops.PROFILE = ops.PROFILE, None
        mlp_ms = sum(a.elapsed_time(b) for (_, _, a, b) in prof)
        mlp_flops = sum(rows * (FLOP_PROP if n_layers == 4 else FLOP_NERF) for (n_layers, rows, _, _) in prof)

    # ---------------- end to end through the public API, host buffers ----------------
    def e2e_step():
        batch = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        return step(batch).cpu()
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        out = e2e_step()
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    # the same K host batches through the chunk-streaming call: every step still copies its inputs in from pinned
    # host memory and its rgb out to host memory inside the timed region, but the host does not wait for chunk i
    # before enqueueing chunk i+1
    for _ in lit.render_rays_stream(host for _ in range(3)):
        pass
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for out in lit.render_rays_stream(host for _ in range(K)):
        checksum += float(out[0, 0])          # touch every step's result on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out.numel() * out.element_size()

    tt = torch.tensor([dev_ms, e2e_s * 1e3, e2e_sync_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, e2e_sync_ms_max = float(tt[0]), float(tt[1]), float(tt[2])

    if rank == 0:
        hbm, tf_burst, tf_sus, src = peaks()
        samples = N_RAYS * SAMPLES_PER_RAY * K * world
        value = samples / (dev_ms_max * 1e-3)
        achieved = mlp_flops / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
        line = {
            "metric": "ray_samples_per_s", "value": value, "unit": "ray-samples/s",
            "rays_per_s": N_RAYS * K * world / (dev_ms_max * 1e-3),
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16" if args.precision == "fp16" else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": N_RAYS, "samples_per_ray": SAMPLES_PER_RAY,
                       "parallelism": f"rays sharded over {world} GPU(s), no data-path collective",
                       "l2": "device-resident loop: a 256 MiB buffer is written between timed steps (untimed) to evict L2; e2e loop: no flush, every step's inputs arrive from pinned host memory",
                       "timing": "CUDA events per step on the launch stream, summed; max over ranks",
                       "call": "hos_render_bkg: one library call per batch" if one_call else "level loop in Python"},
            "e2e": {"value": samples / (e2e_ms_max * 1e-3), "unit": "ray-samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_max / K,
                    "api": "LitMipNeRF360.render_rays_stream(host batches): pinned host tensors in, rgb in pinned host memory out, "
                           "every step; chunk i+1 is enqueued before chunk i's rgb is awaited; each chunk is one CUDA-graph "
                           "replay of hos_render_bkg on static input buffers",
                    "blocking_ms_per_step": e2e_sync_ms_max / K,
                    "blocking_api": "LitMipNeRF360.render_rays(batch) with pinned host tensors in, rgb.cpu() out, one call at a time"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": MLP_KERNEL[args.mlp_variant] + " (tcgen05 fused MLP, 2 launches/step)",
                         "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s", "frac": achieved / tf_burst,
                         "peak_source": f"{src} bf16_tflops (burst)", "traffic": ncu_traffic(),
                         "traffic_note": "DRAM read+write bytes of the two MLP launches of one step, ncu --set full (profiles/r1_mlp_pair_ncu.json)",
                         "kernel_share_of_step": mlp_ms / dev_ms if dev_ms > 0 else None,
                         "flop_per_launch": [N_RAYS * S_PROP * FLOP_PROP, N_RAYS * S_NERF * FLOP_NERF]},
            "clocks": sampler.summary(),
            "wall_ms_per_step_incl_flush": t_wall * 1e3 / K,
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            n = 2048
            sec = cpu_reference_steps(n, 6, 1, threads)          # about 10 s of host work on the box's cores
            line["cpu_baseline"] = {"value": n * SAMPLES_PER_RAY / sec, "unit": "ray-samples/s", "cores": threads,
                                    "kind": "port", "sample": f"{n} of {N_RAYS} rays x {SAMPLES_PER_RAY} samples, 1 warm-up + 6 timed passes, torch CPU fp32 oracle"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

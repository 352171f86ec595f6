#!/usr/bin/env python
"""Headline benchmark: EOFluxVAE.encode_spatial_normalized on synthetic S2L2A 12-band 256x256 patches
(BASELINE.json configs[1]: batch 64 per GPU, bf16 operands, -> 32x32x32 latents), patches/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch (hypernetwork included - nothing is cached across steps).
* value   : device-resident inputs, CUDA events on the launching stream, max over ranks
* e2e     : the same call with HOST (pinned) inputs and a host copy of the latents, H2D/D2H inside the timed region
* roofline: the tcgen05 implicit-GEMM family (all conv / attention-GEMM launches of a step), algorithmic FLOPs
            (2*MACs of the reference formulation) / summed CUDA-event launch durations
* cpu_baseline / --impl reference: the CPU oracle port of the reference path on the box's host cores (bounded sample)
Multi-GPU: patches are independent, each rank encodes its own batch, no data-path collective ("weak" scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))

BATCH = 64
BANDS, SIZE = 12, 256
GF_PER_PATCH_ENCODE = 274.62  # SURVEY.md section 8d (algorithmic, 2*MACs), + ~1.5 GF hypernet per call
METRIC = "S2L2A 256x256 patches/sec (EOFluxVAE encode_spatial_normalized)"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tensor=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), hbm=float(p["hbm_gbs"]),
                    source="MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)")
    return dict(tensor=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md: ~1.4 PF sustained, 6.65 TB/s)")


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples taken DURING the timed region."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []  # upper half = samples under load
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


class CpuReference:
    """Reference path on the host: the CPU oracle port (fp32, all host threads) on bounded samples of the workload."""

    def __init__(self):
        import torch
        from oracle import eovae_oracle as O
        from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
        self.torch, self.O, self.synth = torch, O, synthetic_patches
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.sd = make_state_dict(FULL_CONFIG, 0)
        self.wvs = torch.tensor(WAVELENGTHS["S2L2A"])
        with torch.no_grad():
            x1 = synthetic_patches(1, BANDS, SIZE, seed=11)
            O.encode_spatial_normalized(self.sd, x1, self.wvs)  # warm-up (thread pools, allocator)
            t0 = time.perf_counter()
            O.encode_spatial_normalized(self.sd, x1, self.wvs)
            self.one = time.perf_counter() - t0

    def sample(self, target_seconds: float):
        """Encode one batch sized for ~target_seconds; returns (patches/s, n_patches)."""
        n = int(max(2, min(16, target_seconds / max(self.one, 1e-3))))
        x = self.synth(n, BANDS, SIZE, seed=12)
        with self.torch.no_grad():
            t0 = time.perf_counter()
            self.O.encode_spatial_normalized(self.sd, x, self.wvs)
            dt = time.perf_counter() - t0
        return n / dt, n

    def describe(self, n):
        return (f"{n} patches of 12x256x256 in one batch per sample, fp32, torch {self.torch.__version__} CPU, "
                f"{self.cores} threads, 1 warm-up")


def cpu_reference_throughput(target_seconds: float = 15.0):
    ref = CpuReference()
    v, n = ref.sample(target_seconds)
    return v, ref.cores, ref.describe(n)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference()
    vals, n = [], 0
    for i in range(args.warmup + args.steps):  # one step = one bounded sample (~2 s) of the workload
        v, n = ref.sample(2.0)
        if i >= args.warmup:
            vals.append(v)
    v = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "patches/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * BATCH / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "EOFluxVAE.encode_spatial_normalized, S2L2A 12x256x256 (bounded CPU sample per step; "
                               "ms_per_step is scaled to the 64-patch batch of the GPU arm)"},
        "cpu_baseline": {"value": v, "unit": "patches/s", "cores": ref.cores, "kind": "port", "sample": ref.describe(n)},
        "e2e": {"value": v, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g
    from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL writes its banner ("NCCL version ...") and warnings to its debug file,
        # which defaults to stdout - send it to a per-process file instead
        os.environ["NCCL_DEBUG"] = os.environ.get("EOVAE_NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/eovae_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)
    if not os.path.exists(g.LIB):
        raise RuntimeError("libeovae_sm100.so missing - run __graft_entry__.build() first")
    import eo_vae
    from eo_vae import ops
    eo_vae.set_compute_dtype(torch.bfloat16)

    model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), dev)
    wvs = torch.tensor(WAVELENGTHS["S2L2A"], dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x_dev = torch.randn((BATCH, BANDS, SIZE, SIZE), generator=gen, device=dev).clamp_(-2.0, 6.0)  # 201 MB > 126 MB L2
    x_host = x_dev.cpu().pin_memory()
    z_host = torch.empty((BATCH, 32, SIZE // 8, SIZE // 8), dtype=torch.float32).pin_memory()

    def step_device():
        return model.encode_spatial_normalized(x_dev, wvs)

    from eo_vae.pipeline import encode_stream

    def run_e2e(steps):
        # public host-fed API: every step copies its 201 MB batch from pinned host memory and its latents back; the
        # copies of neighbouring steps overlap the kernels (two device buffers), all inside the timed region
        encode_stream(model, [x_host] * steps, wvs, [z_host] * steps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.no_grad():
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()  # started before the warm-up so that samples exist inside a short timed region
        for _ in range(max(args.warmup, 3)):
            step_device()
        launches0 = ops.launch_count()
        ms_total = timed(step_device, args.steps)
        launches = ops.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        run_e2e(2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(args.steps)          # returns with every stream drained (latents of the last step are on the host)
        e1.record()
        barrier()
        ms_e2e_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms_e2e_t, op=dist.ReduceOp.MAX)
        ms_e2e = float(ms_e2e_t.item())

        # --- kernel-family timing for the roofline (same process, after the headline loop; events per launch)
        roof = None
        if rank == 0:
            ops.PROFILE = []
            for _ in range(min(args.steps, 5)):
                step_device()
            torch.cuda.synchronize()
            fam = {}
            for family, flops, s, e in ops.PROFILE:
                t, f, c = fam.get(family, (0.0, 0.0, 0))
                fam[family] = (t + s.elapsed_time(e), f + flops, c + 1)
            prof = ops.PROFILE
            ops.PROFILE = None
            if os.environ.get("EOVAE_BENCH_DUMP"):
                nsteps = min(args.steps, 5)
                per = len(prof) // nsteps
                rows = []
                for i in range(per):
                    ms = sum(prof[k * per + i][2].elapsed_time(prof[k * per + i][3]) for k in range(nsteps)) / nsteps
                    rows.append({"i": i, "family": prof[i][0], "gflop": prof[i][1] / 1e9, "ms": ms,
                                 "tflops": prof[i][1] / ms / 1e9})
                with open(os.environ["EOVAE_BENCH_DUMP"], "w") as f:
                    json.dump(rows, f, indent=1)
            peaks = load_peaks()
            igemm = {k: v for k, v in fam.items() if k in ("conv", "attn_gemm")}   # launches of igemm_kernel only
            other = {k: {"ms_per_step": v[0] / min(args.steps, 5), "tflops": v[1] / (v[0] * 1e-3) / 1e12, "launches_per_step":
                         v[2] / min(args.steps, 5)} for k, v in fam.items() if k not in igemm}
            t_ms = sum(v[0] for v in igemm.values())
            flops = sum(v[1] for v in igemm.values())
            n_launch = sum(v[2] for v in igemm.values())
            achieved = flops / (t_ms * 1e-3) / 1e12
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "igemm_traffic.json")
            if os.path.exists(tpath):
                with open(tpath) as f:
                    traffic = json.load(f).get("dram_bytes_per_launch")
            roof = {"bound": "tensor", "kernel": "igemm_kernel (tcgen05 implicit GEMM: every convolution of the step)",
                    "achieved": achieved, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": achieved / peaks["tensor"],
                    "traffic": traffic, "peak_source": peaks["source"],
                    "avg_launch_ms": t_ms / n_launch, "launches_per_step": n_launch / min(args.steps, 5),
                    "algorithmic_gflop_per_launch": flops / n_launch / 1e9,
                    "share_of_step": (t_ms / min(args.steps, 5)) / (ms_total / args.steps),
                    "other_tensor_kernels": other}

    train = None
    if not os.environ.get("EOVAE_BENCH_NO_TRAIN"):
        del x_dev
        torch.cuda.empty_cache()
        try:
            train = measure_train_step(g, dev, world, rank)
        except Exception as exc:  # noqa: BLE001 - the secondary figure must never take the headline line down
            train = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank == 0:
        patches = BATCH * world * args.steps
        value = patches / (ms_total * 1e-3)
        e2e_v = patches / (ms_e2e * 1e-3)
        cpu_v, cores, sample = (None, None, None)
        if world == 1:
            cpu_v, cores, sample = cpu_reference_throughput()
        line = {
            "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "EOFluxVAE.encode_spatial_normalized, S2L2A 12x256x256, batch 64 per GPU, "
                                   "random-init reference architecture (ch128, mult 1-2-4-4, z32), -> 32x32x32 latents",
                       "batch_per_gpu": BATCH, "parallelism": f"dp{world} (patches sharded by rank, no collective)",
                       "l2": "inputs (201 MB/step) and every level-0/1 activation exceed the 126 MB L2"},
            "tensor_tflops": value * (GF_PER_PATCH_ENCODE / 1000.0),
            "e2e": {"value": e2e_v, "unit": "patches/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": z_host.numel() * 4},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "train_step": train,
        }
        if cpu_v is not None:
            line["cpu_baseline"] = {"value": cpu_v, "unit": "patches/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


TRAIN_BATCH = 16
GF_PER_PATCH_TRAIN = 2695.0  # 3 x (encode 274.62 + decode 623.81) GFLOP per 12x256x256 patch


def measure_train_step(g, dev, world, rank, steps: int = 5):
    """Secondary figure (BASELINE configs[2]): EOFluxVAE.training_step, S2L2A batch 16 per GPU, Charbonnier + MS-SSIM loss,
    clip 1.0, Adam; gradients averaged over the ranks (bucketed NCCL all-reduce overlapped with backward).  Reported as
    the extra key ``train_step``; the headline metric above is unaffected."""
    import torch
    import torch.distributed as dist
    from eo_vae.graphs import GraphedTrainStep
    from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
    from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict

    def build():
        m = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), dev)
        m.train()
        m.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char", msssim_weight=1.0, msssim_start_step=0).to(dev)
        m.clip_grad = 1.0
        if world > 1:
            m.enable_ddp()
        return m

    wvs = torch.tensor(WAVELENGTHS["S2L2A"], dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    batch = {"image": torch.randn((TRAIN_BATCH, BANDS, SIZE, SIZE), generator=gen, device=dev).clamp_(-2.0, 6.0), "wvs": wvs}

    def timed(fn):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss = fn(3 + i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), float(loss.detach())

    model = build()
    ms_eager, loss = timed(lambda i: model.training_step(batch, i))
    del model
    model = build()
    graphed = GraphedTrainStep(model, batch)
    ms_graph, _ = timed(lambda i: graphed(batch))
    pps = lambda ms: world * TRAIN_BATCH / ms * 1e3  # noqa: E731
    return {"workload": "EOFluxVAE.training_step, S2L2A 12x256x256, batch 16 per GPU, Charbonnier + MS-SSIM, clip 1.0, Adam",
            "patches_per_s": pps(ms_eager), "ms_per_step": ms_eager, "tensor_tflops": pps(ms_eager) * GF_PER_PATCH_TRAIN / 1000.0,
            "graphed_patches_per_s": pps(ms_graph), "graphed_ms_per_step": ms_graph,
            "graphed_tensor_tflops": pps(ms_graph) * GF_PER_PATCH_TRAIN / 1000.0, "loss_after_8_steps": loss, "steps": steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmarks of the EOFluxVAE hot path on synthetic patches (BASELINE.json configs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--impl ours|reference|reference-cuda]

--config (default 2 = the configuration BASELINE.json's metric is quoted on):
  1  EOFluxVAE.reconstruct, S2RGB 1x3x256x256                          patches/s
  2  encode_spatial_normalized, S2L2A 12x256x256, batch 64 per GPU     patches/s   <- headline
  3  training_step (Charbonnier + MS-SSIM, clip 1.0, Adam), S2L2A batch 16 per GPU, gradients averaged over the ranks
  4  the same step with the modality drawn per step and per rank from {S2L2A, S1RTC, S2RGB}
  5  encode_spatial_normalized, S2L1C 13x512x512, batch 32 per GPU

One "step" = one pass of the path over one batch (train configs: forward incl. both hypernetworks, loss, backward, gradient
exchange, clip, Adam; inference configs: the hypernetwork output depends on the wavelength vector and the parameters only,
so after the first call it is served from the module's operand cache - EOVAE_BENCH_NO_OPERAND_CACHE=1 regenerates it
every step, +~0.5 ms).
* value        : device-resident inputs, CUDA events on the launching stream, max over ranks
* e2e          : the same call fed from pinned HOST memory, result read back to the host, copies inside the timed region
* roofline     : the tcgen05 implicit-GEMM family (every convolution / attention GEMM launch of a step), algorithmic
                 FLOPs (2*MACs of the reference formulation) / summed CUDA-event launch durations
* cpu_baseline : the reference's CPU path on the box's host cores, bounded sample (N = 1 only)
* gpu_eager    : the UNMODIFIED reference modules in torch eager on the same B200 (cuDNN / ATen; TF32 'medium' as
                 train.py:66, and bf16 autocast) - the incumbent (N = 1 only; needs baseline/_ref, see build())
--impl reference       : the reference's CPU implementation alone (unmodified modules from baseline/_ref when present,
                         else the oracle port), all host threads, bounded sample per step
--impl reference-cuda  : the gpu_eager measurement alone, as its own JSON line
Multi-GPU: patches are independent; encode configs shard by rank with no data-path collective, train configs add the
one gradient all-reduce ("weak" scaling in both cases).
"""
from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "eo-vae_b200"))

# algorithmic GFLOP per patch, SURVEY.md section 8d (2*MACs of the reference formulation, no credit for recompute)
GF = {("enc", 12, 256): 274.62, ("enc", 3, 256): 273.26, ("enc", 2, 256): 273.11, ("enc", 13, 512): 1124.9,
      ("dec", 12, 256): 623.81, ("dec", 3, 256): 619.85 + 1.81 * 3 / 12, ("dec", 2, 256): 619.85 + 1.81 * 2 / 12,
      ("dec", 13, 512): 2521.6}

CONFIGS = {
    1: dict(kind="reconstruct", modality="S2RGB", size=256, batch=1,
            metric="S2RGB 256x256 patches/sec (EOFluxVAE reconstruct)",
            workload="EOFluxVAE.reconstruct, S2RGB 1x3x256x256 (BASELINE configs[0])"),
    2: dict(kind="encode", modality="S2L2A", size=256, batch=64,
            metric="S2L2A 256x256 patches/sec (EOFluxVAE encode_spatial_normalized)",
            workload="EOFluxVAE.encode_spatial_normalized, S2L2A 12x256x256, batch 64 per GPU, random-init reference "
                     "architecture (ch128, mult 1-2-4-4, z32), -> 32x32x32 latents (BASELINE configs[1])"),
    3: dict(kind="train", modality="S2L2A", size=256, batch=16,
            metric="S2L2A 256x256 patches/sec (EOFluxVAE training_step)",
            workload="EOFluxVAE.training_step, S2L2A 12x256x256, batch 16 per GPU, Charbonnier + MS-SSIM, clip 1.0, Adam, "
                     "gradients averaged over the ranks (BASELINE configs[2])"),
    4: dict(kind="train_mixed", modality="mixed", size=256, batch=16,
            metric="mixed-modality 256x256 patches/sec (EOFluxVAE training_step)",
            workload="EOFluxVAE.training_step, modality drawn per step and per rank from S2L2A(12)/S1RTC(2)/S2RGB(3) with "
                     "random.Random(1234 + rank), batch 16 per GPU, Charbonnier + MS-SSIM, clip 1.0, Adam (BASELINE configs[3])"),
    5: dict(kind="encode", modality="S2L1C", size=512, batch=32,
            metric="S2L1C 512x512 patches/sec (EOFluxVAE encode_spatial_normalized)",
            workload="EOFluxVAE.encode_spatial_normalized, S2L1C 13x512x512, batch 32 per GPU -> 32x64x64 latents "
                     "(BASELINE configs[4])"),
}
MIXED = ["S2L2A", "S1RTC", "S2RGB"]


# The three Upsample convs run in sub-pixel form (four 2x2 convs on the low-resolution input): 16 instead of 36 MAC units, i.e.
# 173.9 -> 77.3 GFLOP per 256x256 patch actually executed (the achieved figure must use the MACs executed).
UPSAMPLE_SAVED_GF = {256: 173.9 * (1.0 - 16.0 / 36.0), 512: 4 * 173.9 * (1.0 - 16.0 / 36.0)}


def gf_per_patch(kind: str, bands: int, size: int) -> float:
    """GFLOP per patch that the step EXECUTES (= the reference formulation's 2*MACs, minus the upsample MACs the sub-pixel
    form does not perform)."""
    enc, dec = GF[("enc", bands, size)], GF[("dec", bands, size)] - UPSAMPLE_SAVED_GF[size]
    return {"encode": enc, "reconstruct": enc + dec, "train": 3.0 * (enc + dec), "train_mixed": 3.0 * (enc + dec)}[kind]


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


# ------------------------------------------------------------------------------------------------- reference arms
class Reference:
    """The reference's own implementation of the path: the UNMODIFIED modules (baseline/_ref or /root/reference, through
    oracle/ref_shim.py) when they are present, otherwise the oracle port.  ``device`` 'cpu' (fp32, all host threads) or a
    CUDA device (torch eager: cuDNN / ATen kernels)."""

    def __init__(self, cfg: dict, device="cpu"):
        import torch
        from oracle import eovae_oracle as O
        from oracle import ref_shim
        from oracle.weights import FULL_CONFIG, WAVELENGTHS, make_state_dict, synthetic_patches
        self.torch, self.O, self.synth, self.cfg, self.WV = torch, O, synthetic_patches, cfg, WAVELENGTHS
        self.device = torch.device(device)
        self.cores = os.cpu_count() or 1
        if self.device.type == "cpu":
            torch.set_num_threads(self.cores)
        self.sd = make_state_dict(FULL_CONFIG, 0)
        self.model = None
        if ref_shim.reference_root() is not None:
            self.model = ref_shim.build_reference_model(FULL_CONFIG, self.sd, train=False).to(self.device)
            self.kind, self.where = "reference", ref_shim.reference_root()
        else:
            if self.device.type != "cpu":
                raise RuntimeError("the GPU-eager incumbent needs the unmodified reference modules (baseline/_ref)")
            self.kind, self.where = "port", "oracle/eovae_oracle.py"
        self.heads = FULL_CONFIG["hyper_heads"]

    def _modality(self, rng=None):
        return self.cfg["modality"] if self.cfg["modality"] != "mixed" else (rng or random).choice(MIXED)

    def run(self, x, wvs):
        """One pass of the configured path over batch x (forward configs)."""
        t, kind = self.torch, self.cfg["kind"]
        with t.no_grad():
            if self.model is not None:
                return self.model.encode_spatial_normalized(x, wvs) if kind == "encode" else self.model.reconstruct(x, wvs)
            if kind == "encode":
                return self.O.encode_spatial_normalized(self.sd, x, wvs, self.heads)
            return self.O.reconstruct(self.sd, x, wvs, self.heads)

    def make_trainer(self):
        """Reference training step by hand (no Lightning here): forward(sampled) -> Charbonnier + (1 - MS-SSIM) -> backward ->
        clip 1.0 -> torch.optim.Adam, as new_autoencoder.py:587-657.  torchmetrics is not installed: the MS-SSIM term is
        the oracle's torch restatement of it (same sequence of depthwise convolutions / pools)."""
        t, O = self.torch, self.O
        if self.model is None:
            raise RuntimeError("training incumbents need the unmodified reference modules")
        self.model.train()
        params = [p for p in self.model.parameters() if p.requires_grad]
        opt = t.optim.Adam(params, lr=1e-4)

        def step(x, wvs, autocast=None):
            ctx = t.autocast(self.device.type, dtype=autocast) if autocast is not None else t.autocast(self.device.type, enabled=False)
            with ctx:
                recon, _ = self.model(x, wvs)
                loss = O.charbonnier_loss(recon.float(), x) + (1.0 - O.ms_ssim(recon.float(), x))
            opt.zero_grad()
            loss.backward()
            t.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            return loss
        return step

    def sample_cpu(self, target_seconds: float, one: float):
        """One bounded sample of the workload on the host; returns (patches/s, patches, seconds)."""
        bands = len(self.WV[self._modality()])
        n = int(max(1, min(self.cfg["batch"], target_seconds / max(one, 1e-3))))
        x = self.synth(n, bands, self.cfg["size"], seed=12)
        wvs = self.torch.tensor(self.WV[self._modality()])
        t0 = time.perf_counter()
        if self.cfg["kind"].startswith("train"):
            self._cpu_step(x, wvs)
        else:
            self.run(x, wvs)
        dt = time.perf_counter() - t0
        return n / dt, n, dt

    def calibrate_cpu(self) -> float:
        """Seconds per patch (one warm-up + one timed single-patch pass)."""
        mod = self._modality()
        x1 = self.synth(1, len(self.WV[mod]), self.cfg["size"], seed=11)
        wvs = self.torch.tensor(self.WV[mod])
        if self.cfg["kind"].startswith("train"):
            self._cpu_step = self.make_trainer() if self.model is not None else None
            if self._cpu_step is None:
                raise RuntimeError("CPU training baseline needs the unmodified reference modules")
            self._cpu_step(x1, wvs)
            t0 = time.perf_counter()
            self._cpu_step(x1, wvs)
        else:
            self.run(x1, wvs)
            t0 = time.perf_counter()
            self.run(x1, wvs)
        return time.perf_counter() - t0

    def describe(self, n):
        bands = "mixed" if self.cfg["modality"] == "mixed" else len(self.WV[self.cfg["modality"]])
        return (f"{n} patches of {bands}x{self.cfg['size']}x{self.cfg['size']} in one batch per sample, fp32, "
                f"{'unmodified reference modules (' + self.where + ')' if self.kind == 'reference' else 'oracle port'}, "
                f"torch {self.torch.__version__} CPU, {self.cores} threads, 1 warm-up")


def cpu_baseline(cfg, target_seconds: float = 15.0):
    ref = Reference(cfg, "cpu")
    one = ref.calibrate_cpu()
    v, n, _ = ref.sample_cpu(target_seconds, one)
    return {"value": v, "unit": "patches/s", "cores": ref.cores, "kind": ref.kind, "sample": ref.describe(n)}


def run_reference(args, cfg):
    """--impl reference: the reference's CPU implementation; each step is a bounded sample (~2 s) of the workload, timed for
    real (nothing extrapolated): value = patches of the samples / their wall time."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    ref = Reference(cfg, "cpu")
    one = ref.calibrate_cpu()
    patches, secs, n = 0, 0.0, 0
    for i in range(args.warmup + args.steps):
        _, n, dt = ref.sample_cpu(2.0, one)
        if i >= args.warmup:
            patches += n
            secs += dt
    v = patches / secs
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": v, "unit": "patches/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"] + f" - CPU arm: each step is a bounded sample of {n} patch(es) of that workload",
                   "patches_per_step": n},
        "cpu_baseline": {"value": v, "unit": "patches/s", "cores": ref.cores, "kind": ref.kind, "sample": ref.describe(n)},
        "e2e": {"value": v, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def gpu_eager(cfg, dev, warmup: int = 5, steps: int = 10):
    """The incumbent: unmodified reference modules, torch eager on this GPU (method of benchmark_compute.py:135-234: warm-up,
    CUDA events).  fp32 with TF32 ('medium', train.py:66) and bf16 autocast."""
    import torch
    from oracle import ref_shim
    if ref_shim.reference_root() is None:
        return {"unavailable": "reference modules not present (baseline/_ref is created by __graft_entry__.build() where "
                               "/root/reference exists)"}
    prev = torch.get_float32_matmul_precision()
    prev_cudnn = torch.backends.cudnn.allow_tf32
    out = {"source": ref_shim.reference_root(), "batch": cfg["batch"], "steps": steps, "warmup": warmup,
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    try:
        torch.set_float32_matmul_precision("medium")
        torch.backends.cudnn.allow_tf32 = True
        ref = Reference(cfg, dev)
        rng = random.Random(1234)
        gen = torch.Generator(device=dev).manual_seed(99)
        mods = MIXED if cfg["modality"] == "mixed" else [cfg["modality"]]
        data = {m: (torch.randn((cfg["batch"], len(ref.WV[m]), cfg["size"], cfg["size"]), generator=gen, device=dev).clamp_(-2, 6),
                    torch.tensor(ref.WV[m], device=dev)) for m in mods}
        train = cfg["kind"].startswith("train")
        step = ref.make_trainer() if train else None
        for name, ac in (("tf32", None), ("bf16_autocast", torch.bfloat16)):
            def once():
                x, wvs = data[rng.choice(mods)]
                if train:
                    return step(x, wvs, ac)
                if ac is None:
                    return ref.run(x, wvs)
                with torch.autocast("cuda", dtype=ac):
                    return ref.run(x, wvs)
            try:
                for _ in range(warmup):
                    once()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    once()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                out[name] = {"patches_per_s": cfg["batch"] / ms * 1e3, "ms_per_step": ms}
            except Exception as exc:  # noqa: BLE001
                out[name] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        out["peak_mem_gib"] = torch.cuda.max_memory_allocated(dev) / 2**30
        del ref, data, step
    except Exception as exc:  # noqa: BLE001 - a reported baseline must never take the headline line down
        out["error"] = f"{type(exc).__name__}: {exc}"[:300]
    finally:
        torch.set_float32_matmul_precision(prev)
        torch.backends.cudnn.allow_tf32 = prev_cudnn
        torch.cuda.empty_cache()
    return out


def run_reference_cuda(args, cfg):
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    ge = gpu_eager(cfg, dev, warmup=max(args.warmup, 3), steps=args.steps)
    best = max((v["patches_per_s"] for k, v in ge.items() if isinstance(v, dict) and "patches_per_s" in v), default=None)
    print(json.dumps({"impl": "reference-cuda", "metric": cfg["metric"], "value": best, "unit": "patches/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True, "dtype": "tf32 / bf16 autocast",
                      "data": "synthetic", "config": {"workload": cfg["workload"] + " - unmodified reference modules, torch eager"},
                      "gpu_eager": ge}), flush=True)


# ------------------------------------------------------------------------------------------------- our arm
def dtype_name(dt) -> str:
    import torch
    return {torch.bfloat16: "bf16", torch.float16: "fp16", torch.float32: "fp32"}[dt]


def run_ours(args, cfg):
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
        dist.init_process_group("nccl", device_id=dev)   # NCCL_DEBUG & co. are left exactly as the caller set them
    if not os.path.exists(g.LIB):
        raise RuntimeError("libeovae_sm100.so missing - run __graft_entry__.build() first")
    import eo_vae
    from eo_vae import ops
    if args.dtype:
        eo_vae.set_compute_dtype({"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype])

    kind, batch, size = cfg["kind"], cfg["batch"], cfg["size"]
    train = kind.startswith("train")
    mods = MIXED if cfg["modality"] == "mixed" else [cfg["modality"]]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    rng = random.Random(1234 + rank)
    data = {m: (torch.randn((batch, len(WAVELENGTHS[m]), size, size), generator=gen, device=dev).clamp_(-2.0, 6.0),
                torch.tensor(WAVELENGTHS[m], dtype=torch.float32, device=dev)) for m in mods}
    host = {m: data[m][0].cpu().pin_memory() for m in mods}
    bands_ref = len(WAVELENGTHS[mods[0]])
    gf = statistics.mean(gf_per_patch(kind, len(WAVELENGTHS[m]), size) for m in mods)

    model = g._model(FULL_CONFIG, make_state_dict(FULL_CONFIG, 0), dev)
    if os.environ.get("EOVAE_BENCH_NO_OPERAND_CACHE"):
        model.encoder.conv_in.CACHE_EVAL_OPERANDS = model.decoder.conv_out.CACHE_EVAL_OPERANDS = False
    sync_state = {}
    if train:
        from eo_vae.models.modules.consistency_loss import EOConsistencyLoss
        model.train()
        model.loss_fn = EOConsistencyLoss(pixel_weight=1.0, rec_loss_type="char", msssim_weight=1.0, msssim_start_step=0).to(dev)
        model.clip_grad = 1.0
        if world > 1:
            model.enable_ddp()
        seq = [rng.choice(mods) for _ in range(4096)]
        counter = [0]

        def pick():
            counter[0] += 1
            return seq[counter[0] % len(seq)]

        def step_device():
            m = pick()
            return model.training_step({model.image_key: data[m][0], "wvs": data[m][1]}, counter[0])

        dev_in = {m: torch.empty_like(data[m][0]) for m in mods}

        def run_e2e(steps):
            # host-fed step: the batch comes from pinned host memory, the loss goes back to the host, every step
            for _ in range(steps):
                m = pick()
                dev_in[m].copy_(host[m], non_blocking=True)
                loss = model.training_step({model.image_key: dev_in[m], "wvs": data[m][1]}, counter[0])
                sync_state["loss"] = float(loss.detach().cpu())
        h2d = statistics.mean(host[m].numel() * 4 for m in mods)
        d2h = 4
    else:
        x_dev, wvs = data[mods[0]]
        x_host = host[mods[0]]
        if kind == "encode":
            out_shape = (batch, FULL_CONFIG["z_channels"], size // 8, size // 8)
            fn = model.encode_spatial_normalized
        else:
            out_shape = tuple(x_dev.shape)
            fn = model.reconstruct
        z_host = torch.empty(out_shape, dtype=torch.float32).pin_memory()

        def step_device():
            with torch.no_grad():
                return fn(x_dev, wvs)

        from eo_vae.pipeline import encode_stream

        def run_e2e(steps):
            # public host-fed API: every step copies its batch from pinned host memory and its result back; the copies of
            # neighbouring steps overlap the kernels (two device buffers), all inside the timed region
            if kind == "encode":
                encode_stream(model, [x_host] * steps, wvs, [z_host] * steps)
            else:
                for _ in range(steps):
                    with torch.no_grad():
                        z_host.copy_(fn(x_host.to(dev, non_blocking=True), wvs), non_blocking=True)
                torch.cuda.synchronize()
        h2d, d2h = x_host.numel() * 4, z_host.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn_, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn_(steps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def loop(steps):
        for _ in range(steps):
            step_device()

    warm = max(args.warmup, 3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up so that samples exist inside a short timed region
    loop(warm)
    launches0 = ops.launch_count()
    ms_total = timed(loop, args.steps)
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    run_e2e(2)
    ms_e2e = timed(run_e2e, args.steps)

    # --- replicas must hold identical parameters after the exchanged steps; then the exposed cost of the exchange: the same
    #     step with the collective switched off (each rank alone; the replicas drift apart from here on, nothing later
    #     depends on them)
    comm = None
    if train and world > 1:
        chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        in_sync = bool(all(torch.equal(allc[0], c) for c in allc))
        gs = model._grad_sync
        model._grad_sync = None
        gs.remove()
        loop(2)
        ms_local = timed(loop, args.steps)
        comm = {"ms_per_step_without_exchange": ms_local / args.steps,
                "exposed_exchange_ms": (ms_total - ms_local) / args.steps,
                "scaling_eff_vs_no_exchange": ms_local / ms_total, "replicas_in_sync": in_sync,
                "buckets": len(gs.buckets), "bucket_mib": [b.flat.numel() * 4 >> 20 for b in gs.buckets]}

    # --- kernel-family timing for the roofline (same process, after the headline loop; events per launch)
    roof = None
    if rank == 0:
        nprof = min(args.steps, 5)
        ops.PROFILE = []
        loop(nprof)
        torch.cuda.synchronize()
        fam = {}
        for family, flops, s, e in ops.PROFILE:
            t, f, c = fam.get(family, (0.0, 0.0, 0))
            fam[family] = (t + s.elapsed_time(e), f + flops, c + 1)
        prof = ops.PROFILE
        ops.PROFILE = None
        if os.environ.get("EOVAE_BENCH_DUMP") and not train:
            per = len(prof) // nprof
            rows = []
            for i in range(per):
                ms = sum(prof[k * per + i][2].elapsed_time(prof[k * per + i][3]) for k in range(nprof)) / nprof
                rows.append({"i": i, "family": prof[i][0], "gflop": prof[i][1] / 1e9, "ms": ms, "tflops": prof[i][1] / ms / 1e9})
            with open(os.environ["EOVAE_BENCH_DUMP"], "w") as f:
                json.dump(rows, f, indent=1)
        peaks = load_peaks()
        igemm = {k: v for k, v in fam.items() if k in ("conv", "attn_gemm")}   # launches of igemm_kernel only
        other = {k: {"ms_per_step": v[0] / nprof, "tflops": v[1] / (v[0] * 1e-3) / 1e12, "launches_per_step": v[2] / nprof}
                 for k, v in fam.items() if k not in igemm}
        t_ms = sum(v[0] for v in igemm.values())
        flops = sum(v[1] for v in igemm.values())
        n_launch = sum(v[2] for v in igemm.values())
        achieved = flops / (t_ms * 1e-3) / 1e12
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "igemm_traffic.json")
        if os.path.exists(tpath) and args.config == 2:
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        roof = {"bound": "tensor", "kernel": "igemm_kernel (tcgen05 implicit GEMM: every convolution of the step)",
                "achieved": achieved, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": achieved / peaks["tensor"],
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks["source"],
                "avg_launch_ms": t_ms / n_launch, "launches_per_step": n_launch / nprof,
                "algorithmic_gflop_per_launch": flops / n_launch / 1e9,
                "share_of_step": (t_ms / nprof) / (ms_total / args.steps),
                "other_tensor_kernels": other}

    # secondary figures of the default line: the other BASELINE configs at this N (their own lines: --config 3 / 4 / 5)
    extra_train = extra_mixed = extra_512 = None
    if kind == "encode" and args.config == 2 and not os.environ.get("EOVAE_BENCH_NO_TRAIN"):
        data.clear()
        if not train:
            del x_dev
        torch.cuda.empty_cache()
        def guarded(fn, *a, **k):
            try:
                return fn(*a, **k)
            except Exception as exc:  # noqa: BLE001 - a secondary figure must never take the headline line down
                return {"error": f"{type(exc).__name__}: {exc}"[:300]}
        extra_512 = guarded(measure_encode, model, CONFIGS[5], dev, world, rank)
        extra_train = guarded(measure_train_step, g, dev, world, rank)
        extra_mixed = guarded(measure_train_step, g, dev, world, rank, mods=MIXED, graphed=False)

    eager = None
    if rank == 0 and world == 1 and not os.environ.get("EOVAE_BENCH_NO_EAGER"):
        del model
        data.clear()
        torch.cuda.empty_cache()
        eager = gpu_eager(cfg, dev)
        if extra_train is not None and "error" not in extra_train:
            extra_train["gpu_eager"] = gpu_eager(CONFIGS[3], dev, warmup=3, steps=5)

    if rank == 0:
        patches = batch * world * args.steps
        value = patches / (ms_total * 1e-3)
        e2e_v = patches / (ms_e2e * 1e-3)
        line = {
            "metric": cfg["metric"], "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype_name(eo_vae.grad_dtype() if train else eo_vae.inference_dtype()),
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "batch_per_gpu": batch,
                       "parallelism": f"dp{world} (patches sharded by rank, " + ("one gradient all-reduce per step)" if train else "no collective)"),
                       "numerics": eo_vae.numerics_description(),
                       "l2": f"inputs ({int(h2d) >> 20} MiB/step) and the level-0/1 activations exceed the 126 MB L2"
                             if batch * size * size >= 64 * 256 * 256 // 4 else
                             "small working set: steps are separated by the ~100 intermediate tensors of the pass, not by an L2 flush"},
            "tensor_tflops": value * (gf / 1000.0),
            "e2e": {"value": e2e_v, "unit": "patches/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
        }
        if comm is not None:
            line["gradient_exchange"] = comm
        if extra_train is not None:
            line["train_step"] = extra_train
        if extra_mixed is not None:
            line["train_step_mixed_modality"] = extra_mixed
        if extra_512 is not None:
            line["encode_512px_13band"] = extra_512
        if eager is not None:
            line["gpu_eager"] = eager
        if world == 1 and not os.environ.get("EOVAE_BENCH_NO_CPU"):
            try:
                line["cpu_baseline"] = cpu_baseline(cfg)
            except Exception as exc:  # noqa: BLE001
                line["cpu_baseline"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


TRAIN_BATCH = 16


def measure_encode(model, cfg, dev, world, rank, steps: int = 5):
    """Secondary figure of the default line: another encode configuration (BASELINE configs[4]: S2L1C 13x512x512, batch 32 per
    GPU) with the model of the headline run; ``--config 5`` measures it as a line of its own."""
    import torch
    import torch.distributed as dist
    from oracle.weights import WAVELENGTHS
    wvs = torch.tensor(WAVELENGTHS[cfg["modality"]], dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(777 + rank)
    x = torch.randn((cfg["batch"], wvs.numel(), cfg["size"], cfg["size"]), generator=gen, device=dev).clamp_(-2.0, 6.0)
    with torch.no_grad():
        for _ in range(3):
            model.encode_spatial_normalized(x, wvs)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            model.encode_spatial_normalized(x, wvs)
        e1.record()
        torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    pps = world * cfg["batch"] / ms * 1e3
    return {"workload": cfg["workload"], "patches_per_s": pps, "ms_per_step": ms,
            "tensor_tflops": pps * gf_per_patch("encode", wvs.numel(), cfg["size"]) / 1000.0, "steps": steps}


def measure_train_step(g, dev, world, rank, steps: int = 5, mods=("S2L2A",), graphed: bool = True):
    """Secondary figure of the default line (BASELINE configs[2], or configs[3] with ``mods`` = the three modalities drawn per
    step and per rank): EOFluxVAE.training_step, batch 16 per GPU, Charbonnier + MS-SSIM loss, clip 1.0, Adam; gradients
    averaged over the ranks.  ``--config 3`` / ``--config 4`` measure the same steps as lines of their own."""
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

    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    batches = {m: {"image": torch.randn((TRAIN_BATCH, len(WAVELENGTHS[m]), 256, 256), generator=gen, device=dev).clamp_(-2.0, 6.0),
                   "wvs": torch.tensor(WAVELENGTHS[m], dtype=torch.float32, device=dev)} for m in mods}
    rng = random.Random(1234 + rank)
    seq = [rng.choice(list(mods)) for _ in range(64)]
    batch = batches[mods[0]]

    def timed(fn):
        for i in range(max(3, len(mods))):
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

    def checksum(m):
        """Replicas must hold bit-identical parameters after exchanged steps."""
        if world == 1:
            return None
        chk = torch.stack([p.detach().double().sum() for p in m.parameters()]).sum().reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        return bool(all(torch.equal(allc[0], c) for c in allc))

    def eager(model):
        # warm-up iterations 0..len(mods)-1 visit every modality once (weight-operand caches, allocator)
        return lambda i: model.training_step(batches[mods[i] if i < len(mods) else seq[i % len(seq)]], i)

    model = build()
    ms_eager, loss = timed(eager(model))
    sync_eager = checksum(model)
    ms_local = None
    if world > 1:   # exposed cost of the exchange: the same step with the collective switched off
        gs = model._grad_sync
        model._grad_sync = None
        gs.remove()
        ms_local, _ = timed(eager(model))
    del model
    pps = lambda ms: world * TRAIN_BATCH / ms * 1e3  # noqa: E731
    gf = statistics.mean(gf_per_patch("train", len(WAVELENGTHS[m]), 256) for m in mods)
    out = {"workload": CONFIGS[3 if len(mods) == 1 else 4]["workload"],
           "patches_per_s": pps(ms_eager), "ms_per_step": ms_eager, "tensor_tflops": pps(ms_eager) * gf / 1000.0,
           "loss_after_8_steps": loss, "steps": steps}
    sync_graph = True
    if graphed:
        model = build()
        graph_step = GraphedTrainStep(model, batch)
        ms_graph, _ = timed(lambda i: graph_step(batch))
        sync_graph = checksum(model)
        out.update({"graphed_patches_per_s": pps(ms_graph), "graphed_ms_per_step": ms_graph,
                    "graphed_tensor_tflops": pps(ms_graph) * gf / 1000.0})
        del model, graph_step
    if world > 1:
        out.update({"ms_per_step_without_exchange": ms_local, "exposed_exchange_ms": ms_eager - ms_local,
                    "scaling_eff": min(1.0, ms_local / ms_eager), "replicas_in_sync": bool(sync_eager and sync_graph)})
        if graphed:
            out["graphed_scaling_eff"] = min(1.0, ms_local / ms_graph)
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--dtype", default=None, choices=["bf16", "fp16", "fp32"], help="override the package's default numerics")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.impl == "reference-cuda":
        run_reference_cuda(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()

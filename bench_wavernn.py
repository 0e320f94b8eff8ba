"""WaveRNN workload of bench.py (--workload wavernn_cfg3): BASELINE.json configs[2],
WaveRNN batched generate (target=8000, overlap=400) on a 30 s random mel -> 58 folds x 8800 steps.

A "step" is one whole generate() of that utterance.  value = delivered audio samples/s (487,600 per
step; the 510,400 raw draws/s are reported beside it), device-resident mel, built-in counter-based
noise.  e2e = wavernn.inference.infer_waveform() with a host numpy mel and the host-side float64
post-processing (cross-fade, mu-law, de-emphasis) inside the timed region, same noise source;
"e2e_torch_rng" additionally reports the reference-compatible mode whose Exp(1) stream is drawn by
the host torch generator (that is what makes the integer samples equal the reference's).
Under torchrun every rank vocodes its own utterance (weak scaling, no collective on the data path).
"""
from __future__ import annotations

import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
FRAMES, TARGET, OVERLAP = 2400, 8000, 400


def _geometry():
    total = FRAMES * 200
    folds = (total - OVERLAP) // (TARGET + OVERLAP)
    if total - (folds * (TARGET + OVERLAP) + OVERLAP) != 0:
        folds += 1
    steps = TARGET + 2 * OVERLAP
    delivered = folds * (TARGET + OVERLAP) + OVERLAP
    return folds, steps, delivered


def cpu_twin(nsteps: int, threads: int):
    """CPU port (the C twin, OpenMP over the fold rows) on `nsteps` steps of the cfg-3 batch."""
    sys.path.insert(0, str(ROOT / "oracle"))
    sys.path.insert(0, str(ROOT / "synth_weights"))
    import numpy as np
    import torch
    import ref_init as ri
    import wavernn_oracle as wo

    os.environ["OMP_NUM_THREADS"] = str(threads)
    folds, steps, _ = _geometry()
    sd = ri.wavernn_state_dict(0, randomize_bn=True)
    twin = wo.Twin({k: v.numpy() for k, v in sd.items() if v.dtype == torch.float32})
    mel = (torch.rand(1, 80, FRAMES, generator=torch.Generator().manual_seed(3)) * 2 - 1)[0].numpy()
    aux, melup = twin.condition(mel)
    starts = np.arange(folds, dtype=np.int32) * (TARGET + OVERLAP)
    t0 = time.perf_counter()
    twin.generate(aux, melup, starts, nsteps, None, seed=1)
    dt = time.perf_counter() - t0
    return folds * nsteps / dt, dt


def run_reference(args, threads):
    folds, steps, delivered = _geometry()
    per = []
    n = 40
    for s in range(args.warmup + args.steps):
        v, dt = cpu_twin(n, threads)
        if s >= args.warmup:
            per.append(dt)
    secs = sum(per)
    value = folds * n * args.steps / secs * (delivered / (folds * steps))
    print(json.dumps({
        "impl": "reference", "metric": "vocoder audio samples/sec", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "wavernn_cfg3: batched generate, 30 s mel, target 8000 / overlap 400 (58 folds x 8800 steps)"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{n} of 8800 steps x 58 folds per step, C twin (OpenMP over folds); the reference's "
                                   "own torch-CPU path measured 11.5k samples/s on 8 cores (SURVEY.md section 6)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT / "oracle"))

    sys.path.insert(0, str(ROOT / "synth_weights"))
    import ref_init as ri
    from bench import ClockSampler
    from mockingbird_b200 import _lib
    from mockingbird_b200.vocoder.wavernn import inference as rnn_vocoder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sd = ri.wavernn_state_dict(0, randomize_bn=True)
    model = rnn_vocoder.load_state(sd, rng="device", seed=1234 + rank)
    if world > 1:
        dist.broadcast(model.packed_arena(), src=0)
    folds, steps, delivered = _geometry()
    mel = torch.rand(1, 80, FRAMES, generator=torch.Generator().manual_seed(3 + rank)) * 2 - 1
    mel_dev = mel.to(dev)
    mel_np = (mel[0] * 4.0).numpy()
    lib = _lib.lib()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([max(e0.elapsed_time(e1), wall)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    def step_resident():
        model.generate_indices(mel_dev, True, TARGET, OVERLAP, None)

    def step_e2e():
        rnn_vocoder.infer_waveform(mel_np, batched=True, target=TARGET, overlap=OVERLAP, progress_callback=lambda *a: None)

    for _ in range(max(1, min(args.warmup, 3))):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.mb_launch_count()
    k = max(1, args.steps)
    ms = timed(step_resident, k)
    launches = int(lib.mb_launch_count() - l0)
    clocks = sampler.stop()
    step_e2e()  # warm-up of the host path (first call imports scipy.signal for the de-emphasis filter)
    ms_e2e = timed(step_e2e, k)
    model.rng = "torch"
    torch.manual_seed(1234)
    ms_torch = timed(step_e2e, 1)
    model.rng = "device"
    value = world * delivered * k / (ms * 1e-3)
    if rank == 0:
        us_per_step = ms / k / steps * 1e3
        flops = 8.14e6 * folds * steps
        cpu = None
        if not args.no_cpu_baseline:
            from bench import cpu_child, host_threads

            threads = host_threads()
            r = cpu_child("wavernn_cfg3", 1200, threads, 240.0)
            v, dt = (r["value"], r["seconds"]) if r else (float("nan"), float("nan"))
            cpu = {"value": v * delivered / (folds * steps), "unit": "samples/s", "cores": threads, "kind": "port",
                   "sample": f"1200 of 8800 steps x 58 folds ({dt:.1f} s), C twin with OpenMP over folds; reference "
                             "torch-CPU path: 11.5k samples/s on 8 cores (SURVEY.md section 6)"}
        print(json.dumps({
            "metric": "vocoder audio samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": k,
            "warmup": max(1, min(args.warmup, 3)), "ms_per_step": ms / k, "rtf": (ms / k * 1e-3) / (delivered / 16000.0),
            "raw_draws_per_s": world * folds * steps * k / (ms * 1e-3), "us_per_sample_step": us_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "wavernn_cfg3: batched generate, 30 s mel (2400 frames), target 8000 / overlap 400 "
                                   "-> 58 folds x 8800 sequential steps per GPU", "rng": "device (counter-based)",
                       "parallelism": f"dp{world}", "l2": "weights are shared-memory resident by design; exchange "
                                                          "buffers are L2 resident by design (no flush applicable)"},
            "e2e": {"value": world * delivered * k / (ms_e2e * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": int(mel_np.nbytes), "d2h_bytes_per_step": folds * steps * 2,
                    "ms_per_step": ms_e2e / k},
            "e2e_torch_rng": {"value": world * delivered / (ms_torch * 1e-3), "unit": "samples/s",
                              "note": "Exp(1) noise drawn by the host torch generator (reference-identical samples); "
                                      "h2d 1.05 GB of noise per step"},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "latency", "kernel": "k_sample_loop", "achieved": flops / (ms / k * 1e-3) / 1e12,
                         "peak": 72.0, "unit": "TFLOP/s", "frac": flops / (ms / k * 1e-3) / 1e12 / 72.0, "traffic": None,
                         "note": "FP32 FFMA; 8800 dependent steps x 6 grid barriers bound the time, not HBM (4 B/draw)"},
            "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()

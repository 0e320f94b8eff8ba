#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 vocoder hot path (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload hifigan_cfg2|fregan_cfg2|wavernn_cfg3|tacotron_cfg4|e2e_cfg5] [--precision f16tc|fp32]

Default workload = BASELINE.json configs[1]: HiFi-GAN Generator forward, batch 32 random mels of
256 frames x 80 bins, per GPU.  A "step" is one forward over one batch.  Prints ONE JSON line
(rank 0).  Under torchrun every rank runs its own batch of the same size (weak scaling: utterance
batches shard across GPUs with no data-path collective; NCCL only broadcasts the packed weights).

  value      samples/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the drop-in surface with HOST (pinned) buffers: H2D of the mel
             batch and D2H of the waveforms inside the timed region
  roofline   dominant kernel (the tcgen05 conv kernel): algorithmic FLOPs / CUDA-event time of
             those launches (separate profiled pass) vs MEASURED_PEAKS.json bf16 peak; plus the
             north-star's layer-granular HBM view for the whole step ("roofline_hbm_step")
  cpu_baseline  the oracle (torch-CPU restatement, bit-identical to the reference forward) on the
             host cores, bounded sample
--impl reference: times that CPU implementation on the same config (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hifigan_cfg2", choices=["hifigan_cfg2", "fregan_cfg2", "wavernn_cfg3", "tacotron_cfg4", "e2e_cfg5"])
    ap.add_argument("--precision", default=os.environ.get("MOCKINGBIRD_B200_GAN_PRECISION", "f16tc"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-child", nargs=3, default=None, help=argparse.SUPPRESS)
    return ap.parse_args()


def host_threads() -> int:
    """CPU threads this job may really use: affinity mask capped by the cgroup CPU quota (a container
    on a 200-core host with an 8-core quota must not spawn 200 spinning OpenMP threads)."""
    n = len(os.sched_getaffinity(0))
    try:
        txt = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if txt[0] != "max":
            n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
    except Exception:
        try:
            q = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
            per = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            if q > 0:
                n = min(n, max(1, int(q / per + 0.5)))
        except Exception:
            pass
    return max(1, min(n, 64))


def log(msg: str) -> None:
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def cpu_child(workload: str, amount: int, threads: int, timeout: float):
    """Run the CPU baseline in a child process with CUDA hidden and a hard
    time limit; returns the child's JSON dict or None."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS=str(threads), MKL_NUM_THREADS=str(threads))
    try:
        out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--cpu-child", workload, str(amount), str(threads)],
                             env=env, capture_output=True, text=True, timeout=timeout)
        for line in reversed(out.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        log(f"cpu child produced no result: {out.stderr[-500:]}")
    except subprocess.TimeoutExpired:
        log(f"cpu child exceeded {timeout}s")
    return None


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_hifigan(batch_rows: int, passes: int, threads: int, fregan: bool = False):
    """The CPU implementation (oracle port of the reference forward), batch-1 calls like
    hifigan/inference.py:66-70.  Returns (samples_per_s, seconds, samples)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    sys.path.insert(0, str(ROOT / "synth_weights"))
    import torch
    import gan_oracle as go
    import ref_init as ri

    torch.set_num_threads(threads)
    cfg = ri.FREGAN_CONFIG if fregan else ri.HIFIGAN_CONFIG_16K
    sd = go.fold_weight_norm(ri.fregan_state_dict(cfg, 0)) if fregan else ri.hifigan_state_dict(cfg, 0)
    fwd = go.fregan_forward if fregan else go.hifigan_forward
    mel = torch.rand(32, 80, 256, generator=torch.Generator().manual_seed(2)) * 8 - 4
    with torch.no_grad():
        fwd(sd, cfg, mel[:1])  # warm-up
        t0 = time.perf_counter()
        n = 0
        for _ in range(passes):
            for i in range(batch_rows):
                y = fwd(sd, cfg, mel[i:i + 1])
                n += y.numel()
        dt = time.perf_counter() - t0
    return n / dt, dt, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    if args.workload == "wavernn_cfg3":
        import bench_wavernn

        return bench_wavernn.run_reference(args, threads)
    if args.workload == "tacotron_cfg4":
        import bench_tacotron

        return bench_tacotron.run_reference(args, threads)
    if args.workload == "e2e_cfg5":
        import bench_e2e

        return bench_e2e.run_reference(args, threads)
    per_step = []
    total = 0
    for s in range(args.warmup + args.steps):
        log(f"reference step {s}")
        v, dt, n = cpu_hifigan(32, 1, threads, fregan=(args.workload == "fregan_cfg2"))
        if s >= args.warmup:
            per_step.append(dt)
            total += n
    secs = sum(per_step)
    value = total / secs
    line = {
        "impl": "reference", "metric": "vocoder audio samples/sec", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "rtf": (secs / args.steps) / (32 * 51200 / 16000.0), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: Generator fwd, batch 32 x 256 frames x 80 mels (batch-1 calls)",
                   "per_gpu_batch": 32, "frames": 256},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x 32 utterances x 256 frames, torch-CPU oracle "
                                   "(bit-identical to the reference forward)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_ours_hifigan(args):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT / "oracle"))  # ref_init only: seeded random-init weights (no checkpoints exist)

    sys.path.insert(0, str(ROOT / "synth_weights"))
    import ref_init as ri
    from mockingbird_b200 import _lib
    from mockingbird_b200.vocoder.fregan.models import FreGAN
    from mockingbird_b200.vocoder.hifigan.models import Generator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    fre = args.workload == "fregan_cfg2"  # SURVEY.md row C1 on the same batch shape (not a BASELINE.json config)
    cfg = ri.FREGAN_CONFIG if fre else ri.HIFIGAN_CONFIG_16K
    make_sd = (lambda: ri.fregan_state_dict(cfg, 0)) if fre else (lambda: ri.hifigan_state_dict(cfg, 0))
    B, T = 32, 256
    g = (FreGAN(cfg, precision=args.precision) if fre else Generator(cfg, precision=args.precision)).to(dev)
    # rank 0 builds + packs the weights, NCCL broadcasts the packed arena (the only collective)
    if rank == 0:
        g.load_state_dict(make_sd())
        g.eval()
        g.remove_weight_norm()
    else:
        g.load_state_dict(make_sd())  # shapes only; contents overwritten below
        g.eval()
        g.remove_weight_norm()
    if world > 1:
        dist.broadcast(g.packed_arena(), src=0)
    hop = g.hop
    mel = (torch.rand(B, 80, T, generator=torch.Generator().manual_seed(2 + rank)) * 8 - 4)
    mel_dev = mel.to(dev)
    mel_pin = mel.pin_memory()
    wav_pin = torch.empty(B, 1, T * hop).pin_memory()
    samples_per_step = B * T * hop
    lib = _lib.lib()

    def barrier():
        torch.cuda.synchronize()
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
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    def step_resident():
        g(mel_dev)

    def step_e2e():
        x = mel_pin.to(dev, non_blocking=True)
        y = g(x)
        wav_pin.copy_(y, non_blocking=True)

    log("weights packed; warm-up")
    for _ in range(max(3, args.warmup)):
        step_resident()
    torch.cuda.synchronize()
    log("timed region (resident)")
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.mb_launch_count()
    ms_total = timed(step_resident, args.steps)
    launches = int(lib.mb_launch_count() - l0)
    clocks = sampler.stop()
    log(f"resident: {ms_total / args.steps:.3f} ms/step; e2e pass")
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    log("profiled pass")

    value = world * samples_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = world * samples_per_step * args.steps / (ms_e2e * 1e-3)
    ms_step = ms_total / args.steps

    # ---- roofline of the dominant kernel: separate profiled pass (events around every launch)
    pk = peaks()
    roof = None
    roof_hbm = None
    roof_tensor = None
    if rank == 0:
        n = g.num_layers()
        acc = {}
        reps = 3
        for _ in range(reps):
            _, ms_layers = g.forward_profiled(mel_dev)
            for i, t in enumerate(ms_layers):
                info = g.layer_info(i)
                cls = "resblock_conv" if "resblocks" in info else ("ups" if " ups." in info else info.split()[1])
                macs, lbytes = g.layer_work(i, B, T)
                a = acc.setdefault(cls, [0.0, 0.0, 0.0, 0])
                a[0] += t
                a[1] += 2 * macs
                a[2] += lbytes
                a[3] += 1
        dom = max(acc, key=lambda k: acc[k][0])
        t_ms, flops, lbytes, cnt = acc[dom]
        tf = flops / (t_ms * 1e-3) / 1e12
        roof_tensor = None
        if args.precision == "f16tc":
            peak = pk["tflops_sustained"]
            roof_tensor = {"bound": "tensor", "kernel": f"tc_conv / tc_pair ({dom})", "achieved": tf, "peak": peak,
                           "unit": "TFLOP/s", "frac": tf / peak, "peak_source": pk["source"] + " bf16 sustained (fp16 same rate)",
                           "launches_timed": cnt, "flop_per_launch": flops / cnt, "ms_per_launch": t_ms / cnt,
                           "share_of_step": (t_ms / reps) / ms_step}
            # north-star roofline of the dominant kernel family (tcgen05 conv kernels: conv_pre, ups, resblocks):
            # ALGORITHMIC layer-granular fp32 bytes of those layers (13 190 B per output sample over
            # the whole generator) / their event-timed duration, vs the measured HBM copy bandwidth.  traffic =
            # DRAM bytes the same launches really move (ncu dram__bytes_read+write, profiles/, per launch).
            fam = [k for k in acc if k != "conv_post"]
            fam_ms = sum(acc[k][0] for k in fam) / reps
            fam_bytes = sum(acc[k][2] for k in fam) / reps
            traffic = None
            launches_per_fw = None
            tj = ROOT / "profiles" / "r01_hifigan_dram_traffic.json"
            if tj.exists() and not fre:
                tr = json.loads(tj.read_text())
                traffic = tr["tc_dram_bytes_per_launch"]
                launches_per_fw = tr["tc_launches_per_forward"]
            gbs = fam_bytes / (fam_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "tc_conv_kernel + tc_pair_kernel (tcgen05 tap convs; all layers but conv_post)",
                    "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "traffic": traffic,
                    "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r01_hifigan_dram_traffic.json",
                    "launches_per_step": launches_per_fw,
                    "algorithmic_bytes_per_launch": (fam_bytes / launches_per_fw) if launches_per_fw else None,
                    "algorithmic_bytes_per_step": fam_bytes, "ms_per_step_in_kernel": fam_ms,
                    "share_of_step": fam_ms / ms_step, "peak_source": pk["source"] + " HBM copy bandwidth",
                    "definition": "layer-granular fp32 bytes (inputs + outputs of every conv layer + weights once) / time"}
        else:
            peak = 72.0  # 148 SM x 128 lanes x 2 x ~1.9 GHz FP32 FFMA, nominal
            roof = {"bound": "tensor", "kernel": f"tapconv_f32 ({dom})", "achieved": tf, "peak": peak, "unit": "TFLOP/s",
                    "frac": tf / peak, "traffic": None, "peak_source": "nominal fp32 FFMA (parity-anchor path)",
                    "launches_timed": cnt, "share_of_step": (t_ms / reps) / ms_step}
        # north-star view: layer-granular fp32 bytes of the whole step / step time vs measured HBM copy BW
        step_bytes = sum(v[2] for v in acc.values()) / reps + (68_926_660 if fre else 51_902_980)  # + fp32 weights once
        gbs = step_bytes / (ms_step * 1e-3) / 1e9
        roof_hbm = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                    "bytes_per_step": step_bytes, "definition": "layer-granular fp32 bytes (SURVEY.md 8d) / step time"}
        step_tf = sum(v[1] for v in acc.values()) / reps / (ms_step * 1e-3) / 1e12
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = host_threads()
        log(f"cpu baseline on {threads} threads")
        r = cpu_child(args.workload, 4, threads, 240.0)
        if r is not None:
            cpu = {"value": r["value"], "unit": "samples/s", "cores": threads, "kind": "port",
                   "sample": f"4 passes x 32 utterances x 256 frames ({r['seconds']:.1f} s), torch-CPU oracle "
                             "(bit-identical to the reference forward), batch-1 calls like hifigan/inference.py"}
    if rank == 0:
        line = {
            "metric": "vocoder audio samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
            "rtf": (ms_step * 1e-3) / (samples_per_step / 16000.0), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 operands / f32 accumulate + f32 residual" if args.precision == "f16tc" else "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: Generator fwd, batch 32 x 256 frames x 80 mels per GPU",
                       "per_gpu_batch": B, "frames": T, "precision": args.precision, "parallelism": f"dp{world}",
                       "l2": "per-step working set (2.9 GB of activations) >> 126 MB L2; no explicit flush",
                       "weights": "random init, torch.manual_seed(0) order of the reference constructor"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": mel_pin.numel() * 4,
                    "d2h_bytes_per_step": wav_pin.numel() * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_tensor": roof_tensor,
            "roofline_hbm_step": roof_hbm,
            "step_tflops": step_tf, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.cpu_child is not None:
        workload, amount, threads = args.cpu_child[0], int(args.cpu_child[1]), int(args.cpu_child[2])
        if workload in ("hifigan_cfg2", "fregan_cfg2"):
            v, dt, n = cpu_hifigan(32, amount, threads, fregan=(workload == "fregan_cfg2"))
        elif workload == "tacotron_cfg4":
            import bench_tacotron

            v, dt = bench_tacotron.cpu_oracle(amount, threads)
        elif workload == "e2e_cfg5":
            import bench_e2e

            v, dt = bench_e2e.cpu_oracle(amount, threads)
        else:
            import bench_wavernn

            v, dt = bench_wavernn.cpu_twin(amount, threads)
        print(json.dumps({"value": v, "seconds": dt}))
        return
    if args.impl == "reference":
        return run_reference(args)
    if args.workload in ("hifigan_cfg2", "fregan_cfg2"):
        return run_ours_hifigan(args)
    if args.workload == "tacotron_cfg4":
        import bench_tacotron

        return bench_tacotron.run_ours(args)
    if args.workload == "e2e_cfg5":
        import bench_e2e

        return bench_e2e.run_ours(args)
    import bench_wavernn

    return bench_wavernn.run_ours(args)


if __name__ == "__main__":
    main()

"""Per-layer CUDA-event timing of the GAN generator plan (mb_gan_forward_profiled).
usage: python tools/profile_layers.py [--model hifigan|fregan] [--precision f16tc] [--batch 32] [--frames 256] [--reps 5] [--out file.tsv]"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "synth_weights"))

import torch  # noqa: E402

import ref_init as ri  # noqa: E402
from mockingbird_b200.vocoder.hifigan.models import Generator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="hifigan", choices=["hifigan", "fregan"])
    ap.add_argument("--precision", default="f16tc")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.model == "fregan":
        from mockingbird_b200.vocoder.fregan.models import FreGAN

        cfg = ri.FREGAN_CONFIG
        g = FreGAN(cfg, precision=a.precision).cuda()
        g.load_state_dict(ri.fregan_state_dict(cfg, 0))
    else:
        cfg = ri.HIFIGAN_CONFIG_16K
        g = Generator(cfg, precision=a.precision).cuda()
        g.load_state_dict(ri.hifigan_state_dict(cfg, 0))
    g.eval()
    g.remove_weight_norm()
    mel = (torch.rand(a.batch, 80, a.frames, generator=torch.Generator().manual_seed(2)) * 8 - 4).cuda()
    for _ in range(3):
        g(mel)
    n = g.num_layers()
    acc = [0.0] * n
    for _ in range(a.reps):
        _, ms = g.forward_profiled(mel)
        for i, t in enumerate(ms):
            acc[i] += t / a.reps
    lines = ["idx\tname\tcin\tcout\tk\tdil\tstride\tms\tTFLOPs\tlayerGB/s"]
    tot = 0.0
    for i in range(n):
        info = g.layer_info(i)
        name = info.split()[1]
        d = dict(tok.split("=") for tok in info.split() if "=" in tok)
        macs, lbytes = g.layer_work(i, a.batch, a.frames)
        t = acc[i]
        tot += t
        lines.append(f"{i}\t{name}\t{d['cin']}\t{d['cout']}\t{d['k']}\t{d['dil']}\t{d['stride']}\t{t:.4f}\t"
                     f"{2 * macs / (t * 1e-3) / 1e12:.1f}\t{lbytes / (t * 1e-3) / 1e9:.0f}")
    lines.append(f"total\t\t\t\t\t\t\t{tot:.3f}")
    text = "\n".join(lines)
    print(text)
    if a.out:
        Path(a.out).write_text(text + "\n")


if __name__ == "__main__":
    main()

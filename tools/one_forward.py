"""Minimal driver for ncu: N forwards of the cfg-2 HiFi-GAN batch (default 2) inside a cudaProfilerStart/Stop range
(run ncu with --profile-from-start off so that weight packing and the warm-up forward are not captured)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "synth_weights"))
import torch  # noqa: E402

import ref_init as ri  # noqa: E402
from mockingbird_b200.vocoder.hifigan.models import Generator  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
prec = sys.argv[2] if len(sys.argv) > 2 else "f16tc"
cfg = ri.HIFIGAN_CONFIG_16K
g = Generator(cfg, precision=prec).cuda()
g.load_state_dict(ri.hifigan_state_dict(cfg, 0))
g.eval()
g.remove_weight_norm()
mel = (torch.rand(32, 80, 256, generator=torch.Generator().manual_seed(2)) * 8 - 4).cuda()
g(mel)  # warm-up (outside the profiled range)
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: only the forwards below are captured
for _ in range(n):
    g(mel)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

import sys, os
from pathlib import Path
ROOT = Path("/root/repo") if Path("/root/repo/tools").exists() else Path.cwd()
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "synth_weights"))
import numpy as np, torch
import ref_init as ri
from mockingbird_b200.vocoder.hifigan.models import Generator
cfg = ri.HIFIGAN_CONFIG_16K
g = Generator(cfg, precision="f16tc").cuda(); g.load_state_dict(ri.hifigan_state_dict(cfg, 0)); g.eval(); g.remove_weight_norm()
mel = (torch.rand(5, 80, 97, generator=torch.Generator().manual_seed(3)) * 8 - 4).cuda()
lens = torch.tensor([97, 50, 1, 96, 13], dtype=torch.int32).cuda()
w = g(mel, lengths=lens).cpu().numpy()
np.save(sys.argv[1], w)

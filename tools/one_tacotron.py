"""Minimal driver for ncu: one Tacotron generate of the cfg-4 batch with a reduced number of steps."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "synth_weights"))
import torch  # noqa: E402

import ref_init as ri  # noqa: E402
import bench_tacotron as bt  # noqa: E402
from mockingbird_b200.synthesizer.inference import Synthesizer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
syn = Synthesizer("unused.pt", verbose=False)
model = syn.load_state(ri.tacotron_state_dict(0, r=2, randomize_bn=True))
chars, emb, _ = bt.make_inputs()
model.generate(chars.cuda(), emb.cuda(), steps=steps, style_idx=-1, min_stop_token=10)  # warm-up: packs images, builds the graph
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off
model.generate(chars.cuda(), emb.cuda(), steps=steps, style_idx=-1, min_stop_token=10)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

"""quick timing of the cfg-3 WaveRNN call in the three noise modes (development tool)"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "synth_weights"))
import numpy as np, torch
import ref_init as ri
from mockingbird_b200.vocoder.wavernn import inference as rnn_vocoder
model = rnn_vocoder.load_state(ri.wavernn_state_dict(0, randomize_bn=True))
mel = torch.rand(1, 80, 2400, generator=torch.Generator().manual_seed(3)) * 2 - 1
mel_np = (mel[0] * 4.0).numpy()
for mode, reps in (("device", 3), ("torch", 4), ("torch_host", 1)):
    model.rng = mode
    for r in range(reps):
        torch.manual_seed(1234)
        torch.cuda.synchronize(); t = time.perf_counter()
        wav, sr = rnn_vocoder.infer_waveform(mel_np, batched=True, target=8000, overlap=400, progress_callback=lambda *a: None)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"{mode:10s} rep {r}: {dt*1e3:8.1f} ms  {len(wav)/dt/1e6:.3f} M samples/s  checksum {np.abs(wav).sum():.6f}")

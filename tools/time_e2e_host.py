"""Where does the host side of hifigan.infer_waveforms spend its time?  (cfg 2: 32 utterances x 256 frames)"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "synth_weights"))
import numpy as np, torch
import ref_init as ri
from mockingbird_b200.vocoder.hifigan import inference as gan

cfg = ri.HIFIGAN_CONFIG_16K
from mockingbird_b200.vocoder.hifigan.models import Generator
g = Generator(cfg).cuda(); g.load_state_dict(ri.hifigan_state_dict(cfg, 0)); g.eval(); g.remove_weight_norm()
gan.generator = g; gan._device = torch.device("cuda")
rng = np.random.default_rng(0)
mels = [(rng.random((80, 256), dtype=np.float32) * 8 - 4) for _ in range(32)]
for _ in range(5): gan.infer_waveforms(mels)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    t0 = time.perf_counter(); gan.infer_waveforms(mels); ts.append(time.perf_counter() - t0)
print("infer_waveforms ms: median %.3f min %.3f" % (np.median(ts) * 1e3, min(ts) * 1e3))
# pieces
n_out = 32 * 256 * 200
host = torch.empty(n_out, dtype=torch.float32).pin_memory()
# (torch's multi-threaded copy_ looks fastest here, 0.02-0.17 ms, but under a container CPU quota - 16 of 64 visible cores on the
# GPU boxes - its 64 OpenMP threads stall a soaked bench loop to 20 ms per call: profiles/r02_e2e_host_path.txt)
for name, fn in [("np.array copy", lambda: np.array(host.numpy(), copy=True)),
                 ("torch copy_", lambda: torch.empty(n_out).copy_(host).numpy()),
                 ("np.empty+copyto", lambda: np.copyto(np.empty(n_out, np.float32), host.numpy()))]:
    fn(); t0 = time.perf_counter()
    for _ in range(20): fn()
    print("%-18s %.3f ms" % (name, (time.perf_counter() - t0) / 20 * 1e3))
hin = torch.empty(32, 80, 256).pin_memory()
t0 = time.perf_counter()
for _ in range(20):
    for r in range(32): hin[r, :, :256] = torch.as_tensor(mels[r])
print("fill pinned input   %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
dev = hin.cuda(); lens = torch.full((32,), 256, dtype=torch.int32).cuda()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
outp = torch.empty(32, 1, 51200).pin_memory()
torch.cuda.synchronize(); e0.record(); w = g(dev, lengths=lens); e1.record(); outp.copy_(w, non_blocking=True); e2.record(); torch.cuda.synchronize()
print("forward %.3f ms  D2H %.3f ms" % (e0.elapsed_time(e1), e1.elapsed_time(e2)))
print("threads", torch.get_num_threads())

"""debug: stage-by-stage comparison of mb_wavernn_postprocess vs the host path on the batched golden"""
import sys, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "synth_weights"))
import numpy as np, torch
import wavernn_oracle as wo
from mockingbird_b200 import _lib
z = np.load(ROOT / "tests/golden/wavernn_seed0.npz"); idx = z["idx2"]; folds, steps = idx.shape; target, overlap = 1000, 100
L = _lib.lib()
d_idx = torch.from_numpy(np.ascontiguousarray(idx)).cuda()
total = folds * (target + overlap) + overlap
def run(mu, pre, fade_len, wave_len):
    ws = torch.empty(int(L.mb_wavernn_postprocess_workspace_bytes(folds, steps, 1, target, overlap)), dtype=torch.uint8, device="cuda")
    out = torch.empty(total, dtype=torch.float64, device="cuda"); n = C.c_int64()
    _lib.check(L.mb_wavernn_postprocess(C.c_void_p(d_idx.data_ptr()), folds, steps, 1, target, overlap, 512, mu, pre, wave_len, fade_len,
               C.c_void_p(out.data_ptr()), C.byref(n), C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize(); return out[:n.value].cpu().numpy()
y = (2 * idx.astype(np.float32) / np.float32(511.) - np.float32(1.)).astype(np.float64)
unf = wo.xfade_and_unfold(y, target, overlap)
a = run(0, 0.0, 0, total); d = np.abs(a - unf); print("unfold      max", d.max(), "first bad", np.flatnonzero(d > 1e-12)[:5], a[np.flatnonzero(d>1e-12)[:3]], unf[np.flatnonzero(d>1e-12)[:3]])
mu = np.sign(unf) / 511 * (512 ** np.abs(unf) - 1)
b = run(1, 0.0, 0, total); d = np.abs(b - mu); print("mulaw       max", d.max(), np.flatnonzero(d > 1e-12)[:5])
from scipy.signal import lfilter
de = lfilter([1], [1, -0.97], mu)
c = run(1, 0.97, 0, total); d = np.abs(c - de); print("deemph      max", d.max(), np.flatnonzero(d > 1e-12)[:5])

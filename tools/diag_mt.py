"""diagnostic: where do the 47 ms between rng='device' (436 ms) and rng='torch' (483 ms) of cfg 3 go?"""
import sys, time, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "synth_weights"))
import numpy as np, torch
import ref_init as ri
from mockingbird_b200 import _lib
from mockingbird_b200.vocoder.wavernn import inference as rnn_vocoder
from mockingbird_b200.vocoder.wavernn.models.fatchord_version import torch_cpu_generator_position
L = _lib.lib()
model = rnn_vocoder.load_state(ri.wavernn_state_dict(0, randomize_bn=True))
mel = (torch.rand(1, 80, 2400, generator=torch.Generator().manual_seed(3)) * 2 - 1).cuda()
B, steps, CH = 58, 8800, 100
def t(fn, n=3):
    out = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); out.append((time.perf_counter() - t0) * 1e3)
    return [round(x, 1) for x in out]
# (1) kernel fed from a device-resident noise tensor (no stream machinery)
noise = torch.empty(steps, B, 512, device="cuda").exponential_(1)
print("device-resident injected noise:", t(lambda: model.generate_indices_device(mel, True, 8000, 400, None, noise=noise)))
model.rng = "device"; print("rng=device                    :", t(lambda: model.generate_indices_device(mel, True, 8000, 400, None)))
model.rng = "torch";  print("rng=torch (mtstream)          :", t(lambda: model.generate_indices_device(mel, True, 8000, 400, None)))
# (2) the stream alone: host MT19937 + H2D + convert, no consumer kernel
ms = C.c_void_p(); words = CH * B * 512 * 2
_lib.check(L.mb_mtstream_create(words, 3, C.byref(ms)))
raw = [torch.empty(words, dtype=torch.int32, device="cuda") for _ in range(2)]
out = [torch.empty(CH * B * 512, dtype=torch.float32, device="cuda") for _ in range(2)]
st = torch.cuda.current_stream().cuda_stream
def stream_only():
    s, l, n = torch_cpu_generator_position()
    _lib.check(L.mb_mtstream_begin(ms, s.ctypes.data, l, n, steps * B * 512 * 2, words))
    for c in range(steps // CH):
        _lib.check(L.mb_mtstream_next(ms, CH * B * 512, C.c_void_p(raw[c & 1].data_ptr()), C.c_void_p(out[c & 1].data_ptr()), C.c_void_p(st)))
        _lib.check(L.mb_mtstream_consumed(ms, C.c_void_p(st)))
    lc, nc = C.c_int32(), C.c_int32()
    _lib.check(L.mb_mtstream_finish(ms, s.ctypes.data, C.byref(lc), C.byref(nc)))
print("mtstream alone (88 chunks)    :", t(stream_only))
# (3) host generator alone
buf = np.empty(words, np.uint32); s, l, n = torch_cpu_generator_position(); lc, nc = C.c_int32(l), C.c_int32(n)
t0 = time.perf_counter()
for c in range(88): L.mb_mt19937_fill(s.ctypes.data, C.byref(lc), C.byref(nc), buf.ctypes.data, words)
print("host mt19937_fill (tempered), 88 chunks: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
# (4) H2D of one chunk from pinned memory, and the convert kernel alone
pin = torch.empty(words, dtype=torch.int32).pin_memory()
print("H2D 23.7 MB pinned x88:", t(lambda: [raw[0].copy_(pin, non_blocking=True) for _ in range(88)]))
print("convert kernel x88    :", t(lambda: [L.mb_mt_to_exp(C.c_void_p(raw[0].data_ptr()), C.c_void_p(out[0].data_ptr()), CH * B * 512, C.c_void_p(st)) for _ in range(88)]))

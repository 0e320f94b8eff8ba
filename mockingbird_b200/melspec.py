"""Host wrapper of the mel-spectrogram front-end kernels (mb_melspec_*): one object per parameter set."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class MelSpectrogram:
    def __init__(self, sample_rate, n_fft, hop_length, win_length, n_mels, fmin, fmax, pad_mode="reflect", preemphasis=0.0,
                 power=2, to_db=False, min_level_db=-100.0, ref_level_db=20.0, normalize=False, max_abs_value=4.0,
                 symmetric=True, transpose_out=False, device=None):
        cfg = _lib.MelSpecConfig()
        cfg.sample_rate, cfg.n_fft, cfg.hop_length, cfg.win_length, cfg.n_mels = sample_rate, n_fft, hop_length, win_length, n_mels
        cfg.fmin, cfg.fmax = float(fmin), float(fmax)
        cfg.pad_mode = {"reflect": 0, "constant": 1}[pad_mode]
        cfg.preemphasis, cfg.power, cfg.to_db = float(preemphasis), int(power), int(bool(to_db))
        cfg.min_level_db, cfg.ref_level_db = float(min_level_db), float(ref_level_db)
        cfg.normalize, cfg.max_abs_value, cfg.symmetric = int(bool(normalize)), float(max_abs_value), int(bool(symmetric))
        cfg.transpose_out = int(bool(transpose_out))
        self._cfg = cfg
        self.n_mels, self.transpose_out = n_mels, bool(transpose_out)
        self._handle = C.c_void_p()
        L = _lib.lib()
        _lib.check(L.mb_melspec_create(C.byref(cfg), C.byref(self._handle)))
        self._device = torch.device(device) if device is not None else _lib.require_cuda()
        nbytes = int(L.mb_melspec_arena_bytes(self._handle))
        with torch.cuda.device(self._device):
            self._arena = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _lib.check(L.mb_melspec_set_arena(self._handle, C.c_void_p(self._arena.data_ptr()), nbytes, C.c_void_p(stream)))

    def __call__(self, wav) -> torch.Tensor:
        """wav: 1-D float array / tensor (host or device) -> device tensor [frames, n_mels] or [n_mels, frames]"""
        L = _lib.lib()
        x = torch.as_tensor(np.asarray(wav, dtype=np.float32) if not torch.is_tensor(wav) else wav, dtype=torch.float32)
        x = x.to(self._device).contiguous()
        n = int(x.numel())
        frames = int(L.mb_melspec_num_frames(self._handle, n))
        shape = (frames, self.n_mels) if self.transpose_out else (self.n_mels, frames)
        out = torch.empty(shape, dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _lib.check(L.mb_melspec_forward(self._handle, C.c_void_p(x.data_ptr()), n, C.c_void_p(out.data_ptr()),
                                            C.c_void_p(stream)))
        return out

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None and self._handle.value:
                _lib.lib().mb_melspec_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

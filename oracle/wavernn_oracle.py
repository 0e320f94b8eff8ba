"""WaveRNN oracle (TEST INFRASTRUCTURE): ctypes wrapper of the C twin (oracle/wavernn_twin.c) plus a
numpy restatement of the host-side post-processing of WaveRNN.generate:

  fold_with_overlap      fatchord_version.py:288-338  (only the fold start offsets are needed)
  xfade_and_unfold       fatchord_version.py:340-402
  decode_mu_law          wavernn/audio.py:102-107
  de_emphasis            wavernn/audio.py:92-93 (scipy.signal.lfilter([1],[1,-0.97]))
  trim / fade-out        fatchord_version.py:251-253
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np

import build_oracle

_lib = None


class TwinWeights(C.Structure):
    _fields_ = [
        ("conv_in_w", C.c_void_p), ("bn0", C.c_void_p * 4),
        ("res_conv1", C.c_void_p * 10), ("res_bn1", (C.c_void_p * 4) * 10),
        ("res_conv2", C.c_void_p * 10), ("res_bn2", (C.c_void_p * 4) * 10),
        ("conv_out_w", C.c_void_p), ("conv_out_b", C.c_void_p), ("up_w", C.c_void_p * 3),
        ("I_w", C.c_void_p), ("I_b", C.c_void_p),
        ("rnn1_wih", C.c_void_p), ("rnn1_whh", C.c_void_p), ("rnn1_bih", C.c_void_p), ("rnn1_bhh", C.c_void_p),
        ("rnn2_wih", C.c_void_p), ("rnn2_whh", C.c_void_p), ("rnn2_bih", C.c_void_p), ("rnn2_bhh", C.c_void_p),
        ("fc1_w", C.c_void_p), ("fc1_b", C.c_void_p), ("fc2_w", C.c_void_p), ("fc2_b", C.c_void_p),
        ("fc3_w", C.c_void_p), ("fc3_b", C.c_void_p),
        ("n_res", C.c_int32), ("up_scales", C.c_int32 * 3),
    ]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build_oracle.build()))
        for name in ("twin_expf", "twin_logf", "twin_sigmoidf", "twin_tanhf"):
            getattr(_lib, name).restype = C.c_float
            getattr(_lib, name).argtypes = [C.c_float]
        _lib.twin_noise.restype = C.c_float
        _lib.twin_noise.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
    return _lib


class Twin:
    def __init__(self, sd: Dict[str, "np.ndarray"]):
        self._keep = {}

        def arr(name):
            a = np.ascontiguousarray(np.asarray(sd[name], dtype=np.float32))
            self._keep[name] = a
            return a.ctypes.data

        def bn(prefix):
            return (C.c_void_p * 4)(arr(prefix + ".weight"), arr(prefix + ".bias"), arr(prefix + ".running_mean"),
                                    arr(prefix + ".running_var"))

        w = TwinWeights()
        w.conv_in_w = arr("upsample.resnet.conv_in.weight")
        w.bn0 = bn("upsample.resnet.batch_norm")
        for i in range(10):
            w.res_conv1[i] = arr(f"upsample.resnet.layers.{i}.conv1.weight")
            w.res_conv2[i] = arr(f"upsample.resnet.layers.{i}.conv2.weight")
            w.res_bn1[i] = bn(f"upsample.resnet.layers.{i}.batch_norm1")
            w.res_bn2[i] = bn(f"upsample.resnet.layers.{i}.batch_norm2")
        w.conv_out_w, w.conv_out_b = arr("upsample.resnet.conv_out.weight"), arr("upsample.resnet.conv_out.bias")
        for j in range(3):
            w.up_w[j] = arr(f"upsample.up_layers.{2 * j + 1}.weight")
        w.I_w, w.I_b = arr("I.weight"), arr("I.bias")
        for n in ("rnn1", "rnn2"):
            setattr(w, n + "_wih", arr(n + ".weight_ih_l0"))
            setattr(w, n + "_whh", arr(n + ".weight_hh_l0"))
            setattr(w, n + "_bih", arr(n + ".bias_ih_l0"))
            setattr(w, n + "_bhh", arr(n + ".bias_hh_l0"))
        for n in ("fc1", "fc2", "fc3"):
            setattr(w, n + "_w", arr(n + ".weight"))
            setattr(w, n + "_b", arr(n + ".bias"))
        w.n_res = 10
        w.up_scales = (C.c_int32 * 3)(5, 5, 8)
        self.w = w

    def condition(self, mel: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """mel [80, T] (normalised) -> aux [T,128], melup [200T, 80]"""
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        T = mel.shape[1]
        aux = np.empty((T, 128), np.float32)
        melup = np.empty((200 * T, 80), np.float32)
        lib().twin_condition(C.byref(self.w), mel.ctypes.data_as(C.c_void_p), C.c_int32(T),
                             aux.ctypes.data_as(C.c_void_p), melup.ctypes.data_as(C.c_void_p))
        return aux, melup

    def generate(self, aux, melup, fold_starts, steps: int, noise: Optional[np.ndarray] = None, seed: int = 0,
                 want_logits: bool = False):
        T = aux.shape[0]
        fs = np.ascontiguousarray(fold_starts, dtype=np.int32)
        B = fs.shape[0]
        out = np.empty((B, steps), np.int16)
        logits = np.empty((B, 512), np.float32) if want_logits else None
        if noise is not None:
            noise = np.ascontiguousarray(noise, dtype=np.float32)
            assert noise.shape == (steps, B, 512)
        lib().twin_generate(C.byref(self.w), aux.ctypes.data_as(C.c_void_p), melup.ctypes.data_as(C.c_void_p),
                            C.c_int32(T), fs.ctypes.data_as(C.c_void_p), C.c_int32(B), C.c_int32(steps),
                            noise.ctypes.data_as(C.c_void_p) if noise is not None else None, C.c_uint64(seed),
                            out.ctypes.data_as(C.c_void_p),
                            logits.ctypes.data_as(C.c_void_p) if logits is not None else None)
        return (out, logits) if want_logits else out

    def teacher_forced(self, aux, melup, fold_starts, steps, noise, ref_idx):
        T = aux.shape[0]
        fs = np.ascontiguousarray(fold_starts, dtype=np.int32)
        B = fs.shape[0]
        out = np.empty((B, steps), np.int16)
        margin = np.empty((B, steps), np.float32)
        noise = np.ascontiguousarray(noise, dtype=np.float32)
        ref_idx = np.ascontiguousarray(ref_idx, dtype=np.int16)
        lib().twin_teacher_forced(C.byref(self.w), aux.ctypes.data_as(C.c_void_p), melup.ctypes.data_as(C.c_void_p),
                                  C.c_int32(T), fs.ctypes.data_as(C.c_void_p), C.c_int32(B), C.c_int32(steps),
                                  noise.ctypes.data_as(C.c_void_p), ref_idx.ctypes.data_as(C.c_void_p),
                                  out.ctypes.data_as(C.c_void_p), margin.ctypes.data_as(C.c_void_p))
        return out, margin


# ---- host-side post-processing restatement (float64 numpy, like the reference) -------------------
def fold_geometry(total_len: int, target: int, overlap: int) -> Tuple[int, np.ndarray]:
    """fold_with_overlap (:314-336): number of folds and their start offsets"""
    num_folds = (total_len - overlap) // (target + overlap)
    extended_len = num_folds * (overlap + target) + overlap
    remaining = total_len - extended_len
    if remaining != 0:
        num_folds += 1
    starts = np.arange(num_folds, dtype=np.int32) * (target + overlap)
    return num_folds, starts


def xfade_and_unfold(y: np.ndarray, target: int, overlap: int) -> np.ndarray:
    """(:340-402)"""
    y = y.astype(np.float64).copy()
    num_folds, length = y.shape
    target = length - 2 * overlap
    total_len = num_folds * (target + overlap) + overlap
    silence_len = overlap // 2
    fade_len = overlap - silence_len
    silence = np.zeros((silence_len), dtype=np.float64)
    t = np.linspace(-1, 1, fade_len, dtype=np.float64)
    fade_in = np.concatenate([silence, np.sqrt(0.5 * (1 + t))])
    fade_out = np.concatenate([np.sqrt(0.5 * (1 - t)), silence])
    y[:, :overlap] *= fade_in
    y[:, -overlap:] *= fade_out
    unfolded = np.zeros((total_len), dtype=np.float64)
    for i in range(num_folds):
        start = i * (target + overlap)
        unfolded[start:start + target + 2 * overlap] += y[i]
    return unfolded


def postprocess(idx: np.ndarray, frames: int, batched: bool, target: int, overlap: int, hp: dict) -> np.ndarray:
    """indices [B, steps] -> waveform exactly as WaveRNN.generate returns it (:236-257)."""
    from scipy.signal import lfilter

    n_classes = 2 ** hp["bits"]
    out = (2 * idx.astype(np.float32) / np.float32(n_classes - 1.0) - np.float32(1.0)).astype(np.float64)
    out = xfade_and_unfold(out, target, overlap) if batched else out[0]
    if hp["mu_law"]:
        mu = n_classes - 1
        out = np.sign(out) / mu * ((1 + mu) ** np.abs(out) - 1)
    if hp["apply_preemphasis"]:
        out = lfilter([1], [1, -hp["preemphasis"]], out)
    wave_len = (frames - 1) * hp["hop_length"]
    fade_out = np.linspace(1, 0, 20 * hp["hop_length"])
    out = out[:wave_len]
    out[-20 * hp["hop_length"]:] *= fade_out
    return out

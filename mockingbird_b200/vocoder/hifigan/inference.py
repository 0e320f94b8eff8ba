"""Drop-in for ``models.vocoder.hifigan.inference`` (reference: models/vocoder/hifigan/inference.py:22-73).

Same module-level protocol: ``load_model(weights_fpath, config_fpath=None, verbose=True)``,
``is_loaded()``, ``infer_waveform(mel, progress_callback=None) -> (wav, sample_rate)`` with module
globals ``generator``, ``_device``, ``output_sample_rate``.  Extension: ``infer_waveforms`` vocodes
a list of utterances as padded batches (the reference vocodes one utterance per call).
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np
import torch

from ... import _lib
from ...utils import AttrDict
from .models import DEFAULT_CONFIG_16K, Generator

generator = None  # type: Optional[Generator]
output_sample_rate = None
_device = None
_precision = os.environ.get("MOCKINGBIRD_B200_GAN_PRECISION", "auto")  # auto | f16tc | f16x3 | fp32 (see vocoder/_gan.py)


def set_precision(precision: str) -> None:
    """'auto' (default: f16tc if a load-time probe shows it within 5e-4 of f16x3 for the checkpoint, else f16x3), 'f16tc',
    'f16x3' or 'fp32'; takes effect at the next load_model."""
    global _precision
    if precision != "auto" and precision not in _lib.PRECISIONS:
        raise ValueError(precision)
    _precision = precision


def load_checkpoint(filepath, device):
    assert os.path.isfile(filepath)
    print("Loading '{}'".format(filepath))
    checkpoint_dict = torch.load(filepath, map_location=device)
    print("Complete.")
    return checkpoint_dict


def load_model(weights_fpath, config_fpath=None, verbose=True):
    global generator, _device, output_sample_rate

    if verbose:
        print("Building hifigan")
    weights_fpath = Path(weights_fpath)
    if config_fpath is None:
        model_config_fpaths = list(weights_fpath.parent.rglob("*.json"))
        if len(model_config_fpaths) > 0:
            config_fpath = model_config_fpaths[0]
    if config_fpath is not None:
        with open(config_fpath) as f:
            h = AttrDict(json.loads(f.read()))
    else:
        h = AttrDict(DEFAULT_CONFIG_16K)
    output_sample_rate = h.sampling_rate
    torch.manual_seed(h.seed)  # side effect kept: reference reseeds the global RNG here (inference.py:39)

    _device = _lib.require_cuda()
    generator = Generator(h, precision=_precision).to(_device)
    state_dict_g = load_checkpoint(weights_fpath, "cpu")
    generator.load_state_dict(state_dict_g["generator"])
    generator.eval()
    generator.remove_weight_norm()


def load_state(state_dict, h=None, precision: Optional[str] = None):
    """Install a generator from an in-memory ``ckpt['generator']`` dict (no file)."""
    global generator, _device, output_sample_rate
    h = AttrDict(h or DEFAULT_CONFIG_16K)
    output_sample_rate = h.sampling_rate
    _device = _lib.require_cuda()
    generator = Generator(h, precision=precision or _precision).to(_device)
    generator.load_state_dict(state_dict)
    generator.eval()
    generator.remove_weight_norm()
    return generator


def is_loaded():
    return generator is not None


def infer_waveform(mel, progress_callback=None):
    if generator is None:
        raise Exception("Please load hifi-gan in memory before using it")

    mel = torch.as_tensor(np.asarray(mel), dtype=torch.float32)
    mel = mel.to(_device, non_blocking=True).unsqueeze(0)
    with torch.no_grad():
        y_g_hat = generator(mel)
        audio = y_g_hat.squeeze()
    audio = audio.cpu().numpy()
    return audio, output_sample_rate


_pinned: dict = {}


def _pinned_buffer(name: str, numel: int) -> torch.Tensor:
    """grow-only pinned staging buffers (cudaHostAlloc per call costs more than the copies it serves)"""
    buf = _pinned.get(name)
    if buf is None or buf.numel() < numel:
        buf = torch.empty(max(numel, 1), dtype=torch.float32).pin_memory()
        _pinned[name] = buf
    return buf


def infer_waveforms(mels: Sequence[np.ndarray], batch_size: int = 32) -> List[np.ndarray]:
    """Vocode many utterances as length-sorted padded batches; each result equals the per-utterance
    call (padding is masked at every layer on the device).  All batches are enqueued back to back
    (pinned H2D -> forward -> pinned D2H) and the host waits once at the end."""
    if generator is None:
        raise Exception("Please load hifi-gan in memory before using it")
    order = sorted(range(len(mels)), key=lambda i: -mels[i].shape[1])
    out: List[Optional[np.ndarray]] = [None] * len(mels)
    hop = generator.hop
    batches = []
    n_in = n_out = 0
    for s in range(0, len(order), batch_size):
        idx = order[s:s + batch_size]
        tmax = max(mels[i].shape[1] for i in idx)
        batches.append((idx, tmax, n_in, n_out))
        n_in += len(idx) * 80 * tmax
        n_out += len(idx) * tmax * hop
    host_in = _pinned_buffer("in", n_in)
    host_out = _pinned_buffer("out", n_out)
    for idx, tmax, o_in, o_out in batches:
        if tmax == 0:
            continue
        hin = host_in[o_in:o_in + len(idx) * 80 * tmax].view(len(idx), 80, tmax)
        lens = torch.empty(len(idx), dtype=torch.int32)
        for r, i in enumerate(idx):
            t = mels[i].shape[1]
            hin[r, :, :t] = torch.as_tensor(np.asarray(mels[i]), dtype=torch.float32)
            if t < tmax:
                hin[r, :, t:] = 0.0
            lens[r] = t
        dev = hin.to(_device, non_blocking=True)
        wav = generator(dev, lengths=lens.to(_device))
        host_out[o_out:o_out + len(idx) * tmax * hop].view(len(idx), 1, tmax * hop).copy_(wav, non_blocking=True)
    torch.cuda.current_stream(_device).synchronize()
    block = np.array(host_out[:n_out].numpy(), copy=True)  # ONE copy out of the reused pinned staging; results are views of it
    for idx, tmax, o_in, o_out in batches:
        if tmax == 0:
            for i in idx:
                out[i] = np.zeros(0, np.float32)
            continue
        wav = block[o_out:o_out + len(idx) * tmax * hop].reshape(len(idx), tmax * hop)
        for r, i in enumerate(idx):
            out[i] = wav[r, : mels[i].shape[1] * hop]
    return out  # type: ignore[return-value]

"""Drop-in for ``models.vocoder.wavernn.inference`` (reference: models/vocoder/wavernn/inference.py:8-64).

``load_model(weights_fpath, verbose=True)`` (the toolbox also passes a config path positionally,
control/toolbox/__init__.py:471, hence ``*ignored``), ``is_loaded()``,
``infer_waveform(mel, normalize=True, batched=True, target=8000, overlap=800, progress_callback=None)``.
"""
from __future__ import annotations

import torch

from ... import _lib
from . import hparams as hp
from .models.fatchord_version import WaveRNN

_model = None  # type: WaveRNN
_device = None


def _build():
    return WaveRNN(
        rnn_dims=hp.voc_rnn_dims,
        fc_dims=hp.voc_fc_dims,
        bits=hp.bits,
        pad=hp.voc_pad,
        upsample_factors=hp.voc_upsample_factors,
        feat_dims=hp.num_mels,
        compute_dims=hp.voc_compute_dims,
        res_out_dims=hp.voc_res_out_dims,
        res_blocks=hp.voc_res_blocks,
        hop_length=hp.hop_length,
        sample_rate=hp.sample_rate,
        mode=hp.voc_mode,
    )


def load_model(weights_fpath, *ignored, verbose=True):
    global _model, _device

    if verbose:
        print("Building Wave-RNN")
    _model = _build()
    _device = _lib.require_cuda()
    _model = _model.cuda()
    if verbose:
        print("Loading model weights at %s" % weights_fpath)
    checkpoint = torch.load(weights_fpath, "cpu")
    _model.load_state_dict(checkpoint['model_state'])
    _model.eval()


def load_state(model_state, rng: str = "torch", seed: int = 0):
    """Install a model from an in-memory ``ckpt['model_state']`` dict (no file)."""
    global _model, _device
    _model = _build()
    _device = _lib.require_cuda()
    _model = _model.cuda()
    _model.load_state_dict(model_state)
    _model.eval()
    _model.rng = rng
    _model.seed = seed
    return _model


def is_loaded():
    return _model is not None


def infer_waveform(mel, normalize=True, batched=True, target=8000, overlap=800, progress_callback=None):
    """Infers the waveform of a mel spectrogram output by the synthesizer (inference.py:45-64)."""
    if _model is None:
        raise Exception("Please load Wave-RNN in memory before using it")

    if normalize:
        mel = mel / hp.mel_max_abs_value
    mel = torch.from_numpy(mel[None, ...])
    wav = _model.generate(mel, batched, target, overlap, hp.mu_law, progress_callback)
    return wav, hp.sample_rate

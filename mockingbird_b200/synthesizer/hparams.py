"""Synthesizer hyper-parameters (reference: models/synthesizer/hparams.py:3-78, utils/hparams.py:63-108):
the inference-relevant subset, same attribute names, same ``loadJson`` override hook."""
from __future__ import annotations

import json


class HParams:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.__dict__

    def keys(self):
        return self.__dict__.keys()

    def loadJson(self, fpath):
        """utils/hparams.py:91-101: override from a *.json found beside the checkpoint"""
        with open(fpath, "r", encoding="utf-8") as f:
            data = json.load(f)
        for k, v in data.items():
            if k not in ["tts_schedule", "tts_finetune_layers"]:
                self.__dict__[k] = v
        return self


hparams = HParams(
    sample_rate=16000, n_fft=1024, num_mels=80, hop_size=256, win_size=1024, fmin=55, min_level_db=-100,
    ref_level_db=20, max_abs_value=4., preemphasis=0.97, preemphasize=True,
    # mel front-end (models/synthesizer/hparams.py:16-70)
    frame_shift_ms=None, fmax=7600, signal_normalization=True, allow_clipping_in_normalization=True, symmetric_mels=True,
    use_lws=False, rescale=True, rescaling_max=0.9, power=1.5, griffin_lim_iters=60,
    tts_embed_dims=512, tts_encoder_dims=256, tts_decoder_dims=128, tts_postnet_dims=512, tts_encoder_K=5,
    tts_lstm_dims=1024, tts_postnet_K=5, tts_num_highways=4, tts_dropout=0.5, tts_cleaner_names=["basic_cleaners"],
    tts_stop_threshold=-3.4, synthesis_batch_size=16, speaker_embedding_size=256, use_gst=True, use_ser_for_gst=True,
)

# models/synthesizer/gst_hyperparameters.py
gst_hparams = HParams(E=512, ref_enc_filters=[32, 32, 64, 64, 128, 128], token_num=10, num_heads=8, n_mels=256)

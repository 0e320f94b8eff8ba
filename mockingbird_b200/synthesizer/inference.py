"""Drop-in for ``models.synthesizer.inference.Synthesizer`` (reference: models/synthesizer/inference.py:15-143).

Same class surface: ``Synthesizer(model_fpath, verbose=True)``, lazy ``load()``, ``is_loaded()``,
``synthesize_spectrograms(texts, embeddings, return_alignments=False, style_idx=0, min_stop_token=5,
steps=2000)``, class attributes ``sample_rate`` / ``hparams``.  The CPU audio front-end helpers
(load_preprocess_wav, make_spectrogram, griffin_lim, inference.py:144-181) are out of scope.
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Union

import numpy as np
import torch

from .. import _lib
from .hparams import hparams
from .models.tacotron import Tacotron
from .utils.symbols import symbols
from .utils.text import text_to_sequence

try:  # the reference romanises Chinese text with pypinyin (inference.py:100); optional here
    from pypinyin import Style, lazy_pinyin
except Exception:  # pragma: no cover
    lazy_pinyin = None
    Style = None


def pad1d(x, max_len, pad_value=0):
    return np.pad(x, (0, max_len - len(x)), mode="constant", constant_values=pad_value)


class Synthesizer:
    sample_rate = hparams.sample_rate
    hparams = hparams

    def __init__(self, model_fpath: Path, verbose=True):
        self.model_fpath = Path(model_fpath)
        self.verbose = verbose
        self.device = _lib.require_cuda()
        if self.verbose:
            print("Synthesizer using device:", self.device)
        self._model = None

    def is_loaded(self):
        return self._model is not None

    def _build(self) -> Tacotron:
        return Tacotron(embed_dims=hparams.tts_embed_dims, num_chars=len(symbols), encoder_dims=hparams.tts_encoder_dims,
                        decoder_dims=hparams.tts_decoder_dims, n_mels=hparams.num_mels, fft_bins=hparams.num_mels,
                        postnet_dims=hparams.tts_postnet_dims, encoder_K=hparams.tts_encoder_K,
                        lstm_dims=hparams.tts_lstm_dims, postnet_K=hparams.tts_postnet_K,
                        num_highways=hparams.tts_num_highways, dropout=hparams.tts_dropout,
                        stop_threshold=hparams.tts_stop_threshold,
                        speaker_embedding_size=hparams.speaker_embedding_size).to(self.device)

    def load(self):
        model_config_fpaths = list(self.model_fpath.parent.rglob("*.json"))
        if len(model_config_fpaths) > 0 and model_config_fpaths[0].exists():
            hparams.loadJson(model_config_fpaths[0])
        self._model = self._build()
        self._model.load(self.model_fpath, self.device)
        self._model.eval()
        if self.verbose:
            print("Loaded synthesizer \"%s\" trained to step %d" % (self.model_fpath.name, self._model.get_step()))

    def load_state(self, model_state):
        """install weights from an in-memory ``ckpt['model_state']`` dict (no file)"""
        self._model = self._build()
        self._model.load_state_dict(model_state, strict=False)
        self._model.eval()
        return self._model

    def synthesize_spectrograms(self, texts: List[str], embeddings: Union[np.ndarray, List[np.ndarray]],
                                return_alignments=False, style_idx=0, min_stop_token=5, steps=2000):
        if not self.is_loaded():
            self.load()
        if lazy_pinyin is not None:
            texts = [" ".join(lazy_pinyin(v, style=Style.TONE3, neutral_tone_with_five=True)) for v in texts]
        inputs = [text_to_sequence(text, hparams.tts_cleaner_names) for text in texts]
        return self.synthesize_from_sequences(inputs, embeddings, return_alignments, style_idx, min_stop_token, steps)

    @staticmethod
    def load_preprocess_wav(fpath):
        """inference.py:143-158 (host utility): load at the synthesizer rate and rescale.  The reference additionally
        runs the third-party `logmmse` denoiser when the clip is longer than 0.4 s; it is applied here only when that
        optional package is importable."""
        from . import audio_host

        wav = audio_host.load_wav(fpath, hparams.sample_rate)[0]
        if hparams.rescale:
            wav = wav / np.abs(wav).max() * hparams.rescaling_max
        if len(wav) > hparams.sample_rate * (0.3 + 0.1):
            try:
                import logmmse

                noise_wav = np.concatenate([wav[:int(hparams.sample_rate * 0.15)], wav[-int(hparams.sample_rate * 0.15):]])
                wav = logmmse.denoise(wav, logmmse.profile_noise(noise_wav, hparams.sample_rate))
            except ImportError:
                pass
        return wav

    @staticmethod
    def make_spectrogram(fpath_or_wav):
        """inference.py:160-172: mel spectrogram [80, M] as fed to the synthesizer in training (GPU front-end,
        synthesizer/audio.py)"""
        from . import audio

        wav = Synthesizer.load_preprocess_wav(fpath_or_wav) if isinstance(fpath_or_wav, (str, Path)) else fpath_or_wav
        return audio.melspectrogram(wav, hparams).astype(np.float32)

    @staticmethod
    def griffin_lim(mel):
        """inference.py:174-180: invert a mel spectrogram with Griffin-Lim (host, numpy/torch-CPU)"""
        from . import audio_host

        return audio_host.inv_mel_spectrogram(mel, hparams)

    def synthesize_from_sequences(self, inputs, embeddings, return_alignments=False, style_idx=0, min_stop_token=5,
                                  steps=2000, dropout_masks=None):
        """the part of synthesize_spectrograms after the text front-end (inference.py:104-142)"""
        if not isinstance(embeddings, list):
            embeddings = [embeddings]
        bs = hparams.synthesis_batch_size
        batched_inputs = [inputs[i:i + bs] for i in range(0, len(inputs), bs)]
        batched_embeds = [embeddings[i:i + bs] for i in range(0, len(embeddings), bs)]
        specs = []
        alignments = None
        for i, batch in enumerate(batched_inputs, 1):
            if self.verbose:
                print(f"\n| Generating {i}/{len(batched_inputs)}")
            max_text_len = max(len(text) for text in batch)
            chars = np.stack([pad1d(text, max_text_len) for text in batch])
            speaker_embeds = np.stack(batched_embeds[i - 1])
            chars = torch.tensor(chars).long()
            speaker_embeddings = torch.tensor(speaker_embeds).float()
            _, mels, alignments = self._model.generate(chars, speaker_embeddings, style_idx=style_idx,
                                                       min_stop_token=min_stop_token, steps=steps,
                                                       dropout_masks=dropout_masks)
            mels = mels.detach().cpu().numpy()
            for m in mels:
                # Trim silence from end of each spectrogram (inference.py:135-138)
                while m.shape[1] > 0 and np.max(m[:, -1]) < hparams.tts_stop_threshold:
                    m = m[:, :-1]
                specs.append(m)
        if self.verbose:
            print("\n\nDone.\n")
        return (specs, alignments) if return_alignments else specs

"""On-disk formats either side of the vocoder path, read-only:

* ``train.txt`` / ``synthesized.txt``: one utterance per line, pipe separated -
  ``wav_fname|mel_fname|embed_fname|n_samples|n_mel_frames|text`` (models/synthesizer/preprocess_audio.py:83,
  preprocess.py:89; read back by vocoder_dataset.py:13-19 and synthesizer_dataset.py:13, which keep the rows whose
  frame count is non-zero);
* mel files: ``np.save`` of float32 ``[n_frames, 80]`` at the synthesizer's +-4 scale (preprocess_audio.py:79,
  synthesize.py:86-93) - transposed to ``[80, n_frames]`` for the vocoders (vocoder_dataset.py:28 divides by
  ``mel_max_abs_value`` for WaveRNN; wavernn.inference.infer_waveform does the same).
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Iterable, List, Sequence

import numpy as np


@dataclass
class MetaRow:
    wav_fname: str
    mel_fname: str
    embed_fname: str
    n_samples: int
    n_frames: int
    text: str


def read_metadata(metadata_fpath) -> List[MetaRow]:
    """parse train.txt / synthesized.txt; rows with n_frames == 0 are dropped like the reference datasets do"""
    rows = []
    with Path(metadata_fpath).open("r", encoding="utf-8") as f:
        for line in f:
            x = line.rstrip("\n").split("|")
            if len(x) < 5:
                continue
            n_frames = int(x[4])
            if not n_frames:
                continue
            rows.append(MetaRow(x[0], x[1], x[2], int(float(x[3])) if x[3] else 0, n_frames, "|".join(x[5:])))
    return rows


def load_mel(mel_fpath) -> np.ndarray:
    """``[n_frames, 80]`` .npy -> float32 ``[80, n_frames]`` (what infer_waveform takes)"""
    mel = np.load(mel_fpath, allow_pickle=False)
    if mel.ndim != 2:
        raise ValueError(f"{mel_fpath}: expected a 2-D mel array, got shape {mel.shape}")
    return np.ascontiguousarray(mel.T.astype(np.float32))


def vocode_files(mel_fpaths: Sequence, vocoder, batch_size: int = 32) -> Iterable[np.ndarray]:
    """vocode .npy mels with a loaded GAN vocoder module (hifigan / fregan inference): batched through
    ``infer_waveforms`` when the module has it, else one call per file"""
    mels = [load_mel(p) for p in mel_fpaths]
    if hasattr(vocoder, "infer_waveforms"):
        return vocoder.infer_waveforms(mels, batch_size=batch_size)
    return [vocoder.infer_waveform(m)[0] for m in mels]

"""GPU parity tests of the mel front-ends against the numpy-float64 restatement of
librosa's algorithm (oracle/melspec_oracle.py - parity unpinned, see its header).

Tolerances: the kernel computes the DFT directly in fp32 (relative error ~1e-6 per bin);
  encoder mel (power, linear): <= 2e-4 of the spectrogram's maximum;
  synthesizer mel (dB, normalised to +-4): <= 2e-3 absolute (one unit = 12.5 dB)."""
import numpy as np
import pytest

import melspec_oracle as mo

pytestmark = pytest.mark.gpu


def _wav(n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    return (0.3 * np.sin(2 * np.pi * 220 * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)


@pytest.mark.parametrize("mode", ["reflect", "constant"])
@pytest.mark.parametrize("n", [201, 1600, 16000, 48123])
def test_encoder_mel(n, mode):
    from mockingbird_b200.encoder import audio

    audio.pad_mode = mode
    try:
        wav = _wav(n, n)
        got = audio.wav_to_mel_spectrogram(wav)
        ref = mo.encoder_mel(wav, mode)
        assert got.shape == ref.shape == (1 + n // 160, 40) and got.dtype == np.float32
        assert np.abs(got - ref).max() <= 2e-4 * ref.max()
    finally:
        audio.pad_mode = "reflect"


@pytest.mark.parametrize("mode", ["reflect", "constant"])
@pytest.mark.parametrize("n", [513, 4000, 32000])
def test_synthesizer_mel(n, mode):
    from mockingbird_b200.synthesizer import audio
    from mockingbird_b200.synthesizer.hparams import hparams

    audio.pad_mode = mode
    try:
        wav = _wav(n, n + 1)
        got = audio.melspectrogram(wav, hparams)
        hp = dict(mo.SYNTH_HP, n_fft=hparams.n_fft, hop_size=hparams.hop_size, win_size=hparams.win_size, fmin=hparams.fmin,
                  fmax=hparams.fmax, num_mels=hparams.num_mels)
        ref = mo.synthesizer_mel(wav, hp, mode)
        assert got.shape == ref.shape == (80, 1 + n // hp["hop_size"]) and got.dtype == np.float32
        assert np.abs(got - ref).max() <= 2e-3
        assert got.min() >= -4.0 and got.max() <= 4.0
    finally:
        audio.pad_mode = "reflect"


def test_embed_utterance_from_wav_end_to_end():
    """wav -> 40-mel frames -> partial slicing -> LSTM encoder: the whole embed_utterance surface on the device path"""
    import torch

    import encoder_oracle as eo
    import ref_init as ri
    from mockingbird_b200.encoder import inference as enc
    from mockingbird_b200.encoder.model import SpeakerEncoder

    m = SpeakerEncoder()
    m.load_state_dict(ri.encoder_state_dict(0))
    m.eval()
    enc.set_model(m)
    wav = _wav(16000 * 3, 7)
    e, pe, ws = enc.embed_utterance(wav, return_partials=True)
    assert e.shape == (256,) and abs(float(np.linalg.norm(e)) - 1.0) < 1e-5 and pe.shape[0] == len(ws) == 3
    # oracle chain on the same signal
    wave_slices, mel_slices = enc.compute_partial_slices(len(wav))
    wpad = np.pad(wav, (0, max(0, wave_slices[-1].stop - len(wav))), "constant")
    frames = mo.encoder_mel(wpad)
    ref = eo.embed_utterance_partials(ri.encoder_state_dict(0), torch.from_numpy(np.stack([frames[s] for s in mel_slices])))
    assert np.abs(e - ref).max() < 2e-4


def test_short_signal_is_rejected():
    from mockingbird_b200 import _lib
    from mockingbird_b200.encoder import audio

    with pytest.raises(_lib.MbError, match="reflect padding"):
        audio.wav_to_mel_spectrogram(np.zeros(100, np.float32))

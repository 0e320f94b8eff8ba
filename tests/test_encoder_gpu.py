"""GPU parity tests of the speaker-encoder path (drop-in surface -> C ABI) against the golden vectors
generated from the live reference and against the CPU oracle on other shapes.

Tolerance: FP32 arithmetic; differences come from summation order and expf/tanhf only.  Embeddings are
unit vectors; asserted at 1e-4 max-abs (north-star bar for float outputs is 1e-3 relative)."""
import numpy as np
import pytest
import torch

import encoder_oracle as eo
import ref_init as ri

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def enc():
    from mockingbird_b200.encoder import inference as enc_inf
    from mockingbird_b200.encoder.model import SpeakerEncoder

    m = SpeakerEncoder()
    m.load_state_dict(ri.encoder_state_dict(0))
    m.eval()
    enc_inf.set_model(m)
    return enc_inf


@pytest.mark.parametrize("name", ["a", "b"])
def test_embed_frames_batch_matches_golden(enc, golden_dir, name):
    z = np.load(golden_dir / "encoder_seed0.npz")
    got = enc.embed_frames_batch(z[f"{name}_frames"])
    assert got.shape == z[f"{name}_embeds"].shape and got.dtype == np.float32
    assert np.abs(got - z[f"{name}_embeds"]).max() < TOL


def test_utterance_embedding_matches_golden(enc, golden_dir):
    z = np.load(golden_dir / "encoder_seed0.npz")
    got = enc.embed_utterances_frames([z["a_frames"], z["a_frames"][:2]])
    assert np.abs(got[0] - z["a_utterance"]).max() < TOL
    raw = z["a_embeds"][:2].mean(0)
    assert np.abs(got[1] - raw / np.linalg.norm(raw)).max() < TOL
    # the single-utterance surface from frames: 3 s of audio -> 3 partials with 50 % overlap
    frames = np.concatenate([z["a_frames"][0], z["a_frames"][1], z["a_frames"][2]])[:400]
    e, pe, ws = enc.embed_utterance_frames(frames, wav_len=48000, return_partials=True)
    assert pe.shape == (3, 256) and len(ws) == 3
    ref = eo.embed_utterance_partials(ri.encoder_state_dict(0), torch.from_numpy(np.stack([frames[s] for s in
                                                                                           (slice(0, 160), slice(80, 240), slice(160, 320))])))
    assert np.abs(e - ref).max() < TOL


def test_large_batch_matches_oracle(enc):
    """cfg-5 shape (many utterances x 6 partials x 160 frames), rows > one library call's chunk"""
    g = torch.Generator().manual_seed(11)
    frames = torch.rand(1030, 160, 40, generator=g) * 0.3
    got = enc.embed_frames_batch(frames.numpy())
    idx = [0, 1, 511, 1023, 1024, 1029]
    ref = eo.embed_frames(ri.encoder_state_dict(0), frames[idx]).numpy()
    assert np.abs(got[idx] - ref).max() < TOL


def test_not_loaded_error():
    from mockingbird_b200.encoder import inference as enc_inf

    saved = enc_inf._model
    enc_inf._model = None
    try:
        with pytest.raises(Exception, match="Model was not loaded"):
            enc_inf.embed_frames_batch(np.zeros((1, 160, 40), np.float32))
    finally:
        enc_inf._model = saved


def test_reupload_invalidates_lazily_packed_images():
    """the whole-sequence input-projection images (packed at first use with >= 512 rows) must be re-packed
    after .to() / a second load_state_dict"""
    from mockingbird_b200.encoder.model import SpeakerEncoder

    sd = ri.encoder_state_dict(0)
    frames = torch.rand(8, 160, 40, generator=torch.Generator().manual_seed(3)) * 0.3
    m = SpeakerEncoder()
    m.load_state_dict(sd)
    a = m.forward(frames).cpu()
    m.to(torch.device("cuda", torch.cuda.current_device()))
    b = m.forward(frames).cpu()
    assert torch.equal(a, b)
    sd2 = {k: (v * 1.25 if "weight_ih" in k else v) for k, v in sd.items()}
    m.load_state_dict(sd2)
    c = m.forward(frames).cpu()
    ref = eo.embed_frames(sd2, frames)
    assert float((c - ref).abs().max()) < TOL and not torch.equal(a, c)

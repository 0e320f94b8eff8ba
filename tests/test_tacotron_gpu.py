"""GPU parity tests of the Tacotron path (drop-in surface -> C ABI) against the golden vectors generated
from the live reference (dropout masks injected) and against the CPU oracle on other shapes.

Tolerance: all arithmetic is FP32; differences come from summation order only.  mel / postnet outputs
within 1e-3 relative (BASELINE.json north_star) - asserted at 2e-4 max-norm relative here."""
import numpy as np
import pytest
import torch

import ref_init as ri
import tacotron_oracle as to

pytestmark = pytest.mark.gpu
TOL = 2e-4


@pytest.fixture(scope="module")
def sd():
    return ri.tacotron_state_dict(0, r=2, randomize_bn=True)


@pytest.fixture(scope="module")
def model(sd):
    from mockingbird_b200.synthesizer.inference import Synthesizer

    s = Synthesizer("unused.pt", verbose=False)
    return s.load_state(sd)


def _masks(z, name):
    enc = np.unpackbits(z[f"{name}_enc_masks"], axis=-1)
    dec = np.unpackbits(z[f"{name}_dec_masks"], axis=-1)
    return torch.from_numpy(enc), torch.from_numpy(dec.reshape(-1, 2, dec.shape[-2], dec.shape[-1]))


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("name", ["a", "b"])
def test_golden(golden_dir, model, name):
    """a: style_idx=-1 (speaker-conditioned GST attention), b: fixed style token 3; ragged char lengths"""
    z = np.load(golden_dir / "tacotron_seed0.npz")
    chars, emb = torch.from_numpy(z[f"{name}_chars"]), torch.from_numpy(z[f"{name}_emb"])
    steps, style, mst, r = [int(v) for v in z[f"{name}_cfg"]]
    assert model.r == r
    mel, lin, attn = model.generate(chars, emb, steps=steps, style_idx=style, min_stop_token=mst,
                                    dropout_masks=_masks(z, name))
    for got, key in ((mel, "mel"), (lin, "linear"), (attn, "attn")):
        ref = torch.from_numpy(z[f"{name}_{key}"])
        assert got.shape == ref.shape, key
        assert _rel(got.cpu(), ref) <= TOL, (key, _rel(got.cpu(), ref))


def test_vs_oracle_longer_and_early_stop(sd, model):
    """B=5, 33 chars, 60 steps with the reference's own min_stop_token=4: the random-weight stop
    rule fires early - frame count and values must match the oracle"""
    g = torch.Generator().manual_seed(123)
    B, Tc, steps = 5, 33, 60
    chars = torch.randint(2, 75, (B, Tc), generator=g)
    chars[1, 20:] = 0
    chars[3, 7:] = 0
    emb = torch.rand(B, 256, generator=g)
    emb = emb / emb.norm(dim=1, keepdim=True)
    nst = steps // 2
    enc = (torch.rand(2, B, Tc, 256, generator=g) < 0.5)
    dec = (torch.rand(nst, 2, B, 256, generator=g) < 0.5)
    masks = [enc[0], enc[1]] + [dec[i, j] for i in range(nst) for j in range(2)]
    for mst in (10, 4):
        mel_r, lin_r, attn_r = to.generate(sd, chars, emb, steps, 0, mst, masks, r=2)
        mel, lin, attn = model.generate(chars, emb, steps=steps, style_idx=0, min_stop_token=mst, dropout_masks=(enc, dec))
        assert mel.shape == mel_r.shape, (mst, mel.shape, mel_r.shape)
        assert _rel(mel.cpu(), mel_r) <= TOL and _rel(lin.cpu(), lin_r) <= TOL and _rel(attn.cpu(), attn_r) <= TOL
    assert mel.shape[2] < steps  # the min_stop_token=4 run stopped early


def test_large_batch_tensor_core_cbhg(sd, model):
    """B = 12 x 48 chars (576 encoder rows) x 96 frames (1152 postnet rows): the CBHG convolutions, highways and GRU
    input projections take the large-M tensor-core route (3-term fp16 split) - same FP32-level tolerance vs the oracle"""
    g = torch.Generator().manual_seed(321)
    B, Tc, steps = 12, 48, 96
    chars = torch.randint(2, 75, (B, Tc), generator=g)
    chars[2, 30:] = 0
    chars[7, 11:] = 0
    emb = torch.rand(B, 256, generator=g)
    emb = emb / emb.norm(dim=1, keepdim=True)
    nst = steps // 2
    enc = (torch.rand(2, B, Tc, 256, generator=g) < 0.5)
    dec = (torch.rand(nst, 2, B, 256, generator=g) < 0.5)
    masks = [enc[0], enc[1]] + [dec[i, j] for i in range(nst) for j in range(2)]
    mel_r, lin_r, attn_r = to.generate(sd, chars, emb, steps, -1, 10, masks, r=2)
    mel, lin, attn = model.generate(chars, emb, steps=steps, style_idx=-1, min_stop_token=10, dropout_masks=(enc, dec))
    assert mel.shape == mel_r.shape == (B, 80, steps)
    assert _rel(mel.cpu(), mel_r) <= TOL, _rel(mel.cpu(), mel_r)
    assert _rel(lin.cpu(), lin_r) <= TOL, _rel(lin.cpu(), lin_r)
    assert _rel(attn.cpu(), attn_r) <= TOL, _rel(attn.cpu(), attn_r)


def test_device_dropout_is_seeded(model):
    chars = torch.randint(2, 75, (2, 10), generator=torch.Generator().manual_seed(1))
    emb = torch.rand(2, 256, generator=torch.Generator().manual_seed(2))
    model.seed = 5
    a, _, _ = model.generate(chars, emb, steps=8, style_idx=0, min_stop_token=10)
    b, _, _ = model.generate(chars, emb, steps=8, style_idx=0, min_stop_token=10)
    model.seed = 6
    c, _, _ = model.generate(chars, emb, steps=8, style_idx=0, min_stop_token=10)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert torch.isfinite(a).all()


def test_synthesizer_surface(tmp_path, sd):
    """Synthesizer class protocol from a checkpoint file: lazy load, batching by 16, trailing trim"""
    from mockingbird_b200.synthesizer.inference import Synthesizer

    torch.save({"model_state": sd}, tmp_path / "taco.pt")
    s = Synthesizer(tmp_path / "taco.pt", verbose=False)
    assert not s.is_loaded() and s.sample_rate == 16000
    texts = ["hello world", "a b c"] * 9  # 18 -> two batches (16 + 2)
    embeds = [np.full(256, 1 / 16.0, np.float32)] * 18
    specs = s.synthesize_spectrograms(texts, embeds, steps=6, min_stop_token=10)
    assert s.is_loaded() and len(specs) == 18
    assert all(m.shape[0] == 80 and m.shape[1] <= 6 and m.dtype == np.float32 for m in specs)
    specs2, align = s.synthesize_spectrograms(texts[:2], embeds[:2], return_alignments=True, steps=6, min_stop_token=10)
    assert len(specs2) == 2 and align.shape[0] == 2


def test_reupload_invalidates_lazily_packed_images(sd):
    """after one large-M generate (CBHG tensor-core images packed on first use) a re-upload - .to(device) or a
    second load_state_dict with DIFFERENT weights - must not reuse the stale images / scales"""
    from mockingbird_b200.synthesizer.inference import Synthesizer

    g = torch.Generator().manual_seed(77)
    B, Tc, steps = 12, 48, 64
    chars = torch.randint(2, 75, (B, Tc), generator=g)
    emb = torch.rand(B, 256, generator=g)
    emb = emb / emb.norm(dim=1, keepdim=True)
    nst = steps // 2
    enc = (torch.rand(2, B, Tc, 256, generator=g) < 0.5)
    dec = (torch.rand(nst, 2, B, 256, generator=g) < 0.5)
    m = Synthesizer("unused.pt", verbose=False).load_state(sd)
    a = m.generate(chars, emb, steps=steps, style_idx=-1, min_stop_token=10, dropout_masks=(enc, dec))[1].cpu()
    m.to(torch.device("cuda", torch.cuda.current_device()))  # forces a fresh arena
    b = m.generate(chars, emb, steps=steps, style_idx=-1, min_stop_token=10, dropout_masks=(enc, dec))[1].cpu()
    assert torch.equal(a, b)
    sd2 = {k: (v * 1.5 if k.endswith("conv.weight") and v.dtype == torch.float32 else v) for k, v in sd.items()}
    m.load_state_dict(sd2)
    c = m.generate(chars, emb, steps=steps, style_idx=-1, min_stop_token=10, dropout_masks=(enc, dec))[1].cpu()
    fresh = Synthesizer("unused.pt", verbose=False).load_state(sd2)
    d = fresh.generate(chars, emb, steps=steps, style_idx=-1, min_stop_token=10, dropout_masks=(enc, dec))[1].cpu()
    assert torch.equal(c, d) and not torch.equal(a, c)


def test_batch_128_rows(sd, model):
    """B = 128 (the full 128-row tensor-core tile; the cfg-5 pipeline batches 128 utterances per generate): same tolerance vs
    the oracle on a short run"""
    g = torch.Generator().manual_seed(4242)
    B, Tc, steps = 128, 24, 12
    chars = torch.randint(2, 75, (B, Tc), generator=g)
    for b in range(0, B, 3):
        chars[b, 8 + (b % 11):] = 0
    emb = torch.rand(B, 256, generator=g)
    emb = emb / emb.norm(dim=1, keepdim=True)
    nst = steps // 2
    enc = (torch.rand(2, B, Tc, 256, generator=g) < 0.5)
    dec = (torch.rand(nst, 2, B, 256, generator=g) < 0.5)
    masks = [enc[0], enc[1]] + [dec[i, j] for i in range(nst) for j in range(2)]
    mel_r, lin_r, attn_r = to.generate(sd, chars, emb, steps, -1, 10, masks, r=2)
    mel, lin, attn = model.generate(chars, emb, steps=steps, style_idx=-1, min_stop_token=10, dropout_masks=(enc, dec))
    assert mel.shape == mel_r.shape == (B, 80, steps)
    assert _rel(mel.cpu(), mel_r) <= TOL and _rel(lin.cpu(), lin_r) <= TOL and _rel(attn.cpu(), attn_r) <= TOL

"""GPU parity tests of the GAN vocoder path: CUDA library (through the C ABI / the drop-in Python
surfaces) vs the CPU oracle and the committed golden vectors.

Tolerances (BASELINE.json north_star: "within 1e-3 relative for HiFi-GAN float waveforms"):
  both  max_rel = |y - ref|_inf / |ref|_inf   and   rms_rel = rms(y - ref) / rms(ref)
  fp32  path : <= 2e-5   (FFMA, differs from the oracle only by summation order)
  f16tc path : <= 1e-3   fp16 operands in the resblocks, 3-term split on the serial layers
  f16x3 path : <= 1e-4   3-term fp16 split on every layer (FP32-equivalent on the tensor cores)
  auto       : <= 1e-3   the drop-in default: f16tc when a load-time probe shows it within 5e-4 of f16x3, else f16x3
No assertion in this file is looser than the north star's 1e-3.
"""
import numpy as np
import pytest
import torch

import gan_oracle as go
import ref_init as ri

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-5, "f16tc": 1e-3, "f16x3": 1e-4, "auto": 1e-3}
PRECISIONS = ["fp32", "f16tc", "f16x3"]


def _hifigan(precision, sd=None, cfg=None):
    from mockingbird_b200.vocoder.hifigan.models import Generator

    cfg = cfg or ri.HIFIGAN_CONFIG_16K
    g = Generator(cfg, precision=precision).cuda()
    g.load_state_dict(sd if sd is not None else ri.hifigan_state_dict(cfg, 0))
    g.eval()
    g.remove_weight_norm()
    return g


@pytest.fixture(scope="module")
def sd0():
    return ri.hifigan_state_dict(ri.HIFIGAN_CONFIG_16K, 0)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_hifigan_golden_small(golden_dir, sd0, precision):
    z = np.load(golden_dir / "hifigan_seed0.npz")
    g = _hifigan(precision, sd0)
    wav = g(torch.from_numpy(z["mel_small"]).cuda()).cpu()
    e = go.rel_errors(wav, torch.from_numpy(z["wav_small"]))
    assert e["max_rel"] <= TOL[precision] and e["rms_rel"] <= TOL[precision], e


@pytest.mark.parametrize("precision", PRECISIONS)
def test_hifigan_cfg2_full_size_rows(golden_dir, sd0, precision):
    """BASELINE.json configs[1]: batch 32 x 256 frames; rows 0 and 31 checked against the golden
    reference output, all rows against a size-independent property (batch invariance)."""
    z = np.load(golden_dir / "hifigan_seed0.npz")
    g = _hifigan(precision, sd0)
    mel = torch.rand(32, 80, 256, generator=torch.Generator().manual_seed(2)) * 8 - 4
    wav = g(mel.cuda())
    assert wav.shape == (32, 1, 51200)
    pick = [int(i) for i in z["full_pick"]]
    e = go.rel_errors(wav[pick].cpu(), torch.from_numpy(z["wav_full"]))
    assert e["max_rel"] <= TOL[precision] and e["rms_rel"] <= TOL[precision], e
    # batch invariance: the same utterance alone gives the same samples
    alone = g(mel[5:6].cuda())
    assert torch.equal(alone[0], wav[5])


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("B,T", [(1, 1), (1, 7), (3, 33), (2, 130)])
def test_hifigan_vs_oracle_shapes(sd0, precision, B, T):
    """edge cases: single frame, ragged tile sizes"""
    g = _hifigan(precision, sd0)
    mel = torch.rand(B, 80, T, generator=torch.Generator().manual_seed(100 + T)) * 8 - 4
    with torch.no_grad():
        ref = go.hifigan_forward(sd0, ri.HIFIGAN_CONFIG_16K, mel)
    wav = g(mel.cuda()).cpu()
    e = go.rel_errors(wav, ref)
    assert e["max_rel"] <= TOL[precision] and e["rms_rel"] <= TOL[precision], e


@pytest.mark.parametrize("precision", PRECISIONS)
def test_hifigan_ragged_batch_equals_per_utterance(sd0, precision):
    """variable-length batch: padding is masked at every layer, each row equals its own batch-1 call"""
    g = _hifigan(precision, sd0)
    lens = [40, 17, 33, 1]
    mel = torch.full((4, 80, 40), -4.0)
    for i, t in enumerate(lens):
        mel[i, :, :t] = torch.rand(80, t, generator=torch.Generator().manual_seed(i)) * 8 - 4
    wav = g(mel.cuda(), lengths=torch.tensor(lens, dtype=torch.int32)).cpu()
    for i, t in enumerate(lens):
        with torch.no_grad():
            ref = go.hifigan_forward(sd0, ri.HIFIGAN_CONFIG_16K, mel[i:i + 1, :, :t])
        e = go.rel_errors(wav[i:i + 1, :, : t * 200], ref)
        assert e["max_rel"] <= TOL[precision], (i, e)
        assert float(wav[i, :, t * 200:].abs().max()) == 0.0 if t < 40 else True


def test_hifigan_empty_batch(sd0):
    g = _hifigan("fp32", sd0)
    assert g(torch.zeros(0, 80, 5).cuda()).shape == (0, 1, 1000)
    assert g(torch.zeros(2, 80, 0).cuda()).shape == (2, 1, 0)


@pytest.mark.parametrize("precision", ["fp32", "f16x3", "auto"])
def test_hifigan_rescaled_init(precision):
    """second, harder initialisation (activations O(1) through the stack, like a trained checkpoint).  fp16 operand
    rounding alone does not hold 1e-3 there (2.4e-3 max), so: the FP32-equivalent tensor-core mode must,
    and the drop-in default (`auto`) must DETECT the case at load time and select it."""
    sd = ri.rescale_variance_preserving(ri.hifigan_state_dict(ri.HIFIGAN_CONFIG_16K, 0), 1.0)
    g = _hifigan(precision, sd)
    mel = torch.rand(2, 80, 64, generator=torch.Generator().manual_seed(9)) * 8 - 4
    with torch.no_grad():
        ref = go.hifigan_forward(sd, ri.HIFIGAN_CONFIG_16K, mel)
    e = go.rel_errors(g(mel.cuda()).cpu(), ref)
    if precision == "fp32":
        assert e["max_rel"] <= 5e-5 and e["rms_rel"] <= 2e-5, e
    else:
        assert e["max_rel"] <= TOL[precision] and e["rms_rel"] <= TOL[precision], e
    if precision == "auto":
        assert g.precision == "f16x3" and g.calibration["max_rel"] > g.AUTO_TOLERANCE, g.calibration


def test_auto_precision_selects_fast_mode_on_reference_init(sd0):
    """on the reference's own initialisation (the BASELINE configs) the probe shows f16tc well inside the tolerance: the
    drop-in default runs the fast mode; the global torch RNG is not consumed by the calibration"""
    torch.manual_seed(123)
    g = _hifigan("auto", sd0)
    assert g.precision == "f16tc" and g.calibration["selected"] == "f16tc", g.calibration
    assert g.calibration["max_rel"] <= g.AUTO_TOLERANCE and g.calibration["rms_rel"] <= g.AUTO_TOLERANCE
    assert torch.equal(torch.rand(3), torch.rand(3, generator=torch.Generator().manual_seed(123)))


def test_hifigan_resblock2_variant():
    cfg = dict(ri.HIFIGAN_CONFIG_16K, resblock="2", resblock_dilation_sizes=[[1, 3], [1, 3], [1, 3]],
               upsample_initial_channel=128)
    torch.manual_seed(3)
    sd = {}
    C0 = 128
    sd["conv_pre.weight"], sd["conv_pre.bias"] = torch.randn(C0, 80, 7) * 0.05, torch.randn(C0) * 0.05
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        sd[f"ups.{i}.weight"] = torch.randn(C0 >> i, C0 >> (i + 1), k) * 0.05
        sd[f"ups.{i}.bias"] = torch.randn(C0 >> (i + 1)) * 0.05
        for j, kk in enumerate(cfg["resblock_kernel_sizes"]):
            for m in range(2):
                ch = C0 >> (i + 1)
                sd[f"resblocks.{i * 3 + j}.convs.{m}.weight"] = torch.randn(ch, ch, kk) * 0.03
                sd[f"resblocks.{i * 3 + j}.convs.{m}.bias"] = torch.randn(ch) * 0.05
    sd["conv_post.weight"], sd["conv_post.bias"] = torch.randn(1, C0 >> 4, 7) * 0.1, torch.randn(1) * 0.05
    g = _hifigan("fp32", sd, cfg)
    mel = torch.rand(2, 80, 21, generator=torch.Generator().manual_seed(1)) * 8 - 4
    with torch.no_grad():
        ref = go.hifigan_forward(sd, cfg, mel)
    e = go.rel_errors(g(mel.cuda()).cpu(), ref)
    assert e["max_rel"] <= 2e-5, e


def test_hifigan_weight_norm_checkpoint_keys(sd0):
    """load_state_dict accepts the checkpoint's weight_g / weight_v layout (hifigan/inference.py:51)"""
    raw = {}
    for k, v in sd0.items():
        if k.endswith(".weight"):
            base = k[: -len(".weight")]
            norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
            raw[base + ".weight_v"] = v * 3.0
            raw[base + ".weight_g"] = norm
        else:
            raw[k] = v
    g = _hifigan("fp32", raw)
    mel = torch.rand(1, 80, 9, generator=torch.Generator().manual_seed(4)) * 8 - 4
    with torch.no_grad():
        ref = go.hifigan_forward(sd0, ri.HIFIGAN_CONFIG_16K, mel)
    assert go.rel_errors(g(mel.cuda()).cpu(), ref)["max_rel"] <= 2e-5


def test_hifigan_inference_module_protocol(tmp_path, sd0):
    """the reference's module-level protocol (hifigan/inference.py:22-73) end to end from files"""
    import json

    from mockingbird_b200.vocoder.hifigan import inference as gan_vocoder

    gan_vocoder.generator = None
    with pytest.raises(Exception, match="Please load hifi-gan"):
        gan_vocoder.infer_waveform(np.zeros((80, 4), np.float32))
    torch.save({"generator": sd0}, tmp_path / "g_hifigan.pt")
    (tmp_path / "config.json").write_text(json.dumps(ri.HIFIGAN_CONFIG_16K))
    gan_vocoder.set_precision("fp32")
    gan_vocoder.load_model(tmp_path / "g_hifigan.pt", verbose=False)
    assert gan_vocoder.is_loaded() and gan_vocoder.output_sample_rate == 16000
    assert torch.initial_seed() == 1234  # load_model reseeds the global RNG like the reference
    mel = (torch.rand(80, 30, generator=torch.Generator().manual_seed(8)) * 8 - 4).numpy()
    wav, sr = gan_vocoder.infer_waveform(mel)
    assert sr == 16000 and wav.shape == (6000,) and wav.dtype == np.float32
    with torch.no_grad():
        ref = go.hifigan_forward(sd0, ri.HIFIGAN_CONFIG_16K, torch.from_numpy(mel)[None])
    assert go.rel_errors(torch.from_numpy(wav)[None, None], ref)["max_rel"] <= 2e-5
    many = gan_vocoder.infer_waveforms([mel[:, :11], mel, mel[:, :25]], batch_size=2)
    assert [w.shape[0] for w in many] == [2200, 6000, 5000]
    assert np.array_equal(many[1], wav)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_fregan_vs_golden_and_oracle(golden_dir, precision):
    from mockingbird_b200.vocoder.fregan.models import FreGAN

    z = np.load(golden_dir / "fregan_seed0.npz")
    sd = ri.fregan_state_dict(ri.FREGAN_CONFIG, 0)
    g = FreGAN(ri.FREGAN_CONFIG, precision=precision).cuda()
    g.load_state_dict(sd)
    g.eval()
    g.remove_weight_norm()
    wav = g(torch.from_numpy(z["mel"]).cuda()).cpu()
    e = go.rel_errors(wav, torch.from_numpy(z["wav"]))
    assert e["max_rel"] <= TOL[precision] and e["rms_rel"] <= TOL[precision], e

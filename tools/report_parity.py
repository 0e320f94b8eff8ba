"""Parity numbers at the BASELINE configuration sizes against the live-reference fixtures (the same comparisons tests/ assert),
printed as one JSON document for profiles/."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "synth_weights"))
import numpy as np, torch
import gan_oracle as go, ref_init as ri
from mockingbird_b200.vocoder.hifigan.models import Generator
from mockingbird_b200.vocoder.fregan.models import FreGAN
G = ROOT / "tests" / "golden"
out = {}
mel = torch.rand(32, 80, 256, generator=torch.Generator().manual_seed(2)) * 8 - 4
for name, cls, cfg, sdf, gold in (("hifigan_cfg2", Generator, ri.HIFIGAN_CONFIG_16K, ri.hifigan_state_dict, "hifigan_seed0.npz"),
                                  ("fregan_cfg2", FreGAN, ri.FREGAN_CONFIG, ri.fregan_state_dict, "fregan_cfg2.npz")):
    z = np.load(G / gold)
    pick = [int(i) for i in z["full_pick"]]
    ref = torch.from_numpy(z["wav_full"])
    for prec in ("auto", "f16tc", "f16x3", "fp32"):
        g = cls(cfg, precision=prec).cuda(); g.load_state_dict(sdf(cfg, 0)); g.eval(); g.remove_weight_norm()
        e = go.rel_errors(g(mel.cuda())[pick].cpu(), ref)
        out[f"{name}/{prec}"] = {**{k: float(f"{v:.3e}") for k, v in e.items()}, "selected": g.precision, "calibration": g.calibration}
# harder init (activations O(1))
sd = ri.rescale_variance_preserving(ri.hifigan_state_dict(ri.HIFIGAN_CONFIG_16K, 0), 1.0)
m2 = torch.rand(2, 80, 64, generator=torch.Generator().manual_seed(9)) * 8 - 4
with torch.no_grad():
    ref2 = go.hifigan_forward(sd, ri.HIFIGAN_CONFIG_16K, m2)
for prec in ("auto", "f16tc", "f16x3", "fp32"):
    g = Generator(ri.HIFIGAN_CONFIG_16K, precision=prec).cuda(); g.load_state_dict(sd); g.eval(); g.remove_weight_norm()
    e = go.rel_errors(g(m2.cuda()).cpu(), ref2)
    out[f"hifigan_variance_preserving_init/{prec}"] = {**{k: float(f"{v:.3e}") for k, v in e.items()}, "selected": g.precision, "calibration": g.calibration}
# Tacotron cfg 4
from mockingbird_b200.synthesizer.inference import Synthesizer
z = np.load(G / "tacotron_cfg4.npz")
model = Synthesizer("unused.pt", verbose=False).load_state(ri.tacotron_state_dict(0, r=2, randomize_bn=True))
enc = torch.from_numpy(np.unpackbits(z["enc_masks"], axis=-1)); dec = np.unpackbits(z["dec_masks"], axis=-1)
dec = torch.from_numpy(dec.reshape(-1, 2, dec.shape[-2], dec.shape[-1]))
model.r = 2
mel_o, lin, attn = model.generate(torch.from_numpy(z["chars"].astype(np.int64)), torch.from_numpy(z["emb"]), steps=400, style_idx=-1, min_stop_token=10, dropout_masks=(enc, dec))
rows = [int(v) for v in z["rows"]]
out["tacotron_cfg4"] = {"mel_max_rel": float((mel_o.cpu()[rows] - torch.from_numpy(z["mel"])).abs().max()) / float(z["mel_absmax"]),
                        "postnet_max_rel": float((lin.cpu()[rows] - torch.from_numpy(z["linear"])).abs().max()) / float(z["linear_absmax"]),
                        "attention_max_abs": float((attn.cpu()[rows] - torch.from_numpy(z["attn"])).abs().max())}
print(json.dumps(out, indent=1))

"""Speaker-encoder golden vectors from the LIVE reference (needs /root/reference; run in the build
container):  python oracle/make_golden_encoder.py  ->  tests/golden/encoder_seed0.npz

Weights: torch.manual_seed(0); SpeakerEncoder(cpu, cpu) (checked equal to ref_init.encoder_state_dict(0)).
Inputs: non-negative "mel power" frames rand(seed) * 0.2 of the shapes the inference path feeds
(partials of 160 frames x 40 channels), plus the utterance-level reduction of inference.py:160-166."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "synth_weights"))
import ref_harness as rh  # noqa: E402
import ref_init as ri  # noqa: E402


def main():
    rh.install()
    model = rh.build_encoder(seed=0)
    sd = ri.encoder_state_dict(0)
    ref_sd = model.state_dict()
    for k, v in ref_sd.items():
        assert torch.equal(v, sd[k]), f"ref_init.encoder_state_dict drifted from the reference constructor at {k}"
    out = {}
    g = torch.Generator().manual_seed(50)
    frames_a = torch.rand(5, 160, 40, generator=g) * 0.2
    frames_b = torch.rand(3, 37, 40, generator=g) * 0.05
    with torch.no_grad():
        out["a_frames"], out["a_embeds"] = frames_a.numpy(), model.forward(frames_a).numpy()
        out["b_frames"], out["b_embeds"] = frames_b.numpy(), model.forward(frames_b).numpy()
    # utterance-level: inference.embed_utterance's reduction over the partial embeddings of `a`
    raw = np.mean(out["a_embeds"], axis=0)
    out["a_utterance"] = raw / np.linalg.norm(raw, 2)
    out["torch_version"] = np.array(torch.__version__)
    dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "encoder_seed0.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()

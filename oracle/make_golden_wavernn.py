"""WaveRNN golden vectors from the LIVE reference (called by oracle/make_golden.py)."""
from __future__ import annotations

import numpy as np
import torch

import ref_harness as rh
import ref_init as ri


def _run(model, fv, mel, batched, target, overlap, seed):
    cap = {}
    orig = fv.decode_mu_law

    def wrap(y, mu, from_labels=True):
        cap["y"] = np.array(y, copy=True)
        return orig(y, mu, from_labels)

    fv.decode_mu_law = wrap
    try:
        torch.manual_seed(seed)
        wav = model.generate(mel, batched, target, overlap, True, progress_callback=lambda *a: None)
    finally:
        fv.decode_mu_law = orig
    return wav, cap["y"]


def main(golden_dir, meta):
    rh.install()
    import models.vocoder.wavernn.models.fatchord_version as fv

    model = rh.build_wavernn(seed=0)
    sd = ri.wavernn_state_dict(0, randomize_bn=True)
    model.load_state_dict(sd)
    # (1) unbatched, 27 frames -> 5400 sequential samples (cfg-1 shaped: batched=False)
    mel1 = torch.rand(1, 80, 27, generator=torch.Generator().manual_seed(1)) * 2 - 1
    wav1, y1 = _run(model, fv, mel1, False, 8000, 400, 1234)
    idx1 = np.rint((y1 + 1) * 511 / 2).astype(np.int16)[None]
    noise1 = ri.wavernn_noise(1234, 1, 128).numpy()
    # (2) batched, 30 frames, target 1000 / overlap 100 -> 6 folds x 1200 steps (cfg-3 shaped)
    mel2 = torch.rand(1, 80, 30, generator=torch.Generator().manual_seed(3)) * 2 - 1
    # the unfolded float array hides the per-fold samples: capture them by wrapping xfade_and_unfold
    cap = {}
    orig_x = model.xfade_and_unfold

    def wrap_x(y, target, overlap):
        cap["folds"] = np.array(y, copy=True)
        return orig_x(y, target, overlap)

    model.xfade_and_unfold = wrap_x
    wav2, _ = _run(model, fv, mel2, True, 1000, 100, 1234)
    idx2 = np.rint((cap["folds"] + 1) * 511 / 2).astype(np.int16)
    np.savez_compressed(golden_dir / "wavernn_seed0.npz", mel1=mel1.numpy(), idx1=idx1, wav1=wav1, noise1_head=noise1,
                        mel2=mel2.numpy(), idx2=idx2, wav2=wav2,
                        meta=meta(weights="ref_init.wavernn_state_dict(0, randomize_bn=True)", gen_seed=1234,
                                  case1="rand(1,80,27;seed 1)*2-1, batched=False",
                                  case2="rand(1,80,30;seed 3)*2-1, batched=True target=1000 overlap=100"))

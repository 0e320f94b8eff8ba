"""Tacotron golden vectors from the LIVE reference (called by oracle/make_golden.py): the reference's
PreNet dropout masks are captured by wrapping torch.nn.functional.dropout in the harness
(reference files untouched), SURVEY.md appendix C item 4."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

import ref_harness as rh
import ref_init as ri


def capture_generate(model, chars, embeds, steps, style_idx, min_stop_token):
    masks = []
    orig = F.dropout

    def wrapped(x, p=0.5, training=True, inplace=False):
        out = orig(x, p, training, inplace)
        masks.append((out != 0) | (x == 0))  # where x == 0 the mask is irrelevant; record "keep"
        return out

    F.dropout = wrapped
    try:
        with torch.no_grad():
            mel, linear, attn = model.generate(chars, embeds, steps=steps, style_idx=style_idx,
                                               min_stop_token=min_stop_token)
    finally:
        F.dropout = orig
    return mel, linear, attn, masks


def make_inputs(B, Tc, seed):
    g = torch.Generator().manual_seed(seed)
    chars = torch.randint(2, 75, (B, Tc), generator=g)
    lens = torch.randint(max(2, Tc // 3), Tc + 1, (B,), generator=g)
    lens[0] = Tc
    for b in range(B):
        chars[b, lens[b]:] = 0
    emb = torch.rand(B, 256, generator=torch.Generator().manual_seed(seed + 1))
    emb = emb / emb.norm(dim=1, keepdim=True)
    return chars, emb, lens


def pack_masks(masks):
    """encoder masks [B,Tc,256] x2, then decoder masks [B,256] x 2 per step -> two uint8 arrays"""
    enc = np.stack([m.numpy().astype(np.uint8) for m in masks[:2]])
    dec = np.stack([m.numpy().astype(np.uint8) for m in masks[2:]])
    return np.packbits(enc, axis=-1), np.packbits(dec, axis=-1)


def main(golden_dir, meta):
    rh.install()
    model = rh.build_tacotron(seed=0)
    sd = ri.tacotron_state_dict(0, r=2, randomize_bn=True)
    model.load_state_dict(sd, strict=True)
    model.eval()
    out = {}
    for name, (B, Tc, steps, style) in {"a": (3, 14, 24, -1), "b": (2, 9, 16, 3)}.items():
        chars, emb, lens = make_inputs(B, Tc, seed=40 + len(out))
        torch.manual_seed(77)
        mel, linear, attn, masks = capture_generate(model, chars, emb, steps, style, 10)
        enc_m, dec_m = pack_masks(masks)
        out.update({f"{name}_chars": chars.numpy(), f"{name}_emb": emb.numpy(), f"{name}_mel": mel.numpy(),
                    f"{name}_linear": linear.numpy(), f"{name}_attn": attn.numpy(), f"{name}_enc_masks": enc_m,
                    f"{name}_dec_masks": dec_m, f"{name}_cfg": np.array([steps, style, 10, 2])})
    np.savez_compressed(golden_dir / "tacotron_seed0.npz", **out,
                        meta=meta(weights="ref_init.tacotron_state_dict(0, r=2, randomize_bn=True)",
                                  cases="a: B=3,Tc=14,steps=24,style_idx=-1; b: B=2,Tc=9,steps=16,style_idx=3; "
                                        "min_stop_token=10, r=2, dropout masks captured (packbits)"))

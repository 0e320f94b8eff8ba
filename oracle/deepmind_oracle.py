"""torch-CPU restatement of the DeepMind-style dual-softmax WaveRNN (TEST INFRASTRUCTURE).

Restates models/vocoder/wavernn/models/deepmind_version.py:75-162 (`WaveRNN.generate`) and :36-72 (`forward`).  The
reference file cannot be imported as shipped (it star-imports `utils.display` / `utils.dsp`, which do not exist in the
tree, and calls `.cuda()` unconditionally); oracle/ref_harness.load_deepmind() imports it UNMODIFIED behind stub modules that
provide the five names it takes from those imports (time, np, stream, combine_signal + a CPU no-op `.cuda()`), and
tests/test_oracle_pinned.py::test_deepmind_oracle_matches_reference pins this restatement to it: identical coarse / fine
integers under the same seed.  combine_signal: wavernn/audio.py:34-35.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

HIDDEN, QUANT = 896, 256
SPLIT = HIDDEN // 2


def combine_signal(coarse, fine):
    return coarse * 256 + fine - 2 ** 15


def generate(sd: Dict[str, torch.Tensor], seq_len: int, noise: Optional[torch.Tensor] = None):
    """-> (output int [seq_len], coarse [seq_len], fine [seq_len]).  Sampling = Categorical(softmax(logits)).sample() from the
    global torch generator (argmax(p/q), q = exponential_(1) over the 256 classes: coarse draw first, then fine), or from an
    injected `noise` [seq_len, 2, 256] of Exp(1) draws (tests)."""
    R, O1w, O1b, O2w, O2b = sd["R.weight"], sd["O1.weight"], sd["O1.bias"], sd["O2.weight"], sd["O2.bias"]
    O3w, O3b, O4w, O4b = sd["O3.weight"], sd["O3.bias"], sd["O4.weight"], sd["O4.bias"]
    Ic, If = sd["I_coarse.weight"], sd["I_fine.weight"]
    b_cu, b_fu = torch.split(sd["bias_u"], SPLIT)
    b_cr, b_fr = torch.split(sd["bias_r"], SPLIT)
    b_ce, b_fe = torch.split(sd["bias_e"], SPLIT)

    def draw(logits, i, which):
        p = F.softmax(logits, dim=1)
        if noise is None:
            return torch.distributions.Categorical(p).sample()
        p = p / p.sum(-1, keepdim=True)  # Categorical renormalises its probs
        return torch.argmax(p / noise[i, which].unsqueeze(0), dim=1)

    c_out, f_out = [], []
    out_coarse = torch.LongTensor([0])
    out_fine = torch.LongTensor([0])
    hidden = torch.zeros(1, HIDDEN)
    with torch.no_grad():
        for i in range(seq_len):
            hidden_coarse, hidden_fine = torch.split(hidden, SPLIT, dim=1)
            oc = out_coarse.unsqueeze(0).float() / 127.5 - 1.
            of = out_fine.unsqueeze(0).float() / 127.5 - 1.
            prev = torch.cat([oc, of], dim=1)
            Icu, Icr, Ice = torch.split(F.linear(prev, Ic), SPLIT, dim=1)
            Rh = F.linear(hidden, R)
            Rcu, Rfu, Rcr, Rfr, Rce, Rfe = torch.split(Rh, SPLIT, dim=1)
            u = torch.sigmoid(Rcu + Icu + b_cu)
            r = torch.sigmoid(Rcr + Icr + b_cr)
            e = torch.tanh(r * Rce + Ice + b_ce)
            hidden_coarse = u * hidden_coarse + (1. - u) * e
            out_coarse = draw(F.linear(F.relu(F.linear(hidden_coarse, O1w, O1b)), O2w, O2b), i, 0)
            c_out.append(out_coarse)
            cp = out_coarse.float() / 127.5 - 1.
            fin = torch.cat([prev, cp.unsqueeze(0)], dim=1)
            Ifu, Ifr, Ife = torch.split(F.linear(fin, If), SPLIT, dim=1)
            u = torch.sigmoid(Rfu + Ifu + b_fu)
            r = torch.sigmoid(Rfr + Ifr + b_fr)
            e = torch.tanh(r * Rfe + Ife + b_fe)
            hidden_fine = u * hidden_fine + (1. - u) * e
            out_fine = draw(F.linear(F.relu(F.linear(hidden_fine, O3w, O3b)), O4w, O4b), i, 1)
            f_out.append(out_fine)
            hidden = torch.cat([hidden_coarse, hidden_fine], dim=1)
    coarse = torch.stack(c_out).squeeze(1).numpy()
    fine = torch.stack(f_out).squeeze(1).numpy()
    return combine_signal(coarse, fine), coarse, fine

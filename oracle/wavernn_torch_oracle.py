"""torch-CPU restatement of the fatchord WaveRNN (TEST INFRASTRUCTURE: the CPU baseline leg of bench.py and an
extra pin of the oracle; never imported by the product).

Follows the reference call for call so that it costs what the reference costs on the host cores:
  UpsampleNetwork / MelResNet / Stretch2d   fatchord_version.py:9-85   (eval-mode BatchNorm)
  WaveRNN.generate                          fatchord_version.py:153-257
  get_gru_cell                              :265-271  (nn.GRUCell built from the GRU layer's weights; consumes the
                                            global RNG exactly like the reference's two constructions)
  pad_tensor / fold_with_overlap            :273-338
Post-processing is wavernn_oracle.postprocess (numpy float64, :236-253).  Pinned by
tests/test_fullsize.py::test_torch_oracle_prefix_matches_reference against the cfg-1 / cfg-3 goldens generated
from the live reference (identical integer samples under torch.manual_seed)."""
from __future__ import annotations

import time
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HOP_TOTAL = 200  # prod(upsample_factors) (5,5,8)
PAD = 2


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.1, 1e-5)


def upsample(sd: Dict[str, torch.Tensor], m: torch.Tensor, scales=(5, 5, 8)):
    """m [1, 80, T+2*pad] -> (mels [1, 200T, 80], aux [1, 200T, 128])  (:60-85)"""
    x = F.conv1d(m, sd["upsample.resnet.conv_in.weight"])
    x = F.relu(_bn(x, sd, "upsample.resnet.batch_norm"))
    i = 0
    while f"upsample.resnet.layers.{i}.conv1.weight" in sd:
        p = f"upsample.resnet.layers.{i}"
        r = x
        x = F.conv1d(x, sd[p + ".conv1.weight"])
        x = F.relu(_bn(x, sd, p + ".batch_norm1"))
        x = F.conv1d(x, sd[p + ".conv2.weight"])
        x = _bn(x, sd, p + ".batch_norm2") + r
        i += 1
    aux = F.conv1d(x, sd["upsample.resnet.conv_out.weight"], sd["upsample.resnet.conv_out.bias"])
    aux = aux.repeat_interleave(HOP_TOTAL, dim=2)  # Stretch2d(total_scale, 1) on [B,1,C,T]
    m = m.unsqueeze(1)
    for j, s in enumerate(scales):
        m = m.repeat_interleave(s, dim=3)
        m = F.conv2d(m, sd[f"upsample.up_layers.{2 * j + 1}.weight"], padding=(0, s))
    indent = PAD * HOP_TOTAL
    m = m.squeeze(1)[:, :, indent:-indent]
    return m.transpose(1, 2), aux.transpose(1, 2)


def fold_with_overlap(x, target, overlap):
    """(:288-338)"""
    _, total_len, features = x.size()
    num_folds = (total_len - overlap) // (target + overlap)
    extended_len = num_folds * (overlap + target) + overlap
    remaining = total_len - extended_len
    if remaining != 0:
        num_folds += 1
        padding = target + 2 * overlap - remaining
        x = torch.cat([x, torch.zeros(1, padding, features)], dim=1)
    folded = torch.zeros(num_folds, target + 2 * overlap, features)
    for i in range(num_folds):
        start = i * (target + overlap)
        end = start + target + 2 * overlap
        folded[i] = x[:, start:end, :]
    return folded


def _gru_cell(sd, name, in_dim, hid):
    cell = nn.GRUCell(in_dim, hid)  # consumes the global RNG like get_gru_cell (:265-271)
    cell.weight_hh.data = sd[name + ".weight_hh_l0"]
    cell.weight_ih.data = sd[name + ".weight_ih_l0"]
    cell.bias_hh.data = sd[name + ".bias_hh_l0"]
    cell.bias_ih.data = sd[name + ".bias_ih_l0"]
    return cell


def generate_indices(sd: Dict[str, torch.Tensor], mels: torch.Tensor, batched: bool, target: int, overlap: int,
                     max_steps: Optional[int] = None):
    """the loop of WaveRNN.generate; returns (int16 [folds, steps_run], seconds spent in the sample loop)"""
    rnn1 = _gru_cell(sd, "rnn1", 512, 512)
    rnn2 = _gru_cell(sd, "rnn2", 544, 512)
    with torch.no_grad():
        m = F.pad(mels, (PAD, PAD))  # pad_tensor(side='both') on the time axis (:273-286)
        mel_up, aux = upsample(sd, m)
        if batched:
            mel_up = fold_with_overlap(mel_up, target, overlap)
            aux = fold_with_overlap(aux, target, overlap)
        b, seq_len, _ = mel_up.size()
        h1 = torch.zeros(b, 512)
        h2 = torch.zeros(b, 512)
        x = torch.zeros(b, 1)
        aux_split = [aux[:, :, 32 * i:32 * (i + 1)] for i in range(4)]
        n = seq_len if max_steps is None else min(seq_len, max_steps)
        out = []
        t0 = time.perf_counter()
        for i in range(n):
            m_t = mel_up[:, i, :]
            a1, a2, a3, a4 = (a[:, i, :] for a in aux_split)
            x = torch.cat([x, m_t, a1], dim=1)
            x = F.linear(x, sd["I.weight"], sd["I.bias"])
            h1 = rnn1(x, h1)
            x = x + h1
            h2 = rnn2(torch.cat([x, a2], dim=1), h2)
            x = x + h2
            x = F.relu(F.linear(torch.cat([x, a3], dim=1), sd["fc1.weight"], sd["fc1.bias"]))
            x = F.relu(F.linear(torch.cat([x, a4], dim=1), sd["fc2.weight"], sd["fc2.bias"]))
            logits = F.linear(x, sd["fc3.weight"], sd["fc3.bias"])
            posterior = F.softmax(logits, dim=1)
            s = torch.distributions.Categorical(posterior).sample()
            out.append(s)
            x = (2 * s.float() / 511. - 1.).unsqueeze(-1)
        dt = time.perf_counter() - t0
    return torch.stack(out).transpose(0, 1).numpy().astype(np.int16), dt

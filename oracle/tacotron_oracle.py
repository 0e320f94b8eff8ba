"""CPU oracle of Tacotron inference (TEST INFRASTRUCTURE, not product code).

torch-CPU fp32 restatement, against torch.nn.functional only, of
  Tacotron.forward / generate     models/synthesizer/models/tacotron.py:199-298
  Encoder / Decoder.forward       tacotron.py:11-44, :71-138
  CBHG, BatchNormConv, Highway    models/sublayer/cbhg.py:42-79, common/batch_norm_conv.py:11-14,
                                  common/highway_network.py:12-17
  LSA                             models/sublayer/lsa.py:21-42
  PreNet (dropout ALWAYS on)      models/sublayer/pre_net.py:11-27
  GlobalStyleToken / STL / MHA    models/sublayer/global_style_token.py:9-145
working from a flat state_dict.  The PreNet dropout masks are INJECTED (``masks``: list of bool
tensors in call order - 2 for the encoder PreNet, then 2 per decoder step), which is what makes a
float parity statement possible; oracle/make_golden_tacotron.py captures them from
the live reference by wrapping torch.nn.functional.dropout in the harness.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.1, 1e-5)


def _bnconv(x, sd, p, k, relu=True):
    y = F.conv1d(x, sd[p + ".conv.weight"], None, 1, k // 2)
    if relu:
        y = F.relu(y)
    return _bn(y, sd, p + ".bnorm")


def _gru_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """one direction of nn.GRU(batch_first), zero initial state; x [B,T,C] -> [B,T,H]"""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = torch.zeros(B, H)
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        gi = F.linear(x[:, t], w_ih, b_ih)
        gh = F.linear(h, w_hh, b_hh)
        i_r, i_z, i_n = gi.chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        # ATen gru_cell (RNN.cpp): r = s(h_r + i_r), z = s(h_z + i_z), n = tanh(i_n + h_n * r), h' = (h - n) * z + n
        r = torch.sigmoid(h_r + i_r)
        z = torch.sigmoid(h_z + i_z)
        n = torch.tanh(i_n + h_n * r)
        h = (h - n) * z + n
        outs[t] = h
    return torch.stack(outs, 1)


def cbhg(x: torch.Tensor, sd: Dict[str, torch.Tensor], p: str, K: int) -> torch.Tensor:
    """x [B, C_in, T] -> [B, T, channels]"""
    residual = x
    T = x.size(-1)
    bank = [_bnconv(x, sd, f"{p}.conv1d_bank.{i}", k)[:, :, :T] for i, k in enumerate(range(1, K + 1))]
    y = torch.cat(bank, dim=1)
    y = F.max_pool1d(y, 2, 1, 1)[:, :, :T]
    y = _bnconv(y, sd, f"{p}.conv_project1", 3)
    y = _bnconv(y, sd, f"{p}.conv_project2", 3, relu=False)
    y = y + residual
    y = y.transpose(1, 2)
    if f"{p}.pre_highway.weight" in sd:
        y = F.linear(y, sd[f"{p}.pre_highway.weight"])
    i = 0
    while f"{p}.highways.{i}.W1.weight" in sd:
        x1 = F.linear(y, sd[f"{p}.highways.{i}.W1.weight"], sd[f"{p}.highways.{i}.W1.bias"])
        x2 = F.linear(y, sd[f"{p}.highways.{i}.W2.weight"], sd[f"{p}.highways.{i}.W2.bias"])
        g = torch.sigmoid(x2)
        y = g * F.relu(x1) + (1.0 - g) * y
        i += 1
    fw = _gru_dir(y, sd[f"{p}.rnn.weight_ih_l0"], sd[f"{p}.rnn.weight_hh_l0"], sd[f"{p}.rnn.bias_ih_l0"],
                  sd[f"{p}.rnn.bias_hh_l0"], False)
    bw = _gru_dir(y, sd[f"{p}.rnn.weight_ih_l0_reverse"], sd[f"{p}.rnn.weight_hh_l0_reverse"],
                  sd[f"{p}.rnn.bias_ih_l0_reverse"], sd[f"{p}.rnn.bias_hh_l0_reverse"], True)
    return torch.cat([fw, bw], dim=2)


def _prenet(x, sd, p, masks: List[torch.Tensor]):
    x = F.relu(F.linear(x, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"]))
    x = x * masks.pop(0).to(x.dtype) * 2.0  # F.dropout(p=0.5, training=True): keep * 1/(1-p)
    x = F.relu(F.linear(x, sd[p + ".fc2.weight"], sd[p + ".fc2.bias"]))
    x = x * masks.pop(0).to(x.dtype) * 2.0
    return x


def gst_constant_encoder_output(sd) -> torch.Tensor:
    """ReferenceEncoder applied to an all-zero [N,1,256] input (tacotron.py:251): input independent."""
    out = torch.zeros(1, 1, 1, 256)
    i = 0
    while f"gst.encoder.convs.{i}.weight" in sd:
        out = F.conv2d(out, sd[f"gst.encoder.convs.{i}.weight"], sd[f"gst.encoder.convs.{i}.bias"], 2, 1)
        p = f"gst.encoder.bns.{i}"
        out = F.batch_norm(out, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                           False, 0.1, 1e-5)
        out = F.relu(out)
        i += 1
    out = out.transpose(1, 2).contiguous().view(1, 1, -1)
    h = _gru_dir(out, sd["gst.encoder.gru.weight_ih_l0"], sd["gst.encoder.gru.weight_hh_l0"],
                 sd["gst.encoder.gru.bias_ih_l0"], sd["gst.encoder.gru.bias_hh_l0"], False)
    return h[:, -1]  # [1, 256]


def _mha(query, key, sd, num_heads=8, key_dim=64):
    q = F.linear(query, sd["gst.stl.attention.W_query.weight"])
    k = F.linear(key, sd["gst.stl.attention.W_key.weight"])
    v = F.linear(key, sd["gst.stl.attention.W_value.weight"])
    split = q.shape[-1] // num_heads
    q = torch.stack(torch.split(q, split, dim=2), dim=0)
    k = torch.stack(torch.split(k, split, dim=2), dim=0)
    v = torch.stack(torch.split(v, split, dim=2), dim=0)
    scores = torch.matmul(q, k.transpose(2, 3)) / (key_dim ** 0.5)
    scores = F.softmax(scores, dim=3)
    out = torch.matmul(scores, v)
    return torch.cat(torch.split(out, 1, dim=0), dim=3).squeeze(0)


def style_embedding(sd, speaker_embedding: torch.Tensor, style_idx: int) -> torch.Tensor:
    """[B, 1, 512] (or [1,1,512] for a fixed token) as tacotron.py:238-253 computes it in eval mode"""
    if 0 <= style_idx < 10:
        query = torch.zeros(1, 1, 512)
        key = torch.tanh(sd["gst.stl.embed"])[style_idx].unsqueeze(0).expand(1, -1, -1)
        return _mha(query, key, sd)
    enc = gst_constant_encoder_output(sd).expand(speaker_embedding.shape[0], -1)
    q = torch.cat([enc, speaker_embedding], dim=-1).unsqueeze(1)
    keys = torch.tanh(sd["gst.stl.embed"]).unsqueeze(0).expand(speaker_embedding.shape[0], -1, -1)
    return _mha(q, keys, sd)


def encoder_outputs(sd, chars, speaker_embedding, style_idx, masks) -> Tuple[torch.Tensor, torch.Tensor]:
    x = F.embedding(chars, sd["encoder.embedding.weight"])
    x = _prenet(x, sd, "encoder.pre_net", masks)
    enc = cbhg(x.transpose(1, 2), sd, "encoder.cbhg", 5)
    B, Tc, _ = enc.shape
    seq = torch.cat([enc, speaker_embedding.unsqueeze(1).expand(B, Tc, -1)], dim=2)
    style = style_embedding(sd, speaker_embedding, style_idx).expand(B, Tc, -1)
    seq = torch.cat([seq, style], dim=2)
    proj = F.linear(seq, sd["encoder_proj.weight"])
    return seq, proj


def _lstm_cell(x, h, c, sd, p):
    # ATen lstm_cell: gates = linear_hh(h) + linear_ih(x); c' = f*c + i*g; h' = o * tanh(c')
    g = F.linear(h, sd[p + ".weight_hh"], sd[p + ".bias_hh"]) + F.linear(x, sd[p + ".weight_ih"], sd[p + ".bias_ih"])
    i, f, gg, o = g.chunk(4, 1)
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    h = torch.sigmoid(o) * torch.tanh(c)
    return h, c


def generate(sd: Dict[str, torch.Tensor], chars: torch.Tensor, speaker_embedding: torch.Tensor, steps: int,
             style_idx: int, min_stop_token: float, masks: List[torch.Tensor], r: Optional[int] = None):
    """-> (mel_outputs [B,80,M], postnet/linear [B,80,M], attn [B, M/r, Tc]); consumes ``masks``."""
    with torch.no_grad():
        masks = list(masks)
        r = int(sd["decoder.r"]) if r is None else r
        B = chars.shape[0]
        seq, proj = encoder_outputs(sd, chars, speaker_embedding, style_idx, masks)
        Tc = seq.shape[1]
        attn_h = torch.zeros(B, 128)
        h1 = torch.zeros(B, 1024)
        h2 = torch.zeros(B, 1024)
        c1 = torch.zeros(B, 1024)
        c2 = torch.zeros(B, 1024)
        ctx = torch.zeros(B, seq.shape[2])
        cumulative = torch.zeros(B, Tc)
        char_mask = (chars != 0).float()
        prenet_in = torch.zeros(B, 80)
        mel_out, attn_out = [], []
        for t in range(0, steps, r):
            p = _prenet(prenet_in, sd, "decoder.prenet", masks)
            gi = F.linear(torch.cat([ctx, p], dim=-1), sd["decoder.attn_rnn.weight_ih"], sd["decoder.attn_rnn.bias_ih"])
            gh = F.linear(attn_h, sd["decoder.attn_rnn.weight_hh"], sd["decoder.attn_rnn.bias_hh"])
            i_r, i_z, i_n = gi.chunk(3, 1)
            h_r, h_z, h_n = gh.chunk(3, 1)
            rg, zg = torch.sigmoid(h_r + i_r), torch.sigmoid(h_z + i_z)
            ng = torch.tanh(i_n + h_n * rg)
            attn_h = (attn_h - ng) * zg + ng
            # LSA
            pq = F.linear(attn_h, sd["decoder.attn_net.W.weight"], sd["decoder.attn_net.W.bias"]).unsqueeze(1)
            loc = F.conv1d(cumulative.unsqueeze(1), sd["decoder.attn_net.conv.weight"], sd["decoder.attn_net.conv.bias"],
                           1, 15)
            ploc = F.linear(loc.transpose(1, 2), sd["decoder.attn_net.L.weight"])
            u = F.linear(torch.tanh(pq + proj + ploc), sd["decoder.attn_net.v.weight"]).squeeze(-1)
            u = u * char_mask
            scores = F.softmax(u, dim=1)
            cumulative = cumulative + scores
            ctx = torch.bmm(scores.unsqueeze(1), seq).squeeze(1)
            x = F.linear(torch.cat([ctx, attn_h], dim=1), sd["decoder.rnn_input.weight"], sd["decoder.rnn_input.bias"])
            h1, c1 = _lstm_cell(x, h1, c1, sd, "decoder.res_rnn1")
            x = x + h1
            h2, c2 = _lstm_cell(x, h2, c2, sd, "decoder.res_rnn2")
            x = x + h2
            mels = F.linear(x, sd["decoder.mel_proj.weight"]).view(B, 80, 20)[:, :, :r]
            stop = torch.sigmoid(F.linear(torch.cat([x, ctx], dim=1), sd["decoder.stop_proj.weight"],
                                          sd["decoder.stop_proj.bias"]))
            mel_out.append(mels)
            attn_out.append(scores.unsqueeze(1))
            prenet_in = mels[:, :, -1]
            if bool((stop * 10 > min_stop_token).all()) and t > 10:
                break
        mel = torch.cat(mel_out, dim=2)
        post = cbhg(mel, sd, "postnet", 5)
        linear = F.linear(post, sd["post_proj.weight"]).transpose(1, 2)
        return mel, linear, torch.cat(attn_out, 1)

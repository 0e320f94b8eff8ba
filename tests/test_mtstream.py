"""The reference-identical noise path (csrc/mt_stream.cu, csrc/mt19937_host.cpp): the library's MT19937 continuation of the
global torch CPU generator must equal ATen's own `exponential_` stream draw for draw, for any generator position and any
chunking, and must hand back the generator state ATen would be left in (WaveRNN.generate, fatchord_version.py:223-226)."""
import ctypes as C

import numpy as np
import pytest
import torch

from mockingbird_b200 import _lib
from mockingbird_b200.vocoder.wavernn.models.fatchord_version import (set_torch_cpu_generator_position,
                                                                        torch_cpu_generator_position)


def _fill(state, left, nxt, n):
    out = np.empty(n, np.uint32)
    l, x = C.c_int32(left), C.c_int32(nxt)
    _lib.check(_lib.lib().mb_mt19937_fill(state.ctypes.data, C.byref(l), C.byref(x), out.ctypes.data, n))
    return out, l.value, x.value


def _to_exp(raw):
    """ATen CPU: uniform_real_distribution<double>(random64()) then -log1p(-u), cast to float"""
    r64 = (raw[0::2].astype(np.uint64) << np.uint64(32)) | raw[1::2].astype(np.uint64)
    u = (r64 & np.uint64((1 << 53) - 1)).astype(np.float64) * 2.0 ** -53
    return (-np.log1p(-u)).astype(np.float32)


@pytest.mark.parametrize("burn", [0, 1, 311, 623, 624, 625, 5000])
def test_fill_equals_aten_exponential_stream(burn):
    """any starting position inside / at the edge of a 624-word block; chunked fills == one fill == ATen"""
    torch.manual_seed(99)
    if burn:
        torch.empty(burn, dtype=torch.float32).uniform_()  # float uniform consumes one 32-bit draw per element
    state, left, nxt = torch_cpu_generator_position()
    n_el = 58 * 512 * 3 + 17
    ref = torch.empty(n_el).exponential_(1).numpy()
    s_ref, l_ref, n_ref = torch_cpu_generator_position()
    st = state.copy()
    raw, l1, n1 = _fill(st, left, nxt, 2 * n_el)
    assert np.array_equal(_to_exp(raw), ref)
    assert (l1, n1) == (l_ref, n_ref) and np.array_equal(st, s_ref)
    # the same draws in ragged chunks
    st2, l2, n2, parts = state.copy(), left, nxt, []
    for m in (1, 2, 623, 624, 625, 10000):
        r, l2, n2 = _fill(st2, l2, n2, m)
        parts.append(r)
    cat = np.concatenate(parts)
    assert np.array_equal(cat, raw[:len(cat)])


def test_set_generator_position_roundtrip():
    torch.manual_seed(5)
    torch.rand(1000)
    state, left, nxt = torch_cpu_generator_position()
    a = torch.rand(7)
    torch.manual_seed(6)
    set_torch_cpu_generator_position(state, left, nxt)
    assert torch.equal(torch.rand(7), a)


@pytest.mark.gpu
def test_device_conversion_equals_aten():
    """mt_to_exp_kernel (CUDA double log1p) == ATen's host values on 3 M draws: identical up to the documented
    double-rounding corner (probability ~2^-28 per element of a 1-ulp difference)"""
    torch.manual_seed(4321)
    state, left, nxt = torch_cpu_generator_position()
    n_el = 3_000_000
    ref = torch.empty(n_el).exponential_(1)
    raw, _, _ = _fill(state.copy(), left, nxt, 2 * n_el)
    d_raw = torch.from_numpy(raw.view(np.int32)).cuda()
    d_out = torch.empty(n_el, device="cuda")
    _lib.check(_lib.lib().mb_mt_to_exp(C.c_void_p(d_raw.data_ptr()), C.c_void_p(d_out.data_ptr()), n_el,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    got = d_out.cpu()
    diff = (got != ref).nonzero().flatten()
    assert len(diff) <= 1, len(diff)
    if len(diff):
        i = int(diff[0])
        assert abs(float(got[i]) - float(ref[i])) <= abs(float(ref[i])) * 1.2e-7


@pytest.mark.gpu
def test_generate_fast_torch_rng_equals_host_replay():
    """rng='torch' (MT19937 stream in the library) == rng='torch_host' (ATen exponential_ per step): same integer
    samples, and the global generator ends in the same state (ragged: 8 folds, last chunk partial)"""
    import ref_init as ri
    from mockingbird_b200.vocoder.wavernn import inference as rnn_vocoder

    model = rnn_vocoder.load_state(ri.wavernn_state_dict(0, randomize_bn=True))
    mel = torch.rand(1, 80, 13, generator=torch.Generator().manual_seed(21)) * 2 - 1
    outs, tails = [], []
    for mode in ("torch_host", "torch", "torch"):
        model.rng = mode
        torch.manual_seed(777)
        outs.append(model.generate_indices(mel, True, 300, 35, None))
        tails.append(torch.rand(5))
    model.rng = "torch"
    assert outs[0].shape == (8, 370)  # 13 frames -> 2600 samples -> 7 full folds of 335 + a remainder fold
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[1], outs[2])
    assert torch.equal(tails[0], tails[1]) and torch.equal(tails[1], tails[2])

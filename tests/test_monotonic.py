"""monotonic_align.maximum_path: numpy restatement == the reference's own compiled core.pyx (oracle/_ref,
built by oracle/build_oracle.build_ref() in the build container; travels to the GPU box) == the CUDA wavefront DP."""
import numpy as np
import pytest
import torch

import monotonic_oracle as mo


def _case(seed, b, ty, tx):
    rng = np.random.RandomState(seed)
    v = (rng.randn(b, ty, tx) * 3).astype(np.float32)
    t_ys = rng.randint(max(1, ty // 2), ty + 1, size=b).astype(np.int32)
    t_xs = np.minimum(rng.randint(1, tx + 1, size=b), t_ys).astype(np.int32)  # a monotonic path needs t_x <= t_y
    t_ys[0], t_xs[0] = ty, min(tx, ty)
    return v, t_ys, t_xs


@pytest.mark.parametrize("shape", [(3, 37, 11), (2, 1, 1), (4, 64, 64), (2, 300, 75)])
def test_restatement_equals_compiled_reference(shape):
    core = mo.reference_core()
    if core is None:
        pytest.skip("oracle/_ref not built (needs /root/reference once)")
    v, t_ys, t_xs = _case(1, *shape)
    p_ref, v_ref = np.zeros(v.shape, np.int32), v.copy()
    core.maximum_path_c(p_ref, v_ref, t_ys, t_xs)
    p, vv = mo.maximum_path_numpy(v, t_ys, t_xs)
    assert np.array_equal(p, p_ref) and np.array_equal(vv, v_ref)
    assert np.array_equal(p.sum(axis=(1, 2)), t_ys)  # one cell per frame


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 37, 11), (2, 1, 1), (4, 64, 64), (16, 1000, 200), (2, 300, 75)])
def test_cuda_equals_reference(shape):
    from mockingbird_b200.monotonic_align import maximum_path

    v, t_ys, t_xs = _case(2, *shape)
    b, ty, tx = shape
    mask = np.zeros((b, ty, tx), np.float32)
    for i in range(b):
        mask[i, : t_ys[i], : t_xs[i]] = 1
    got = maximum_path(torch.from_numpy(v).cuda(), torch.from_numpy(mask).cuda())
    assert got.dtype == torch.float32 and got.shape == (b, ty, tx)
    core = mo.reference_core()
    if core is not None:
        p_ref, v_ref = np.zeros(v.shape, np.int32), v.copy()
        core.maximum_path_c(p_ref, v_ref, t_ys, t_xs)
    else:
        p_ref, _ = mo.maximum_path_numpy(v, t_ys, t_xs)
    assert np.array_equal(got.cpu().numpy().astype(np.int32), p_ref)
    # size-independent properties: exactly one cell per valid frame, monotone non-decreasing column, ends at the corners
    g = got.cpu().numpy()
    for i in range(b):
        cols = g[i, : t_ys[i]].argmax(axis=1)
        assert g[i].sum() == t_ys[i] and cols[0] == 0 and cols[-1] == t_xs[i] - 1
        assert np.all(np.diff(cols) >= 0) and np.all(np.diff(cols) <= 1)

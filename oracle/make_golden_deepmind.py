"""DeepMind dual-softmax WaveRNN golden vectors from the LIVE reference (TEST INFRASTRUCTURE; container only).

    CUDA_VISIBLE_DEVICES="" python oracle/make_golden_deepmind.py

models/vocoder/wavernn/models/deepmind_version.py is imported UNMODIFIED through ref_harness.load_deepmind() (stub modules
for the two imports that do not exist in the tree, CPU no-op `.cuda()`).  Weights: ref_init.deepmind_state_dict(0) (constructor
order, gate biases ~ N(0, 0.1) so they are exercised); torch.manual_seed(1234) before generate(4000)."""
from __future__ import annotations

import os
import sys
from pathlib import Path

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent / "synth_weights"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import ref_harness as rh  # noqa: E402
import ref_init as ri  # noqa: E402
from make_golden import GOLDEN, meta  # noqa: E402

STEPS = 4000

if __name__ == "__main__":
    W = rh.load_deepmind()
    torch.manual_seed(0)
    m = W()
    m.load_state_dict(ri.deepmind_state_dict(0))
    torch.manual_seed(1234)
    out, coarse, fine = m.generate(STEPS)
    np.savez_compressed(GOLDEN / "deepmind_seed0.npz", coarse=coarse.astype(np.int16), fine=fine.astype(np.int16),
                        output=out.astype(np.int32),
                        meta=meta(weights="ref_init.deepmind_state_dict(0)", gen_seed=1234, steps=STEPS,
                                  call="WaveRNN(896, 256).generate(4000) on CPU through ref_harness.load_deepmind()"))
    print("deepmind", coarse.shape, out[:6])

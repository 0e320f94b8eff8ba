"""Generate the golden fixtures under tests/golden/ from the LIVE reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):

    CUDA_VISIBLE_DEVICES="" python oracle/make_golden.py

Every fixture stores the seeds, torch version and CPU capability it was made with (SURVEY.md
appendix C item 6).  Weights are never stored: tests rebuild them with synth_weights/ref_init.py, which
tests/test_oracle_pinned.py proves bit-identical to the reference constructors.
"""
from __future__ import annotations

import json
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

GOLDEN = HERE.parent / "tests" / "golden"


def meta(**kw):
    cap = [l for l in torch.__config__.show().splitlines() if "CPU capability" in l]
    d = dict(torch=torch.__version__, cpu_capability=cap[0].strip() if cap else "?",
             threads=torch.get_num_threads(), reference_commit="28dc5e1")
    d.update(kw)
    return json.dumps(d)


def golden_hifigan():
    """cfg-2-shaped small case (B=2, T=48) and two full-length utterances of cfg 2 (T=256)."""
    g = rh.build_hifigan(seed=0)
    with torch.no_grad():
        mel_small = torch.rand(2, 80, 48, generator=torch.Generator().manual_seed(2)) * 8 - 4
        wav_small = g(mel_small)
        mel_full = torch.rand(32, 80, 256, generator=torch.Generator().manual_seed(2)) * 8 - 4
        pick = [0, 31]
        wav_full = torch.cat([g(mel_full[i:i + 1]) for i in pick])
    np.savez_compressed(GOLDEN / "hifigan_seed0.npz", mel_small=mel_small.numpy(), wav_small=wav_small.numpy(),
                        full_pick=np.array(pick), wav_full=wav_full.numpy(),
                        meta=meta(weights_seed=0, mel_small="rand(2,80,48;seed 2)*8-4",
                                  mel_full="rand(32,80,256;seed 2)*8-4 rows 0 and 31"))


def golden_fregan():
    g = rh.build_fregan(seed=0)
    with torch.no_grad():
        mel = torch.rand(2, 80, 40, generator=torch.Generator().manual_seed(12)) * 8 - 4
        wav = g(mel)
    np.savez_compressed(GOLDEN / "fregan_seed0.npz", mel=mel.numpy(), wav=wav.numpy(),
                        meta=meta(weights_seed=0, mel="rand(2,80,40;seed 12)*8-4"))


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    rh.install()
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    golden_hifigan()
    golden_fregan()
    try:
        import make_golden_wavernn  # noqa: F401  (added with the WaveRNN path)

        make_golden_wavernn.main(GOLDEN, meta)
    except ImportError:
        pass
    try:
        import make_golden_tacotron

        make_golden_tacotron.main(GOLDEN, meta)
    except ImportError:
        pass
    for p in sorted(GOLDEN.glob("*.npz")):
        print(p.name, p.stat().st_size)


if __name__ == "__main__":
    main()

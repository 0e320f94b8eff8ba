"""smoke(): one tiny WaveRNN generate on cuda:0 checked against the CPU twin (bit-exact)."""
import numpy as np
import torch

import ref_init as ri
import wavernn_oracle as wo


def run():
    from mockingbird_b200.vocoder.wavernn import inference as rnn_vocoder

    sd = ri.wavernn_state_dict(0, randomize_bn=True)
    model = rnn_vocoder.load_state(sd, rng="device", seed=3)
    mel = torch.rand(1, 80, 2, generator=torch.Generator().manual_seed(1)) * 2 - 1
    idx = model.generate_indices(mel, False, 8000, 400, None)
    twin = wo.Twin({k: v.numpy() for k, v in sd.items() if v.dtype == torch.float32})
    aux, melup = twin.condition(mel[0].numpy())
    ref = twin.generate(aux, melup, [0], 400, None, seed=3)
    assert np.array_equal(idx, ref), "WaveRNN kernel != CPU twin"
    print("smoke wavernn: 400 samples bit-exact vs twin")

"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: utterance sharding, weight-arena
broadcast, host gather - the N>1 path of bench.py / the serving recipe in INTEGRATION.md."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mockingbird_b200 import distributed as mbd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lengths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        arena = torch.arange(4096, dtype=torch.int64).to(torch.uint8) if rank == 0 else torch.zeros(4096, dtype=torch.uint8)
        mbd.broadcast_weights([arena], src=0)
        ok_arena = bool(torch.equal(arena, torch.arange(4096, dtype=torch.int64).to(torch.uint8)))
        mine = mbd.shard_utterances(lengths, rank, world)
        local = [("utt", i, lengths[i] * 200) for i in mine]  # stand-in for vocoded waveforms
        gathered = mbd.gather_object_lists(local, dst=0)
        if rank == 0:
            shards = [mbd.shard_utterances(lengths, r, world) for r in range(world)]
            merged = mbd.merge_sharded(gathered, shards, len(lengths))
            q.put((ok_arena, merged, [len(s) for s in shards], [sum(lengths[i] for i in s) for s in shards]))
        else:
            q.put((ok_arena,))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_broadcast_gather():
    lengths = [120, 33, 75, 20, 99, 64, 101, 47, 58]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = [o for o in outs if len(o) == 4][0]
    assert all(o[0] for o in outs)                      # both ranks hold rank 0's arena bytes
    ok, merged, counts, work = full
    assert [m[1] for m in merged] == list(range(len(lengths)))   # every utterance exactly once, in order
    assert [m[2] for m in merged] == [l * 200 for l in lengths]
    assert abs(counts[0] - counts[1]) <= 1
    assert abs(work[0] - work[1]) <= max(lengths)        # length-sorted dealing balances padded work


def test_shard_is_a_partition_single_process():
    lengths = list(range(1, 38))
    for ws in (1, 2, 4, 8):
        shards = [mbd.shard_utterances(lengths, r, ws) for r in range(ws)]
        assert sorted(i for s in shards for i in s) == list(range(len(lengths)))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def test_encoder_module_protocol_without_gpu():
    """host-side surface of the speaker-encoder drop-in: names, defaults and error behaviour need no GPU"""
    import numpy as np
    import pytest

    from mockingbird_b200.encoder import audio, inference

    assert inference.partials_n_frames == 160 and inference.sampling_rate == 16000 and inference.mel_n_channels == 40
    assert not inference.is_loaded() or inference._model is not None
    saved = inference._model
    inference._model = None
    try:
        with pytest.raises(Exception, match="Model was not loaded"):
            inference.embed_frames_batch(np.zeros((1, 160, 40), np.float32))
    finally:
        inference._model = saved
    import torch

    if not torch.cuda.is_available():  # the front-end is a CUDA kernel: it must fail loudly, never fall back to the CPU
        from mockingbird_b200 import _lib

        with pytest.raises(_lib.MbError, match="no CPU fallback"):
            audio.wav_to_mel_spectrogram(np.zeros(16000, np.float32))
    # every partial slice has the requested length and the wav / mel slices stay aligned (160 samples per frame)
    w, m = inference.compute_partial_slices(40000, overlap=0.25)
    assert all(s.stop - s.start == 160 for s in m)
    assert all(ws.start == ms.start * 160 and ws.stop == ms.stop * 160 for ws, ms in zip(w, m))


def test_vocoder_on_disk_formats(tmp_path):
    """train.txt rows and [frames, 80] .npy mels"""
    import numpy as np

    from mockingbird_b200.vocoder import formats

    mel = np.arange(7 * 80, dtype=np.float32).reshape(7, 80)
    np.save(tmp_path / "mel-a.npy", mel, allow_pickle=False)
    (tmp_path / "train.txt").write_text("audio-a.npy|mel-a.npy|embed-a.npy|1400|7|ni3 hao3\n"
                                        "audio-b.npy|mel-b.npy|embed-b.npy|0|0|dropped\n"
                                        "audio-c.npy|mel-c.npy|embed-c.npy|2000|10|text with | pipe\n", encoding="utf-8")
    rows = formats.read_metadata(tmp_path / "train.txt")
    assert [r.mel_fname for r in rows] == ["mel-a.npy", "mel-c.npy"]
    assert rows[0].n_samples == 1400 and rows[0].n_frames == 7 and rows[0].text == "ni3 hao3"
    assert rows[1].text == "text with | pipe"
    m = formats.load_mel(tmp_path / "mel-a.npy")
    assert m.shape == (80, 7) and m.dtype == np.float32 and m.flags["C_CONTIGUOUS"] and m[3, 2] == mel[2, 3]

    class FakeVocoder:
        def infer_waveform(self, mel):
            return np.zeros(mel.shape[1] * 200, np.float32), 16000

    wavs = formats.vocode_files([tmp_path / "mel-a.npy"], FakeVocoder())
    assert len(wavs) == 1 and wavs[0].shape == (1400,)


def _fold_worker(rank, world, port, num_folds, steps, q):
    import numpy as np

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = mbd.fold_range(num_folds, rank, world)
        # stand-in for the device result: row g holds g*1000 + step
        local = (np.arange(lo, hi, dtype=np.int16)[:, None] * 100 + np.arange(steps, dtype=np.int16)[None, :]).astype(np.int16)
        full = mbd.gather_fold_rows(local, num_folds, dst=0)
        q.put((rank, lo, hi, None if full is None else full.tolist()))
    finally:
        dist.destroy_process_group()


def test_fold_range_partition():
    """58 folds over 8 ranks -> 8,8,7,... contiguous, disjoint, complete; more ranks than folds ok"""
    for n, w in ((58, 8), (58, 1), (58, 4), (3, 8), (0, 2), (64, 2)):
        r = [mbd.fold_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(sizes) <= 1
    assert [hi - lo for lo, hi in (mbd.fold_range(58, k, 8) for k in range(8))] == [8, 8, 7, 7, 7, 7, 7, 7]


def test_two_rank_fold_gather():
    """the WaveRNN fold-sharding host logic on gloo, world_size 2: 7 folds -> 4 + 3 rows, gathered in fold order"""
    import numpy as np

    num_folds, steps = 7, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fold_worker, args=(r, 2, port, num_folds, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (outs[0][1], outs[0][2], outs[1][1], outs[1][2]) == (0, 4, 4, 7)
    assert outs[1][3] is None
    full = np.array(outs[0][3], dtype=np.int16)
    want = (np.arange(num_folds, dtype=np.int16)[:, None] * 100 + np.arange(steps, dtype=np.int16)[None, :]).astype(np.int16)
    assert np.array_equal(full, want)

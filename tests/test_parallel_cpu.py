"""Data-parallel host logic on CPU: LPT sharding and the waveform gather over a 2-rank gloo group
(the only collective on the path; NCCL on the GPU box uses the same code)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from artspeech_b200 import engine


def test_lpt_sharding_balances_and_covers():
    g = torch.Generator().manual_seed(0)
    costs = torch.randint(15, 600, (512,), generator=g).tolist()
    for world in (1, 2, 4, 8):
        shards = engine.shard_utterances(costs, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(512))                      # every utterance exactly once
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)          # LPT bound
    assert engine.shard_utterances([], 4) == [[], [], [], []]
    assert engine.shard_utterances([5.0], 2) == [[0], []]


def test_length_contiguous_sharding_keeps_padding_flat():
    """engine.shard_by_length on the config-5 length distribution: a partition, cost-balanced within a few per cent,
    and -- unlike LPT on frame counts -- the padded work per frame does not grow with the number of ranks."""
    g = torch.Generator().manual_seed(1)
    frames = (torch.randint(15, 595, (512,), generator=g) * 5.3).long().tolist()
    q = lambda f: (f + 79) // 80 * 80

    def padded(shard):
        f = [frames[i] for i in shard]
        return sum(len(b) * q(max(f[i] for i in b)) for b in engine.bucket_utterances(f, 16, 25600, quantum=80))
    base = padded(list(range(512))) / sum(frames)
    for world in (1, 2, 4, 8):
        shards = engine.shard_by_length(frames, world)
        assert sorted(i for s in shards for i in s) == list(range(512))
        cost = [padded(s) + 280 * len(s) for s in shards]
        assert max(cost) <= 1.08 * sum(cost) / world
        assert sum(padded(s) for s in shards) / sum(frames) <= base + 0.02
        lpt = engine.shard_utterances(frames, world)
        if world == 8:
            assert sum(padded(s) for s in lpt) / sum(frames) >= base + 0.10      # what it replaces (uniform lengths: +15 %; the long-tailed config-5 lengths: +28 %)
        # contiguous in length: every utterance of rank r is at least as long as every utterance of rank r + 1
        for a, b in zip(shards, shards[1:]):
            if a and b:
                assert min(frames[i] for i in a) >= max(frames[i] for i in b)
    assert engine.shard_by_length([], 3) == [[], [], []]
    assert sorted(i for s in engine.shard_by_length([100, 50], 4) for i in s) == [0, 1]


def test_length_bucketed_micro_batches():
    """engine.bucket_utterances on the BASELINE config-5 length distribution: every utterance exactly once, batch and
    padded-frame budgets respected, little padding."""
    g = torch.Generator().manual_seed(1)
    frames = (torch.randint(15, 595, (512,), generator=g) * 5.3).long().tolist()
    for max_batch, budget in ((16, 25600), (8, 12800), (16, None), (1, None)):
        batches = engine.bucket_utterances(frames, max_batch, budget)
        assert sorted(i for b in batches for i in b) == list(range(512))
        assert all(1 <= len(b) <= max_batch for b in batches)
        if budget is not None:
            assert all(len(b) == 1 or len(b) * max(frames[i] for i in b) <= budget for b in batches)
        padded = sum(len(b) * max(frames[i] for i in b) for b in batches)
        assert padded <= 1.06 * sum(frames)                  # sorted neighbours: < 6 % padding
    # quantised frame counts (the engine's shape buckets): still a partition, the budget holds for the bucketed size,
    # and the set of (batch size, bucket) keys is small
    batches = engine.bucket_utterances(frames, 16, 25600, quantum=128)
    assert sorted(i for b in batches for i in b) == list(range(512))
    q = lambda f: (f + 127) // 128 * 128
    assert all(len(b) == 1 or len(b) * q(max(frames[i] for i in b)) <= 25600 for b in batches)
    keys = {(len(b), q(max(frames[i] for i in b))) for b in batches}
    assert len(keys) <= len(batches) and sum(len(b) * q(max(frames[i] for i in b)) for b in batches) <= 1.10 * sum(frames)
    assert engine.bucket_utterances([], 16) == []
    assert engine.bucket_utterances([40000], 16, 25600) == [[0]]          # over-budget utterance: its own batch
    assert engine.bucket_utterances([10, 30, 20], 2) == [[1, 2], [0]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        costs = [800, 300, 650, 120, 480]
        mine = engine.shard_utterances(costs, world)[rank]
        # stand-in "waveforms": utterance i has costs[i]*3 samples filled with i+1
        S = max([costs[i] * 3 for i in mine], default=1)
        wav = torch.zeros(len(mine), S)
        lens = torch.zeros(len(mine), dtype=torch.int64)
        for r, i in enumerate(mine):
            wav[r, :costs[i] * 3] = i + 1
            lens[r] = costs[i] * 3
        wavs, lns = engine.gather_waveforms(wav, lens, dst=0)
        # host-known shapes (from the same deterministic sharding): no size exchange, same result
        shards = engine.shard_utterances(costs, world)
        shapes = [(len(sh), max([costs[i] * 3 for i in sh], default=1)) for sh in shards]
        wavs2, lns2 = engine.gather_waveforms(wav, lens, dst=0, shapes=shapes)
        # int16 PCM payload (Synthesizer(pcm16=True)): travels as bytes, NCCL has no int16
        wavs3, lns3 = engine.gather_waveforms(wav.to(torch.int16), lens, dst=0, shapes=shapes)
        # the pipelined gatherer: ring buffers + one collective per submit; two submits, the second one wins
        gat = engine.WaveformGatherer("cpu", shapes, torch.float32, dst=0)
        gat.submit(wav * 0, lens)
        gat.submit(wav, lens)
        wavs4, lns4 = gat.results()
        if rank == 0:
            ok = all(torch.equal(a, b) for a, b in zip(wavs, wavs2)) and all(torch.equal(a, b) for a, b in zip(lns, lns2))
            ok &= all(w.dtype == torch.int16 and torch.equal(w.float(), a) for w, a in zip(wavs3, wavs))
            ok &= all(torch.equal(a, b) for a, b in zip(lns, lns3))
            ok &= all(torch.equal(a, b) for a, b in zip(wavs, wavs4)) and all(torch.equal(a, b) for a, b in zip(lns, lns4))
            for r in range(world):
                for row, i in enumerate(shards[r]):
                    n = int(lns[r][row])
                    ok &= n == costs[i] * 3 and bool((wavs[r][row, :n] == i + 1).all()) and \
                        bool((wavs[r][row, n:] == 0).all())
            q.put(ok)
        else:
            assert wavs is None and lns is None and wavs2 is None and wavs3 is None and wavs4 is None
    finally:
        dist.destroy_process_group()


def test_gather_waveforms_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True

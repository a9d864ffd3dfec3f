"""BASELINE config 5: 512 mixed-length utterances (seeded LibriTTS-like token lengths: min 15, p50 125, p90 314,
max 594; frames ~ 5.3 x tokens; 240-frame reference mels) sharded over the GPUs of one box.
    python tools/box_sweep.py                                  (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/box_sweep.py
LPT sharding by frame count -> per-rank length-bucketed ragged micro-batches (eager launches: every micro-batch has
its own shape) -> NCCL gather of the waveforms on rank 0.  Timed on the device, max over ranks; prints one JSON line."""
import json, math, os, sys, time, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from artspeech_b200 import checkpoint, engine

N_UTT = int(os.environ.get("ASB_SWEEP_UTTS", "512"))


def make_workload(n=N_UTT, seed=0):
    g = torch.Generator().manual_seed(seed)
    sigma = math.log(314 / 125) / 1.2816                     # lognormal through p50 = 125, p90 = 314
    tl = (125 * torch.exp(sigma * torch.randn(n, generator=g))).round().clamp(15, 594).long()
    toks = [torch.randint(1, 178, (int(t),), generator=g) for t in tl]
    durs = [2 + (torch.rand(int(t), generator=g) < 0.67).long() for t in tl]     # mean 2.67 half-rate frames per token
    return toks, durs


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    syn = engine.Synthesizer(checkpoint.build_random_artsspeech(0), checkpoint.build_random_generator(0), device=dev,
                             use_cuda_graph=False)
    toks, durs = make_workload()
    frames = [2 * int(d.sum()) for d in durs]
    mine = engine.shard_utterances(frames, world)[rank]
    g = torch.Generator().manual_seed(100 + rank)
    mels = [(torch.randn(80, 240, generator=g) * 0.5).clamp(-2, 2) for _ in mine]
    my_t, my_d = [toks[i] for i in mine], [durs[i] for i in mine]
    engine.synthesize_many(syn, my_t[:4], mels[:4], my_d[:4])          # warm-up: weight packing, kernel attributes
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    wavs, fr = engine.synthesize_many(syn, my_t, mels, my_d)
    S = engine.HOP * max(fr)
    pad = torch.zeros(len(wavs), S, device=dev)
    for j, w in enumerate(wavs):
        pad[j, :w.numel()] = w
    lens = torch.tensor([w.numel() for w in wavs], device=dev)
    if world > 1:
        got, _ = engine.gather_waveforms(pad, lens, dst=0)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        audio_s = sum(frames) * engine.HOP / engine.SAMPLE_RATE
        batches = engine.bucket_utterances(fr, 16, 25600)
        padded = sum(len(b) * max(fr[i] for i in b) for b in batches)
        print(json.dumps({"workload": f"box_sweep_{N_UTT}_mixed_length", "n_gpus": world, "utterances": N_UTT,
                          "audio_s": round(audio_s, 1), "ms": round(float(ms), 2),
                          "audio_s_per_s": round(audio_s / (float(ms) / 1e3), 1),
                          "rank0_micro_batches": len(batches), "rank0_padding_overhead": round(padded / max(sum(fr), 1) - 1, 4),
                          "rank0_wall_s": round(time.time() - t0, 3), "cuda_graph": False,
                          "gathered_ranks": (len(got) if world > 1 else 1)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

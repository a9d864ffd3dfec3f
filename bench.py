#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: ArtSpeech batched synthesis (text/phonemes +
reference mel -> mel -> waveform) on B200.

    python bench.py --gpus N --steps K --warmup W            # ours (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU PyTorch path

Workload (BASELINE.json configs[1]): per GPU, 16 utterances x 150 phoneme
tokens, a 3 s reference mel (240 frames) each, forced integer durations summing to 400 per
utterance -> 800 mel frames = 10.0 s of 24 kHz audio per utterance, 160 audio-seconds per step per
GPU.  Synthetic seeded inputs, random-init conditioned weights (artspeech_b200.checkpoint): the
reference's checkpoints are not redistributable.

One JSON line on stdout (rank 0).  `value` = audio-seconds synthesised per wall-second with inputs
resident in HBM; `e2e` = the same through the public API with pinned HOST inputs (H2D of tokens /
reference mels, D2H of the waveforms inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import warnings

warnings.filterwarnings("ignore", category=FutureWarning)
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, TT, TR, SUM_DUR = int(os.environ.get("ASB_BENCH_B", "16")), 150, 240, 400   # env override: experiments only
FRAMES = 2 * SUM_DUR                                   # mel frames per utterance
AUDIO_S_PER_UTT = FRAMES * 300 / 24000.0               # 10.0 s
VOCODER_FLOP_PER_FRAME = 623.7e6
WORKLOAD = "libritts_batch16x10s_text2wave"


def make_inputs(rank: int, B: int = B_PER_GPU):
    g = torch.Generator().manual_seed(1234 + rank)
    tokens = torch.randint(1, 178, (B, TT), generator=g)
    mels = (torch.randn(B, 80, TR, generator=g) * 0.5).clamp(-2, 2)
    base = SUM_DUR // TT
    dur = torch.full((B, TT), base, dtype=torch.int64)
    dur[:, : SUM_DUR - base * TT] += 1                 # sums to exactly 400
    tok_lens = torch.full((B,), TT, dtype=torch.int64)
    mel_lens = torch.full((B,), TR, dtype=torch.int64)
    return tokens, tok_lens, mels, mel_lens, dur


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([p.strip() for p in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max([int(s[1]) for s in self.samples if s[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(self.samples)}


def vocoder_traffic():
    """DRAM bytes (read + write) of the vocoder's conv launches per step, from the committed ncu capture
    (profiles/r01_vocoder_traffic.json); None when no capture is available."""
    path = os.path.join(ROOT, "profiles", "r01_vocoder_traffic.json")
    if os.path.isfile(path):
        try:
            return json.load(open(path)).get("dram_bytes_per_step")
        except Exception:
            return None
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU PyTorch path (oracle/restate.py is its pinned restatement;
# the reference tree itself does not exist on the GPU box)
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(n_utts: int, threads: int):
    """Time the fp32 CPU path on ``n_utts`` utterances of the bench workload (batch-1 acoustic loop,
    as the reference's step='test' is batch-1 only; vocoder on the same utterance).  Returns
    (audio_seconds, seconds)."""
    from artspeech_b200 import checkpoint
    from oracle import restate
    torch.set_num_threads(threads)
    model = checkpoint.build_random_artsspeech(0)
    gen = checkpoint.build_random_generator(0)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    gsd = {k: v.detach() for k, v in gen.state_dict().items()}
    dist = {k: v.cpu() for k, v in model.distribution.items()}
    tokens, _, mels, _, dur = make_inputs(0, max(n_utts, 1))
    # warm-up on a short utterance (thread pools, oneDNN primitive caches)
    restate.generator_forward(gsd, restate.artsspeech_test(sd, tokens[:1, :20], mels[:1, :, :100], dist,
                                                           durations=torch.ones(20, dtype=torch.long)))
    t0 = time.perf_counter()
    for i in range(n_utts):
        mel = restate.artsspeech_test(sd, tokens[i:i + 1], mels[i:i + 1], dist, durations=dur[i])
        wav = restate.generator_forward(gsd, mel)
        assert wav.shape[-1] == FRAMES * 300
    dt = time.perf_counter() - t0
    return n_utts * AUDIO_S_PER_UTT, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n_per_step = 2
    for _ in range(args.warmup):
        pass                                           # warm-up happens inside cpu_reference_sample
    times = []
    audio = 0.0
    for _ in range(args.steps):
        a, t = cpu_reference_sample(n_per_step, threads)
        audio += a
        times.append(t)
    total = sum(times)
    value = audio / total
    line = {"impl": "reference", "metric": "synthesized_audio_seconds_per_second", "value": value,
            "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tokens": TT, "ref_frames": TR, "mel_frames": FRAMES,
                       "utterances_per_step": n_per_step},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port",
                             "sample": f"{n_per_step} utterances x 10 s per step, batch-1 acoustic + vocoder, fp32 torch"},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from artspeech_b200 import checkpoint, engine, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = checkpoint.build_random_artsspeech(0)
    gen = checkpoint.build_random_generator(0)
    syn = engine.Synthesizer(model, gen, device=dev, use_cuda_graph=not args.no_graph,
                             pipeline_depth=1 if args.no_graph else args.pipeline)

    tokens, tok_lens, mels, mel_lens, dur = make_inputs(rank)
    tok_d, mel_d = tokens.to(dev), mels.to(dev)
    # pinned host staging for the end-to-end arm
    tok_h, mel_h = tokens.pin_memory(), mels.pin_memory()
    wav_hs = [torch.empty(B_PER_GPU, FRAMES * 300, dtype=torch.float32).pin_memory() for _ in range(max(args.pipeline, 1))]
    wav_h = wav_hs[0]
    calls = [0]
    wav_lens_d = torch.full((B_PER_GPU,), FRAMES * 300, device=dev)

    def step_resident():
        return syn.synthesize(tok_d, tok_lens, mel_d, mel_lens, dur)

    def step_e2e():
        t = tok_h.to(dev, non_blocking=True)
        m = mel_h.to(dev, non_blocking=True)
        wav, _, _ = syn.synthesize(t, tok_lens, m, mel_lens, dur)
        calls[0] += 1
        # the waveforms are produced on the engine's stream for this call: gather / D2H follow on that stream
        with torch.cuda.stream(syn.last_stream):
            if world > 1:
                gathered, _ = engine.gather_waveforms(wav, wav_lens_d, shapes=[tuple(wav.shape)] * world)
            wav_hs[calls[0] % len(wav_hs)].copy_(wav, non_blocking=True)
        return wav

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        syn.join()                 # pipelined calls run on the engine's streams: the end event waits for all of them
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()
    barrier()

    # inputs (tokens 19 KB + mels 1.2 MB) are tiny, but every step streams > 10 GB of activations
    # through HBM, far beyond the 126 MB L2: no explicit flush needed between timed iterations.
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = ops.launch_count
    ms = timed(step_resident, args.steps)
    launches = ops.launch_count - l0
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None

    # extra (not the headline): the same step with the style-encoder outputs cached per reference voice
    # (engine.encode_voice) — the serving case where many utterances share a voice
    voice = syn.encode_voice(mel_d, mel_lens)

    def step_voice_cached():
        return syn.synthesize(tok_d, tok_lens, mel_d, mel_lens, dur, voice=voice)
    for _ in range(3):
        step_voice_cached()
    ms_cached = timed(step_voice_cached, args.steps)

    # dominant kernel family: the vocoder's tensor-core convolutions (27 fused resblock-pair launches + 24
    # implicit-GEMM launches, nothing else).
    # Timed alone, replayed from a CUDA graph (no launch gaps), with CUDA events on the launch stream.
    _, lens_m, mel_out = step_resident()
    syn.join()
    mel_static = mel_out.clone()
    lens_static = lens_m.clone()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            syn.generator(mel_static, lens_static)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    vg = torch.cuda.CUDAGraph()
    l_before = ops.launch_count
    with torch.cuda.graph(vg):
        syn.generator(mel_static, lens_static)
    voc_launches = ops.launch_count - l_before
    for _ in range(3):
        vg.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(5, args.steps)
    e0.record()
    for _ in range(reps):
        vg.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    voc_ms = e0.elapsed_time(e1) / reps
    voc_flops = VOCODER_FLOP_PER_FRAME * B_PER_GPU * FRAMES

    # MAS (BASELINE config 5): 64 x 200 tokens x 1000 frames, fp32, bit-exact path
    mas_info = None
    if rank == 0:
        from artspeech_b200 import mas
        gm = torch.Generator().manual_seed(7)
        val = torch.randn(64, 200, 1000, generator=gm).to(dev)
        xl = torch.randint(100, 201, (64,), generator=gm)
        yl = torch.maximum(torch.randint(500, 1001, (64,), generator=gm), xl)
        xl[0], yl[0] = 200, 1000
        xl_d, yl_d = xl.to(dev, torch.int32), yl.to(dev, torch.int32)
        for _ in range(3):
            mas.maximum_path_lens(val, xl_d, yl_d)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(20):
            mas.maximum_path_lens(val, xl_d, yl_d)
        e1.record()
        torch.cuda.synchronize(dev)
        mas_ms = e0.elapsed_time(e1) / 20
        mas_bytes = 2.0 * 64 * 200 * 1000 * 4
        mas_info = {"shape": "64x200x1000 fp32", "ms": mas_ms, "algorithmic_GBps": mas_bytes / mas_ms / 1e6,
                    "frac_of_hbm_peak": mas_bytes / mas_ms / 1e6 / peaks()["hbm_gbs"],
                    "note": "serial dependency chain over Ty: latency-bound by construction, not roofline-graded"}

    # HBM-bound kernel family: fused InstanceNorm + AdaIN + LeakyReLU (as_adain_norm_apply) at the decoder's
    # widest block, fp32 in -> f16 out.  Six rotating input sets (> 126 MB L2), launches replayed back to back
    # from a CUDA graph, algorithmic bytes = one read + one write of the tensor.
    hbm_info = None
    if rank == 0:
        nset, Cd = 6, 1024
        xs = [torch.randn(B_PER_GPU, FRAMES, Cd, device=dev) * 0.7 + 3.0 for _ in range(nset)]
        gbv = torch.randn(B_PER_GPU, 2 * Cd, device=dev) * 0.3
        lens_f = torch.full((B_PER_GPU,), FRAMES, dtype=torch.int32, device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for xv in xs:
                ops.adain_norm(xv, gbv, 0.2, lens_f, torch.float16)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        ag = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ag):
            keep = [ops.adain_norm(xv, gbv, 0.2, lens_f, torch.float16) for xv in xs]
        ag.replay()
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(5):
            ag.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        ad_us = e0.elapsed_time(e1) / (5 * nset) * 1e3
        ad_bytes = B_PER_GPU * FRAMES * Cd * (4 + 2)
        hbm_info = {"bound": "hbm", "kernel": "adain_ring_kernel (InstanceNorm + AdaIN + LeakyReLU, 16x800x1024 fp32 -> f16)",
                    "achieved": ad_bytes / ad_us / 1e3, "peak": peaks()["hbm_gbs"], "unit": "GB/s",
                    "frac": ad_bytes / ad_us / 1e3 / peaks()["hbm_gbs"], "us_per_launch": ad_us,
                    "bytes_per_launch": ad_bytes, "l2": "6 rotating input sets (474 MB), graph replay"}
        del xs, keep

    # the step before the path: log-mel front-end of the 16 reference recordings (3 s each)
    fe_info = None
    if rank == 0:
        from artspeech_b200 import frontend
        fe = frontend.LogMel().to(dev)
        wv = torch.randn(B_PER_GPU, TR * 300, device=dev) * 0.1
        for _ in range(3):
            fe(wv)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(20):
            fe(wv)
        e1.record()
        torch.cuda.synchronize(dev)
        fe_info = {"kernel": "log_mel_kernel (STFT 2048/1200/300 + 80 mel + log, fp32)", "recordings": B_PER_GPU,
                   "seconds_each": TR * 300 / 24000.0, "us": e0.elapsed_time(e1) / 20 * 1e3,
                   "note": "not part of the timed step (reference mels are the step's inputs, as in BASELINE configs)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    audio_per_step = world * B_PER_GPU * AUDIO_S_PER_UTT
    value = audio_per_step * args.steps / (ms / 1e3)
    e2e_value = audio_per_step * args.steps / (ms_e2e / 1e3)
    pk = peaks()
    achieved = voc_flops / (voc_ms / 1e3) / 1e12
    line = {
        "metric": "synthesized_audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 (vocoder) / f16 (acoustic) operands, f32 accumulate", "data": "synthetic",
        "config": {"workload": WORKLOAD, "utterances_per_gpu": B_PER_GPU, "tokens": TT, "ref_frames": TR,
                   "mel_frames": FRAMES, "audio_s_per_step_per_gpu": B_PER_GPU * AUDIO_S_PER_UTT,
                   "weights": "random-init conditioned (seed 0); vocoder checkpoint g_00935000 not available",
                   "cuda_graph": not args.no_graph, "batches_in_flight": 1 if args.no_graph else args.pipeline,
                   "l2": "activations streamed per step >> 126 MB L2, no explicit flush"},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(tok_h.numel() * 8 + mel_h.numel() * 4),
                "d2h_bytes_per_step": int(wav_h.numel() * 4)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": f"resblock_pair + conv_igemm kernels (vocoder, {voc_launches} launches per step)",
                     "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16_tflops_sustained"], "traffic": vocoder_traffic(),
                     "peak_source": pk["source"] + " sustained bf16 (kernel family timed inside a long step)",
                     "flops_per_step": voc_flops, "vocoder_ms": voc_ms,
                     "vocoder_share_of_step": voc_ms / (ms / args.steps)},
        "roofline_hbm": hbm_info,
        "frontend": fe_info,
        "voice_cached": {"value": world * B_PER_GPU * AUDIO_S_PER_UTT * args.steps / (ms_cached / 1e3), "unit": "audio-s/s",
                         "ms_per_step": ms_cached / args.steps,
                         "note": "style-encoder outputs cached per reference voice (not the BASELINE config: informational)"},
        "mas": mas_info,
    }
    if world == 1 and not args.no_cpu_baseline:
        a, t = cpu_reference_sample(args.cpu_utts, os.cpu_count() or 1)
        line["cpu_baseline"] = {"value": a / t, "unit": "audio-s/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": f"{args.cpu_utts} utterances of the same workload (10 s each), batch-1 "
                                          f"acoustic + vocoder, fp32 torch restatement of the reference"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline", type=int, default=2,
                    help="batches in flight per GPU (2: the acoustic model of step i+1 overlaps the vocoder of step i)")
    ap.add_argument("--cpu-utts", type=int, default=4)
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): libraries that write to fd 1 (NCCL prints its version banner there
    # when NCCL_DEBUG is set) are sent to stderr for the whole run, the JSON line goes to the saved descriptor
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

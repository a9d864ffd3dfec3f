"""Batched text -> waveform synthesis (the glue of the reference's ``test.py:90-119``).

``Synthesizer`` owns an ``ArtsSpeech`` (second stage) and a vocoder ``Generator`` on one GPU and
runs ``tokens + reference mel -> mel -> waveform`` for a batch of utterances, each with batch-1
semantics.  With host-side lengths and durations the whole pass has no host synchronisation and is
captured into a CUDA graph per shape signature (launch-bound otherwise: ~600 kernels per step).

Multi-GPU: utterances are independent; ``shard_utterances`` assigns them to ranks
longest-first (LPT) and ``gather_waveforms`` collects the results on rank 0 over NCCL — the only
collective on the path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

import os

from . import _lib, ops

SAMPLE_RATE = 24000
HOP = 300


class Synthesizer:
    """``pipeline_depth`` > 1 keeps that many calls in flight: call ``i`` runs on side stream ``i % depth``
    with its own CUDA-graph instance and buffers, so the latency-bound acoustic model of one batch
    overlaps the throughput-bound vocoder of the previous one.  The outputs of a pipelined call are
    produced on ``last_stream``: consume them there (``with torch.cuda.stream(syn.last_stream)``) or
    after ``join()``, which makes the current stream wait for every call issued so far.
    ``pcm16=True`` returns int16 samples (``rint(32767 * wav)``, the on-disk format of ``soundfile.write`` at
    test.py:119) written directly by the vocoder's last kernel: half the device->host bytes per utterance."""

    def __init__(self, model, generator, device="cuda:0", use_cuda_graph: bool = True, pipeline_depth: int = 1,
                 pcm16: bool = False, acoustic_sms: Optional[int] = None):
        self.device = torch.device(device)
        self.pcm16 = bool(pcm16)
        # SM split between the two phases when batches overlap: the acoustic model's persistent kernels are sized
        # for ``acoustic_sms`` SMs and the vocoder's for the rest, so that the latency-bound acoustic chain of batch
        # i+1 always finds free SMs beside the vocoder of batch i (0 / None: every kernel sizes for the whole GPU)
        env = os.environ.get("ASB_ACOUSTIC_SMS")
        self.acoustic_sms = int(env) if env is not None else (acoustic_sms or 0)
        self.pipeline_depth = max(1, int(pipeline_depth))
        self._slot_streams = None
        self._slot_done = {}
        self._calls = 0
        self.last_stream = None
        self.model = model.to(self.device).eval()
        self.model.distribution = {k: v.to(self.device) for k, v in self.model.distribution.items()}
        self.generator = generator.to(self.device).eval()
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}
        self.launches_per_call = None

    # ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode_voice(self, mels, mel_lens):
        """Style-encoder pass for a batch of reference recordings: ``mels`` fp32 [B,80,Tr] (device), ``mel_lens``
        int64 [B] (host) -> ``(f0, n, ema, Style)``, reusable as ``synthesize(..., voice=...)`` for every
        utterance spoken with these voices (item i of the batch uses voice i)."""
        ml = [int(v) for v in mel_lens.tolist()]
        return tuple(self.model.style_encoder(mels, mel_lens.to(self.device), "second", self.model.distribution,
                                              host_lengths=ml))

    def _forward(self, tokens, tok_lens, mels, mel_lens, durations, host_meta=None, voice=None):
        split = self.acoustic_sms if self.pipeline_depth > 1 else 0
        lib = _lib.load()
        total = torch.cuda.get_device_properties(self.device).multi_processor_count
        try:
            if split:
                lib.as_set_sm_limit(split)
            mel, aux = self.model([tokens, tok_lens, mels, mel_lens], step="test", durations=durations, return_aux=True,
                                  host_meta=host_meta, voice=voice)
            if split:
                lib.as_set_sm_limit(max(2, total - split))
            wav = self.generator(mel, aux["mel_lengths"], pcm16=self.pcm16)
        finally:
            if split:
                lib.as_set_sm_limit(0)
        return wav.view(wav.shape[0], -1), aux["mel_lengths"], mel

    @torch.no_grad()
    def synthesize(self, tokens, tok_lens, mels, mel_lens, durations=None, voice=None):
        """``tokens`` int64 [B,Tt] (device), ``tok_lens`` / ``mel_lens`` int64 [B] (HOST tensors keep the
        pass sync-free), ``mels`` fp32 [B,80,Tr] (device), ``durations`` int64 [B,Tt] (HOST) or None
        (use the duration predictor; costs one device->host sync).  ``voice``: cached ``encode_voice`` result
        (the style encoder is then skipped).
        Returns (wav fp32 [B, 300*Tm_max], mel_lengths int32 [B] (device), mel fp32 [B,80,Tm_max])."""
        self.last_stream = torch.cuda.current_stream(self.device)
        host_side = durations is not None and not durations.is_cuda and not tok_lens.is_cuda and not mel_lens.is_cuda
        if not host_side:
            return self._forward(tokens, tok_lens, mels, mel_lens, durations, voice=voice)
        tl, ml = [int(v) for v in tok_lens.tolist()], [int(v) for v in mel_lens.tolist()]
        Lmax = max(int(durations[b, :tl[b]].sum()) for b in range(len(tl)))
        meta = {"mel_lens": ml, "Lmax": Lmax}
        graphable = self.use_cuda_graph and all(v == mels.shape[2] for v in ml)
        if not graphable:
            return self._forward(tokens, tok_lens.to(self.device), mels, mel_lens.to(self.device),
                                 durations.to(self.device), meta, voice=voice)
        slot = 0
        cur = torch.cuda.current_stream(self.device)
        run_stream = cur
        if self.pipeline_depth > 1:
            if self._slot_streams is None:
                self._slot_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.pipeline_depth)]
            slot = self._calls % self.pipeline_depth
            run_stream = self._slot_streams[slot]
        self._calls += 1
        key = (slot, voice is not None, tuple(tokens.shape), tuple(mels.shape), tuple(tl), tuple(ml),
               hash(durations.numpy().tobytes()))
        entry = self._graphs.get(key)
        if entry is None:
            static_tok = tokens.clone()
            static_mel = mels.clone()
            static_voice = None if voice is None else tuple(v.clone() for v in voice)
            tok_lens, mel_lens = tok_lens.to(self.device), mel_lens.to(self.device)
            durations = durations.to(self.device)
            # warm-up on a side stream (weight packing, cudaFuncSetAttribute, allocator)
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._forward(static_tok, tok_lens, static_mel, mel_lens, durations, meta, voice=static_voice)
            cur.wait_stream(s)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            before = ops.launch_count
            with torch.cuda.graph(g):
                out = self._forward(static_tok, tok_lens, static_mel, mel_lens, durations, meta, voice=static_voice)
            self.launches_per_call = ops.launch_count - before
            # the captured kernels read these device tensors on every replay: keep them alive with the graph
            entry = (g, static_tok, static_mel, out, (tok_lens, mel_lens, durations), static_voice)
            self._graphs[key] = entry
        g, static_tok, static_mel, out, _keepalive, static_voice = entry
        if run_stream is not cur:
            # the slot's stream picks up after whatever produced the inputs on the caller's stream;
            # the caller's stream never waits for the slot (that would serialise the pipeline)
            ready = torch.cuda.Event()
            ready.record(cur)
            run_stream.wait_event(ready)
            tokens.record_stream(run_stream)
            mels.record_stream(run_stream)
            for v in (voice or ()):
                v.record_stream(run_stream)
        with torch.cuda.stream(run_stream):
            static_tok.copy_(tokens, non_blocking=True)
            if static_voice is None:
                static_mel.copy_(mels, non_blocking=True)
            else:
                for dst, src in zip(static_voice, voice):
                    dst.copy_(src, non_blocking=True)
            g.replay()
            if run_stream is not cur:
                done = torch.cuda.Event()
                done.record(run_stream)
                self._slot_done[slot] = done
        self.last_stream = run_stream
        ops._count(self.launches_per_call)
        return out

    def join(self):
        """Make the current stream wait for every pipelined call issued so far."""
        cur = torch.cuda.current_stream(self.device)
        for ev in self._slot_done.values():
            cur.wait_event(ev)


# -------------------------------------------------------------------------------------------
# data-parallel helpers
# -------------------------------------------------------------------------------------------
def shard_utterances(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of utterances to ranks (cost ~ frames to synthesise)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += costs[i]
    return shards


def bucket_utterances(frames: Sequence[int], max_batch: int = 16, max_padded_frames: Optional[int] = None) -> List[List[int]]:
    """Length-bucketed micro-batches for a mixed-length shard: utterances sorted by frame count,
    longest first, cut into runs of at most ``max_batch`` whose PADDED size ``len(run) * max(frames in run)`` stays
    within ``max_padded_frames`` -- neighbours in the sorted order have similar lengths, so padding stays small and
    every micro-batch costs about the same.  Every index appears exactly once; a single utterance longer than the
    budget gets a micro-batch of its own."""
    order = sorted(range(len(frames)), key=lambda i: (-int(frames[i]), i))
    batches: List[List[int]] = []
    cur: List[int] = []
    cur_max = 0
    for i in order:
        f = max(int(frames[i]), 1)
        longest = max(cur_max, f)
        over = max_padded_frames is not None and cur and (len(cur) + 1) * longest > max_padded_frames
        if cur and (len(cur) >= max_batch or over):
            batches.append(cur)
            cur, longest = [], f
        cur.append(i)
        cur_max = longest
    if cur:
        batches.append(cur)
    return batches


@torch.no_grad()
def synthesize_many(syn: "Synthesizer", tokens: Sequence[torch.Tensor], ref_mels: Sequence[torch.Tensor],
                    durations: Sequence[torch.Tensor], max_batch: int = 16, max_padded_frames: Optional[int] = 25600):
    """Mixed-length utterances through ``syn`` in length-bucketed ragged micro-batches (BASELINE config 5).

    ``tokens[i]`` int64 [Tt_i], ``ref_mels[i]`` fp32 [80, Tr_i], ``durations[i]`` int64 [Tt_i] (host tensors).
    Returns ``(wavs, frames)``: ``wavs[i]`` is a device tensor with the ``300 * frames[i]`` samples of utterance
    ``i`` (fp32, or int16 when ``syn.pcm16``), independent of what it was batched with (batch-1 semantics)."""
    n = len(tokens)
    frames = [2 * int(d.sum()) for d in durations]
    wavs: List[Optional[torch.Tensor]] = [None] * n
    for idx in bucket_utterances(frames, max_batch, max_padded_frames):
        Tt = max(int(tokens[i].shape[0]) for i in idx)
        Tr = max(int(ref_mels[i].shape[1]) for i in idx)
        tok = torch.zeros(len(idx), Tt, dtype=torch.long)
        dur = torch.zeros(len(idx), Tt, dtype=torch.long)
        mel = torch.zeros(len(idx), ref_mels[idx[0]].shape[0], Tr)
        for j, i in enumerate(idx):
            tok[j, :tokens[i].shape[0]] = tokens[i]
            dur[j, :durations[i].shape[0]] = durations[i]
            mel[j, :, :ref_mels[i].shape[1]] = ref_mels[i]
        tl = torch.tensor([int(tokens[i].shape[0]) for i in idx])
        ml = torch.tensor([int(ref_mels[i].shape[1]) for i in idx])
        wav, _, _ = syn.synthesize(tok.to(syn.device, non_blocking=True), tl, mel.to(syn.device, non_blocking=True), ml, dur)
        for j, i in enumerate(idx):
            wavs[i] = wav[j, :HOP * frames[i]].clone()
    return wavs, frames


def gather_waveforms(wav: torch.Tensor, lengths: torch.Tensor, dst: int = 0, group=None, shapes=None):
    """Gather per-rank ``wav`` [B_r, S_r] (+ sample ``lengths`` [B_r]) on ``dst``.

    Ranks may hold different batch sizes / paddings: sizes are all-gathered first, payloads are
    padded to the common maximum for one ``gather``.  ``shapes`` (list of ``(B_r, S_r)`` per rank, known on
    the host — e.g. from ``shard_utterances``) skips that exchange and its host synchronisation, which keeps
    a pipelined engine asynchronous.  Works with NCCL (GPU) and gloo (CPU tests).
    Returns (list of [B_r, S_r] tensors, list of lengths) on ``dst``, (None, None) elsewhere."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if shapes is None:
        shape = torch.tensor([wav.shape[0], wav.shape[1]], dtype=torch.int64, device=wav.device)
        gathered = [torch.zeros_like(shape) for _ in range(world)]
        dist.all_gather(gathered, shape, group=group)
        shapes = [tuple(int(v) for v in s.tolist()) for s in gathered]
    else:
        shapes = [tuple(int(v) for v in s) for s in shapes]
        assert len(shapes) == world and shapes[rank] == tuple(wav.shape)
    Bm, Sm = max(s[0] for s in shapes), max(s[1] for s in shapes)
    if tuple(wav.shape) == (Bm, Sm) and wav.is_contiguous():
        pad = wav
    else:
        pad = torch.zeros(Bm, Sm, dtype=wav.dtype, device=wav.device)
        pad[:wav.shape[0], :wav.shape[1]] = wav
    lpad = torch.zeros(Bm, dtype=torch.int64, device=wav.device)
    lpad[:lengths.shape[0]] = lengths.to(torch.int64)
    if rank == dst:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        lbufs = [torch.empty_like(lpad) for _ in range(world)]
    else:
        bufs = lbufs = None
    dist.gather(pad, bufs, dst=dst, group=group)
    dist.gather(lpad, lbufs, dst=dst, group=group)
    if rank != dst:
        return None, None
    return ([b[:s[0], :s[1]] for b, s in zip(bufs, shapes)], [l[:s[0]] for l, s in zip(lbufs, shapes)])

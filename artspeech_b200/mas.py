"""Monotonic alignment search behind the reference's function names.

Drop-in for ``S_monotonic_align.py`` (``maximum_path1`` :5-47, ``maximum_path2`` :50-95,
``mask_from_len(s)`` :99-133) and ``S_monotonic_align_Triton.py`` (``maximum_path`` :56-71).
All three entry points run the same sm_100a kernel (``as_mas_maximum_path``); they differ only in
the tie rule the reference versions use:
  * ``maximum_path1``            — ties move to ``x-1``   (``where(a > b, 0, -1)``, :40)
  * ``maximum_path2``/``maximum_path`` — ties stay          (strict ``>`` / ``<``, :91 / Triton :37)
Unlike the Triton version the input ``value`` is never modified.
"""
from __future__ import annotations

import torch

from . import _lib

TIE_STAY, TIE_MOVE = 0, 1


@torch.no_grad()
def mask_from_len(lens: torch.Tensor, max_len=None):
    """``[B] -> [B, max_len]`` boolean mask, ``index < len`` (S_monotonic_align.py:99-114)."""
    if max_len is None:
        max_len = int(lens.max())
    index = torch.arange(max_len, device=lens.device).to(lens).view(1, -1)
    return index < lens.unsqueeze(1)


@torch.no_grad()
def mask_from_lens(similarity: torch.Tensor, symbol_lens: torch.Tensor, mel_lens: torch.Tensor):
    """``[B,S,T]`` mask of valid (symbol, frame) cells in ``similarity``'s dtype (:117-133)."""
    _, S, T = similarity.size()
    mask_s = mask_from_len(symbol_lens, S)
    mask_t = mask_from_len(mel_lens, T)
    return (mask_s.unsqueeze(2) & mask_t.unsqueeze(1)).to(similarity)


@torch.no_grad()
def maximum_path_lens(value: torch.Tensor, x_len: torch.Tensor, y_len: torch.Tensor,
                      tie_mode: int = TIE_STAY, dtype=torch.float32) -> torch.Tensor:
    """MAS on explicit lengths.  ``value`` [B,Tx,Ty] (cuda); returns the 0/1 path [B,Tx,Ty]."""
    if not value.is_cuda:
        raise _lib.AsError("artspeech_b200.mas requires CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    v = value.detach()
    if v.dtype != torch.float32:
        v = v.float()
    v = v.contiguous()
    B, Tx, Ty = v.shape
    xl = x_len.to(device=v.device, dtype=torch.int32).contiguous()
    yl = y_len.to(device=v.device, dtype=torch.int32).contiguous()
    path = torch.empty_like(v)
    ws_bytes = lib.as_mas_workspace_bytes(B, Tx, Ty)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=v.device)
    stream = torch.cuda.current_stream(v.device).cuda_stream
    with torch.cuda.device(v.device):
        rc = lib.as_mas_maximum_path(v.data_ptr(), xl.data_ptr(), yl.data_ptr(), path.data_ptr(),
                                     B, Tx, Ty, int(tie_mode), ws.data_ptr(), ws_bytes, stream)
    _lib.check(rc, "as_mas_maximum_path")
    return path if dtype == torch.float32 else path.to(dtype)


@torch.no_grad()
def align_lens(value: torch.Tensor, x_len: torch.Tensor, y_len: torch.Tensor, tie_mode: int = TIE_STAY):
    """MAS plus what its training callers derive from the path, from ONE kernel launch
    (train_second.py:181-187, train_first.py:176-181): returns ``(path, durations, token_of_frame)`` with
    ``durations[B,Tx] int32 == path.sum(-1)`` (``d_gt``) and ``token_of_frame[B,Ty] int32`` = the row the path
    occupies in each column (-1 for ``y >= y_len``)."""
    if not value.is_cuda:
        raise _lib.AsError("artspeech_b200.mas requires CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    v = value.detach()
    if v.dtype != torch.float32:
        v = v.float()
    v = v.contiguous()
    B, Tx, Ty = v.shape
    xl = x_len.to(device=v.device, dtype=torch.int32).contiguous()
    yl = y_len.to(device=v.device, dtype=torch.int32).contiguous()
    path = torch.empty_like(v)
    dur = torch.empty(B, Tx, dtype=torch.int32, device=v.device)
    tok = torch.empty(B, Ty, dtype=torch.int32, device=v.device)
    ws_bytes = lib.as_mas_workspace_bytes(B, Tx, Ty)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=v.device)
    stream = torch.cuda.current_stream(v.device).cuda_stream
    with torch.cuda.device(v.device):
        rc = lib.as_mas_align(v.data_ptr(), xl.data_ptr(), yl.data_ptr(), path.data_ptr(), dur.data_ptr(),
                              tok.data_ptr(), B, Tx, Ty, int(tie_mode), ws.data_ptr(), ws_bytes, stream)
    _lib.check(rc, "as_mas_align")
    return path, dur, tok


@torch.no_grad()
def expand_tokens(x: torch.Tensor, token_of_frame: torch.Tensor) -> torch.Tensor:
    """``x[B,C,Tx] @ path[B,Tx,Ty]`` for the 0/1 path that ``token_of_frame`` describes, as a gather
    (models.py:296,323-324: ``T_en @ s2s_attn_mono``, ``A_en @ s2s_attn_mono``).  Bit-identical to the fp32
    matmul: every output is one input value (or 0 past ``y_len``)."""
    if not x.is_cuda:
        raise _lib.AsError("artspeech_b200.mas requires CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    xf = x.detach().float().contiguous()
    B, C, Tx = xf.shape
    tok = token_of_frame.to(device=xf.device, dtype=torch.int32).contiguous()
    Ty = tok.shape[1]
    out = torch.empty(B, C, Ty, dtype=torch.float32, device=xf.device)
    stream = torch.cuda.current_stream(xf.device).cuda_stream
    with torch.cuda.device(xf.device):
        rc = lib.as_expand_tokens(xf.data_ptr(), tok.data_ptr(), out.data_ptr(), B, C, Tx, Ty, stream)
    _lib.check(rc, "as_expand_tokens")
    return out.to(x.dtype)


@torch.no_grad()
def align(value: torch.Tensor, mask: torch.Tensor, tie_mode: int = TIE_STAY):
    """``align_lens`` on the reference's mask argument (``mask_from_lens(...)``, S_monotonic_align.py:117-133)."""
    x_len, y_len = _lens_from_mask(mask)
    return align_lens(value, x_len, y_len, tie_mode)


def _lens_from_mask(mask: torch.Tensor):
    # x_len = mask[:, :, 0].sum(1), y_len = mask[:, 0, :].sum(1) (S_monotonic_align.py:14-15)
    m = mask != 0
    return m[:, :, 0].sum(dim=1), m[:, 0, :].sum(dim=1)


@torch.no_grad()
def maximum_path1(logp: torch.Tensor, attn_mask: torch.Tensor) -> torch.Tensor:
    x_len, y_len = _lens_from_mask(attn_mask)
    return maximum_path_lens(logp, x_len, y_len, TIE_MOVE).to(logp.dtype)


@torch.no_grad()
def maximum_path2(logp: torch.Tensor, attn_mask: torch.Tensor) -> torch.Tensor:
    x_len, y_len = _lens_from_mask(attn_mask)
    return maximum_path_lens(logp, x_len, y_len, TIE_STAY).to(logp.dtype)


@torch.no_grad()
def maximum_path(value: torch.Tensor, mask: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """Signature of the Triton wrapper (S_monotonic_align_Triton.py:56-71)."""
    x_len, y_len = _lens_from_mask(mask)
    return maximum_path_lens(value, x_len, y_len, TIE_STAY, dtype=dtype)

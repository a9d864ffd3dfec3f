"""TEST INFRASTRUCTURE — numpy restatement of the reference's MAS (``maximum_path``).

Follows S_monotonic_align.py: the cumulative pass of ``maximum_path1`` (:22-30) /
``cumulative_logp`` (:52-77) — identical in both versions — and the two back-tracking rules:
``tie='stay'`` = ``maximum_path2`` (:84-93, also the Triton kernel S_monotonic_align_Triton.py:31-38),
``tie='move'`` = ``maximum_path1`` (:32-46).

Pinned against the reference itself by ``tests/test_oracle_pinning.py`` (where /root/reference is
mounted) and against ``tests/golden/mas_*.npz`` (generated from the reference by
``oracle/make_golden.py``) everywhere else.
"""
from __future__ import annotations

import numpy as np

NEG = np.float32(-1e32)


def cumulative(value: np.ndarray, x_len: np.ndarray, y_len: np.ndarray) -> np.ndarray:
    """fp32 cumulative scores, exactly S_monotonic_align.py:11-30 (masking + column sweep)."""
    B, Tx, Ty = value.shape
    ix = np.arange(Tx)[None, :, None] < x_len[:, None, None]
    iy = np.arange(Ty)[None, None, :] < y_len[:, None, None]
    q = (value.astype(np.float32) * (ix & iy).astype(np.float32)).astype(np.float32)  # logp * mask
    q[:, 1:, 0] = NEG
    for ty in range(1, Ty):
        prev1 = q[:, :, ty - 1]
        prev2 = np.roll(prev1, 1, axis=1)
        prev2[:, 0] = NEG
        q[:, :, ty] += np.where(prev1 > prev2, prev1, prev2)
    return q


def maximum_path(value: np.ndarray, x_len, y_len, tie: str = "stay") -> np.ndarray:
    assert tie in ("stay", "move")
    x_len = np.asarray(x_len).astype(np.int64)
    y_len = np.asarray(y_len).astype(np.int64)
    B, Tx, Ty = value.shape
    q = cumulative(value, x_len, y_len)
    path = np.zeros((B, Tx, Ty), dtype=np.float32)
    for b in range(B):
        if x_len[b] <= 0 or y_len[b] <= 0:
            continue
        idx = int(x_len[b]) - 1
        path[b, idx, y_len[b] - 1] = 1
        for ty in range(int(y_len[b]) - 1, 0, -1):
            if idx != 0:
                same, diag = q[b, idx, ty - 1], q[b, idx - 1, ty - 1]
                if tie == "stay":
                    move = diag > same          # maximum_path2 :91
                else:
                    move = not (same > diag)    # maximum_path1 :40 (where(a > b, 0, -1))
                if move:
                    idx -= 1
            path[b, idx, ty - 1] = 1
    return path

"""Ragged-list wire format of the dataset tensors (``utils/misc.py:56-89``).

``seen_entities_tensor`` stores, per prefix, the list of answers where each answer is itself a list
of alternative mention ids, packed as  [offsets..., 0, values...]  with absolute offsets into the
packed array: [[5],[6,7],[8]] <-> [5,6,8,9,0,5,6,7,8]."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def pack_list_of_lists(lol) -> List[int]:
    lens = [len(x) if isinstance(x, (list, tuple)) else 1 for x in lol]
    header = len(lens) + 2                     # n+1 offsets and the 0 terminator
    offsets = (np.concatenate([[0], np.cumsum(lens, dtype=np.int64)]) + header).astype(np.int64).tolist()
    values: List[int] = []
    for x in lol:
        if isinstance(x, (list, tuple)):
            values.extend(int(v) for v in x)
        else:
            values.append(int(x))
    return offsets + [0] + values


def unpack_list_of_lists(packed: Sequence[int]) -> Tuple[List[List[int]], List[int]]:
    arr = np.asarray(packed, dtype=np.int64).reshape(-1)
    if arr.size == 0:
        return [], []
    zero = int(np.argmax(arr == 0))            # terminator follows the offsets
    offsets = arr[:zero]
    lol = [arr[offsets[i]:offsets[i + 1]].tolist() for i in range(len(offsets) - 1)]
    flat = arr[offsets[0]:offsets[-1]].tolist() if len(offsets) > 1 else []
    return lol, flat


def unpack_offsets(packed: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Array form: (lengths of each alternative list, flat values)."""
    arr = np.asarray(packed, dtype=np.int64).reshape(-1)
    if arr.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    zero = int(np.argmax(arr == 0))
    offsets = arr[:zero]
    if len(offsets) < 2:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.diff(offsets), arr[offsets[0]:offsets[-1]]

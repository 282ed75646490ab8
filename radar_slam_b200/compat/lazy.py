"""Column-backed stand-ins for the reference's list-of-dict stage outputs.

The reference hands results from stage to stage as Python lists of dicts -- 'peaks' (dechirp.py:265-272, 6 keys),
'targets' (angle_estimation.py:289-300, 10 keys) -- and the pipeline script pickles them into .npz files
(run_ego_motion_pipeline.py:164-169, 218-219).  At ~24 000 detections per frame, building and re-parsing those
dicts costs more than every kernel of the path together.  LazyRecords keeps the columns the CUDA path produced
(one numpy array per key) and only builds a dict when somebody indexes or iterates it, so

  * code written against the reference keeps working: len(), [i], slices, iteration, ==, 'key' in rec[i];
  * the drop-in classes read the columns directly (column(name)) and never touch a dict;
  * np.savez(..., peaks=lazy) stores the columns: __array__ wraps the object in a 0-d object array and pickling goes
    through __reduce__.  np.load(..., allow_pickle=True)[key] then yields that 0-d array; records_of() unwraps it
    (and also accepts the reference's 1-D object arrays and plain lists).  Set RADAR_SLAM_B200_EAGER_FILES=1 to make
    __array__ materialise the reference's 1-D object array of dicts instead, for consumers outside this package.
"""
from __future__ import annotations

import os
from collections.abc import Sequence
from typing import Dict, Iterable, Iterator, List, Optional, Tuple

import numpy as np


class LazyRecords(Sequence):
    __slots__ = ("_cols", "_keys", "_n", "_none")

    def __init__(self, columns: Dict[str, Optional[np.ndarray]], keys: Optional[Iterable[str]] = None):
        self._keys: Tuple[str, ...] = tuple(keys) if keys is not None else tuple(columns)
        self._cols = {k: columns[k] for k in self._keys}
        self._none = tuple(k for k in self._keys if self._cols[k] is None)     # keys whose value is None in every record
        lens = {len(v) for v in self._cols.values() if v is not None}
        if len(lens) > 1:
            raise ValueError(f"columns of different lengths: {sorted(lens)}")
        self._n = lens.pop() if lens else 0

    # ---- sequence protocol
    def __len__(self) -> int:
        return self._n

    def _record(self, i: int) -> dict:
        return {k: (None if v is None else v[i]) for k, v in self._cols.items()}

    def __getitem__(self, i):
        if isinstance(i, slice):
            return LazyRecords({k: (None if v is None else v[i]) for k, v in self._cols.items()}, self._keys)
        if isinstance(i, (np.ndarray, list)):
            idx = np.asarray(i)
            return LazyRecords({k: (None if v is None else v[idx]) for k, v in self._cols.items()}, self._keys)
        i = int(i)
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError("record index out of range")
        return self._record(i)

    def __iter__(self) -> Iterator[dict]:
        for i in range(self._n):
            yield self._record(i)

    def __bool__(self) -> bool:
        return self._n > 0

    def __eq__(self, other) -> bool:
        if not isinstance(other, (Sequence, np.ndarray)) or len(other) != self._n:
            return False
        for mine, theirs in zip(self, other):
            if set(mine) != set(theirs):
                return False
            for k, v in mine.items():
                w = theirs[k]
                if v is None or w is None:
                    if v is not w:
                        return False
                elif not np.array_equal(np.asarray(v), np.asarray(w)):
                    return False
        return True

    def __repr__(self) -> str:
        return f"LazyRecords({self._n} records, keys={list(self._keys)})"

    # ---- column access (the fast path of the drop-in classes)
    def keys(self) -> Tuple[str, ...]:
        return self._keys

    def column(self, name: str) -> Optional[np.ndarray]:
        return self._cols[name]

    def has_column(self, name: str) -> bool:
        return name in self._cols

    def tolist(self) -> List[dict]:
        return [self._record(i) for i in range(self._n)]

    # ---- numpy / pickle interop
    def __array__(self, dtype=None, copy=None):
        if os.environ.get("RADAR_SLAM_B200_EAGER_FILES") == "1":
            out = np.empty(self._n, dtype=object)
            for i in range(self._n):
                out[i] = self._record(i)
            return out
        out = np.empty((), dtype=object)
        out[()] = self
        return out

    def __reduce__(self):
        return (LazyRecords, (self._cols, self._keys))


def records_of(obj):
    """Whatever a stage received as 'peaks' / 'targets' -> a sequence of records: LazyRecords (possibly wrapped in the
    0-d object array np.load returns), the reference's 1-D object array of dicts, or a plain list."""
    if isinstance(obj, np.ndarray):
        if obj.ndim == 0:
            obj = obj.item()
        else:
            return obj.tolist()
    if isinstance(obj, LazyRecords):
        return obj
    return list(obj)


def column_of(records, name: str, dtype=None) -> np.ndarray:
    """One key of every record as an array: free for LazyRecords, a Python loop for lists of dicts."""
    if isinstance(records, LazyRecords) and records.has_column(name) and records.column(name) is not None:
        col = records.column(name)
        return col if dtype is None else np.asarray(col, dtype=dtype)
    vals = [r[name] for r in records]
    return np.array(vals, dtype=dtype) if dtype is not None else np.array(vals)

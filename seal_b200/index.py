"""Drop-in for ``seal.index`` (/root/reference/seal/index.py:20-204): class ``FMIndex``.

Same attributes (beginnings, occurring, occurring_distinct, occurring_counts, labels), same
methods, same return values; every FM-index query is answered by the CUDA kernels behind
include/sealfm.h.  Extra ``*_batch`` methods expose the batched kernels to callers that want them.
"""
import bisect
import pickle
import struct
import tempfile
from typing import Iterable, Iterator, List, Optional, Set, Tuple

import numpy as np

from ._lib import lib, check
from .cpp_modules.fm_index import FMIndex as _FMIndex
from .cpp_modules.fm_index import load_FMIndex

SHIFT = 10            # index.py:16
FORMAT = "<l"         # index.py:18


class FMIndex(_FMIndex):

    beginnings: List[int]
    occurring: Set[int]
    occurring_distinct: List[int]
    occurring_counts: List[int]
    labels: Optional[List[str]]

    def __init__(self):
        super().__init__()
        self.beginnings = [0]
        self.occurring = set()
        self.occurring_distinct = []
        self.occurring_counts = []
        self.labels = None

    def initialize(self, sequences: Iterable[List[int]], in_memory: bool = False) -> None:   # index.py:39-66
        occurring = set()
        if in_memory:
            data = []
            for seq in sequences:
                self.beginnings.append(self.beginnings[-1] + len(seq))
                occurring |= set(seq)
                data.extend(x + SHIFT for x in seq[::-1])
            self.occurring = list(occurring)
            super().initialize(data)
        else:
            with tempfile.NamedTemporaryFile() as tmp:
                for seq in sequences:
                    self.beginnings.append(self.beginnings[-1] + len(seq))
                    occurring |= set(seq)
                    arr = np.asarray(seq[::-1], dtype="<i4") + SHIFT
                    tmp.write(arr.astype("<i4").tobytes())
                tmp.flush()
                self.occurring = list(occurring)
                super().initialize_from_file(tmp.name, 4)
        self._sync_beginnings()
        self.occurring_distinct, self.occurring_counts = self.get_distinct_count(0, len(self))

    def _sync_beginnings(self):
        b = np.asarray(self.beginnings, dtype=np.uint64)
        check(lib.sealfm_set_beginnings(self._handle(), b.ctypes.data, len(b)))

    def get_doc(self, doc_index: int) -> List[int]:                                      # index.py:68-75
        doc = self.extract_text(self.beginnings[doc_index], self.beginnings[doc_index + 1])
        return [x - SHIFT for x in doc]

    def get_doc_index(self, token_index: int) -> int:                                    # index.py:77-82
        return bisect.bisect_right(self.beginnings, token_index) - 1

    def get_doc_length(self, doc_index: int) -> int:                                     # index.py:84-88
        return self.beginnings[doc_index + 1] - self.beginnings[doc_index]

    def get_token_index_from_row(self, row: int) -> int:                                 # index.py:90-94
        return self.locate(row)

    def get_doc_index_from_row(self, row: int) -> int:                                   # index.py:96-100
        return self.get_doc_index(self.locate(row))

    def get_range(self, sequence: List[int]) -> Tuple[int, int]:                         # index.py:102-111
        # the reference folds backward_search_step over the tokens starting from (0, size());
        # backward_search_multi is that very fold (fm_index.cpp:55-65) in one kernel launch.
        lo, hi = self.backward_search_multi([t + SHIFT for t in sequence])
        return lo, hi

    def get_count(self, sequence: List[int]) -> int:                                     # index.py:113-118
        start, end = self.get_range(sequence)
        return end - start

    def get_doc_indices(self, sequence: List[int]) -> Iterator[int]:                     # index.py:120-126
        start, end = self.get_range(sequence)
        if end > start:
            rows = np.arange(start, end, dtype=np.uint64)
            out = np.zeros(len(rows), dtype=np.uint64)
            check(lib.sealfm_doc_index_from_rows(self._dev(), len(rows), rows.ctypes.data, out.ctypes.data))
            for d in out.tolist():
                yield d

    def get_continuations(self, sequence: List[int]) -> List[int]:                       # index.py:128-134
        start, end = self.get_range(sequence)
        return self.get_distinct(start, end)

    def get_distinct(self, low: int, high: int) -> List[int]:                            # index.py:136-141
        return [c - SHIFT for c in self.distinct(low, high) if c > 0]

    def get_distinct_count(self, low: int, high: int) -> Tuple[List[int], List[int]]:    # index.py:143-156
        return self.get_distinct_count_multi([low], [high])[0]

    def get_distinct_count_multi(self, lows: List[int], highs: List[int]):               # index.py:158-171
        ret = []
        for data in self.distinct_count_multi(lows, highs):
            distinct, counts = [], []
            for d, c in zip(data[0::2], data[1::2]):
                if d > 0:
                    distinct.append(d - SHIFT)
                    counts.append(c)
            ret.append((distinct, counts))
        return ret

    def __len__(self) -> int:                                                            # index.py:173-177
        return self.beginnings[-1]

    @property
    def n_docs(self) -> int:                                                             # index.py:179-184
        return len(self.beginnings) - 1

    def save(self, path: str) -> None:                                                   # index.py:186-193
        with open(path + ".oth", "wb") as f:
            pickle.dump((self.beginnings, self.occurring, self.labels), f)
        return super().save(path + ".fmi")

    @classmethod
    def load(cls, path: str) -> "FMIndex":                                               # index.py:195-204
        index = load_FMIndex(path + ".fmi")
        index.__class__ = cls
        with open(path + ".oth", "rb") as f:
            index.beginnings, index.occurring, index.labels = pickle.load(f)
        index._sync_beginnings()
        index.occurring_distinct, index.occurring_counts = index.get_distinct_count(0, len(index))
        return index

    # ---- batched extensions ------------------------------------------------------------------------
    def get_range_batch(self, sequences):
        lo, hi = self.backward_search_multi_batch([[t + SHIFT for t in s] for s in sequences])
        return lo, hi

    def get_doc_index_from_rows(self, rows):
        r = np.ascontiguousarray(np.asarray(rows, dtype=np.uint64)); out = np.zeros(len(r), dtype=np.uint64)
        check(lib.sealfm_doc_index_from_rows(self._dev(), len(r), r.ctypes.data, out.ctypes.data))
        return out

    def locate_rows(self, rows):
        """rows -> (token positions, document ids) as uint64 arrays: locate + bisect over `beginnings`
        (index.py:77-82,90-100) for a whole batch in one launch."""
        pos = self.locate_batch(rows)
        b = np.asarray(self.beginnings, dtype=np.uint64)
        return pos, (np.searchsorted(b, pos, side="right").astype(np.int64) - 1)

    def get_docs(self, doc_indices):
        """[get_doc(d) for d in doc_indices] (index.py:68-75) with one extraction launch."""
        b = [self.beginnings[d] for d in doc_indices]; e = [self.beginnings[d + 1] for d in doc_indices]
        if not b:
            return []
        return [(t.astype(np.int64) - SHIFT).tolist() for t in self.extract_text_batch(b, e)]

    def get_docs_arrays(self, doc_indices):
        """get_docs without the list conversion: int64 arrays of token ids."""
        b = [self.beginnings[d] for d in doc_indices]; e = [self.beginnings[d + 1] for d in doc_indices]
        if not b:
            return []
        return [t.astype(np.int64) - SHIFT for t in self.extract_text_batch(b, e)]

    def prefix_allowed_tokens_fn(self):
        """fairseq/GENRE-style hook named by BASELINE.json:north_star:
        prefix_allowed_tokens_fn(batch_id, input_ids) -> List[int] (index.py:128-134 semantics)."""
        def fn(batch_id, input_ids):
            ids = input_ids.tolist() if hasattr(input_ids, "tolist") else list(input_ids)
            return self.get_continuations(ids[1:])
        return fn

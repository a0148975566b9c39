"""TEST INFRASTRUCTURE ONLY — never imported by the product package (seal_b200/).

ctypes front-ends for the two CPU oracles of the FM-index path:

* ``PortFM``  -> oracle/liboracle_fm.so, the plain-C restatement (oracle/fm_oracle.c).
* ``RefFM``   -> oracle/_ref/libseal_ref*.so, the UNMODIFIED reference class
  (seal/cpp_modules/fm_index.cpp + vendored sdsl-lite) behind oracle/ref_shim.cpp.  Only exists
  where it was built (this container; it travels to the GPU box as a prebuilt .so).

``OracleIndex`` restates seal/index.py (SHIFT, reversal, get_range, get_count,
get_distinct_count[_multi], locate, get_doc ...) on top of either back-end, so decode-loop oracles
and tests can talk to "the reference FMIndex" without /root/reference being present.
"""
import bisect
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIFT = 10  # seal/index.py:16

_u64 = C.c_uint64
_vp = C.c_void_p


def _np_u64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


def build_port():
    """Compile oracle/fm_oracle.c -> oracle/liboracle_fm.so (gcc only)."""
    so = os.path.join(_HERE, "liboracle_fm.so")
    src = os.path.join(_HERE, "fm_oracle.c")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle_fm.so"], stdout=subprocess.DEVNULL)
    return so


def ref_available(popcnt=False):
    return os.path.exists(os.path.join(_HERE, "_ref", "libseal_ref_popcnt.so" if popcnt else "libseal_ref.so"))


class _Base:
    """Common numpy-level interface. Symbols are raw (already SHIFTed) u64."""

    def size(self):
        raise NotImplementedError

    def backward_search_step(self, sym, lo, hi):
        raise NotImplementedError

    def backward_search_step_batch(self, sym, lo, hi):
        sym, lo, hi = _np_u64(sym), _np_u64(lo), _np_u64(hi)
        ol = np.empty_like(sym); oh = np.empty_like(sym)
        for i in range(len(sym)):
            ol[i], oh[i] = self.backward_search_step(int(sym[i]), int(lo[i]), int(hi[i]))
        return ol, oh


class PortFM(_Base):
    def __init__(self, text):
        self.L = C.CDLL(build_port())
        L = self.L
        L.fmo_build.restype = _vp; L.fmo_build.argtypes = [_vp, _u64]
        L.fmo_free.argtypes = [_vp]
        for nm in ("fmo_size", "fmo_sigma"):
            getattr(L, nm).restype = _u64; getattr(L, nm).argtypes = [_vp]
        L.fmo_max_level.restype = C.c_uint32; L.fmo_max_level.argtypes = [_vp]
        L.fmo_bv_rank.restype = _u64; L.fmo_bv_rank.argtypes = [_vp, _u64]
        L.fmo_wt_rank.restype = _u64; L.fmo_wt_rank.argtypes = [_vp, _u64, _u64]
        L.fmo_backward_search_step.argtypes = [_vp, _u64, _u64, _u64, _vp]
        L.fmo_backward_search_multi.argtypes = [_vp, _vp, _u64, _vp]
        L.fmo_distinct_count.restype = _u64; L.fmo_distinct_count.argtypes = [_vp, _u64, _u64, _vp, _u64]
        L.fmo_visited_nodes.restype = _u64; L.fmo_visited_nodes.argtypes = [_vp, _u64, _u64]
        L.fmo_locate.restype = _u64; L.fmo_locate.argtypes = [_vp, _u64]
        L.fmo_extract_text.restype = _u64; L.fmo_extract_text.argtypes = [_vp, _u64, _u64, _vp, _u64]
        for nm in ("fmo_tree_words", "fmo_rank_blocks", "fmo_sa_samples", "fmo_isa_samples",
                   "fmo_alphabet", "fmo_C", "fmo_bwt", "fmo_sa"):
            getattr(L, nm).restype = C.POINTER(_u64); getattr(L, nm).argtypes = [_vp, _vp]
        t = _np_u64(text)
        self.h = L.fmo_build(t.ctypes.data, len(t))
        if not self.h:
            raise MemoryError("fmo_build failed")
        self._buf = np.zeros(2 * (int(self.sigma()) + 2), dtype=np.uint64)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.fmo_free(self.h); self.h = None

    def size(self): return int(self.L.fmo_size(self.h))
    def sigma(self): return int(self.L.fmo_sigma(self.h))
    def max_level(self): return int(self.L.fmo_max_level(self.h))
    def bv_rank(self, idx): return int(self.L.fmo_bv_rank(self.h, idx))
    def wt_rank(self, i, c): return int(self.L.fmo_wt_rank(self.h, i, c))

    def backward_search_step(self, sym, lo, hi):
        out = (_u64 * 2)()
        self.L.fmo_backward_search_step(self.h, sym, lo, hi, out)
        return int(out[0]), int(out[1])

    def backward_search_multi(self, q):
        q = _np_u64(q); out = (_u64 * 2)()
        self.L.fmo_backward_search_multi(self.h, q.ctypes.data, len(q), out)
        return int(out[0]), int(out[1])

    def distinct_count(self, lo, hi):
        k = self.L.fmo_distinct_count(self.h, lo, hi, self._buf.ctypes.data, len(self._buf))
        return self._buf[:k].copy()

    def visited_nodes(self, lo, hi): return int(self.L.fmo_visited_nodes(self.h, lo, hi))
    def locate(self, row): return int(self.L.fmo_locate(self.h, row))

    def extract_text(self, begin, end):
        n = max(int(end) - int(begin), 0)
        buf = np.zeros(n + 1, dtype=np.uint64)
        k = self.L.fmo_extract_text(self.h, begin, end, buf.ctypes.data, len(buf))
        return buf[:k].copy()

    def section(self, name):
        n = _u64()
        p = getattr(self.L, "fmo_" + name)(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()


class RefFM(_Base):
    """The reference itself. text=None + path -> load_FMIndex(path)."""

    def __init__(self, text=None, path=None, popcnt=False, from_file=None):
        so = os.path.join(_HERE, "_ref", "libseal_ref_popcnt.so" if popcnt else "libseal_ref.so")
        if not os.path.exists(so):
            raise FileNotFoundError(so + " (build with `make -C oracle ref` where /root/reference exists)")
        self.L = C.CDLL(so)
        L = self.L
        L.ref_new.restype = _vp
        L.ref_free.argtypes = [_vp]
        L.ref_load.restype = _vp; L.ref_load.argtypes = [C.c_char_p]
        L.ref_save.argtypes = [_vp, C.c_char_p]
        L.ref_initialize.argtypes = [_vp, _vp, _u64]
        L.ref_initialize_from_file.argtypes = [_vp, C.c_char_p, C.c_int]
        for nm in ("ref_size", "ref_sigma"):
            getattr(L, nm).restype = _u64; getattr(L, nm).argtypes = [_vp]
        L.ref_max_level.restype = C.c_uint32; L.ref_max_level.argtypes = [_vp]
        L.ref_backward_search_step.argtypes = [_vp, _u64, _u64, _u64, _vp]
        L.ref_backward_search_step_batch.argtypes = [_vp, _u64, _vp, _vp, _vp, _vp, _vp]
        L.ref_backward_search_multi.argtypes = [_vp, _vp, _u64, _vp]
        L.ref_distinct.restype = _u64; L.ref_distinct.argtypes = [_vp, _u64, _u64, _vp, _u64]
        L.ref_distinct_count.restype = _u64; L.ref_distinct_count.argtypes = [_vp, _u64, _u64, _vp, _u64]
        L.ref_distinct_count_multi.restype = _u64
        L.ref_distinct_count_multi.argtypes = [_vp, _u64, _vp, _vp, _vp, _vp, _u64]
        L.ref_locate.restype = _u64; L.ref_locate.argtypes = [_vp, _u64]
        L.ref_extract_text.restype = _u64; L.ref_extract_text.argtypes = [_vp, _u64, _u64, _vp, _u64]
        if path is not None:
            self.h = L.ref_load(path.encode())
        elif from_file is not None:
            self.h = L.ref_new()
            L.ref_initialize_from_file(self.h, from_file[0].encode(), int(from_file[1]))
        else:
            t = _np_u64(text)
            self.h = L.ref_new()
            L.ref_initialize(self.h, t.ctypes.data, len(t))
        self._buf = np.zeros(2 * (int(self.sigma()) + 2), dtype=np.uint64)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_free(self.h); self.h = None

    def save(self, path): self.L.ref_save(self.h, path.encode())
    def size(self): return int(self.L.ref_size(self.h))
    def sigma(self): return int(self.L.ref_sigma(self.h))
    def max_level(self): return int(self.L.ref_max_level(self.h))

    def backward_search_step(self, sym, lo, hi):
        out = (_u64 * 2)()
        self.L.ref_backward_search_step(self.h, sym, lo, hi, out)
        return int(out[0]), int(out[1])

    def backward_search_step_batch(self, sym, lo, hi):
        sym, lo, hi = _np_u64(sym), _np_u64(lo), _np_u64(hi)
        ol = np.empty_like(sym); oh = np.empty_like(sym)
        self.L.ref_backward_search_step_batch(self.h, len(sym), sym.ctypes.data, lo.ctypes.data,
                                              hi.ctypes.data, ol.ctypes.data, oh.ctypes.data)
        return ol, oh

    def backward_search_multi(self, q):
        q = _np_u64(q); out = (_u64 * 2)()
        self.L.ref_backward_search_multi(self.h, q.ctypes.data, len(q), out)
        return int(out[0]), int(out[1])

    def distinct(self, lo, hi):
        k = self.L.ref_distinct(self.h, lo, hi, self._buf.ctypes.data, len(self._buf))
        return self._buf[:k].copy()

    def distinct_count(self, lo, hi):
        k = self.L.ref_distinct_count(self.h, lo, hi, self._buf.ctypes.data, len(self._buf))
        return self._buf[:k].copy()

    def distinct_count_multi(self, lows, highs, want_output=True):
        """The reference's std::async fan-out (fm_index.cpp:111-131)."""
        lows, highs = _np_u64(lows), _np_u64(highs)
        n = len(lows)
        offs = np.zeros(n + 1, dtype=np.uint64)
        if not want_output:
            self.L.ref_distinct_count_multi(self.h, n, lows.ctypes.data, highs.ctypes.data, None, None, 0)
            return None
        tot = self.L.ref_distinct_count_multi(self.h, n, lows.ctypes.data, highs.ctypes.data,
                                              offs.ctypes.data, None, 0)
        out = np.zeros(max(int(tot), 1), dtype=np.uint64)
        self.L.ref_distinct_count_multi(self.h, n, lows.ctypes.data, highs.ctypes.data,
                                        offs.ctypes.data, out.ctypes.data, len(out))
        return [out[int(offs[i]):int(offs[i + 1])].copy() for i in range(n)]

    def locate(self, row): return int(self.L.ref_locate(self.h, row))

    def extract_text(self, begin, end):
        n = max(int(end) - int(begin), 0)
        buf = np.zeros(n + 1, dtype=np.uint64)
        k = self.L.ref_extract_text(self.h, begin, end, buf.ctypes.data, len(buf))
        return buf[:k].copy()


def make_backend(text, prefer_ref=True):
    """RefFM when the compiled reference is present, else the C port."""
    if prefer_ref and ref_available():
        return RefFM(text)
    return PortFM(text)


class OracleIndex:
    """Restatement of seal/index.py:20-204 on an oracle back-end (token-level API)."""

    def __init__(self, sequences=None, backend="auto", _raw=None):
        self.beginnings = [0]
        self.occurring = set()
        self.labels = None
        if _raw is not None:
            self.fm = _raw
            return
        data = []
        occurring = set()
        for seq in sequences:                                  # index.py:46-53
            seq = list(seq)
            self.beginnings.append(self.beginnings[-1] + len(seq))
            occurring |= set(seq)
            data.extend(x + SHIFT for x in seq[::-1])
        self.occurring = list(occurring)
        if backend == "ref" or (backend == "auto" and ref_available()):
            self.fm = RefFM(data)
        else:
            self.fm = PortFM(data)
        self.occurring_distinct, self.occurring_counts = self.get_distinct_count(0, len(self))  # index.py:66

    def __len__(self): return self.beginnings[-1]                # index.py:173-177
    def size(self): return self.fm.size()
    @property
    def n_docs(self): return len(self.beginnings) - 1

    def backward_search_step(self, sym, lo, hi): return self.fm.backward_search_step(sym, lo, hi)

    def get_range(self, sequence):                               # index.py:102-111
        start_row, end_row = 0, self.fm.size()
        for token in sequence:
            start_row, end_row = self.fm.backward_search_step(token + SHIFT, start_row, end_row)
        return start_row, end_row + 1

    def get_count(self, sequence):                               # index.py:113-118
        s, e = self.get_range(sequence)
        return e - s

    def get_distinct_count(self, low, high):                     # index.py:143-156
        data = self.fm.distinct_count(low, high)
        distinct, counts = [], []
        for d, c in zip(data[0::2].tolist(), data[1::2].tolist()):
            if d > 0:
                distinct.append(d - SHIFT); counts.append(c)
        return distinct, counts

    def get_distinct(self, low, high):                           # index.py:136-141
        return self.get_distinct_count(low, high)[0]

    def get_distinct_count_multi(self, lows, highs):             # index.py:158-171
        return [self.get_distinct_count(l, h) for l, h in zip(lows, highs)]

    def get_continuations(self, sequence):                       # index.py:128-134
        s, e = self.get_range(sequence)
        return self.get_distinct(s, e)

    def locate(self, row): return self.fm.locate(row)
    def get_doc_index(self, token_index):                        # index.py:77-82
        return bisect.bisect_right(self.beginnings, token_index) - 1
    def get_doc_index_from_row(self, row):                       # index.py:96-100
        return self.get_doc_index(self.locate(row))
    def get_token_index_from_row(self, row):                     # index.py:90-94
        return self.locate(row)
    def get_doc_indices(self, sequence):                         # index.py:120-126
        s, e = self.get_range(sequence)
        for row in range(s, e):
            yield self.get_doc_index_from_row(row)
    def get_doc(self, doc_index):                                # index.py:68-75
        doc = self.fm.extract_text(self.beginnings[doc_index], self.beginnings[doc_index + 1])
        return [int(x) - SHIFT for x in doc]
    def get_doc_length(self, doc_index):
        return self.beginnings[doc_index + 1] - self.beginnings[doc_index]

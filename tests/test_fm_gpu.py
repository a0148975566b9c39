"""GPU parity: the CUDA FM-index kernels, called through the C ABI (include/sealfm.h) exactly as
the reference's SWIG module would be, against the CPU oracle on the same seeded inputs; bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "fm_golden.npz"))
CASES = ["keeper", "toy", "rand5k", "phrase"]


@pytest.fixture(scope="module", autouse=True)
def need_gpu():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"


def mk(text):
    from seal_b200.cpp_modules.fm_index import FMIndex
    fm = FMIndex(); fm.initialize(text)
    return fm


@pytest.mark.parametrize("name", CASES)
def test_goldens_through_the_abi(name):
    fm = mk(G[f"{name}.text"])
    size = int(G[f"{name}.size"])
    assert fm.size() == size
    fl, fh = G[f"{name}.first_lo"], G[f"{name}.first_hi"]
    syms = np.arange(len(fl), dtype=np.uint64)
    ol, oh = fm.backward_search_step_batch(syms, np.zeros_like(syms), np.full_like(syms, size))
    assert np.array_equal(ol, fl) and np.array_equal(oh, fh)          # incl. the §H1 quirk
    assert fm.backward_search_step(int(syms[-1]), 0, size) == [int(fl[-1]), int(fh[-1])]
    wsym, wlo, whi = G[f"{name}.walk_sym"], G[f"{name}.walk_lo"], G[f"{name}.walk_hi"]
    dc_off, dc = G[f"{name}.dc_off"], G[f"{name}.dc"]
    W, D = wsym.shape
    cl = np.zeros(W, dtype=np.uint64); ch = np.full(W, size, dtype=np.uint64)
    for d in range(D):
        a, b = fm.backward_search_step_batch(wsym[:, d], cl, ch)
        assert np.array_equal(a, wlo[:, d]) and np.array_equal(b, whi[:, d])
        cl, ch = a, b
    # backward_search_multi == fold of steps, returns hi exclusive (fm_index.cpp:55-65)
    lo_m, hi_m = fm.backward_search_multi_batch([wsym[w].tolist() for w in range(W)])
    assert np.array_equal(lo_m, wlo[:, -1]) and np.array_equal(hi_m, whi[:, -1] + 1)
    assert fm.backward_search_multi(wsym[0].tolist()) == [int(wlo[0, -1]), int(whi[0, -1]) + 1]
    # distinct_count over every visited range, one multi call
    lows = wlo.reshape(-1); highs = whi.reshape(-1) + 1
    ok = highs >= lows
    res = fm.distinct_count_multi(lows[ok].tolist(), highs[ok].tolist())
    ks = np.nonzero(ok)[0]
    for r, k in zip(res, ks):
        assert np.array_equal(np.asarray(r, dtype=np.uint64), dc[int(dc_off[k]):int(dc_off[k + 1])]), k
    got = fm.locate_batch(G[f"{name}.loc_rows"])
    assert np.array_equal(got, G[f"{name}.loc"])
    assert fm.locate(int(size + 5)) == 2**64 - 1                        # fm_index.cpp:165
    eo, ex = G[f"{name}.ext_off"], G[f"{name}.ext"]
    for i, (b, e) in enumerate(zip(G[f"{name}.ext_b"], G[f"{name}.ext_e"])):
        assert fm.extract_text(int(b), int(e)) == ex[int(eo[i]):int(eo[i + 1])].tolist()


def test_sdsl_fmi_from_reference_answers_like_the_reference():
    from seal_b200.cpp_modules.fm_index import load_FMIndex
    fm = load_FMIndex(os.path.join(HERE, "golden", "tiny_ref.fmi"))
    size = int(G["phrase.size"])
    syms = np.arange(len(G["phrase.first_lo"]), dtype=np.uint64)
    ol, oh = fm.backward_search_step_batch(syms, np.zeros_like(syms), np.full_like(syms, size))
    assert np.array_equal(ol, G["phrase.first_lo"]) and np.array_equal(oh, G["phrase.first_hi"])
    assert np.array_equal(fm.locate_batch(G["phrase.loc_rows"]), G["phrase.loc"])


def test_corpus_walks_masks_and_docs_vs_oracle(small_corpus):
    """80 k-token phrase corpus: the seal/index.py-level API, the batched kernels and the mask
    expansion against the oracle (compiled reference when shipped, else the C port)."""
    import torch
    from oracle.fm_oracle import OracleIndex
    from seal_b200.index import FMIndex
    docs = small_corpus
    seqs = [d.tolist() for d in docs]
    ora = OracleIndex(seqs)
    idx = FMIndex(); idx.initialize(seqs, in_memory=True)
    idx2 = FMIndex(); idx2.initialize(seqs, in_memory=False)             # file path, '<l' ints
    assert np.array_equal(idx.section(0), idx2.section(0))
    assert len(idx) == len(ora) and idx.n_docs == ora.n_docs and idx.size() == ora.size()
    assert idx.occurring_distinct == ora.occurring_distinct                # §H2 semantics
    assert idx.occurring_counts == ora.occurring_counts
    assert sorted(idx.occurring) == sorted(ora.occurring)
    assert idx.get_range([]) == ora.get_range([])
    rng = np.random.default_rng(3)
    for _ in range(60):
        d = seqs[int(rng.integers(0, len(seqs)))]
        a = int(rng.integers(0, len(d) - 6)); n = int(rng.integers(1, 6))
        seq = d[a:a + n]
        assert idx.get_range(seq) == ora.get_range(seq)
        assert idx.get_count(seq) == ora.get_count(seq)
        assert idx.get_continuations(seq) == ora.get_continuations(seq)
        lo, hi = ora.get_range(seq)
        assert idx.get_distinct_count(lo, hi) == ora.get_distinct_count(lo, hi)
        rows = list(range(lo, min(hi, lo + 5)))
        assert [idx.get_doc_index_from_row(r) for r in rows] == [ora.get_doc_index_from_row(r) for r in rows]
        assert idx.get_doc_index_from_rows(rows).tolist() == [ora.get_doc_index_from_row(r) for r in rows]
        assert list(idx.get_doc_indices(seq))[:5] == [ora.get_doc_index_from_row(r) for r in rows]
    for di in (0, 17, len(seqs) - 1):
        assert idx.get_doc(di) == ora.get_doc(di) == seqs[di]
    # unseen n-gram / unknown token
    assert idx.get_count([4, 4, 4, 4, 4, 4, 4]) == ora.get_count([4, 4, 4, 4, 4, 4, 4])
    assert idx.get_range([50264]) == ora.get_range([50264])
    # device-tensor LF + mask expansion for a batch of live ranges
    R = 512
    toks = np.asarray([seqs[int(rng.integers(0, len(seqs)))][int(rng.integers(0, 30))] for _ in range(R)])
    sym = torch.tensor(toks + 10, dtype=torch.int64, device="cuda")
    lo0 = torch.zeros(R, dtype=torch.int64, device="cuda"); hi0 = torch.full((R,), idx.size(), dtype=torch.int64, device="cuda")
    lo1, hi1 = idx.lf_step_tensors(sym, lo0, hi0)
    exp = [ora.get_range([int(t)]) for t in toks]
    assert lo1.tolist() == [e[0] for e in exp] and (hi1 + 1).tolist() == [e[1] for e in exp]
    V = 50265
    mask = idx.expand_mask_tensors(lo1, hi1 + 1, V).cpu().numpy().view(np.uint32)
    for r in range(0, R, 7):
        allowed = np.nonzero(np.unpackbits(mask[r].view(np.uint8), bitorder="little")[:V])[0].tolist()
        assert allowed == ora.get_distinct(*exp[r]), r
    # empty range and full range rows
    lo2 = torch.tensor([5, 0], dtype=torch.int64, device="cuda"); hi2 = torch.tensor([5, len(idx)], dtype=torch.int64, device="cuda")
    m2 = idx.expand_mask_tensors(lo2, hi2, V).cpu().numpy().view(np.uint32)
    assert m2[0].sum() == 0
    allowed = np.nonzero(np.unpackbits(m2[1].view(np.uint8), bitorder="little")[:V])[0].tolist()
    assert allowed == ora.occurring_distinct
    # save / load round trip keeps answers (index.py:186-204)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        idx.save(os.path.join(d, "ix"))
        back = FMIndex.load(os.path.join(d, "ix"))
        assert back.occurring_distinct == idx.occurring_distinct and back.beginnings == idx.beginnings
        assert back.get_range(seqs[5][2:5]) == ora.get_range(seqs[5][2:5])


def test_properties_at_scale():
    """1 M-token corpus (oracle too slow to enumerate): size-independent properties of the kernels."""
    import torch
    from seal_b200.synthetic import make_corpus, corpus_symbols
    from seal_b200.cpp_modules.fm_index import FMIndex
    docs = make_corpus(n_docs=10_000, doc_len=100, n_phrases=20_000, seed=99)
    text = corpus_symbols(docs)
    fm = FMIndex(); fm.initialize(text)
    m = fm.size()
    assert m == text.size + 1
    rng = np.random.default_rng(5)
    # (1) counts of a range's distinct symbols sum to its width; symbols ascending
    lows = rng.integers(0, m - 1, size=200); highs = np.minimum(lows + rng.integers(1, 5000, size=200), m)
    for (lo, hi), r in zip(zip(lows, highs), fm.distinct_count_multi(lows.tolist(), highs.tolist())):
        s, c = r[0::2], r[1::2]
        assert sum(c) == hi - lo and s == sorted(s) and all(x > 0 for x in c)
    # (2) LF of a range by every one of its distinct symbols partitions it: widths equal the counts
    lo, hi = int(lows[0]), int(highs[0])
    r = fm.distinct_count(lo, hi)
    ol, oh = fm.backward_search_step_batch(r[0::2], [lo] * (len(r) // 2), [hi - 1] * (len(r) // 2))
    assert ((oh + 1 - ol).tolist()) == r[1::2]
    # (3) locate is a bijection rows -> text positions; inverse via extract of the whole text head
    rows = rng.permutation(m)[:4000].astype(np.uint64)
    pos = fm.locate_batch(rows)
    assert len(set(pos.tolist())) == len(rows) and pos.max() < m
    # (4) every document n-gram is found, and locate lands inside a document that contains it
    for di in rng.integers(0, len(docs), size=20):
        d = docs[di].tolist()
        q = [t + 10 for t in d[10:14]]
        lo, hi = fm.backward_search_multi(q)
        assert hi > lo
    # (5) extract_text round trip: reversed-text coordinates (seal/index.py:68-75)
    n_tok = text.size
    for di in (0, 5000, 9999):
        got = fm.extract_text(di * 100, (di + 1) * 100)
        assert [g - 10 for g in got] == docs[di].tolist()


def _build_raw(text, gpu):
    import ctypes as C
    from seal_b200._lib import lib, check
    from seal_b200.cpp_modules.fm_index import FMIndex as RawFM
    a = np.ascontiguousarray(np.asarray(text, dtype=np.uint64))
    out = C.c_void_p()
    if gpu:
        check(lib.sealfm_build_gpu(a.ctypes.data, len(a), 0, C.byref(out)))
    else:
        check(lib.sealfm_build(a.ctypes.data, len(a), C.byref(out)))
    fm = RawFM(); fm._adopt(out.value)
    return fm


def _texts():
    from seal_b200.synthetic import make_corpus, corpus_symbols
    rng = np.random.default_rng(17)
    yield "one symbol", [5]
    yield "two symbols", [9, 9]
    for n in (31, 32, 33, 63, 64, 65, 127, 128, 129, 1000):
        yield f"random n={n}", rng.integers(1, 40, size=n)
    yield "single run (longest possible repeats)", np.full(5000, 7)
    yield "period-3 text", np.tile([3, 1, 2], 3000)
    yield "all distinct", rng.permutation(4000) + 1
    yield "wide alphabet", rng.integers(1, 2 ** 31, size=3000)
    yield "max symbol 2^32-1", np.array([2 ** 32 - 1, 1, 2 ** 32 - 1, 7, 1], dtype=np.uint64)
    docs = make_corpus(n_docs=2000, doc_len=100, n_phrases=1500, seed=8)      # verbatim repeats, duplicate phrases
    yield "phrase corpus 200k", corpus_symbols(docs)
    dup = np.concatenate([corpus_symbols(docs[:50])] * 6)                      # whole documents repeated 6 times
    yield "duplicated documents", dup


def test_gpu_index_builder_matches_host_builder_section_by_section():
    """sealfm_build_gpu (radix-sort prefix doubling, fm_build.cu) vs sealfm_build (host SA-IS): tree bits,
    alphabet, C, SA samples and ISA samples must be identical words -- given the text they are unique, and
    the host builder's sections are pinned against sdsl's own .fmi (tests/golden)."""
    names = ["tree", "alphabet", "C", "sa_samples", "isa_samples"]
    for label, text in _texts():
        h = _build_raw(text, gpu=False); g = _build_raw(text, gpu=True)
        assert (g.size(), ) == (h.size(), ), label
        for w, nm in enumerate(names):
            a, b = h.section(w), g.section(w)
            assert a.shape == b.shape and np.array_equal(a, b), f"{label}: section {nm} differs"
    with pytest.raises(Exception):
        _build_raw([3, 0, 4], gpu=True)                           # symbol 0 is the sentinel


def test_gpu_index_builder_at_benchmark_scale():
    """10 M tokens: same sections as the host builder; prints both build times."""
    import time
    from seal_b200.synthetic import make_corpus, corpus_symbols
    text = corpus_symbols(make_corpus())
    t0 = time.perf_counter(); g = _build_raw(text, gpu=True); tg = time.perf_counter() - t0
    t0 = time.perf_counter(); g2 = _build_raw(text, gpu=True); tg2 = time.perf_counter() - t0
    t0 = time.perf_counter(); h = _build_raw(text, gpu=False); th = time.perf_counter() - t0
    print(f"index build, 10 M tokens: GPU {tg:.3f} s (first call) / {tg2:.3f} s, host SA-IS {th:.3f} s")
    for w in range(5):
        assert np.array_equal(h.section(w), g.section(w)), w
    del g2


def test_rows_beyond_2_pow_32():
    """SURVEY.md section 8 size table: KILT-scale indexes need 33-bit rows ("ranges must be u64",
    /root/reference/seal/cpp_modules/fm_index.hpp:16-18).  A closed-form index with size() = 2^32 + 5 000: the text
    a^n $ has SA[i] = n - i and BWT = a^n $, so every answer is known without building anything: rank counters above
    2^32, LF steps, locate (a walk of up to 31 LF steps ending on a 33-bit SA sample), document ids by bisection."""
    from seal_b200.cpp_modules.fm_index import FMIndex as RawFM
    from seal_b200._lib import lib, check
    n = (1 << 32) + 4999
    m = n + 1
    words = (m + 63) // 64
    tree = np.full(words, np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
    tree[n >> 6] = np.uint64((1 << (n & 63)) - 1)          # bit n (the sentinel's row) and the padding are 0
    tree[(n >> 6) + 1:] = 0
    sa = (np.uint64(n) - np.arange((m + 31) // 32, dtype=np.uint64) * np.uint64(32))
    isa = (np.uint64(n) - np.arange(n // 64 + 1, dtype=np.uint64) * np.uint64(64))
    fm = RawFM.from_sections(m, 1, tree, [0, 1], [0, 1, m], sa, isa)
    del tree, sa, isa
    fm.to_device(0)
    assert fm.size() == m
    beginnings = np.arange(0, n + 1000, 1000, dtype=np.uint64)
    check(lib.sealfm_set_beginnings(fm._handle(), beginnings.ctypes.data, len(beginnings)))
    rng = np.random.default_rng(3)
    big = (1 << 32)
    l = np.concatenate([rng.integers(big - 100, n - 10, size=500), rng.integers(0, n - 10, size=500)]).astype(np.uint64)
    r = np.minimum(l + rng.integers(0, 5000, size=1000).astype(np.uint64), np.uint64(n))      # inclusive, may include the sentinel row
    ol, oh = fm.backward_search_step_batch(np.ones(1000, dtype=np.uint64), l, r)
    assert np.array_equal(ol, 1 + np.minimum(l, n)) and np.array_equal(oh, np.minimum(r + 1, n).astype(np.uint64))
    assert int(ol.max()) > big and int(oh.max()) > big
    # 'aaa' has n - 2 occurrences; its range is [3, n + 1) -- 33-bit bounds out of backward_search_multi
    lo, hi = fm.backward_search_multi([1, 1, 1])
    assert (lo, hi) == (3, n + 1)
    rows = np.concatenate([rng.integers(big, m, size=2000), rng.integers(0, m, size=2000), [0, n, m, big - 1, big, big + 1]]).astype(np.uint64)
    pos = fm.locate_batch(rows)
    exp = np.where(rows < m, np.uint64(n) - np.minimum(rows, np.uint64(n)), np.uint64(0xFFFFFFFFFFFFFFFF))
    assert np.array_equal(pos, exp)
    ok = rows < m
    docs = np.zeros(int(ok.sum()), dtype=np.uint64)
    rr = np.ascontiguousarray(rows[ok])
    check(lib.sealfm_doc_index_from_rows(fm._dev(), len(rr), rr.ctypes.data, docs.ctypes.data))
    assert np.array_equal(docs, (np.searchsorted(beginnings, exp[ok], side="right") - 1).astype(np.uint64))
    # successor sets through the expansion kernels: a range inside the a-run -> {a}; one that ends on the last row -> {a, $}
    out = fm.distinct_count_multi([big + 5, n - 3, big + 5], [big + 4000, n + 1, n + 1])
    assert out[0] == [1, 3995] and out[1] == [0, 1, 1, 3] and out[2] == [0, 1, 1, 4994]

import numpy as np
import pytest

from spacy_ray_b200.native import featurize as native
from spacy_ray_b200.pipeline.doc import featurize_words_py, hash_string, word_shape

WORDS = ["Hello", "world", "U.S.A.", "12345", "x", "", "aaaaaaaaBBBBBB11111", "naïve", "Ünïcode", "can't",
         "A" * 120, "MiXeD-42"]


def test_word_shape():
    assert word_shape("Hello") == "Xxxxx" and word_shape("aaaaaaaa") == "xxxx" and word_shape("U.S.A.") == "X.X.X."
    assert word_shape("12345") == "dddd"


def test_hash_is_stable_and_nonzero():
    assert hash_string("abc") == hash_string("abc") != hash_string("abd")
    assert all(hash_string(w) != 0 for w in WORDS)


@pytest.mark.skipif(not native.available(), reason="native host runtime not built")
def test_native_featurizer_is_bit_identical_to_python():
    a = native.featurize_words(WORDS)
    b = featurize_words_py(WORDS)
    assert a.dtype == np.uint64 and np.array_equal(a, b)


def test_collate_fallback_and_native_agree():
    rng = np.random.default_rng(0)
    lens = [3, 1, 5, 2]
    store = rng.integers(1, 1 << 40, size=(sum(lens), 4)).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ids = np.array([2, 0, 3], dtype=np.int64)
    cap = 20

    def run():
        attrs = np.full((cap, 4), -1, dtype=np.int64)
        mask = np.full((cap,), -1, dtype=np.float32)
        starts = np.zeros(3, dtype=np.int32)
        ls = np.zeros(3, dtype=np.int32)
        used = native.collate(store, off, ids, attrs, mask, starts, ls)
        return used, attrs, mask, starts, ls

    used, attrs, mask, starts, ls = run()
    assert used == 5 + 3 + 2 + 3 + 1 and ls.tolist() == [5, 3, 2] and starts.tolist() == [1, 7, 11]
    assert np.array_equal(attrs[1:6], store[off[2]:off[3]]) and mask[:14].tolist() == [0, 1, 1, 1, 1, 1, 0, 1, 1, 1, 0, 1, 1, 0]
    assert float(np.abs(attrs[0]).sum()) == 0 and float(np.abs(attrs[13:]).sum()) == 0


def test_group_rows_native_matches_python_fallback_on_edge_cases():
    """srb_group_rows (counting sort of a batch's padded rows by dense vocabulary id) vs the numpy
    fallback: single doc, one-token docs, a batch that exactly fills the row capacity, repeated docs."""
    import numpy as np

    import spacy_ray_b200.native.featurize as nat

    rng = np.random.default_rng(3)
    lens = np.array([1, 1, 7, 3, 1, 12, 2], dtype=np.int64)
    doc_off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=doc_off[1:])
    T = int(doc_off[-1])
    n_groups = np.array([5, 3, 9, 2], dtype=np.int32)
    gid = np.ascontiguousarray(np.stack([rng.integers(0, n, size=T) for n in n_groups], axis=1).astype(np.int32))
    scratch = np.zeros(int(n_groups.sum()) + 4, dtype=np.int32)
    cases = [np.array([2]), np.array([0, 1, 4]), np.arange(len(lens)), np.array([5, 5, 3])]
    for ids in cases:
        ids = ids.astype(np.int64)
        rows = int(lens[ids].sum()) + len(ids) + 1
        for rb in (rows, rows + 13):
            got = np.full(4 * rb, -7, dtype=np.int32)
            want = np.full(4 * rb, -7, dtype=np.int32)
            n1 = nat.group_rows(gid, n_groups, doc_off, ids, rb, got, scratch)
            lib, nat._lib = nat._lib, None
            tried, nat._tried = nat._tried, True
            try:
                n2 = nat.group_rows(gid, n_groups, doc_off, ids, rb, want, scratch)
            finally:
                nat._lib, nat._tried = lib, tried
            assert n1 == n2 == int(lens[ids].sum())
            np.testing.assert_array_equal(got, want)
    if nat.available():
        import pytest

        with pytest.raises(ValueError):
            nat.group_rows(gid, n_groups, doc_off, np.arange(len(lens), dtype=np.int64), 5,
                           np.zeros(4 * 5, dtype=np.int32), scratch)

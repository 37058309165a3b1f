"""The oracle's NEAR / BEFORE / NOTNEAR / quorum nodes (pinned on the reference's golden vectors in test_oracle_golden.py) against
brute force over a seeded synthetic corpus whose token positions are known exactly (M.synth_token): matched document sets of
two-keyword operators, plus containment / monotonicity properties for more keywords. CPU only; these operators are not on the CUDA
path yet (it answers MGPU_E_UNSUPPORTED)."""
import random

import pytest

import helpers
import manticoresearch_b200.mgpu as M

N_DOCS = 2500
VOCAB = 400


@pytest.fixture(scope="module")
def corpus(tmp_path_factory):
    prefix = str(tmp_path_factory.mktemp("ops") / "c")
    p = M.SynthParams(N_DOCS, vocab=VOCAB, threads=2, body_min=8, body_max=60, body_mu=3.0, body_sigma=0.5)
    M.build_synthetic(prefix, p)
    hits = {}            # term -> doc -> sorted [(field<<24) | pos]
    for d in range(N_DOCS):
        for f in range(2):
            for k in range(M.synth_field_len(p, d, f)):
                hits.setdefault(M.synth_token(p, d, f, k), {}).setdefault(d, []).append((f << 24) | (k + 1))
    idx = helpers.OracleIndex(prefix)
    yield {"idx": idx, "hits": hits}
    idx.close()


def _rows(idx, node):
    r = idx.search([M.Query(node, ranker=M.RANK_BM25, max_matches=N_DOCS, sort_keys=[M.SortKey(M.KEYPART_ROWID, 0, False)])]).get(0)
    assert r["status"] == 0
    assert r["total_found"] == len(r["rowid"])
    return list(r["rowid"])


def _kw(t, pos):
    return M.kw(M.synth_keyword(t), pos)


def _pairs(corpus, n, seed):
    rng = random.Random(seed)
    terms = [t for t, docs in corpus["hits"].items() if len(docs) >= 20]
    return [tuple(rng.sample(terms, 2)) for _ in range(n)]


def test_near_two_keywords_equals_brute_force(corpus):
    hits = corpus["hits"]
    for a, b in _pairs(corpus, 40, 1):
        for dist in (1, 2, 5):
            exp = [d for d in sorted(set(hits[a]) & set(hits[b]))
                   if any(0 < abs(x - y) <= dist for x in hits[a][d] for y in hits[b][d])]
            got = _rows(corpus["idx"], M.Node(M.OP_NEAR, children=[_kw(a, 1), _kw(b, 2)], oparg=dist))
            assert got == exp, (a, b, dist)


def test_before_two_keywords_equals_brute_force(corpus):
    hits = corpus["hits"]
    for a, b in _pairs(corpus, 40, 2):
        exp = [d for d in sorted(set(hits[a]) & set(hits[b]))
               if any((x >> 24) == (y >> 24) and x < y for x in hits[a][d] for y in hits[b][d])]
        got = _rows(corpus["idx"], M.Node(M.OP_BEFORE, children=[_kw(a, 1), _kw(b, 2)]))
        assert got == exp, (a, b)


def test_notnear_two_keywords_equals_brute_force(corpus):
    hits = corpus["hits"]
    for a, b in _pairs(corpus, 40, 3):
        for dist in (1, 3, 8):
            exp = []
            for d in sorted(hits[a]):
                nots = hits[b].get(d, [])
                ok = False
                for x in hits[a][d]:
                    after = [y for y in nots if y >= x]
                    if not after or x + dist < after[0]:
                        ok = True
                exp.append(d) if ok else None
            got = _rows(corpus["idx"], M.Node(M.OP_NOTNEAR, children=[_kw(a, 1), _kw(b, 2)], oparg=dist))
            assert got == exp, (a, b, dist)


def test_quorum_equals_presence_count(corpus):
    hits = corpus["hits"]
    rng = random.Random(4)
    terms = [t for t, docs in hits.items() if len(docs) >= 20]
    for _ in range(25):
        ts = rng.sample(terms, rng.randint(3, 6))
        thr = rng.randint(2, len(ts) - 1)
        exp = [d for d in range(N_DOCS) if sum(1 for t in ts if d in hits[t]) >= thr]
        node = M.Node(M.OP_QUORUM, words=[M.Keyword(M.synth_keyword(t), i + 1) for i, t in enumerate(ts)], oparg=thr)
        assert _rows(corpus["idx"], node) == exp, (ts, thr)


def test_nway_near_and_before_are_contained_and_monotone(corpus):
    rng = random.Random(5)
    terms = [t for t, docs in corpus["hits"].items() if len(docs) >= 200]
    idx = corpus["idx"]
    for _ in range(15):
        ts = rng.sample(terms, 3)
        kws = lambda: [_kw(t, i + 1) for i, t in enumerate(ts)]     # noqa: E731
        allof = set(_rows(idx, M.AND(*kws())))
        prev = set()
        for dist in (1, 3, 9, 40):
            cur = set(_rows(idx, M.Node(M.OP_NEAR, children=kws(), oparg=dist)))
            assert prev <= cur <= allof, (ts, dist)
            prev = cur
        before = set(_rows(idx, M.Node(M.OP_BEFORE, children=kws())))
        assert before <= allof
        # a << b << c implies a << b
        assert before <= set(_rows(idx, M.Node(M.OP_BEFORE, children=kws()[:2])))


def _float_corpus(tmp_path, n=300):
    import random
    import struct
    rng = random.Random(11)
    special = [0.0, -0.0, 1.5, -1.5, 3.0e38, -3.0e38, 1e-40, -1e-40, float("inf"), float("-inf")]
    vals = [special[i] if i < len(special) else rng.choice([rng.uniform(-100, 100), rng.randint(-3, 3) * 0.5]) for i in range(n)]
    bits = [struct.unpack("<I", struct.pack("<f", v))[0] for v in vals]
    docs = [{"id": 1 + i, "fields": [[("w", 1)] + ([("x", 2)] if i % 3 else [])], "attrs": [bits[i], i % 7]} for i in range(n)]
    prefix = str(tmp_path / "flt")
    M.build_index(prefix, ["body"], docs, attr_names=["price", "grp"])
    return prefix, [struct.unpack("<f", struct.pack("<I", b))[0] for b in bits]


def float_sort_queries():
    price, grp = 1, 2       # attribute 0 is the document id
    root = M.OR(M.kw("w", 1), M.kw("x", 2))
    return [M.Query(root, ranker=M.RANK_BM25, max_matches=1000, sort_keys=[M.SortKey(M.KEYPART_FLOAT, price, True)]),
            M.Query(root, ranker=M.RANK_BM25, max_matches=1000, sort_keys=[M.SortKey(M.KEYPART_FLOAT, price, False)]),
            M.Query(root, ranker=M.RANK_BM25, max_matches=40, sort_keys=[M.SortKey(M.KEYPART_INT, grp, False), M.SortKey(M.KEYPART_FLOAT, price, True)]),
            M.Query(root, ranker=M.RANK_BM25, max_matches=25, sort_keys=[M.SortKey(M.KEYPART_FLOAT, price, False), M.SortKey(M.KEYPART_WEIGHT, 0, True)])]


def test_float_sort_key(tmp_path):
    """SPH_KEYPART_FLOAT (src/sphinxsort.cpp:4690-4696): a 32-bit attribute ordered as an IEEE float, -0 == +0, ties by the next key and
    finally by rowid ascending; checked against Python's own float ordering"""
    prefix, vals = _float_corpus(tmp_path)
    idx = helpers.OracleIndex(prefix)
    try:
        qs = float_sort_queries()
        rs = idx.search(qs)
        n = len(vals)
        desc = sorted(range(n), key=lambda i: (-vals[i], i))
        asc = sorted(range(n), key=lambda i: (vals[i], i))
        assert list(rs.get(0)["rowid"]) == desc and list(rs.get(1)["rowid"]) == asc
        by_grp = sorted(range(n), key=lambda i: (i % 7, -vals[i], i))[:40]
        assert list(rs.get(2)["rowid"]) == by_grp
        w = dict(zip(rs.get(0)["rowid"], rs.get(0)["weight"]))
        assert list(rs.get(3)["rowid"]) == sorted(range(n), key=lambda i: (vals[i], -w[i], i))[:25]
    finally:
        idx.close()

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

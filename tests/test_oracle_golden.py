"""Pins the CPU oracle against the reference's own golden vectors.

test/test_019, test_037, test_322, test_116, test_114 model.bin and gtests_rtstuff.cpp:244-335, transcribed
by tests/golden/make_golden.py.  These are results of the real reference searchd, so a pass here means the
index writer + oracle pair reproduces the reference bit-exactly on these queries: docid set, order,
integer weights, total_found and per-keyword docs/hits.
"""
import helpers
import pytest


def _cases():
    return [(c["name"], i) for c in helpers.load_golden() for i in range(len(c["queries"]))]


@pytest.mark.parametrize("case_name,qi", _cases())
def test_oracle_matches_reference_golden(golden_cases, golden_indexes, case_name, qi):
    case = next(c for c in golden_cases if c["name"] == case_name)
    q = case["queries"][qi]
    idx = helpers.OracleIndex(golden_indexes[case_name])
    try:
        query = helpers.golden_query(case, q)
        rs = idx.search([query])
        r = rs.get(0)
        assert r["status"] == 0, q["text"]
        exp = q["expect"]
        got = list(zip(r["docid"], r["weight"]))
        if q.get("ids_only"):       # SphinxQL `select *` results: the reference's model holds no weights for these
            got = [(d, 0) for d, _ in got]
        limit = q.get("limit")
        if limit:
            got = got[:limit]
        assert got == [tuple(m) for m in exp["matches"]], (q["text"], got[:8], exp["matches"][:8])
        assert r["total_found"] == exp["total_found"], q["text"]
        if not case.get("skip_word_stats"):
            kws = query.keywords()
            stats = rs.word_stats(0, len(kws))
            for k, (docs, hits) in zip(kws, stats):
                if k.word in exp["words"]:
                    assert [docs, hits] == exp["words"][k.word], (q["text"], k.word)
    finally:
        idx.close()


def test_reference_built_index_opens():
    """tests/golden/ref_index: a v57 index written by the reference itself (test/test_406). The oracle's reader must take the real
    writer's bytes: header, keywords dictionary with its checkpoint + dict-header tail, doclists with inlined hits"""
    import os
    import manticoresearch_b200.mgpu as M
    prefix = os.path.join(helpers.ROOT, "tests", "golden", "ref_index", "index.0")
    idx = helpers.OracleIndex(prefix)
    try:
        assert idx.total_docs == 1
        for word in ("doc", "one"):
            assert idx.word_stats(word) == (1, 1)
            rowid, hits, fields, pos = idx.decode_doclist(word)
            assert list(rowid) == [0] and list(hits) == [1] and list(fields) == [1]
            assert int(pos[0]) >> 63 == 1          # the only hit is inlined into the doclist (src/sphinx.cpp:523-530)
        assert idx.word_stats("two") is None
        r = idx.search([M.Query(M.AND(M.kw("doc", 1), M.kw("one", 2)), max_matches=5)]).get(0)
        assert r["total_found"] == 1 and r["rowid"] == [0]
    finally:
        idx.close()

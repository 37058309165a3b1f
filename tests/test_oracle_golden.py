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


def _fnv1a64(b):
    h = 0xcbf29ce484222325
    for c in b:
        h = ((h ^ c) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def _unzip(raw, pos):
    """MSB-first 7-bit groups (src/fileio.cpp:31-45)"""
    v = 0
    while True:
        b = raw[pos]
        pos += 1
        v = (v << 7) | (b & 0x7F)
        if not b & 0x80:
            return v, pos


def test_dict_crc_golden(golden_cases, golden_indexes, golden_indexes_crc):
    """dict=crc (CSphDiskDictTraits, src/sphinx.cpp:18263-18339; CWordlist::GetWord, src/indexformat.cpp:425-473): every golden corpus
    rewritten with FNV64 word ids gives the reference's golden results through the oracle's crc reader, the .spi holds ascending
    FNV-1a ids (published known answers: "a", "foobar"), and the doclist files are the keywords build's, reordered"""
    import struct
    assert _fnv1a64(b"a") == 0xaf63dc4c8601ec8c and _fnv1a64(b"foobar") == 0x85944171f73967e8
    ran = 0
    for case in golden_cases:
        prefix = golden_indexes_crc[case["name"]]
        raw = open(prefix + ".spi", "rb").read()
        hdr = open(prefix + ".sph", "rb").read()
        assert raw[0] == 1
        idx, kw = helpers.OracleIndex(prefix), helpers.OracleIndex(golden_indexes[case["name"]])
        try:
            words = set()
            for d in case["docs"]:
                for t in d["fields"]:
                    for w, _ in helpers.case_tokens(case, t):
                        words.add(w)
            ids = sorted(_fnv1a64(w.encode()) for w in words)
            # walk the chunks from the first one: 64 entries, zero delta + last doclist length, next chunk
            got, pos, n_in_chunk, last = [], 1, 0, 0
            while len(got) < len(ids):
                delta, pos = _unzip(raw, pos)
                if delta == 0:
                    assert n_in_chunk == 64
                    _, pos = _unzip(raw, pos)
                    n_in_chunk, last = 0, 0
                    continue
                last += delta
                got.append(last)
                _, pos = _unzip(raw, pos)
                docs, pos = _unzip(raw, pos)
                _, pos = _unzip(raw, pos)
                if docs > 32:
                    _, pos = _unzip(raw, pos)
                n_in_chunk += 1
            assert got == ids, case["name"]
            for w in sorted(words)[:50]:
                assert idx.word_stats(w) == kw.word_stats(w)
                a, b = idx.decode_doclist(w), kw.decode_doclist(w)
                for x, y in zip(a[:3], b[:3]):
                    assert (x == y).all()
            assert idx.word_stats("nosuchkeywordanywhere") is None
            for q in case["queries"]:
                query = helpers.golden_query(case, q)
                r = idx.search([query]).get(0)
                assert r["status"] == 0, q["text"]
                got_m = list(zip(r["docid"], r["weight"]))
                if q.get("ids_only"):
                    got_m = [(d, 0) for d, _ in got_m]
                if q.get("limit"):
                    got_m = got_m[:q["limit"]]
                assert got_m == [tuple(m) for m in q["expect"]["matches"]], (case["name"], q["text"])
                assert r["total_found"] == q["expect"]["total_found"], q["text"]
                ran += 1
        finally:
            idx.close()
            kw.close()
    assert ran == sum(len(c["queries"]) for c in golden_cases)

"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bar: bit-exact matched rowids, integer weights, order (ties by rowid) and total_found.
Run on the B200 box with:  python -m pytest tests -m gpu
"""
import os

import numpy as np
import pytest

import helpers
import manticoresearch_b200.mgpu as M
from manticoresearch_b200 import workload

pytestmark = pytest.mark.gpu

N_DOCS = 200_000


@pytest.fixture(scope="module")
def synth(tmp_path_factory):
    base = tmp_path_factory.mktemp("synth")
    prefix = str(base / "s200k")
    params = M.SynthParams(N_DOCS)
    M.build_synthetic(prefix, params)
    gpu = M.Index(prefix, device=0)
    cpu = helpers.OracleIndex(prefix)
    yield {"prefix": prefix, "params": params, "gpu": gpu, "cpu": cpu}
    gpu.close()
    cpu.close()


def _compare_batch(synth, queries, allow_unsupported=0.0):
    g = synth["gpu"].search(queries)
    c = synth["cpu"].search(queries)
    unsupported = 0
    for i in range(len(queries)):
        a, b = g.get(i), c.get(i)
        if a["status"] == M.MGPU_E_UNSUPPORTED:
            unsupported += 1
            continue
        helpers.assert_same_results(a, b, ctx="query %d" % i)
    assert unsupported <= allow_unsupported * len(queries), "too many unsupported queries: %d" % unsupported
    return unsupported


def test_library_is_native_cuda():
    lib = M.lib()
    assert lib.mgpu_abi_version() == 1


@pytest.mark.parametrize("word", ["t0000001", "t0000002", "t0000010", "t0000100", "t0001000", "t0010000", "t0100000", "t0500000"])
def test_decode_doclist_matches_oracle(synth, word):
    """K1: warp-per-block VByte decode == DiskIndexQword_c::ReadNext for every posting (rowid, hits, fields, hitlist pos)"""
    exp = synth["cpu"].decode_doclist(word)
    got = synth["gpu"].decode_doclist(word)
    if exp is None:
        assert got is None
        return
    for name, e, g in zip(("rowid", "hits", "fields", "hitlist_pos"), exp, got):
        assert np.array_equal(e, g), (word, name, int(np.argmax(e != g)) if len(e) else -1)


def test_decode_doclist_block_boundaries(synth):
    """terms whose df straddles the 32-doc block size (short lists have no skiplist at all)"""
    cpu, gpu = synth["cpu"], synth["gpu"]
    seen = set()
    for t in range(1000, 400000, 37):
        w = M.synth_keyword(t)
        st = cpu.word_stats(w)
        if st is None:
            continue
        df = st[0]
        bucket = df if df <= 70 else None
        if bucket is None or bucket in seen:
            continue
        seen.add(bucket)
        exp, got = cpu.decode_doclist(w), gpu.decode_doclist(w)
        for e, g in zip(exp, got):
            assert np.array_equal(e, g), (w, df)
    assert {1, 31, 32, 33, 64}.issubset(seen) or len(seen) > 40


def test_cfg2_mix_bm25(synth):
    """config 2 query mix (2-8 term AND / OR / (a b)|(c d)), SPH_RANK_BM25 with field weights, top-100"""
    queries = workload.cfg2_queries(n=300, max_rank=50000, max_matches=100)
    _compare_batch(synth, queries)


def test_or_queries_small_k_and_odd_weights(synth):
    """pure OR queries mixing hot and sparse keywords, K from 1 to 100, negative / swapped field weights, index weights:
    exercises the candidate-pool threshold logic (shared per-query bound, warm-up rounds) of stream_kernel"""
    import random
    rng = random.Random(99)
    qs = []
    for i in range(160):
        ranks = rng.sample(range(1, 40), rng.randint(1, 3)) + [int(10 ** rng.uniform(2, 4.5)) for _ in range(rng.randint(1, 4))]
        rng.shuffle(ranks)
        leaves = [M.kw(M.synth_keyword(r - 1), p + 1) for p, r in enumerate(ranks)]
        fw = rng.choice([[10, 1], [1, 1], [1, 10], [3, -2], None])
        qs.append(M.Query(M.OR(*leaves), ranker=M.RANK_BM25, field_weights=fw, max_matches=rng.choice([1, 3, 20, 100]), index_weight=rng.choice([1, 1, 2])))
    # the same keywords again so that every frequent one is shared by >= 2 queries (hot store)
    assert _compare_batch(synth, qs + qs[:40]) == 0


def test_or_class_bound_pass_edges(synth):
    """stream_kernel<512,ORONLY> (launch class 5): the integer weight bound must never drop a row the exact ranking would keep.
    Edges: stop-word-only queries (negative idf everywhere, weights tie massively), keywords limited to one field, K from 1 to
    3000, index weights up to the bound's validity limit and beyond it (bound switched off), field weights 0 / 250 / 251,
    overridden collection statistics (global IDF inputs of the sharded path: idf signs flip), duplicated keywords, range / values /
    exclude filters and attribute or rowid sort keys"""
    import random
    rng = random.Random(4242)
    title, body = 1, 2
    qs = []
    for i in range(220):
        shape = i % 5
        if shape == 0:      # stop words only: every idf <= 0
            ranks = rng.sample(range(1, 8), rng.randint(2, 4))
        elif shape == 1:    # hot + sparse + very rare
            ranks = rng.sample(range(1, 60), 2) + [int(10 ** rng.uniform(2.5, 5.2)) for _ in range(rng.randint(1, 3))]
        elif shape == 2:    # hot only, mid density
            ranks = rng.sample(range(5, 400), rng.randint(2, 6))
        elif shape == 3:    # a keyword twice (same qpos semantics as two leaves)
            r = rng.randint(2, 200)
            ranks = [r, rng.randint(2, 3000), r]
        else:
            ranks = rng.sample(range(1, 2000), rng.randint(2, 8))
        leaves = []
        for p, r in enumerate(ranks):
            n = M.kw(M.synth_keyword(r - 1), p + 1)
            if rng.random() < 0.3:
                n.fields(rng.choice([title, body]))
            leaves.append(n)
        kwargs = {}
        if i % 7 == 3:      # what apply_global_idf would pass on a shard: other totals, other doc counts
            kwargs = dict(total_docs=rng.choice([N_DOCS // 3, N_DOCS * 5]),
                          word_docs=[max(1, int(N_DOCS * rng.uniform(0.0001, 0.9))) for _ in ranks])
        if i % 4 == 1:      # filters inside the bound pass; attribute / rowid sort keys (whole-key compare, weight only for the result)
            gid, ts = synth["gpu"].attr_index("gid"), synth["gpu"].attr_index("ts")
            kwargs["filters"] = rng.choice([[M.Filter(gid, 100, 299)], [M.Filter(gid, values=[5, 17, 900, 901])],
                                            [M.Filter(gid, 0, 499, exclude=True), M.Filter(ts, 0, 1 << 30)], []])
            kwargs["sort_keys"] = rng.choice([[], [M.SortKey(M.KEYPART_INT, ts, True)], [M.SortKey(M.KEYPART_INT, gid, False), M.SortKey(M.KEYPART_INT, ts, True)],
                                              [M.SortKey(M.KEYPART_ROWID, 0, True)]])
        qs.append(M.Query(M.OR(*leaves), ranker=M.RANK_BM25, field_weights=rng.choice([[10, 1], [1, 1], [0, 5], [250, 1], [251, 7], None]),
                          max_matches=rng.choice([1, 2, 10, 100, 3000]), index_weight=rng.choice([1, 1, 1, 7, 1024, 1025]), **kwargs))
    batch = synth["gpu"].prepare(qs + qs[:60])      # repeated keywords -> hot store
    st = batch.stats()
    batch.free()
    assert st["class_queries"][5] >= 150 and st["class_queries"][5] + st["class_queries"][6] >= 270, st["class_queries"]
    assert _compare_batch(synth, qs + qs[:60]) == 0


def test_hot_dnf_class_edges(synth):
    """stream_kernel<512,2> (launch class 6): OR-of-AND-groups programs whose multi-keyword groups are hot keywords only: all-dense
    ANDs, (a b)|(c d), (a b)|(c SPARSE d), (a b c)|d with a sparse or a hot single, three groups, field limits, negative-idf keywords inside groups,
    filters / attribute sort, K from 1 to 3000"""
    import random
    rng = random.Random(777)
    title, body = 1, 2
    gid, ts = synth["gpu"].attr_index("gid"), synth["gpu"].attr_index("ts")

    def leaf(r, p):
        n = M.kw(M.synth_keyword(r - 1), p)
        if rng.random() < 0.2:
            n.fields(rng.choice([title, body]))
        return n

    qs = []
    for i in range(240):
        shape = i % 6
        pos = [0]

        def group(lo, hi, n):
            out = []
            for r in rng.sample(range(lo, hi), n):
                pos[0] += 1
                out.append(leaf(r, pos[0]))
            return M.AND(*out) if n > 1 else out[0]

        if shape == 0:      # all-dense AND (dense driver: not for the intersection kernel)
            root = group(1, 25, rng.randint(2, 5))
        elif shape == 1:    # two dense groups
            root = M.OR(group(1, 30, rng.randint(2, 3)), group(1, 30, rng.randint(2, 3)))
        elif shape == 2:    # dense group | hot-but-sparser group, or a group driven by ONE sparse keyword (hot keywords probed at its postings)
            if i % 12 == 2:
                root = M.OR(group(1, 20, 2), group(20, 300, 2))
            else:
                pos[0] += 3
                mixed = [leaf(rng.randint(1, 40), pos[0] - 2), leaf(rng.randint(1500, 30000), pos[0] - 1)]
                if rng.random() < 0.5:
                    mixed.append(leaf(rng.randint(1, 200), pos[0]))
                rng.shuffle(mixed)
                root = M.OR(group(1, 20, 2), M.AND(*mixed))
        elif shape == 3:    # dense group | single sparse keyword | single hot keyword
            root = M.OR(group(1, 20, rng.randint(2, 3)), group(2000, 60000, 1), group(1, 200, 1))
        elif shape == 4:    # three groups
            root = M.OR(group(1, 15, 2), group(1, 40, 2), group(5, 60, 3))
        else:               # single first, then groups
            root = M.OR(group(1, 300, 1), group(1, 25, 2), group(1, 25, 2))
        kwargs = {}
        if i % 5 == 2:
            kwargs["filters"] = rng.choice([[M.Filter(gid, 100, 299)], [M.Filter(gid, values=[5, 17, 900, 901])], []])
            kwargs["sort_keys"] = rng.choice([[], [M.SortKey(M.KEYPART_INT, ts, True)], [M.SortKey(M.KEYPART_ROWID, 0, True)]])
        qs.append(M.Query(root, ranker=M.RANK_BM25, field_weights=rng.choice([[10, 1], [1, 1], [0, 5], None]),
                          max_matches=rng.choice([1, 5, 100, 3000]), index_weight=rng.choice([1, 1, 3]), **kwargs))
    batch = synth["gpu"].prepare(qs + qs[:80])
    st = batch.stats()
    batch.free()
    assert st["class_queries"][6] + st["class_queries"][5] >= 150, st["class_queries"]     # hot groups: class 6, or the bitmap path of class 5
    assert _compare_batch(synth, qs + qs[:80]) == 0
    # the same programs with only the groups of dense keywords on the bitmap path of orbits_kernel (normally every all-hot group goes there)
    synth["gpu"].set_option("bits_dnf_div", 8)
    try:
        batch = synth["gpu"].prepare(qs + qs[:80])
        st = batch.stats()
        batch.free()
        assert st["class_queries"][5] >= 100, st["class_queries"]
        assert _compare_batch(synth, qs + qs[:80]) == 0
    finally:
        synth["gpu"].set_option("bits_dnf_div", 0)


def test_cfg4_mix_with_andnot(synth):
    queries = workload.cfg2_queries(n=200, seed=77, max_rank=50000, max_matches=1000, with_andnot=0.2)
    _compare_batch(synth, queries)


def test_negated_group_members_all_routes(synth):
    """`a b -c -d` programs: on the bitmaps when every keyword is hot (the negated one clears the group's presence word), on and_kernel when
    a sparse keyword leads (probe, reject if present), on dense tiles with group_neg=0; field limits on either side; missing keywords"""
    import random
    rng = random.Random(4242)

    def w(lo, hi, pos):
        return M.kw("t%07d" % rng.randint(lo, hi), pos)

    qs = []
    for i in range(240):
        shape = i % 6
        if shape == 0:      # all dense (repeated below, so every keyword is hot)
            pos = [w(1, 12, 1), w(1, 12, 2)]
            neg = [w(1, 20, 3)]
        elif shape == 1:    # two negated keywords
            pos = [w(1, 15, 1), w(1, 30, 2), w(1, 30, 3)]
            neg = [w(1, 10, 4), w(5, 40, 5)]
        elif shape == 2:    # sparse driver, dense negated keyword
            pos = [w(3000, 40000, 1), w(1, 50, 2)]
            neg = [w(1, 8, 3)]
        elif shape == 3:    # sparse negated keyword
            pos = [w(1, 20, 1), w(1, 20, 2)]
            neg = [w(2000, 60000, 3)]
        elif shape == 4:    # field limits
            pos = [w(1, 10, 1).fields(2), w(1, 25, 2)]
            neg = [w(1, 10, 3).fields(rng.choice([1, 2, 3]))]
        else:               # a keyword the index does not hold on the negated side
            pos = [w(1, 10, 1), w(1, 40, 2)]
            neg = [M.kw("nosuchword", 3)]
        left = M.AND(*pos)
        root = left
        for n in neg:
            root = M.ANDNOT(root, n)
        qs.append(M.Query(root, ranker=M.RANK_BM25, field_weights=rng.choice([[10, 1], [1, 1], None]), max_matches=rng.choice([5, 100, 2000])))
    batch = synth["gpu"].prepare(qs + qs[:120])
    st = batch.stats()
    batch.free()
    assert st["class_queries"][5] >= 100 and st["class_queries"][2] >= 40, st["class_queries"]
    assert _compare_batch(synth, qs + qs[:120]) == 0
    for mode in (1, 0):
        synth["gpu"].set_option("group_neg", mode)
        try:
            assert _compare_batch(synth, qs + qs[:120]) == 0
        finally:
            synth["gpu"].set_option("group_neg", 2)


def test_random_boolean_trees(synth):
    """fuzz: random AND/OR/ANDNOT/MAYBE trees, field limits, boosts, negative/zero weights, missing words, index weights"""
    queries = workload.random_boolean_queries(400, seed=1234)
    _compare_batch(synth, queries, allow_unsupported=0.1)


def test_rank_none_and_single_word(synth):
    qs = [M.Query(M.kw("t0000005", 1), ranker=r, max_matches=10, field_weights=[3, 2])
          for r in (M.RANK_PROXIMITY, M.RANK_MATCHANY, M.RANK_FIELDMASK, M.RANK_SPH04, M.RANK_WORDCOUNT)]   # single keyword under every ranker
    qs += [M.Query(M.kw("t0000005", 1), ranker=M.RANK_NONE, max_matches=10),
          M.Query(M.kw("t0000005", 1), ranker=M.RANK_BM25, max_matches=10),
          M.Query(M.kw("t0000005", 1), ranker=M.RANK_PROXIMITY_BM25, max_matches=10),   # single word -> WeightSum ranker
          M.Query(M.kw("nosuchword", 1), ranker=M.RANK_BM25, max_matches=10),
          M.Query(M.AND(M.kw("nosuchword", 1), M.kw("t0000005", 2)), ranker=M.RANK_BM25, max_matches=10)]
    _compare_batch(synth, qs)


def test_stopword_or_filter_sort_topk(synth):
    """config 5 shape: stop-word ORs + gid range filter + ORDER BY ts DESC, sorter-bound top-10k"""
    queries = workload.cfg5_queries(synth["gpu"], n=12, max_matches=10000)
    g = synth["gpu"].search(queries)
    c = synth["cpu"].search(queries)
    for i in range(len(queries)):
        a, b = g.get(i), c.get(i)
        helpers.assert_same_results(a, b, ctx="cfg5 %d" % i)
        assert a["sort_attr"] == b["sort_attr"]
        assert a["sort_attr"] == sorted(a["sort_attr"], reverse=True)


def test_values_filter_and_asc_sort(synth):
    gpu = synth["gpu"]
    gid, ts = gpu.attr_index("gid"), gpu.attr_index("ts")
    root = M.OR(M.kw("t0000003", 1), M.kw("t0000050", 2))
    qs = [M.Query(root, ranker=M.RANK_BM25, max_matches=200, filters=[M.Filter(gid, values=[5, 17, 900])],
                  sort_keys=[M.SortKey(M.KEYPART_INT, ts, False), M.SortKey(M.KEYPART_WEIGHT, 0, True)]),
          M.Query(root, ranker=M.RANK_BM25, max_matches=200, filters=[M.Filter(gid, 0, 499, exclude=True)],
                  sort_keys=[M.SortKey(M.KEYPART_WEIGHT, 0, False), M.SortKey(M.KEYPART_INT, gid, True)]),
          M.Query(root, ranker=M.RANK_BM25, max_matches=50, sort_keys=[M.SortKey(M.KEYPART_ROWID, 0, True)])]
    _compare_batch(synth, qs)


def test_cfg1_two_term_and_proximity_bm25(synth):
    """config 1: two-term AND, SPH_RANK_PROXIMITY_BM25 (LCS over merged hit streams), max_matches 1000"""
    queries = workload.cfg1_queries(n=150)
    assert _compare_batch(synth, queries) == 0


def test_cfg3_phrase_and_proximity(synth):
    """config 3: 2-3 word phrases and "a b c"~5 sampled from the corpus, PROXIMITY_BM25, top-1000"""
    queries = workload.cfg3_queries(synth["params"], n=200)
    assert _compare_batch(synth, queries) == 0
    g = synth["gpu"].search(queries)
    assert all(g.get(i)["total_found"] >= 1 for i in range(0, len(queries), 2))    # phrases are sampled from real docs


def test_hit_rankers_over_random_trees(synth):
    """fuzz of the hit stage: phrase/proximity nodes inside boolean trees, stop-word phrases (long hitlists),
    duplicated keywords (RankerState_Proximity_fn<.., HANDLE_DUPES>), field limits, all four rankers"""
    queries = workload.random_hit_queries(synth["params"], 300, seed=4321)
    assert _compare_batch(synth, queries, allow_unsupported=0.05) <= 15


def test_hit_level_operators_fuzz(synth):
    """NEAR (two keywords), BEFORE, NOTNEAR and quorum nodes over keywords that really sit close together, alone and inside
    AND / OR / ANDNOT trees, with position filters, field limits, repeated keywords, every ranker: the hit stage's FSMs
    (FSMmultinear_c, ExtOrder_c, ExtNotNear_c, ExtQuorum_c restated in hit_stage.cuh) against the oracle's"""
    queries = workload.random_hit_queries(synth["params"], 400, seed=977, with_hitops=True)
    n_ops = sum(1 for q in queries if any(n.op in (M.OP_NEAR, M.OP_BEFORE, M.OP_NOTNEAR, M.OP_QUORUM) for n in _walk(q.root)))
    assert n_ops > 80
    assert _compare_batch(synth, queries, allow_unsupported=0.05) <= 20
    g = synth["gpu"].search(queries)
    assert sum(1 for i in range(len(queries)) if g.get(i)["total_found"] > 0) > len(queries) // 2


def _walk(node):
    yield node
    for c in node.children:
        yield from _walk(c)


def test_hot_store_escape_and_plain_format(tmp_path):
    """dense hot-term store edge cases: documents with >= 255 hits of a shared keyword (escape list), a keyword present in
    every row, ragged last tile; plus the same corpus written with hit_format=plain"""
    import random
    rng = random.Random(7)
    docs = []
    for i in range(4500):
        # >= 255 hits of a hot keyword: every 911th document, and a dense stretch of them (the escape entries are chained per hash bucket)
        body = [("common", p + 1) for p in range(rng.choice([1, 1, 2, 3, 300 if i % 911 == 0 else 4]) if not 2000 <= i < 2600 else 255 + i % 50)]
        n0 = len(body)
        if i % 3 == 0:
            body += [("third", n0 + 1), ("third", n0 + 2)]
        if i % 7 == 0:
            body.append(("seventh", len(body) + 1))
        title = [("common", 1)] if i % 5 == 0 else [("rare%d" % (i % 40), 1)]
        docs.append({"id": 10 + i, "fields": [title, body], "attrs": []})
    for inline in (True, False):
        prefix = str(tmp_path / ("hot_%d" % inline))
        M.build_index(prefix, ["title", "body"], docs, hit_format_inline=inline)
        gpu, cpu = M.Index(prefix, device=0), helpers.OracleIndex(prefix)
        try:
            qs = [M.Query(M.OR(M.kw("common", 1), M.kw("seventh", 2)), ranker=M.RANK_BM25, field_weights=[3, 1], max_matches=5000),
                  M.Query(M.AND(M.kw("common", 1), M.kw("third", 2)), ranker=M.RANK_BM25, max_matches=100),
                  M.Query(M.ANDNOT(M.kw("common", 1), M.kw("third", 2)), ranker=M.RANK_BM25, max_matches=100),
                  M.Query(M.MAYBE(M.kw("seventh", 1), M.kw("common", 2)), ranker=M.RANK_BM25, max_matches=100),
                  M.Query(M.OR(M.AND(M.kw("third", 1), M.kw("rare3", 2)), M.AND(M.kw("common", 3).fields(1), M.kw("seventh", 4))), ranker=M.RANK_BM25, max_matches=100),
                  M.Query(M.AND(M.kw("rare7", 1), M.kw("third", 2), M.kw("common", 3)), ranker=M.RANK_BM25, max_matches=100),
                  M.Query(M.AND(M.kw("common", 1), M.kw("third", 2)), ranker=M.RANK_PROXIMITY_BM25, max_matches=100),
                  M.Query(M.PHRASE([("common", 1), ("third", 2)]), ranker=M.RANK_PROXIMITY_BM25, max_matches=100)]
            g, c = gpu.search(qs), cpu.search(qs)
            for i in range(len(qs)):
                helpers.assert_same_results(g.get(i), c.get(i), ctx="hot store, inline=%s, query %d" % (inline, i))
            assert g.get(0)["total_found"] == 4500
        finally:
            gpu.close()
            cpu.close()


def test_float_sort_key_on_gpu(tmp_path):
    """MGPU_KEYPART_FLOAT: the packed key holds the float's order-preserving integer image; same order as the oracle's float comparator"""
    import test_oracle_operators as TO
    prefix, _ = TO._float_corpus(tmp_path)
    gpu, cpu = M.Index(prefix, device=0), helpers.OracleIndex(prefix)
    try:
        qs = TO.float_sort_queries()
        g, c = gpu.search(qs), cpu.search(qs)
        for i in range(len(qs)):
            helpers.assert_same_results(g.get(i), c.get(i), ctx="float sort key, query %d" % i)
    finally:
        gpu.close()
        cpu.close()


def test_parsed_golden_queries_on_gpu(golden_cases, golden_indexes):
    """query text -> mgpu_parse_query -> mgpu_search_batch: the tree exactly as the restated parser shapes it (one-child AND / NOT
    wrappers of FixupNots, OR form of "..."/1, excluded flags of TagExcluded, n-ary NEAR of TransformNear) gives the reference's
    golden results on the CUDA path, or is refused where the hand-written tree is"""
    import test_query_parser as TQ
    ran = 0
    for case in golden_cases:
        gpu = M.Index(golden_indexes[case["name"]], device=0)
        try:
            for qi, q in enumerate(case["queries"]):
                query = helpers.golden_query(case, q)
                query.root = TQ.parse_case_query(case, qi)[0]
                r = gpu.search([query]).get(0)
                if q.get("gpu_unsupported"):
                    assert r["status"] == M.MGPU_E_UNSUPPORTED, (case["name"], q["text"], r["status"])
                    continue
                assert r["status"] == 0, (case["name"], q["text"], r["status"])
                got = list(zip(r["docid"], r["weight"]))
                if q.get("ids_only"):
                    got = [(d, 0) for d, _ in got]
                if q.get("limit"):
                    got = got[:q["limit"]]
                assert got == [tuple(m) for m in q["expect"]["matches"]], (case["name"], q["text"])
                assert r["total_found"] == q["expect"]["total_found"]
                ran += 1
        finally:
            gpu.close()
    assert ran >= 130


def test_sentence_paragraph_are_refused_on_gpu(synth):
    """the front-end parses SENTENCE / PARAGRAPH; the CUDA path (no index_sp boundary hits) must refuse them per query, never guess"""
    gpu = synth["gpu"]
    qs = [M.Query(M.parse_query("t0000001 SENTENCE t0000002", ["title", "body"])[0], max_matches=10),
          M.Query(M.parse_query("t0000001 t0000002", ["title", "body"])[0], max_matches=10),
          M.Query(M.parse_query('t0000001 PARAGRAPH "t0000002 t0000003"', ["title", "body"])[0], max_matches=10),
          M.Query(M.parse_query("", ["title", "body"])[0], max_matches=10),                         # nothing to search for: the parser's empty node
          M.Query(M.parse_query("@@relaxed @nosuch t0000001 t0000002", ["title", "body"])[0], max_matches=10)]
    r = gpu.search(qs)
    assert r.get(0)["status"] == M.MGPU_E_UNSUPPORTED and r.get(2)["status"] == M.MGPU_E_UNSUPPORTED
    assert r.get(1)["status"] == 0 and r.get(1)["total_found"] > 0
    assert r.get(3)["status"] == 0 and r.get(3)["total_found"] == 0
    c = synth["cpu"].search(qs)
    helpers.assert_same_results(r.get(4), c.get(4), ctx="relaxed query")


def test_golden_vectors_on_gpu_dict_crc(golden_cases, golden_indexes_crc, tmp_path):
    """dict=crc indexes (word ids = sphFNV64 of the keyword, src/sphinx.cpp:18263-18339): mgpu_index_open reads the id-keyed
    dictionary, query keywords are hashed at bind time; same golden results. Plus the sharded handle over two crc shards"""
    test_golden_vectors_on_gpu(golden_cases, golden_indexes_crc)
    docs_a = [{"id": 1 + i, "fields": [[("alpha", 1)], [("beta", 1), ("gamma%d" % (i % 3), 2)]], "attrs": []} for i in range(100)]
    docs_b = [{"id": 101 + i, "fields": [[("beta", 1)], [("alpha", 1), ("gamma%d" % (i % 5), 2)]], "attrs": []} for i in range(70)]
    pa, pb, pf = str(tmp_path / "a"), str(tmp_path / "b"), str(tmp_path / "full")
    M.build_index(pa, ["title", "body"], docs_a, dict_crc=True)
    M.build_index(pb, ["title", "body"], docs_b, dict_crc=True)
    M.build_index(pf, ["title", "body"], docs_a + docs_b, dict_crc=True)
    sh, cpu = M.ShardedIndex([pa, pb], [0, 0]), helpers.OracleIndex(pf)
    try:
        qs = [M.Query(M.AND(M.kw("alpha", 1), M.kw("gamma1", 2)), ranker=M.RANK_BM25, field_weights=[3, 1], max_matches=50),
              M.Query(M.OR(M.kw("gamma4", 1), M.kw("gamma2", 2), M.kw("missing", 3)), ranker=M.RANK_PROXIMITY_BM25, max_matches=50),
              M.Query(M.PHRASE([("beta", 1), ("gamma0", 2)]), ranker=M.RANK_PROXIMITY_BM25, max_matches=50)]
        g, c = sh.search(qs), cpu.search(qs)
        for i, q in enumerate(qs):
            helpers.assert_same_results(g.get(i), c.get(i), ctx="crc shards, query %d" % i)
            assert g.word_stats(i, len(q.keywords())) == c.word_stats(i, len(q.keywords()))
    finally:
        sh.close()
        cpu.close()


def test_golden_vectors_on_gpu(golden_cases, golden_indexes):
    """the reference's own golden results (17 model.bin files + gtest WeightBoundary): every query runs on the CUDA path, quorum /
    NEAR / BEFORE / NOTNEAR over plain keywords included; the trees flagged gpu_unsupported (make_golden.py: n-way NEAR, operator
    children that are not plain keywords) must be refused, never guessed"""
    ran = 0
    for case in golden_cases:
        gpu = M.Index(golden_indexes[case["name"]], device=0)
        try:
            for q in case["queries"]:
                query = helpers.golden_query(case, q)
                r = gpu.search([query]).get(0)
                ran += 1
                if q.get("gpu_unsupported"):
                    # shapes pinned in the oracle only: the CUDA path has to refuse them, never guess
                    assert r["status"] == M.MGPU_E_UNSUPPORTED, (case["name"], q["text"], r["status"])
                    continue
                assert r["status"] == 0, (case["name"], q["text"], r["status"])
                got = list(zip(r["docid"], r["weight"]))
                if q.get("ids_only"):       # SphinxQL `select *` results: the reference's model holds no weights for these;
                    got = [(d, 0) for d, _ in got]      # the weights are checked against the oracle below
                if q.get("limit"):
                    got = got[:q["limit"]]
                assert got == [tuple(m) for m in q["expect"]["matches"]], (case["name"], q["text"])
                assert r["total_found"] == q["expect"]["total_found"]
                if q.get("ids_only"):
                    cpu = helpers.OracleIndex(golden_indexes[case["name"]])
                    try:
                        helpers.assert_same_results(r, cpu.search([query]).get(0), ctx="%s %s" % (case["name"], q["text"]))
                    finally:
                        cpu.close()
        finally:
            gpu.close()
    assert ran == sum(len(c["queries"]) for c in golden_cases)


def test_full_size_properties(synth):
    """size-independent properties: AND subset of OR, total_found additivity over disjoint filters, idempotence"""
    gpu = synth["gpu"]
    a, b = M.kw("t0000004", 1), M.kw("t0000009", 2)
    gid = gpu.attr_index("gid")
    q_and = M.Query(M.AND(a, b), ranker=M.RANK_BM25, max_matches=1000)
    q_or = M.Query(M.OR(M.kw("t0000004", 1), M.kw("t0000009", 2)), ranker=M.RANK_BM25, max_matches=1000)
    q_lo = M.Query(M.OR(M.kw("t0000004", 1), M.kw("t0000009", 2)), ranker=M.RANK_BM25, max_matches=10, filters=[M.Filter(gid, 0, 499)])
    q_hi = M.Query(M.OR(M.kw("t0000004", 1), M.kw("t0000009", 2)), ranker=M.RANK_BM25, max_matches=10, filters=[M.Filter(gid, 500, 999)])
    rs = gpu.search([q_and, q_or, q_lo, q_hi, q_or])
    r_and, r_or, r_lo, r_hi, r_or2 = (rs.get(i) for i in range(5))
    da, db = gpu.word_stats("t0000004")[0], gpu.word_stats("t0000009")[0]
    assert r_and["total_found"] + r_or["total_found"] == da + db     # |A and B| + |A or B| = |A| + |B|
    assert r_lo["total_found"] + r_hi["total_found"] == r_or["total_found"]
    assert r_or == r_or2
    w = r_or["weight"]
    assert all(w[i] > w[i + 1] or (w[i] == w[i + 1] and r_or["rowid"][i] < r_or["rowid"][i + 1]) for i in range(len(w) - 1))


def test_rowid_range_shards_merge_equals_unsharded(tmp_path):
    """config 4 shape on one GPU: 3 logical rowid-range shards searched separately (global IDF inputs, global rowids in the
    keys), K keys per shard merged by shard_merge_kernel == the unsharded oracle, for doc-only and hit-ranked queries"""
    import torch
    from manticoresearch_b200 import distributed as D
    total, world, K = 30000, 3, 64
    full = str(tmp_path / "full")
    M.build_synthetic(full, M.SynthParams(total, vocab=1 << 15))
    cpu = helpers.OracleIndex(full)
    shards = []
    for r in range(world):
        first, n = D.shard_range(total, r, world)
        prefix = str(tmp_path / ("shard%d" % r))
        M.build_synthetic(prefix, M.SynthParams(n, first_doc=first, vocab=1 << 15))
        shards.append(M.Index(prefix, device=0, rowid_base=first))
    try:
        def make():
            return (workload.cfg2_queries(n=120, max_rank=8000, max_matches=K, with_andnot=0.15)
                    + workload.cfg1_queries(n=40, max_matches=K) + workload.cfg3_queries(M.SynthParams(total, vocab=1 << 15), n=40, max_matches=K))
        queries, plain = make(), make()
        words = sorted({k.word for q in queries for k in q.keywords()})
        gdf = {w: sum((s.word_stats(w) or (0, 0))[0] for s in shards) for w in words}
        D.apply_global_idf(queries, total, gdf)
        nq = len(queries)
        dev = torch.device("cuda", 0)
        stream = torch.cuda.current_stream()
        all_keys = torch.zeros((world, nq, K, 2), dtype=torch.int64, device=dev)
        all_counts = torch.zeros((world, nq), dtype=torch.int32, device=dev)
        totals = torch.zeros((world, nq), dtype=torch.int64, device=dev)
        for r, s in enumerate(shards):
            b = s.prepare(queries)
            b.run()
            b.export_keys(all_keys[r].data_ptr(), all_counts[r].data_ptr(), totals[r].data_ptr(), K)
            b.sync()        # export is asynchronous on the index's own stream; the merge below runs on torch's stream
            b.free()
        out_keys = torch.zeros((nq, K, 2), dtype=torch.int64, device=dev)
        out_counts = torch.zeros((nq,), dtype=torch.int32, device=dev)
        rc = M.lib().mgpu_merge_shard_keys(0, all_keys.data_ptr(), all_counts.data_ptr(), world, nq, K, out_keys.data_ptr(), out_counts.data_ptr(), stream.cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
        keys = out_keys.cpu().numpy().astype("uint64")
        counts = out_counts.cpu().tolist()
        tot = totals.sum(0).cpu().tolist()
        ref = cpu.search(plain)
        nonempty = 0
        for qi in range(nq):
            e = ref.get(qi)
            got = [D.unpack_key(int(keys[qi, i, 0]), int(keys[qi, i, 1])) for i in range(counts[qi])]
            assert [g[0] for g in got] == e["rowid"], ("rowid/order", qi)
            assert [g[1] for g in got] == e["weight"], ("weight", qi)
            assert tot[qi] == e["total_found"], ("total_found", qi)
            nonempty += e["total_found"] > 0
        assert nonempty > nq // 2
    finally:
        for s in shards:
            s.close()
        cpu.close()


def test_sharded_handle_equals_unsharded_oracle(tmp_path):
    """mgpu_sharded_open / mgpu_sharded_search_batch (C++: plan once, a host thread per shard, key exchange + shard_merge_kernel):
    3 rowid-range shards on one GPU (stream-ordered device copies stand in for ncclSend/ncclRecv) == the UNSHARDED oracle:
    global rowids, weights, order, total_found, document ids, keyword statistics; cfg2 mix incl. ANDNOT, hit-ranked queries,
    attribute sort + filter; keyword order decided from the global df even where a shard's local counts invert it"""
    from manticoresearch_b200 import distributed as D
    total, world, K = 30000, 3, 64
    params = M.SynthParams(total, vocab=1 << 15)
    full = str(tmp_path / "full")
    M.build_synthetic(full, params)
    cpu = helpers.OracleIndex(full)
    prefixes = []
    for r in range(world):
        first, n = D.shard_range(total, r, world)
        prefixes.append(str(tmp_path / ("shard%d" % r)))
        M.build_synthetic(prefixes[-1], M.SynthParams(n, first_doc=first, vocab=1 << 15))
    sh = M.ShardedIndex(prefixes, [0] * world)
    single = M.Index(full, device=0)
    try:
        assert sh.total_docs == total
        queries = (workload.cfg2_queries(n=150, max_rank=8000, max_matches=K, with_andnot=0.15)
                   + workload.cfg1_queries(n=40, max_matches=K) + workload.cfg3_queries(params, n=40, max_matches=K)
                   + workload.cfg5_queries(single, n=20, max_matches=200))
        # AND queries over keywords of nearly equal df: shard-local counts order them differently than the whole index does
        mids = [M.synth_keyword(r) for r in range(400, 440)]
        for i in range(0, 36, 3):
            queries.append(M.Query(M.AND(*[M.kw(w, j + 1) for j, w in enumerate(mids[i:i + 4])]), ranker=M.RANK_BM25, field_weights=[3, 1], max_matches=K))
        got = sh.search(queries)
        ref = cpu.search(queries)
        nonempty = 0
        for qi, q in enumerate(queries):
            g, e = got.get(qi), ref.get(qi)
            helpers.assert_same_results(g, e, ctx="sharded query %d" % qi)
            assert g["docid"] == e["docid"], ("docid", qi)
            nw = len(q.keywords())
            assert got.word_stats(qi, nw) == ref.word_stats(qi, nw), ("word stats", qi)
            nonempty += e["total_found"] > 0
        assert nonempty > len(queries) // 2
        st = sh.stats()
        assert st["n_shards"] == world and st["kernel_launches"] > world
        # a second batch on the same handle (buffers are reused)
        again = sh.search(queries[:50])
        for qi in range(50):
            helpers.assert_same_results(again.get(qi), ref.get(qi), ctx="sharded query %d, second batch" % qi)
    finally:
        sh.close()
        single.close()
        cpu.close()


def test_malformed_queries_are_rejected_not_crashed(synth):
    """the C ABI takes caller-built flattened trees: a child list naming its own node (or an ancestor), counts without arrays and
    negative counts come back as MGPU_E_BAD_QUERY for that query; the rest of the batch still runs"""
    good = M.Query(M.OR(M.kw("t0000100", 1), M.kw("t0002000", 2)), ranker=M.RANK_BM25, max_matches=20)
    cyc = M.Query(M.OR(M.AND(M.kw("t0000100", 1), M.OR(M.kw("t0000200", 2), M.kw("t0000300", 3))), M.kw("t0002000", 4)), ranker=M.RANK_BM25, max_matches=20)
    nofilt = M.Query(M.OR(M.kw("t0000100", 1), M.kw("t0002000", 2)), ranker=M.RANK_BM25, max_matches=20)
    negvals = M.Query(M.kw("t0000100", 1), ranker=M.RANK_BM25, max_matches=20, filters=[M.Filter(0, values=[1, 2])])
    queries = [good, cyc, nofilt, negvals, good]
    arr = M.pack_queries(queries)
    # query 1: make the inner OR node list the root among its children -> a cycle
    kids = arr[1].children
    inner = [i for i in range(arr[1].n_nodes) if arr[1].nodes[i].n_children == 2 and i != arr[1].root][-1]
    kids[arr[1].nodes[inner].first_child] = arr[1].root
    arr[2].n_filters = 2            # a count without an array
    arr[3].filters[0].n_values = -5
    rs = M.ResultSet(queries)
    synth["gpu"].search_packed(arr, len(queries), rs)
    ref = synth["cpu"].search([good])
    assert rs.get(1)["status"] == M.MGPU_E_BAD_QUERY
    assert rs.get(2)["status"] == M.MGPU_E_BAD_QUERY
    assert rs.get(3)["status"] == M.MGPU_E_BAD_QUERY
    for i in (0, 4):
        helpers.assert_same_results(rs.get(i), ref.get(0), ctx="good query %d next to malformed ones" % i)


def test_reference_built_index_on_gpu():
    """the v57 index the reference's own writer produced (tests/golden/ref_index, from test/test_406): mgpu_index_open takes it,
    K1 decodes the same postings as the oracle's reader, and a query over it returns the oracle's result"""
    prefix = os.path.join(helpers.ROOT, "tests", "golden", "ref_index", "index.0")
    gpu = M.Index(prefix, device=0)
    cpu = helpers.OracleIndex(prefix)
    try:
        assert gpu.total_docs == cpu.total_docs == 1
        for word in ("doc", "one"):
            assert gpu.word_stats(word) == cpu.word_stats(word)
            for a, b in zip(gpu.decode_doclist(word), cpu.decode_doclist(word)):
                assert (a == b).all()
        queries = [M.Query(M.AND(M.kw("doc", 1), M.kw("one", 2)), max_matches=5), M.Query(M.OR(M.kw("doc", 1), M.kw("nope", 2)), ranker=M.RANK_BM25, max_matches=5),
                   M.Query(M.PHRASE([("doc", 1), ("one", 2)]), max_matches=5), M.Query(M.PHRASE([("one", 1), ("doc", 2)]), max_matches=5)]
        g, c = gpu.search(queries), cpu.search(queries)
        for i in range(len(queries)):
            helpers.assert_same_results(g.get(i), c.get(i), ctx="reference-built index, query %d" % i)
    finally:
        gpu.close()
        cpu.close()


def test_close_is_refused_while_batches_live(tmp_path):
    """a batch keeps a pointer to its index: mgpu_index_close with a live batch is refused and the handle keeps working"""
    prefix = str(tmp_path / "life")
    docs = [{"id": 1 + i, "fields": [[("aa", 1)] + ([("bb", 2)] if i % 3 == 0 else [])], "attrs": []} for i in range(3000)]
    M.build_index(prefix, ["body"], docs)
    gpu = M.Index(prefix, device=0)
    q = [M.Query(M.OR(M.kw("aa", 1), M.kw("bb", 2)), ranker=M.RANK_BM25, max_matches=10)]
    batch = gpu.prepare(q)
    with pytest.raises(M.MgpuError):
        gpu.close()
    batch.run()
    first = batch.fetch().get(0)
    batch.free()
    again = gpu.search(q).get(0)
    assert first["rowid"] == again["rowid"] and first["weight"] == again["weight"] and len(first["rowid"]) > 0
    gpu.close()


def test_corrupt_dictionary_is_rejected_at_open(tmp_path):
    """mgpu_index_open validates what the kernels will trust: a keyword that claims more documents than its doclist can hold
    (four varints per record) comes back as MGPU_E_FORMAT, not as out-of-bounds device reads"""
    import shutil
    prefix = str(tmp_path / "ok")
    docs = [{"id": 1 + i, "fields": [[("w%d" % (i % 50), 1), ("all", 2)]], "attrs": []} for i in range(2000)]
    M.build_index(prefix, ["body"], docs)
    bad = str(tmp_path / "bad")
    for ext in ("sph", "spi", "spd", "spp", "spe", "spa", "spm"):
        if os.path.exists(prefix + "." + ext):
            shutil.copy(prefix + "." + ext, bad + "." + ext)
    # truncate .spd in the middle of the last doclists: offsets past the end / doclists shorter than 4 bytes per document
    size = os.path.getsize(bad + ".spd")
    with open(bad + ".spd", "r+b") as f:
        f.truncate(size // 2)
    with pytest.raises(M.MgpuError) as e:
        M.Index(bad, device=0)
    assert e.value.code == M.MGPU_E_FORMAT

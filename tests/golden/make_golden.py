#!/usr/bin/env python3
"""Generates tests/golden/golden_vectors.json from the reference's own functional tests.

Reads /root/reference/test/test_NNN/{test.xml,model.bin} (this container only; the GPU box has no
/root/reference, which is why the resulting JSON is committed).  The documents are transcribed from
each test.xml's <db_insert>/<custom_insert>; the expected results come straight out of model.bin
(PHP serialize(), see php_unserialize.py).  Query TREES are written by hand below, in the shape the
reference's parser produces (XQParser_t::AddOp n-ary nodes, one keyword per leaf, phrase/proximity
nodes with word lists, atom positions 1,2,3.. in query order, src/sphinxquery.cpp:1267, 1634-1678),
because the reference's parser needs bison; tests/test_query_parser.py checks mgpu_parse_query (the restated parser) against them.

Tree notation: ["kw", word, atompos, fieldmask?], ["and"|"or"|"andnot"|"maybe", child...],
["phrase", [[word,pos]...], fieldmask?], ["prox", N, [[word,pos]...]], ["quorum", N, [[word,pos]...]], ["near", N, child...], ["before", child...], ["notnear", N, must, not],
["sentence"|"paragraph", child...].
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from php_unserialize import php_unserialize  # noqa: E402

REF = "/root/reference/test"


def model(test):
    m = php_unserialize(open(os.path.join(REF, test, "model.bin"), "rb").read())
    return m[0]


def api_expect(r, want_query=None):
    if want_query is not None:
        assert r["query"] == want_query, (r["query"], want_query)
    matches = [[int(k), int(v["weight"])] for k, v in (r.get("matches") or {}).items()]
    words = {str(k): [int(v["docs"]), int(v["hits"])] for k, v in (r.get("words") or {}).items()}
    return {"matches": matches, "total_found": int(r["total_found"]), "words": words}


def ql_expect(r):
    rows = [[int(row["id"]), int(row["w"])] for row in r["rows"].values()]
    return {"matches": rows, "total_found": int(r["total_rows"]), "words": {}}


ALL = 0xFFFFFFFF
TITLE, BODY = 1, 2

out = {"source": "ravelry/manticoresearch test/test_015,016,017,019,030,037,041,052,054,055,059,094,114,115,116,133,138,322,349 model.bin + src/gtests/gtests_rtstuff.cpp:244-335 + api/libsphinxclient/smoke_ref.txt", "cases": []}

# ---------------------------------------------------------------------------------------------
# test_019 "extended queries", index `test` (min_word_len=2, ngram_len=1 for CJK)
# ---------------------------------------------------------------------------------------------
docs_019 = [
    (111, "", "basic query"),
    (222, "", "phrase query on steroids"),
    (333, "sample program", 'this is a test program that prints out "hello world" to the console'),
    (444, "", "china 吐我"),
    (555, "sample program two", "something written in basic | canon ef 16-35 lens"),
    (666, "sample program three", "something written in perl"),
    (777, "", "77 lies multiplied by 77"),
    (888, "", "agent 0077"),
    (999, "", "1234567812345678"),
    (901, "aaa", "aaa"),
    (902, "aaa", ""),
    (903, "", "aaa"),
    (910, "", "wordbefore\0\0wordafter"),
]
m19 = model("test_019")
q019 = [
    (0, "basic query", ["and", ["kw", "basic", 1], ["kw", "query", 2]]),
    (1, '"phrase query"', ["phrase", [["phrase", 1], ["query", 2]]]),
    (3, "@title sample @body world", ["and", ["kw", "sample", 1, TITLE], ["kw", "world", 2, BODY]]),
    (4, '"quorum query test"/1', ["quorum", 1, [["quorum", 1], ["query", 2], ["test", 3]]]),
    (5, '"quorum query test"/4', ["quorum", 4, [["quorum", 1], ["query", 2], ["test", 3]]]),
    (6, '"hello program"~3', ["prox", 3, [["hello", 1], ["program", 2]]]),
    (7, '"hello program"~4', ["prox", 4, [["hello", 1], ["program", 2]]]),
    (8, "吐", ["kw", "吐", 1]),
    (9, "我", ["kw", "我", 1]),
    (10, "basic | china", ["or", ["kw", "basic", 1], ["kw", "china", 2]]),
    (11, '"test program" | basic', ["or", ["phrase", [["test", 1], ["program", 2]]], ["kw", "basic", 3]]),
    # after "..."~N the parser's position is FixupAtomPos() = last word + 1 and the next keyword advances it again (src/sphinxquery.y:103,
    # src/sphinxquery.cpp:1267): `basic` sits at 4, found when mgpu_parse_query was checked against these trees
    (12, '"test that"~3 | basic', ["or", ["prox", 3, [["test", 1], ["that", 2]]], ["kw", "basic", 4]]),
    (14, "@title sample @body -basic", ["andnot", ["kw", "sample", 1, TITLE], ["kw", "basic", 2, BODY]]),
    (15, "-basic|perl sample", ["andnot", ["kw", "sample", 3], ["or", ["kw", "basic", 1], ["kw", "perl", 2]]]),
    (17, "77", ["kw", "77", 1]),
    (18, "0077", ["kw", "0077", 1]),
    (19, "@title test", ["kw", "test", 1, TITLE]),
    (20, "@!title aaa", ["kw", "aaa", 1, ALL & ~TITLE]),
    (21, "@!(title,body) aaa", ["kw", "aaa", 1, ALL & ~(TITLE | BODY)]),
    (30, "1234567812345678", ["kw", "1234567812345678", 1]),
    (33, "canon 16 35", ["and", ["kw", "canon", 1], ["kw", "16", 2], ["kw", "35", 3]]),
]
case = {"name": "test_019", "fields": ["title", "body"], "min_word_len": 2,
        "docs": [{"id": d[0], "fields": [d[1], d[2]]} for d in docs_019], "queries": []}
for qi, text, tree in q019:
    case["queries"].append({"text": text, "tree": tree, "ranker": "proximity_bm25", "expect": api_expect(m19[qi], text)})
out["cases"].append(case)

# test_019 index `fld` (documents 1000..1010 of the same table): the "regression for ranker fieldmask" query. SPH_RANK_FIELDMASK
# (RankerState_Fieldmask_fn, src/sphinxsearch.cpp:1582-1610): weight = bit mask of the fields the keywords hit.
docs_019f = [(1000, "spec1", "dummy1"), (1001, "spec1 dummy1", ""), (1002, "", "spec1 dummy1"), (1003, "spec2 dummy2 text2", ""),
             (1004, "spec2", "dummy2 text2"), (1005, "spec3", "dummy3 text3"), (1006, "spec3 dummy3 text3", "spec3"),
             (1007, "spec4 dummy4", "text4"), (1008, "spec4", "dummy4 text4"), (1009, "spec5 of my", "dummy5"), (1010, "spec5", "of my text5")]
case = {"name": "test_019_fld", "fields": ["title", "body"], "min_word_len": 1,
        "docs": [{"id": d[0], "fields": [d[1], d[2]]} for d in docs_019f], "queries": []}
r = m19[52]
assert "ranker='fieldmask'" in r["sphinxql"] and "'spec1 | dummy1'" in r["sphinxql"]
case["queries"].append({"text": r["sphinxql"], "tree": ["or", ["kw", "spec1", 1], ["kw", "dummy1", 2]], "ranker": "fieldmask", "sort": "id_asc",
                        "expect": {"matches": [[int(row["id"]), int(row["weight()"])] for row in r["rows"].values()],
                                   "total_found": int(r["total_rows"]), "words": {}}})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_016 "expr sorting vs filters": legacy SPH_MATCH_ANY. PrepareQueryEmulation (src/searchd.cpp:2140-2185) rewrites the raw
# query `test it` into `"test it"/1` and picks SPH_RANK_MATCHANY (RankerState_MatchAny_fn, src/sphinxsearch.cpp:1611-1668); the
# threshold-1 quorum becomes an OR chain over the keywords (src/searchnode.cpp:1638-1688). sortmode=expr "@weight" is relevance order,
# "-@weight" is weight ascending.
# ---------------------------------------------------------------------------------------------
docs_016 = [(111, 1, "this is test"), (222, 1, "just a test"), (333, 2, "for test-ing purposes"), (444, 1, "lets test it")]
m16 = model("test_016")
any_tree = ["quorum", 1, [["test", 1], ["it", 2]]]
case = {"name": "test_016", "fields": ["body"], "attrs": ["group_id"], "min_word_len": 1,
        "docs": [{"id": d[0], "fields": [d[2]], "attrs": [d[1]]} for d in docs_016], "queries": []}
case["queries"].append({"text": m16[0]["query"], "tree": any_tree, "ranker": "matchany", "filters": [["group_id", 1, 1]], "expect": api_expect(m16[0], "test it")})
case["queries"].append({"text": m16[1]["query"], "tree": any_tree, "ranker": "matchany", "sort": "weight_asc", "expect": api_expect(m16[1], "test it")})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_017 "phrase matching vs stop words and short words" (stopwords a/the/and/of, min_word_len=3; both consume a position, in the
# documents and in the query). Legacy SPH_MATCH_PHRASE = `"..."` ranked by SPH_RANK_PROXIMITY (PrepareQueryEmulation,
# src/searchd.cpp:2140-2185; ExtRanker_State_T<RankerState_Proximity_fn<false,...>>: weight = sum of field weight * LCS, no BM25);
# the same phrases in extended2 mode are ranked by the default PROXIMITY_BM25.
# ---------------------------------------------------------------------------------------------
docs_017 = [(1, "walking shoes"), (2, "try walking in my shoes"), (3, "Microsoft. The Office."), (4, "Microsoft Office")]
m17 = model("test_017")
ph17 = [["phrase", [["walking", 1], ["shoes", 2]]], ["phrase", [["walking", 1], ["shoes", 4]]],
        ["phrase", [["microsoft", 1], ["office", 2]]], ["phrase", [["microsoft", 1], ["office", 3]]]]
case = {"name": "test_017", "fields": ["body"], "min_word_len": 3, "stopwords": ["a", "the", "and", "of"],
        "docs": [{"id": d[0], "fields": [d[1]]} for d in docs_017], "queries": []}
for qi in range(8):
    case["queries"].append({"text": m17[qi]["query"], "tree": ph17[qi % 4], "ranker": "proximity" if qi < 4 else "proximity_bm25", "expect": api_expect(m17[qi])})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_015 "phrase matching vs duplicate keywords" (stop word "2" consumes a position; '-' separates): phrases and a long AND with
# repeated keywords (HasQwordDupes: RankerState_Proximity_fn<*,true>), legacy phrase mode (SPH_RANK_PROXIMITY) and extended2
# (PROXIMITY_BM25)
# ---------------------------------------------------------------------------------------------
docs_015 = [(111, "lets test foo bar baz bar stuff"), (222, "bar baz foo"), (333, "foo baz bar"), (444, "i i did it it"),
            (555, "zee lord of zee rings"), (666, "Braun 370-2 3702 370 2 Braun Series 3 - 370 Shaver")]
m15 = model("test_015")
ph15 = [["phrase", [["bar", 1], ["baz", 2], ["bar", 3]]], ["phrase", [["foo", 1], ["bar", 2], ["baz", 3], ["bar", 4]]],
        ["phrase", [["i", 1], ["did", 2], ["it", 3]]], ["phrase", [["zee", 1], ["lord", 2], ["of", 3], ["zee", 4], ["rings", 5]]]]
case = {"name": "test_015", "fields": ["body"], "min_word_len": 1, "stopwords": ["2"],
        "docs": [{"id": d[0], "fields": [d[1]]} for d in docs_015], "queries": []}
for qi in range(8):
    case["queries"].append({"text": m15[qi]["query"], "tree": ph15[qi % 4], "ranker": "proximity" if qi < 4 else "proximity_bm25", "expect": api_expect(m15[qi])})
braun = ["and"] + [["kw", w, p] for w, p in [("braun", 1), ("370", 2), ("3702", 4), ("370", 5), ("braun", 7), ("series", 8), ("3", 9), ("370", 10), ("shaver", 11)]]
case["queries"].append({"text": m15[8]["query"], "tree": braun, "ranker": "proximity_bm25", "expect": api_expect(m15[8])})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_030 "ext2 ranking" (stopwords a/the/and/of, min_word_len=3): AND queries under PROXIMITY_BM25, a stop word inside the query
# keeps its atom position
# ---------------------------------------------------------------------------------------------
docs_030 = [(1, "one two three"), (2, "one two three four"), (3, "one then two then three then four"),
            (4, "senior pastor of Riverside church"), (5, "senior pastor and the Riverside church")]
m30 = model("test_030")
case = {"name": "test_030", "fields": ["body"], "min_word_len": 3, "stopwords": ["a", "the", "and", "of"],
        "docs": [{"id": d[0], "fields": [d[1]]} for d in docs_030], "queries": []}
case["queries"].append({"text": m30[0]["query"], "tree": ["and", ["kw", "one", 1], ["kw", "two", 2], ["kw", "three", 3]],
                        "ranker": "proximity_bm25", "expect": api_expect(m30[0], "one two three")})
case["queries"].append({"text": m30[1]["query"], "tree": ["and", ["kw", "senior", 1], ["kw", "pastor", 2], ["kw", "riverside", 4], ["kw", "church", 5]],
                        "ranker": "proximity_bm25", "expect": api_expect(m30[1], "senior pastor of riverside church")})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_094 "proximity queries": 257 repetitions of `one two` in one document (BYTE LCS wraps: weight 255500), 6-keyword proximities
# in and out of query order
# ---------------------------------------------------------------------------------------------
docs_094 = [(1, "one two " * 257), (2, "two one"), (3, "aa bb cc dd ee ff")]
m94 = model("test_094")
case = {"name": "test_094", "fields": ["body"], "min_word_len": 1,
        "docs": [{"id": d[0], "fields": [d[1]]} for d in docs_094], "queries": []}
for qi, words in [(0, ["one", "two"]), (1, ["aa", "bb", "cc", "dd", "ee", "ff"]), (2, ["aa", "bb", "dd", "cc", "ee", "ff"]), (3, ["aa", "bb", "ee", "ff"])]:
    case["queries"].append({"text": m94[qi]["query"], "tree": ["prox", 10, [[w, i + 1] for i, w in enumerate(words)]],
                            "ranker": "proximity_bm25", "expect": api_expect(m94[qi])})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_059 "phrase boundaries" (phrase_boundary = '.', phrase_boundary_step = 10: `one. two` puts 10 extra positions between the
# words), index `test`: field-start anchor, phrase and proximity across the boundary
# ---------------------------------------------------------------------------------------------
docs_059 = [(1, "first."), (2, " second"), (3, "one. two three"), (4, "one two three")]
m59 = model("test_059")
case = {"name": "test_059", "fields": ["body"], "min_word_len": 1, "phrase_boundary": ".", "phrase_boundary_step": 10,
        "docs": [{"id": d[0], "fields": [d[1]]} for d in docs_059], "queries": []}
for qi, tree in [(0, ["kw", "second", 1, ALL, {"start": 1}]), (1, ["phrase", [["one", 1], ["two", 2]]]),
                 (2, ["prox", 10, [["one", 1], ["two", 2]]]), (3, ["prox", 11, [["one", 1], ["two", 2]]])]:
    case["queries"].append({"text": m59[qi]["query"], "tree": tree, "ranker": "proximity_bm25", "expect": api_expect(m59[qi])})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_054 "quorum", index `test`: real ExtQuorum_c nodes (src/searchnode.cpp:4319-4650) incl. repeated keywords and percent
# thresholds (the caller resolves "/0.59" to an absolute count: ExtQuorum_c::GetThreshold with the parser's truncated percent;
# 0 = "rounds to nothing", which still builds a quorum node with threshold 1). The CUDA path implements only the degenerate
# quorums (threshold 1 -> OR, threshold >= words -> AND) and must answer MGPU_E_UNSUPPORTED for the others: "gpu_unsupported".
# ---------------------------------------------------------------------------------------------
docs_054 = [(1, "hello world"), (2, "one two three four five")]      # source test1: document_id in (1, 2)
m54 = model("test_054")
case = {"name": "test_054", "fields": ["text"], "min_word_len": 1,
        "docs": [{"id": d[0], "fields": [d[1]]} for d in docs_054], "queries": []}
for qi, thr, words in [(0, 1, "hello heaven"), (1, 2, "hello from above"), (2, 3, "one two foo bar"), (3, 3, "one two two bar"),
                       (4, 2, "one two two bar"), (5, 3, "two two one three"), (6, 3, "two two one foo"), (18, 0, "one world"),
                       (19, 2, "five tree oak one two hive"), (20, 3, "five tree oak one two hive"), (21, 4, "five tree oak one two hive")]:
    ws = words.split()
    q = {"text": m54[qi]["query"].strip(), "tree": ["quorum", thr, [[w, i + 1] for i, w in enumerate(ws)]], "ranker": "proximity_bm25",
         "expect": api_expect(m54[qi])}
    if thr != 1 and thr < len(ws):
        q["gpu_unsupported"] = True
    case["queries"].append(q)
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_138 "quorum vs decreased matched word": keywords run out while the quorum node walks on (ExtQuorum_c drops an exhausted
# keyword with RemoveFast and recounts what is left); test2 holds 512 copies of `world space` = several document chunks
# ---------------------------------------------------------------------------------------------
data_138 = {1: "world space", 2: "one", 3: "two", 4: "world", 5: "space", 6: "unused1", 7: "unused2"}
m138 = model("test_138")
q138 = ["quorum", 2, [[w, i + 1] for i, w in enumerate("one two unused1 unused2 space world".split())]]
case = {"name": "test_138_test1", "fields": ["text"], "min_word_len": 1,
        "docs": [{"id": k, "fields": [v]} for k, v in data_138.items()], "queries": []}
case["queries"].append({"text": m138[0]["query"], "tree": q138, "ranker": "proximity_bm25", "gpu_unsupported": True, "expect": api_expect(m138[0])})
out["cases"].append(case)
case = {"name": "test_138_test2", "fields": ["text"], "min_word_len": 1,
        "docs": [{"id": i, "fields": [data_138[1]]} for i in range(1, 513)] + [{"id": 600 + k, "fields": [v]} for k, v in data_138.items()], "queries": []}
case["queries"].append({"text": m138[1]["query"], "tree": q138, "ranker": "proximity_bm25", "gpu_unsupported": True, "expect": api_expect(m138[1])})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_115 "NEAR syntax", index `idx` (documents with id < 100; blend_chars '-' only adds whole-token variants, the parts keep the
# positions a plain split gives them, so '-' is indexed as a separator here). NEAR trees arrive in the shape TransformNear
# (src/sphinx.cpp:15046-15100) leaves them in: `(a b c) NEAR/3 d` is the 4-way `a NEAR/3 b NEAR/3 c NEAR/3 d`; equal distances chain
# into one n-ary node (XQParser_t::AddOp), different ones nest. Oracle only (FSMmultinear_c); the CUDA path must refuse NEAR.
# ---------------------------------------------------------------------------------------------
import re  # noqa: E402
xml115 = open(os.path.join(REF, "test_115", "test.xml"), encoding="utf-8").read()
ins115 = xml115[xml115.index("<db_insert>"):xml115.index("</db_insert>")]
docs_115 = [(int(m.group(1)), m.group(2)) for m in re.finditer(r"\(\s*(\d+),\s*'((?:[^']|'')*)'\s*\)", ins115) if int(m.group(1)) < 100]
docs_115.append((21, "zwei " + "oy vey ho ho ho " * 1024))
docs_115.sort()
assert [d[0] for d in docs_115] == list(range(1, 18)) + [20, 21, 22]
m115 = model("test_115")


def K(w, p):
    return ["kw", w, p]


def PH(ws, p0):
    return ["phrase", [[w, p0 + i] for i, w in enumerate(ws)]]


q115 = [(0, ["near", 2, PH("ab", 1), PH("cd", 3)]),
        (1, ["near", 2, PH("cd", 1), PH("ab", 3)]),
        (2, ["and", K("a", 1), ["near", 2, K("b", 2), K("c", 3)], K("d", 4)]),
        (3, ["near", 2, ["near", 5, ["near", 2, K("a", 1), K("b", 2)], K("c", 3)], K("d", 4)]),
        (4, ["near", 3, K("a", 1), K("b", 2), K("c", 3), K("d", 4)]),
        (5, ["near", 3, K("a", 1), K("d", 2), K("b", 3), K("c", 4)]),
        (6, ["near", 3, K("a", 1), K("b", 2), K("c", 3), K("d", 4)]),
        (7, ["near", 2, K("burden", 1), K("financial", 2), K("share", 3)]),
        (8, ["near", 2, K("burden", 1), K("share", 2), K("financial", 3)]),
        (9, ["near", 2, K("share", 1), K("financial", 2), K("burden", 3)]),
        (10, ["near", 2, K("financial", 1), K("share", 2), K("burden", 3)]),
        (11, ["near", 2, PH("ab", 1), PH("cd", 3), PH("fg", 5)]),
        (12, ["near", 3, K("a", 1), K("b", 2), K("c", 3), K("d", 4)]),
        (18, ["near", 3, K("five", 1), K("one", 2)]),
        (19, ["near", 3, K("six", 1), K("one", 2)]),
        (20, ["near", 3, ["or", K("five", 1), K("six", 2)], K("one", 3)])]
case = {"name": "test_115", "fields": ["title"], "min_word_len": 1,
        "docs": [{"id": d[0], "fields": [d[1].replace("-", " ")]} for d in docs_115], "queries": []}
for qi, tree in q115:
    case["queries"].append({"text": m115[qi]["query"].strip(), "tree": tree, "ranker": "proximity_bm25", "gpu_unsupported": True, "expect": api_expect(m115[qi])})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_052 "before operator", index `test` (documents 1..5, fields title / text): ExtOrder_c (src/searchnode.cpp:4657-4935) over keywords,
# phrases, OR groups, anchored keywords and a degenerate quorum. Note the query position after `"zzz aaa"/1`: the threshold token takes
# a position of its own (the parser only rewinds it for proximity, FixupAtomPos), so `bbb` sits at 4. Oracle only (OrderNode_c); the
# CUDA path must refuse BEFORE.
# ---------------------------------------------------------------------------------------------
docs_052 = [(1, "aaa bbb", "ccc ddd"), (2, "xxx", "ccc ddd eee fff ggg"), (3, "yyy", "one one one two three"),
            (4, "zzz", "one two three one three one two four one two three four"), (5, "", "a b c d e f g")]
m52 = model("test_052")


def KM(w, p, **mods):
    return ["kw", w, p, ALL, mods]


def BF(*children):
    return ["before"] + list(children)


def SEQ(words):
    return BF(*[K(w, i + 1) for i, w in enumerate(words.split())])


q52 = {0: SEQ("aaa ccc"), 1: SEQ("aaa bbb ccc"), 2: SEQ("aaa ccc ddd"), 3: SEQ("ccc ddd"), 4: SEQ("ccc eee fff"), 5: SEQ("ccc ddd ggg"),
       6: SEQ("ccc ddd xxx"), 7: SEQ("eee ddd ggg"), 8: SEQ("one two three"), 9: SEQ("one three"), 10: SEQ("one one three"),
       11: SEQ("one one one three"), 12: SEQ("one one one one three"), 13: SEQ("one two three four"),
       14: BF(PH("abc", 1), K("b", 4), K("c", 5), K("d", 6)), 15: BF(PH("abc", 1), K("c", 4), K("d", 5), K("e", 6)),
       16: BF(PH("abc", 1), K("e", 4), K("f", 5), K("g", 6)), 17: BF(K("a", 1), PH("bcd", 2), K("e", 5)),
       18: BF(PH("abcd", 1), PH("def", 5)), 19: BF(PH("abcd", 1), PH("efg", 5)),
       20: BF(["or", K("ccc", 1), ["phrase", [["ddd", 2], ["eee", 3]]]], ["or", K("ddd", 4), K("ggg", 5)]),
       21: BF(K("ccc", 1), KM("ddd", 2, end=1)), 22: BF(KM("one", 1, start=1), K("two", 2), KM("three", 3, end=1)),
       23: BF(KM("one", 1, start=1), ["phrase", [["one", 2], ["one", 3]]], K("two", 4), KM("three", 5, end=1)),
       24: BF(["quorum", 1, [["zzz", 1], ["aaa", 2]]], K("bbb", 4)), 25: BF(["quorum", 1, [["zzz", 1], ["aaa", 2]]], K("ddd", 4))}
case = {"name": "test_052", "fields": ["title", "text"], "min_word_len": 1,
        "docs": [{"id": d[0], "fields": [d[1], d[2]]} for d in docs_052], "queries": []}
for qi, tree in q52.items():
    case["queries"].append({"text": m52[qi]["query"].strip(), "tree": tree, "ranker": "proximity_bm25", "gpu_unsupported": True, "expect": api_expect(m52[qi])})
out["cases"].append(case)

# test_052 index `test1` (document 6): NEAR nodes over a phrase and an OR group as the operands of BEFORE
case = {"name": "test_052_test1", "fields": ["title", "text"], "min_word_len": 1,
        "docs": [{"id": 6, "fields": ["", "h1 h2 h3 h4 h5"]}], "queries": []}


def NR52(p0):
    return ["near", 5, ["phrase", [["h1", p0], ["h2", p0 + 1]]], ["or", K("h3", p0 + 2), K("h4", p0 + 3), K("h5", p0 + 4)]]


case["queries"].append({"text": m52[26]["query"].strip(), "tree": BF(NR52(1), NR52(6)), "ranker": "proximity_bm25", "gpu_unsupported": True,
                        "expect": api_expect(m52[26])})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_349 "NOTNEAR", index `idx`: `select * ... order by id asc` results carry no weights, so these pin the matched document sets only
# ("ids_only"). ExtNotNear_c (src/searchnode.cpp:5325-5478) over keywords, phrases, proximity, OR groups and a nested NOTNEAR.
# Oracle only (NotNearNode_c); the CUDA path must refuse NOTNEAR.
# ---------------------------------------------------------------------------------------------
xml349 = open(os.path.join(REF, "test_349", "test.xml"), encoding="utf-8").read()
a349 = xml349.index("INSERT INTO `test_table` VALUES")
ins349 = xml349[a349:xml349.index("</db_insert>", a349)]
docs_349 = [(int(m.group(1)), m.group(2)) for m in re.finditer(r"\(\s*(\d+),\s*'((?:[^']|'')*)'\s*\)", ins349)]
docs_349.append((21, "zwei " + "oy vey ho ho ho " * 1023 + "oy vey ho h"))
docs_349.sort()
assert [d[0] for d in docs_349] == list(range(1, 18)) + [20, 21, 22]
m349 = model("test_349")


def NN(n, left, right):
    return ["notnear", n, left, right]


def PW(words, p0):
    return ["phrase", [[w, p0 + i] for i, w in enumerate(words.split())]]


q349 = {0: NN(1, K("a", 1), K("c", 2)), 1: NN(2, K("a", 1), K("c", 2)), 2: NN(3, K("a", 1), K("c", 2)), 3: NN(5, K("a", 1), K("c", 2)),
        4: NN(6, K("a", 1), K("c", 2)), 5: NN(7, K("a", 1), K("c", 2)), 6: NN(15, K("a", 1), K("c", 2)),
        7: NN(1, K("b", 1), K("c", 2)), 8: NN(2, K("b", 1), K("c", 2)), 9: NN(2, PH("ab", 1), K("c", 3)), 10: NN(3, PH("ab", 1), K("c", 3)),
        11: NN(3, K("a", 1), ["or", K("c", 2), K("d", 3)]),
        13: NN(3, K("a", 1), PW("c x d", 2)), 14: NN(9, K("a", 1), PW("c x d", 2)), 15: NN(11, K("a", 1), PW("c x x d", 2)),
        16: NN(1, K("oy", 1), K("ho", 2)), 17: NN(2, K("oy", 1), K("ho", 2)), 18: NN(2, K("zwei", 1), K("ho", 2)), 19: NN(4, K("zwei", 1), K("ho", 2)),
        20: NN(1, K("zwei", 1), K("vey", 2)), 21: NN(2, K("zwei", 1), K("vey", 2)), 22: NN(1, K("vey", 1), K("ho", 2)), 23: NN(1, K("vey", 1), K("oy", 2)),
        24: NN(1, K("d", 1), K("a", 2)), 25: NN(3, K("d", 1), K("a", 2)), 26: NN(1, K("c", 1), K("x", 2)), 27: NN(2, K("c", 1), K("x", 2)),
        28: NN(3, K("c", 1), K("x", 2)), 29: NN(2, K("x", 1), K("c", 2)),
        30: NN(2, NN(3, PH("ab", 1), ["or", K("d", 3), K("e", 4)]), K("c", 5)),
        31: NN(1, ["prox", 4, [["a", 1], ["b", 2]]], K("c", 4)),
        32: NN(1, ["or", PH("ab", 1), PW("a x b", 3)], K("c", 6))}
case = {"name": "test_349", "fields": ["title"], "attrs": ["gid"], "min_word_len": 1,
        "docs": [{"id": d[0], "fields": [d[1].replace("-", " ")], "attrs": [11]} for d in docs_349], "queries": []}
for qi, tree in q349.items():
    r = m349[qi]
    case["queries"].append({"text": r["sphinxql"].strip(), "tree": tree, "ranker": "proximity_bm25", "sort": "id_asc", "ids_only": True, "gpu_unsupported": True,
                            "expect": {"matches": [[int(row["id"]), 0] for row in (r.get("rows") or {}).values()], "total_found": int(r["total_rows"]), "words": {}}})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_037 "rankers", index `test` (stem_ru/stem_en: both doc and query words are stemmed alike, so
# treating the Russian words as opaque tokens gives the same postings; word stats are NOT compared)
# ---------------------------------------------------------------------------------------------
docs_037 = [(1, "зимние шины диски чего то тут зимние шины", ""),
            (2, "test doc two", "second stupid test document with random content")] + [(i, "filler", "filler") for i in range(3, 11)]
m37 = model("test_037")
ph = ["phrase", [["зимние", 1], ["шины", 2]]]
case = {"name": "test_037", "fields": ["title", "body"], "min_word_len": 1,
        "docs": [{"id": d[0], "fields": [d[1], d[2]]} for d in docs_037], "queries": [], "skip_word_stats": True}
for qi, tree, ranker in [(0, ph, "proximity_bm25"), (1, ph, "bm25"), (2, ph, "none"), (3, ph, "wordcount"), (4, ["kw", "test", 1, TITLE], "bm25")]:
    case["queries"].append({"text": m37[qi]["query"], "tree": tree, "ranker": ranker, "expect": api_expect(m37[qi])})
out["cases"].append(case)

# test_037 index `test2` (docs 11..16): `market street` ranked by SPH04 (RankerState_ProximityBM25Exact_fn: LCS*4 + head*2 + exact)
docs_037b = [(11, "market street", ""), (12, "market street west", ""), (13, "north market street", ""),
             (14, "farmers market street north", ""), (15, "flower street market", ""), (16, "market street is so very market street", "")]
case = {"name": "test_037_test2", "fields": ["title", "body"], "min_word_len": 1,
        "docs": [{"id": d[0], "fields": [d[1], d[2]]} for d in docs_037b], "queries": []}
case["queries"].append({"text": m37[5]["query"], "tree": ["and", ["kw", "market", 1], ["kw", "street", 2]], "ranker": "sph04", "expect": api_expect(m37[5])})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_322 "field weights" (fields title, body, spam; default ranker = proximity_bm25; negative weights)
# ---------------------------------------------------------------------------------------------
docs_322 = [(1, "|sample program", "|program flow direct", "|sample program flow"),
            (2, "|one sample program", "|program rev flow", "|one rev flow"),
            (3, "|sample two program", "|sub program flow", "|two sub program"),
            (100, "unsigned", "", "")]
m322 = model("test_322")
pf = ["and", ["kw", "program", 1], ["kw", "flow", 2]]
case = {"name": "test_322", "fields": ["title", "body", "spam"], "min_word_len": 1,
        "docs": [{"id": d[0], "fields": list(d[1:])} for d in docs_322], "queries": []}
for qi, fw, ranker in [(1, [1, 2, 1], "proximity_bm25"), (2, [1, 2, 10], "proximity_bm25"), (3, [1, 2, 10], "wordcount"),
                       (4, [1, 2, 0], "proximity_bm25"), (5, [1, 2, 0], "wordcount"), (6, [1, 2, -2], "proximity_bm25"),
                       (7, [1, 2, -10], "proximity_bm25"), (8, [1, 2, -10], "wordcount")]:
    case["queries"].append({"text": m322[qi]["sphinxql"], "tree": pf, "ranker": ranker, "field_weights": fw, "expect": ql_expect(m322[qi])})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_116 "bound cases of the proximity node" (>32 docs per chunk, 532 matches in one doc)
# ---------------------------------------------------------------------------------------------
docs_116 = []
line = ""
for i in range(10):
    docs_116.append("a %s b" % line)
    line += "x "
docs_116 += ["e x f"] * 510
docs_116.append("e x f x x x e x f x x e x f x x")
docs_116.append(" y y i x j" * 532)
m116 = model("test_116")
case = {"name": "test_116", "fields": ["title"], "min_word_len": 1,
        "docs": [{"id": i + 1, "fields": [t]} for i, t in enumerate(docs_116)], "queries": []}
for qi, tree in [(0, ["prox", 3, [["a", 1], ["b", 2]]]), (1, ["prox", 2, [["e", 1], ["f", 2]]]), (2, ["prox", 2, [["i", 1], ["j", 2]]])]:
    case["queries"].append({"text": m116[qi]["query"], "tree": tree, "ranker": "wordcount", "limit": 20, "expect": api_expect(m116[qi])})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_114 "phrase matching vs hit buffer boundary" (520 phrase hits in one document)
# ---------------------------------------------------------------------------------------------
docs_114 = ["aaaa bbbb cccc dddd"] * 510 + ["aaaa bbbb x aaaa bbbb " + " x cccc dddd" * 520]
m114 = model("test_114")
case = {"name": "test_114", "fields": ["title"], "min_word_len": 1,
        "docs": [{"id": i + 1, "fields": [t]} for i, t in enumerate(docs_114)], "queries": []}
for qi, tree in [(0, ["phrase", [["aaaa", 1], ["bbbb", 2]]]), (1, ["phrase", [["cccc", 1], ["dddd", 2]]])]:
    case["queries"].append({"text": m114[qi]["query"], "tree": tree, "ranker": "wordcount", "limit": 20, "expect": api_expect(m114[qi])})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# test_055 "position anchors": ^word / word$ (ExtTermPos_T, TERM_POS_FIELD_START / FIELD_END), incl. a 600-hit document
# rows in insertion order (the self-joins copy rows with document_id>=10 six times)
# ---------------------------------------------------------------------------------------------
rows_055 = [(1, "", "one"), (2, "", "one and two"), (3, "", "one but not the other one"), (4, "", "two and one"), (9, "", "other three")]
rows_055 += [(i, "", "three") for i in range(10, 20)]
for shift in (10, 20, 40, 80, 160, 320):
    rows_055 += [(d + shift, t, b) for d, t, b in rows_055 if d >= 10]
rows_055.append((2000, "badger " * 600, "badger badger mushroom"))
rows_055 += [(1000, "", "other"), (1001, "", "other three blind mice")]
m55 = model("test_055")
case = {"name": "test_055", "fields": ["title", "body"], "min_word_len": 1,
        "docs": [{"id": d[0], "fields": [d[1], d[2]]} for d in rows_055], "queries": []}
for qi, tree in [(0, ["and", ["kw", "one", 1, ALL, {"start": 1}], ["kw", "two", 2]]),
                 (1, ["and", ["kw", "other", 1, ALL, {"start": 1}], ["kw", "three", 2]]),
                 (2, ["kw", "three", 1, ALL, {"end": 1}]),
                 (3, ["kw", "badger", 1, ALL, {"start": 1}])]:
    case["queries"].append({"text": m55[qi]["query"], "tree": tree, "ranker": "proximity_bm25", "limit": 20, "expect": api_expect(m55[qi])})
out["cases"].append(case)

# ---------------------------------------------------------------------------------------------
# RTN.WeightBoundary, src/gtests/gtests_rtstuff.cpp:244-335: 1 doc, `@title cat` -> rowid 0, weight 1500
# ---------------------------------------------------------------------------------------------
out["cases"].append({
    "name": "gtest_RTN_WeightBoundary", "fields": ["title", "content"], "min_word_len": 1,
    "docs": [{"id": 1, "fields": ["If I were a cat...", "We are the greatest cat"]}],
    "queries": [{"text": "@title cat", "tree": ["kw", "cat", 1, TITLE], "ranker": "proximity_bm25",
                 "expect": {"matches": [[1, 1500]], "total_found": 1, "words": {}}}],
})



# ---------------------------------------------------------------------------------------------
# Which of these trees does the CUDA path still refuse (MGPU_E_UNSUPPORTED, never a guess)? Round 2 runs quorum nodes, and NEAR /
# BEFORE / NOTNEAR over plain keywords in the hit stage. Left to the oracle: NEAR with three and more children (the reference's
# FSMmultinear_c keeps m_uFirstQpos across documents, so the hits of a document depend on the documents before it) and
# NEAR / BEFORE / NOTNEAR whose children are phrases, OR groups or other operators.
# ---------------------------------------------------------------------------------------------
# test_133 "SENTENCE, PARAGRAPH, and ZONE operators", index `test` (html_strip=1, index_sp=1): the eight SENTENCE / PARAGRAPH queries.
# ExtUnit_c (src/searchnode.cpp:4958-5330) over the boundary hits the indexing side writes (BuildZoneHits, src/sphinx.cpp:22232-22270:
# a boundary takes a position of its own; a paragraph mark is also a sentence mark). The index holds 17 documents (301 / 302 are
# commented out in test.xml); the zone documents only matter through N. Oracle only: the CUDA path refuses these operators.
# ---------------------------------------------------------------------------------------------
xml133 = open(os.path.join(REF, "test_133", "test.xml"), encoding="utf-8").read()
xml133 = re.sub(r"<!--.*?-->", "", xml133, flags=re.S)
docs_133 = []
for m in re.finditer(r"<db_insert>\s*(?:<!\[CDATA\[)?\s*insert into test_table values \( (\d+), (\d+), (.*?)\);?\s*(?:\]\]>)?\s*</db_insert>", xml133, re.S):
    doc_id, gid, expr = int(m.group(1)), int(m.group(2)), m.group(3).strip()
    cdata = "<![CDATA[" in m.group(0)
    if doc_id == 5:
        text = "A ram zam zam. " * 171 + "Zam ram!"
    elif doc_id == 400:
        body = re.search(r"CONCAT\('Clock',CHAR\(4\),'(.*)'\)", expr, re.S).group(1)
        text = "Clock\x04" + body.replace("\\'", "'")
    else:
        text = re.match(r"'(.*)'$", expr, re.S).group(1).replace("\\'", "'")
        if not cdata:
            text = text.replace("&lt;", "<").replace("&gt;", ">").replace("&amp;", "&")
    docs_133.append((doc_id, gid, text))
docs_133.sort()
assert [d[0] for d in docs_133] == [1, 2, 3, 4, 5, 6, 100, 101, 200, 201, 202, 211, 300, 310, 311, 400, 500], [d[0] for d in docs_133]
m133 = model("test_133")


def UNIT(kind, *kids):
    return [kind] + list(kids)


q133 = {0: UNIT("sentence", K("one", 1), K("two", 2)),
        1: ["and", UNIT("sentence", K("one", 1), K("two", 2)), K("three", 3)],
        2: UNIT("sentence", K("one", 1), K("two", 2), K("three", 3)),
        3: UNIT("sentence", PH(["one", "two"], 1), K("three", 3)),
        4: UNIT("sentence", K("zam", 1), K("ram", 2)),
        5: UNIT("paragraph", K("fox", 1), K("dog", 2)),
        6: UNIT("sentence", K("sentence", 1), K("paragraph", 2)),
        7: UNIT("paragraph", K("sentence", 1), K("paragraph", 2))}
case = {"name": "test_133", "fields": ["title"], "attrs": ["gid"], "min_word_len": 1, "html_strip": 1, "index_sp": 1,
        "docs": [{"id": d[0], "fields": [d[2]], "attrs": [d[1]]} for d in docs_133], "queries": []}
for qi, tree in q133.items():
    case["queries"].append({"text": m133[qi]["query"], "tree": tree, "ranker": "proximity_bm25", "gpu_unsupported": True, "expect": api_expect(m133[qi])})
out["cases"].append(case)


# ---------------------------------------------------------------------------------------------
# api/libsphinxclient: the C client's own smoke test. smoke_ref.txt is what its test program printed against a real searchd over the
# index of smoke_data.csv (fields title / content, attributes idd / group_id; the multi-value attributes are left out): extended2,
# SPH_RANK_PROXIMITY_BM25, field_weights=(title=100, content=1) (test.c:58-75). The three plain queries and the group_id filter.
# tests/test_api_wire.py replays the same calls through the reference's client and the wire responder on the GPU.
# ---------------------------------------------------------------------------------------------
LSC = os.path.join(os.path.dirname(REF), "api", "libsphinxclient")
smoke_docs = []
for line in open(os.path.join(LSC, "smoke_data.csv"), encoding="utf-8"):
    f = line.rstrip("\n").split("+")
    if len(f) >= 5:
        smoke_docs.append({"id": int(f[0]), "fields": [f[1], f[2]], "attrs": [int(f[3]), int(f[4])]})
ref_txt = open(os.path.join(LSC, "smoke_ref.txt"), encoding="utf-8").read()
blocks = re.findall(r"Query '([^']*)' retrieved (\d+) of (\d+) matches\.\nQuery stats:\n((?:\t.*\n)*)\nMatches:\n((?:\d+\. .*\n)*)", ref_txt)


def smoke_expect(b):
    words = {m.group(1): [int(m.group(3)), int(m.group(2))] for m in re.finditer(r"'([^']*)' found (\d+) times in (\d+) documents", b[3])}
    matches = [[int(m.group(1)), int(m.group(2))] for m in re.finditer(r"doc_id=(\d+), weight=(\d+)", b[4])]
    return {"matches": matches, "total_found": int(b[2]), "words": words}


assert [b[0] for b in blocks[:3]] == ["is", "is test", "test number"]
case = {"name": "libsphinxclient_smoke", "fields": ["title", "content"], "attrs": ["idd", "group_id"], "min_word_len": 1, "docs": smoke_docs, "queries": []}
for b, tree in zip(blocks[:3], [K("is", 1), ["and", K("is", 1), K("test", 2)], ["and", K("test", 1), K("number", 2)]]):
    case["queries"].append({"text": b[0], "tree": tree, "ranker": "proximity_bm25", "field_weights": [100, 1], "expect": smoke_expect(b)})
after_filter = ref_txt[ref_txt.index("* test_filter"):]               # test_filter (test.c:396-416): group_id = 1 first
flt = re.findall(r"Query '([^']*)' retrieved (\d+) of (\d+) matches\.\nQuery stats:\n((?:\t.*\n)*)\nMatches:\n((?:\d+\. .*\n)*)", after_filter)[0]
case["queries"].append({"text": "is", "tree": K("is", 1), "ranker": "proximity_bm25", "field_weights": [100, 1], "filters": [["group_id", 1, 1]],
                        "expect": smoke_expect(flt)})
out["cases"].append(case)


# ---------------------------------------------------------------------------------------------
# test_041 "phrase shift": a separate star inside a phrase stands for any one keyword ("that * box"): XQParser_t::GetToken counts the
# [ * ] between the phrase's tokens and PhraseShiftQpos moves the in-query positions behind them (src/sphinxquery.cpp:1318-1348, 1701-1738);
# the first star right behind the quote does not match the [ * ] pattern, so `"* * * box always"` shifts by two. Indexes `phrase_shift`
# (plain) and `phrase_shift_min_wlen` (min_word_len=2: the document's `a` is overshort and still takes a position). SphinxQL rows carry
# no weights (ids_only). Plain phrases over keywords: the CUDA path runs them.
# ---------------------------------------------------------------------------------------------
m41 = model("test_041")
docs_041 = [(1, "that orange box might be and not yellow"), (2, "however that is a green box always apears here as usual"), (3, "that orange box might be not yellow")]


def PS(*wp):
    return ["phrase", [[w, p] for w, p in wp]]


q041 = {"phrase_shift": {107: PS(("that", 1), ("box", 2)), 108: PS(("that", 1), ("box", 3)), 109: PS(("that", 1), ("box", 5)),
                         110: PS(("that", 1), ("box", 3), ("might", 4), ("not", 7), ("yellow", 8)), 120: PS(("box", 3), ("always", 4))},
        "phrase_shift_min_wlen": {115: PS(("that", 1), ("box", 2)), 116: PS(("that", 1), ("box", 3)), 117: PS(("that", 1), ("box", 5)),
                                  118: PS(("that", 1), ("box", 3), ("might", 4), ("not", 7), ("yellow", 8)), 122: PS(("box", 3), ("always", 4))}}
for index, trees in q041.items():
    case = {"name": "test_041_" + index, "fields": ["title"], "attrs": ["idd"], "min_word_len": 2 if index.endswith("min_wlen") else 1,
            "docs": [{"id": d[0], "fields": [d[1]], "attrs": [11]} for d in docs_041], "queries": []}
    for qi, tree in trees.items():
        r = m41[qi]
        assert (" FROM %s WHERE" % index) in r["sphinxql"], r["sphinxql"]
        case["queries"].append({"text": r["sphinxql"].strip(), "tree": tree, "ranker": "proximity_bm25", "sort": "id_asc", "ids_only": True,
                                "expect": {"matches": [[int(row["id"]), 0] for row in (r.get("rows") or {}).values()], "total_found": int(r["total_rows"]), "words": {}}})
    out["cases"].append(case)


# ---------------------------------------------------------------------------------------------
def gpu_refuses(t):
    kind = t[0]
    if kind in ("kw", "phrase", "prox", "quorum"):
        return False
    if kind in ("sentence", "paragraph"):
        return True
    kids = t[2:] if kind in ("near", "notnear") else t[1:]
    if kind in ("near", "before", "notnear"):
        if any(k[0] != "kw" for k in kids) or (kind == "near" and len(kids) != 2):
            return True
        return False
    return any(gpu_refuses(k) for k in kids)


for case in out["cases"]:
    for q in case["queries"]:
        q.pop("gpu_unsupported", None)
        if gpu_refuses(q["tree"]):
            q["gpu_unsupported"] = True

with open(os.path.join(HERE, "golden_vectors.json"), "w", encoding="utf-8") as f:
    json.dump(out, f, ensure_ascii=False, indent=1)
print("wrote", sum(len(c["queries"]) for c in out["cases"]), "golden queries in", len(out["cases"]), "cases")

# ---------------------------------------------------------------------------------------------
# SHOW PLAN: the reference's OWN parser output (`transformed_tree`, sphExplainQuery) for every query of its test suite whose index has
# no morphology / wordforms / blended characters in play -> tests/golden/show_plan.json; tests/test_query_parser.py requires
# mgpu_parsed_explain to print the same text. test_207's lemmatised keywords are kept for their MODIFIERS (field_end / boost): the
# keyword itself is masked there.
# ---------------------------------------------------------------------------------------------
plans = []


def harvest(test, keep, **settings):
    m = php_unserialize(open(os.path.join(REF, test, "model.bin"), "rb").read())[0]
    prev = None
    for k, q in m.items():
        if isinstance(q, dict) and str(q.get("sphinxql", "")).lower().strip().startswith("show plan") and k in keep:
            tree = [row["Value"] for row in (q.get("rows") or {}).values() if row.get("Variable") == "transformed_tree"][0]
            mm = re.search(r"match\s*\(\s*'(.*)'\s*\)", prev["sphinxql"], re.I | re.S)
            text = mm.group(1).replace("\\\\", "\\")        # the SphinxQL string literal's own escaping
            plans.append(dict({"test": test, "query": text, "plan": tree}, **settings))
        prev = q


harvest("test_022", {24, 27, 30, 33, 36, 39}, fields=["title", "content"])      # multi1 / multi2: their multiforms do not fire on these
harvest("test_022", {68}, fields=["title", "content"])                          # '"foo bar baz"/24 mois': the threshold takes a position
harvest("test_113", {8}, fields=["title"])
harvest("test_115", {43, 46}, fields=["title"])
harvest("test_192", {57}, fields=["title", "content"])
harvest("test_222", {217}, fields=["title"])                                    # escaped modifiers are not modifiers
harvest("test_207", {211, 213, 215, 217, 219}, fields=["title"], mask_words=True)
with open(os.path.join(HERE, "show_plan.json"), "w", encoding="utf-8") as f:
    json.dump({"source": "SHOW PLAN rows of ravelry/manticoresearch test/test_022,113,115,192,207,222 model.bin", "plans": plans}, f, ensure_ascii=False, indent=1)
print("wrote %d SHOW PLAN vectors" % len(plans))

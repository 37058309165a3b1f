"""Query front-end: mgpu_parse_query restates XQParser_t (src/sphinxquery.cpp:1011-1830, grammar
src/sphinxquery.y) + XQParseHelper_c::FixupTree (:343-387) + PrepareQueryEmulation (src/searchd.cpp:2141-2190).

Pinned two ways on the reference's own test queries (the `text` of every golden vector is the string test.xml sends to searchd):
the parsed tree equals the hand-transcribed tree of tests/golden/make_golden.py, and the parsed tree run through the oracle gives
the reference's model.bin results. Host only: no GPU involved."""
import re

import pytest

import helpers
import manticoresearch_b200.mgpu as M

# the legacy match modes of the golden cases (test.xml <query mode="...">); everything else is extended2
LEGACY = {("test_017", 0): M.MATCH_PHRASE, ("test_017", 1): M.MATCH_PHRASE, ("test_017", 2): M.MATCH_PHRASE, ("test_017", 3): M.MATCH_PHRASE,
          ("test_015", 0): M.MATCH_PHRASE, ("test_015", 1): M.MATCH_PHRASE, ("test_015", 2): M.MATCH_PHRASE, ("test_015", 3): M.MATCH_PHRASE,
          ("test_016", 0): M.MATCH_ANY, ("test_016", 1): M.MATCH_ANY}

OPS = {M.OP_AND: "and", M.OP_OR: "or", M.OP_ANDNOT: "andnot", M.OP_MAYBE: "maybe", M.OP_BEFORE: "before", M.OP_SENTENCE: "sentence", M.OP_PARAGRAPH: "paragraph"}


def match_text(text):
    """the MATCH() argument of a SphinxQL golden query, or the text itself"""
    m = re.search(r"match\s*\(\s*'(.*?)'\s*\)", text, re.I)
    return m.group(1) if m else text


def compact(n):
    """Node -> the compact form of make_golden.py; one-child AND / NOT wrappers (FixupNots leaves them, ExtNode_i::Create folds them) dropped"""
    if n.words:
        if len(n.words) == 1 and n.op == M.OP_AND:
            k = n.words[0]
            out = ["kw", k.word, k.atom_pos]
            mods = {}
            if k.field_start:
                mods["start"] = 1
            if k.field_end:
                mods["end"] = 1
            if n.field_max_pos:
                mods["max_pos"] = n.field_max_pos
            if n.field_mask != 0xFFFFFFFF or mods:
                out.append(n.field_mask)
            if mods:
                out.append(mods)
            return out
        ws = [[k.word, k.atom_pos] for k in n.words]
        if n.op == M.OP_PHRASE:
            return ["phrase", ws] + ([n.field_mask] if n.field_mask != 0xFFFFFFFF else [])
        if n.op == M.OP_PROXIMITY:
            return ["prox", n.oparg, ws]
        if n.op == M.OP_QUORUM:
            return ["quorum", n.oparg, ws]
        raise AssertionError(n.op)
    kids = [compact(c) for c in n.children]
    if n.op in (M.OP_AND, M.OP_NOT) and len(kids) == 1:
        return kids[0]
    if n.op == M.OP_NEAR:
        return ["near", n.oparg] + kids
    if n.op == M.OP_NOTNEAR:
        return ["notnear", n.oparg] + kids
    return [OPS[n.op]] + kids


def parse_case_query(case, qi):
    q = case["queries"][qi]
    return M.parse_query(match_text(q["text"]), case["fields"], min_word_len=case.get("min_word_len", 1), stopwords=case.get("stopwords", ()),
                         match_mode=LEGACY.get((case["name"], qi), M.MATCH_EXTENDED))


def _cases():
    return [(c["name"], i) for c in helpers.load_golden() for i in range(len(c["queries"]))]


@pytest.mark.parametrize("case_name,qi", _cases())
def test_parsed_tree_equals_hand_tree_and_reproduces_golden(golden_cases, golden_indexes, case_name, qi):
    case = next(c for c in golden_cases if c["name"] == case_name)
    q = case["queries"][qi]
    root, ranker, _ = parse_case_query(case, qi)
    def transformed(t):     # TransformQuorum (src/sphinx.cpp:14643-14669): "a b"/1 reaches the index as an OR; the planner takes either form
        if t[0] == "quorum" and t[1] == 1:
            return ["or"] + [["kw", w, p] for w, p in t[2]]
        return [transformed(c) if isinstance(c, list) and c and isinstance(c[0], str) and c[0] in
                ("kw", "and", "or", "andnot", "maybe", "phrase", "prox", "quorum", "near", "before", "notnear", "sentence", "paragraph") else c for c in t]

    hand = transformed(q["tree"])
    assert compact(root) == hand, (q["text"], compact(root), hand)
    if ranker is not None:
        assert ranker == helpers.RANKERS[q["ranker"]], q["text"]
    # the parsed tree, exactly as the parser shaped it, through the oracle == the reference's result
    query = helpers.golden_query(case, q)
    query.root = root
    idx = helpers.OracleIndex(golden_indexes[case_name])
    try:
        r = idx.search([query]).get(0)
        assert r["status"] == 0, q["text"]
        got = list(zip(r["docid"], r["weight"]))
        if q.get("ids_only"):
            got = [(d, 0) for d, _ in got]
        if q.get("limit"):
            got = got[:q["limit"]]
        assert got == [tuple(m) for m in q["expect"]["matches"]], q["text"]
        assert r["total_found"] == q["expect"]["total_found"], q["text"]
    finally:
        idx.close()


FIELDS = ["title", "body"]


def P(text, **kw):
    return compact(M.parse_query(text, FIELDS, **kw)[0])


def test_operators_and_fixups():
    assert P("a b c") == ["and", ["kw", "a", 1], ["kw", "b", 2], ["kw", "c", 3]]
    assert P("a | b | c") == ["or", ["kw", "a", 1], ["kw", "b", 2], ["kw", "c", 3]]
    assert P("a MAYBE b") == ["maybe", ["kw", "a", 1], ["kw", "b", 2]]
    assert P("a maybe b") == ["and", ["kw", "a", 1], ["kw", "maybe", 2], ["kw", "b", 3]]      # operators are case sensitive
    assert P("a -b") == ["andnot", ["kw", "a", 1], ["kw", "b", 2]]
    assert P("a !b !c") == ["andnot", ["kw", "a", 1], ["or", ["kw", "b", 2], ["kw", "c", 3]]]
    assert P("(a b) | c") == ["or", ["and", ["kw", "a", 1], ["kw", "b", 2]], ["kw", "c", 3]]
    assert P("a (b | c) d") == ["and", ["kw", "a", 1], ["or", ["kw", "b", 2], ["kw", "c", 3]], ["kw", "d", 4]]
    assert P("Hello WORLD_1") == ["and", ["kw", "hello", 1], ["kw", "world_1", 2]]
    assert P("well-known fact") == ["and", ["kw", "well", 1], ["kw", "known", 2], ["kw", "fact", 3]]   # a dash inside a word separates
    assert P('"a b c"~5 d') == ["and", ["prox", 5, [["a", 1], ["b", 2], ["c", 3]]], ["kw", "d", 5]]     # FixupAtomPos: last word + 1, then the keyword's own step
    assert [k.excluded for k in M.parse_query("a -(b c)", FIELDS)[0].all_keywords()] == [False, True, True]
    assert P('"a b c"/2 d') == ["and", ["quorum", 2, [["a", 1], ["b", 2], ["c", 3]]], ["kw", "d", 5]]   # the threshold token takes one
    assert P('"single"') == ["kw", "single", 1]                                                          # FixupDegenerates
    assert P('"one"/3') == ["kw", "one", 1]
    assert P("a << b << c") == ["before", ["kw", "a", 1], ["kw", "b", 2], ["kw", "c", 3]]
    assert P("a NEAR/3 b NEAR/3 c") == ["near", 3, ["kw", "a", 1], ["kw", "b", 2], ["kw", "c", 3]]
    assert P("a NEAR/3 b NEAR/4 c") == ["near", 4, ["near", 3, ["kw", "a", 1], ["kw", "b", 2]], ["kw", "c", 3]]
    assert P("a NOTNEAR/2 b") == ["notnear", 2, ["kw", "a", 1], ["kw", "b", 2]]
    assert P("a NEAR b") == ["and", ["kw", "a", 1], ["kw", "near", 2], ["kw", "b", 3]]                  # NEAR without /N is a keyword
    assert P("") == ["and"] and P("  .. ,, ") == ["and"]                                                 # nothing to search for: one empty node


def test_field_limits_and_modifiers():
    assert P("@title a b @body c") == ["and", ["kw", "a", 1, 1], ["kw", "b", 2, 1], ["kw", "c", 3, 2]]
    assert P("@(title,body) a") == ["kw", "a", 1, 3]
    assert P("@!body a") == ["kw", "a", 1, 0xFFFFFFFD]
    assert P("@* a") == ["kw", "a", 1]
    assert P("@title[5] a") == ["kw", "a", 1, 1, {"max_pos": 5}]
    assert P("(@title a) b") == ["and", ["kw", "a", 1, 1], ["kw", "b", 2]]         # a limit inside parentheses ends with them
    assert P("@title (a b) c") == ["and", ["kw", "a", 1, 1], ["kw", "b", 2, 1], ["kw", "c", 3, 1]]       # AddOp appends to the AND on its left
    assert P('@body "a b"') == ["phrase", [["a", 1], ["b", 2]], 2]
    assert P("^a b$") == ["and", ["kw", "a", 1, 0xFFFFFFFF, {"start": 1}], ["kw", "b", 2, 0xFFFFFFFF, {"end": 1}]]
    root = M.parse_query("a^1.5 b", FIELDS)[0]
    assert [k.boost for k in root.all_keywords()] == [1.5, 1.0]
    with pytest.raises(M.MgpuError) as e:
        P("@nosuchfield a")
    assert e.value.code == M.MGPU_E_BAD_QUERY and "nosuchfield" in str(e.value)


def test_positions_stopwords_overshort_escapes():
    assert P("walking in my shoes", min_word_len=3) == ["and", ["kw", "walking", 1], ["kw", "shoes", 4]]
    assert P("the lord of the rings", stopwords=("the", "of")) == ["and", ["kw", "lord", 2], ["kw", "rings", 5]]
    assert P("the lord of the rings", stopwords=("the", "of"), stopword_step=0) == ["and", ["kw", "lord", 1], ["kw", "rings", 2]]
    assert P('"the lord of the rings"', stopwords=("the", "of")) == ["phrase", [["lord", 2], ["rings", 5]]]
    assert P(r"aaa\!bbb \-ccc") == ["and", ["kw", "aaa", 1], ["kw", "bbb", 2], ["kw", "ccc", 3]]        # escaped specials separate
    assert P("walking -in shoes", min_word_len=3) == ["and", ["kw", "walking", 1], ["kw", "shoes", 3]]   # 'la !word': the NOT goes with its overshort operand
    assert P("привет МИР") == ["and", ["kw", "привет", 1], ["kw", "мир", 2]]
    assert P("我吐") == ["and", ["kw", "我", 1], ["kw", "吐", 2]]


def test_legacy_match_modes():
    root, ranker, _ = M.parse_query('hello "world" -x', FIELDS, match_mode=M.MATCH_ALL)
    assert compact(root) == ["and", ["kw", "hello", 1], ["kw", "world", 2], ["kw", "x", 3]] and ranker == M.RANK_PROXIMITY
    root, ranker, _ = M.parse_query("hello world", FIELDS, match_mode=M.MATCH_ANY)
    assert compact(root) == ["or", ["kw", "hello", 1], ["kw", "world", 2]] and ranker == M.RANK_MATCHANY
    root, ranker, _ = M.parse_query("hello | world", FIELDS, match_mode=M.MATCH_PHRASE)
    assert compact(root) == ["phrase", [["hello", 1], ["world", 2]]] and ranker == M.RANK_PROXIMITY
    root, ranker, _ = M.parse_query("hello | world", FIELDS, match_mode=M.MATCH_BOOLEAN)
    assert compact(root) == ["or", ["kw", "hello", 1], ["kw", "world", 2]] and ranker == M.RANK_NONE


@pytest.mark.parametrize("text,needle", [
    ("-a", "single NOT"), ("-a -b", "NOT operators only"), ("a | (-b)", "NOT is not allowed within OR"), ("a << -b", "before operand"),
    ('"a b"/0', "quorum threshold too low"), ('"a b"~0', "proximity threshold too low"), ("a (b", "syntax error"), ("a | | b", "syntax error"),
    ('"a b', "syntax error"), ('"a b"/1.5', "quorum threshold out of bounds"),
])
def test_parse_errors(text, needle):
    with pytest.raises(M.MgpuError) as e:
        M.parse_query(text, FIELDS)
    assert e.value.code == M.MGPU_E_BAD_QUERY and needle in str(e.value), str(e.value)


def test_sentence_and_paragraph_parse(tmp_path):
    """sentence / paragraph rules of the grammar (src/sphinxquery.y:125-138): left-associative chains over keywords and quoted phrases.
    The oracle evaluates them (ExtUnit_c, pinned by test/test_133 in the golden set); the CUDA path answers MGPU_E_UNSUPPORTED
    (tests/test_gpu_parity.py)"""
    E = lambda t: " ".join(M.explain_query(t, FIELDS).split())
    assert E("a SENTENCE b") == "SENTENCE( AND(KEYWORD(a, querypos=1)), AND(KEYWORD(b, querypos=2)))"
    assert E('a SENTENCE "b c" SENTENCE d') == ("SENTENCE( AND(KEYWORD(a, querypos=1)), PHRASE(KEYWORD(b, querypos=2), KEYWORD(c, querypos=3)), "
                                                 "AND(KEYWORD(d, querypos=4)))")
    assert E("x | a PARAGRAPH b") == "OR( AND(KEYWORD(x, querypos=1)), PARAGRAPH( AND(KEYWORD(a, querypos=2)), AND(KEYWORD(b, querypos=3))))"
    assert E("a sentence b") == "AND( AND(KEYWORD(a, querypos=1)), AND(KEYWORD(sentence, querypos=2)), AND(KEYWORD(b, querypos=3)))"
    for bad in ("a SENTENCE", "a SENTENCE (b c)", "a SENTENCE b PARAGRAPH c", "SENTENCE a"):
        with pytest.raises(M.MgpuError) as e:
            M.parse_query(bad, FIELDS)
        assert "syntax error" in str(e.value), bad
    prefix = str(tmp_path / "sp")
    M.build_index(prefix, FIELDS, [{"id": 1, "fields": [[("a", 1)], [("b", 1)]], "attrs": []}])
    idx = helpers.OracleIndex(prefix)
    try:
        r = idx.search([M.Query(M.parse_query("a SENTENCE b", FIELDS)[0], max_matches=10)]).get(0)
        assert r["status"] == 0 and list(r["docid"]) == [1]       # no boundary hits in the index: the operator degenerates into AND
    finally:
        idx.close()


@pytest.mark.parametrize("text", ["ZONE:h1 a", "ZONESPAN:(h1,h2) a"])
def test_unsupported_syntax_is_refused_not_guessed(text):
    with pytest.raises(M.MgpuError) as e:
        M.parse_query(text, FIELDS)
    assert e.value.code == M.MGPU_E_UNSUPPORTED


def test_fuzz_never_crashes_and_trees_are_well_formed(tmp_path):
    """random token soup: the parser answers (a tree or a message) and every tree it returns is one the engine's planner takes:
    the oracle runs it (or refuses an operator shape), never reports a malformed tree"""
    import random
    rng = random.Random(20260219)
    vocab = ["a", "b", "c", "dd", "e5", "the", "x", "NEAR/2", "NOTNEAR/3", "MAYBE", "<<", "|", "-", "!", "(", ")", '"', '"', "~2", "/2", "/0.5", "@title", "@body",
             "@(title,body)", "@!title", "@title[3]", "^", "$", "\\", "=", "*", "привет", "我", "  ", "^1.5", "@*", "-a", "b$", "^c"]
    docs = [{"id": 1 + i, "fields": [[(w, k + 1) for k, w in enumerate(rng.choices(["a", "b", "c", "dd", "e5", "x"], k=3))],
                                      [(w, k + 1) for k, w in enumerate(rng.choices(["a", "b", "c", "dd", "e5", "x"], k=6))]], "attrs": []} for i in range(50)]
    prefix = str(tmp_path / "fz")
    M.build_index(prefix, FIELDS, docs)
    idx = helpers.OracleIndex(prefix)
    parsed = failed = 0
    try:
        for _ in range(1500):
            text = " ".join(rng.choices(vocab, k=rng.randint(1, 9))) if rng.random() < 0.8 else "".join(rng.choices(vocab, k=rng.randint(1, 9)))
            try:
                root, _, _ = M.parse_query(text, FIELDS, stopwords=("the",), min_word_len=rng.choice([1, 2]))
            except M.MgpuError as e:
                assert e.code in (M.MGPU_E_BAD_QUERY, M.MGPU_E_UNSUPPORTED) and str(e), text
                failed += 1
                continue
            parsed += 1

            def check(n):
                assert not (n.words and n.children), text
                if n.words and n.op in (M.OP_PHRASE, M.OP_PROXIMITY, M.OP_QUORUM):
                    pos = [k.atom_pos for k in n.words]
                    assert pos == sorted(pos) and len(set(pos)) == len(pos) and len(pos) >= 2, (text, pos)
                for k in n.words:
                    assert k.word and k.atom_pos >= 1, text
                if n.op == M.OP_ANDNOT:
                    assert len(n.children) == 2, text
                if n.op == M.OP_NOT:
                    assert len(n.children) == 1, text
                for c in n.children:
                    check(c)
            check(root)
            r = idx.search([M.Query(root, max_matches=10)]).get(0)
            assert r["status"] in (0, M.MGPU_E_UNSUPPORTED), (text, r["status"])
        assert parsed > 200 and failed > 200
    finally:
        idx.close()


def test_pathological_nesting_is_refused_not_crashed():
    """a daemon must survive any query text: deep parentheses / deep operator chains answer "query too complex" (the reference
    measures its stack for the same purpose), long flat queries parse"""
    with pytest.raises(M.MgpuError) as e:
        M.parse_query("(" * 100000 + "a" + ")" * 100000, FIELDS)
    assert "too complex" in str(e.value)
    with pytest.raises(M.MgpuError) as e:
        M.parse_query(" ".join("w%d NEAR/%d" % (i, i + 1) for i in range(3000)) + " z", FIELDS)
    assert "too complex" in str(e.value)
    root, _, _ = M.parse_query(" ".join("w%d" % i for i in range(20000)), FIELDS)
    assert len(root.children) == 20000
    root, _, _ = M.parse_query(" | ".join("w%d" % i for i in range(20000)), FIELDS)
    assert len(root.children) == 20000
    root, _, _ = M.parse_query("(" * 300 + "a b" + ")" * 300, FIELDS)
    assert compact(root) == ["and", ["kw", "a", 1], ["kw", "b", 2]]


def _show_plans():
    import json
    import os
    return json.load(open(os.path.join(helpers.ROOT, "tests", "golden", "show_plan.json"), encoding="utf-8"))["plans"]


@pytest.mark.parametrize("i", range(17))
def test_explain_equals_the_references_show_plan(i):
    """mgpu_parsed_explain vs the `transformed_tree` rows the reference's searchd printed for its own test queries (SHOW PLAN,
    sphExplainQuery): the reference parser's output itself, not a hand transcription"""
    p = _show_plans()[i]
    got = " ".join(M.explain_query(p["query"], p["fields"]).split())
    want = " ".join(p["plan"].split())
    if p.get("mask_words"):     # a lemmatiser rewrote the keyword there; its modifiers are what this vector pins
        mask = lambda s: re.sub(r"KEYWORD\([^,]+,", "KEYWORD(*,", s).replace(", morphed", "")
        got, want = mask(got), mask(want)
    assert got == want, p["query"]


def test_explain_format():
    assert M.explain_query("@title hello -world", FIELDS) == ("ANDNOT(\n  AND(\n    AND(fields=(title), KEYWORD(hello, querypos=1))), \n  NOT(\n"
                                                               "    AND(fields=(title), KEYWORD(world, querypos=2, excluded))))")
    assert " ".join(M.explain_query('"a b"~3 | ^c$ | @body[5] d', FIELDS).split()) == \
        "OR( PROXIMITY(distance=3, KEYWORD(a, querypos=1), KEYWORD(b, querypos=2)), AND(KEYWORD(c, querypos=4, field_start, field_end)), AND(fields=(body), max_field_pos=5, KEYWORD(d, querypos=5)))"


def test_relaxed_mode():
    """@@relaxed (src/sphinxquery.cpp:1752-1760, AddField :49-74, DeleteNodesWOFields :217-255): unknown fields warn instead of failing and
    the keywords under them leave the tree"""
    root, _, warning = M.parse_query("@@relaxed @nosuch hello @title world", FIELDS)
    assert compact(root) == ["kw", "world", 2, 1] and "no field 'nosuch' found in schema" in warning
    root, _, warning = M.parse_query("@@relaxed ((@title hello) | (@missed world)) @body other terms", FIELDS)
    assert compact(root) == ["and", ["kw", "hello", 1, 1], ["kw", "other", 3, 2], ["kw", "terms", 4, 2]]
    root, _, warning = M.parse_query("@@relaxed @(title,nosuch) hello", FIELDS)
    assert compact(root) == ["kw", "hello", 1, 1]
    with pytest.raises(M.MgpuError):
        M.parse_query("@nosuch hello", FIELDS)
    with pytest.raises(M.MgpuError):
        M.parse_query("@@relaxedx @nosuch hello", FIELDS)      # not the option: a field limit on an unknown field `relaxedx`... and a syntax error before it

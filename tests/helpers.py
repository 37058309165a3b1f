"""Shared test helpers: oracle FFI, golden-corpus tokenizer, tree builders, result comparison.

The oracle (oracle/liboracle.so) is the CHECKER: it is loaded only here, in the tests.
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import manticoresearch_b200.mgpu as M  # noqa: E402

ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "golden_vectors.json")


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)
    return ORACLE_LIB


_olib = None


def oracle_lib():
    global _olib
    if _olib is None:
        build_oracle()
        lib = C.CDLL(ORACLE_LIB)
        lib.oracle_open.restype = C.c_void_p
        lib.oracle_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        lib.oracle_close.argtypes = [C.c_void_p]
        lib.oracle_total_docs.restype = C.c_int64
        lib.oracle_total_docs.argtypes = [C.c_void_p]
        lib.oracle_search_batch.restype = C.c_int
        lib.oracle_search_batch.argtypes = [C.c_void_p, C.POINTER(M.c_query), C.c_int, C.POINTER(M.c_result)]
        lib.oracle_decode_doclist.restype = C.c_int
        lib.oracle_decode_doclist.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                              C.POINTER(C.c_uint64), C.c_int64, C.POINTER(C.c_int64)]
        lib.oracle_decode_hitlist.restype = C.c_int
        lib.oracle_decode_hitlist.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint32), C.c_int]
        lib.oracle_word_stats.restype = C.c_int
        lib.oracle_word_stats.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _olib = lib
    return _olib


class OracleIndex:
    def __init__(self, prefix):
        self.lib = oracle_lib()
        err = C.create_string_buffer(512)
        self.h = self.lib.oracle_open(prefix.encode(), err, 512)
        if not self.h:
            raise RuntimeError("oracle_open failed: " + err.value.decode())

    def close(self):
        if self.h:
            self.lib.oracle_close(self.h)
            self.h = None

    @property
    def total_docs(self):
        return self.lib.oracle_total_docs(self.h)

    def search(self, queries):
        arr = M.pack_queries(queries)
        rs = M.ResultSet(queries)
        self.lib.oracle_search_batch(self.h, arr, len(queries), rs.results)
        return rs

    def word_stats(self, word):
        d, h = C.c_int64(), C.c_int64()
        if not self.lib.oracle_word_stats(self.h, word.encode("utf-8"), C.byref(d), C.byref(h)):
            return None
        return d.value, h.value

    def decode_doclist(self, word):
        import numpy as np
        st = self.word_stats(word)
        if st is None:
            return None
        n = st[0]
        rowid = np.zeros(n, dtype=np.uint32)
        hits = np.zeros(n, dtype=np.uint32)
        fields = np.zeros(n, dtype=np.uint32)
        pos = np.zeros(n, dtype=np.uint64)
        nout = C.c_int64()
        self.lib.oracle_decode_doclist(self.h, word.encode("utf-8"), rowid.ctypes.data_as(C.POINTER(C.c_uint32)),
                                       hits.ctypes.data_as(C.POINTER(C.c_uint32)), fields.ctypes.data_as(C.POINTER(C.c_uint32)),
                                       pos.ctypes.data_as(C.POINTER(C.c_uint64)), n, C.byref(nout))
        assert nout.value == n, (nout.value, n)
        return rowid, hits, fields, pos

    def decode_hitlist(self, word, hitlist_pos, cap=100000):
        buf = (C.c_uint32 * cap)()
        n = self.lib.oracle_decode_hitlist(self.h, word.encode("utf-8"), C.c_uint64(int(hitlist_pos)), buf, cap)
        return list(buf[:n])


# ---------------------------------------------------------------------------------------------
# golden corpora: whitespace/punctuation tokenizer equivalent to the reference's default charset_table
# for the ASCII / Cyrillic / CJK-ngram texts the golden tests use
# ---------------------------------------------------------------------------------------------

def _is_cjk(ch):
    o = ord(ch)
    return 0x2E80 <= o <= 0x9FFF or 0xAC00 <= o <= 0xD7AF or 0xF900 <= o <= 0xFAFF


def _is_word_char(ch):
    o = ord(ch)
    return ch.isascii() and (ch.isalnum() or ch == "_") or 0x410 <= o <= 0x44F


MAGIC_SENTENCE, MAGIC_PARAGRAPH = "\x03sentence", "\x03paragraph"     # MAGIC_WORD_SENTENCE / MAGIC_WORD_PARAGRAPH, src/sphinx.cpp:163-164
_PARA_TAGS = ("p", "div", "h1", "h2", "h3", "h4", "h5", "h6", "li", "ul", "ol", "tr", "td", "th", "table", "br", "hr", "blockquote", "pre", "form", "address", "dl", "dd", "dt")


def html_strip(text):
    """the part of the reference's HTML stripper the golden corpora need: block-level tags become a paragraph mark (\\x03 =
    MAGIC_CODE_PARAGRAPH, src/sphinxint.h:50), every other tag a separator"""
    import re

    def sub(m):
        name = re.match(r"</?\s*([A-Za-z0-9_]+)", m.group(0))
        return "\x03" if name and name.group(1).lower() in _PARA_TAGS else " "
    return re.sub(r"<[^<>]*>?", sub, text)


def _sentence_boundary(text, i, cur):
    """CSphTokenizerBase::CodepointArbitrationI (src/sphinx.cpp:4577-4652) for the character at i, with `cur` accumulated so far"""
    ch = text[i]
    if ch in "?!":
        return True
    if ch != ".":
        return False
    nxt = text[i + 1] if i + 1 < len(text) else "\0"
    nxt2 = text[i + 2] if i + 2 < len(text) else "\0"
    nxt3 = text[i + 3] if i + 3 < len(text) else "\0"
    is_alpha = lambda c: c.isascii() and (c.isalnum() or c in "-_")
    inword = is_alpha(nxt) or ord(nxt) >= 0x80 or nxt == ","
    inphrase = nxt in " \t\r\n" and ("a" <= nxt2 <= "z" or (nxt2 == "(" and "a" <= nxt3 <= "z"))
    cap = lambda c: "A" <= c <= "Z"
    middle = False
    if len(cur) == 1:
        middle = cap(text[i - 1])
    elif len(cur) == 2 and cap(text[i - 2]):
        middle = not cap(text[i - 1]) or text[i - 2:i] in ("MR", "MS", "DR")
    elif len(cur) == 3:
        middle = cur[0] in "md" and cur[1] == "r" and cur[2] == "s"
    return not (inword or inphrase or middle)


def tokenize(text, min_word_len=1, stopwords=(), phrase_boundary="", phrase_boundary_step=0, index_sp=False):
    """-> [(keyword, pos)], pos 1-based; overshort tokens and stop words consume a position (overshort_step=1, stopword_step=1);
    a phrase_boundary character followed by a separator (or the end) advances the position by phrase_boundary_step;
    index_sp: sentence / paragraph boundaries become hits of the magic keywords at a position of their own
    (CSphSource_Document::BuildZoneHits, src/sphinx.cpp:22232-22270)"""
    out, pos, cur = [], 0, ""

    def flush():
        nonlocal cur, pos
        if cur:
            pos += 1
            if len(cur) >= min_word_len and cur not in stopwords:
                out.append((cur, pos))
            cur = ""

    for i, ch in enumerate(text):
        if _is_cjk(ch):
            flush()
            pos += 1
            out.append((ch, pos))
        elif _is_word_char(ch):
            cur += ch.lower()
        elif index_sp and ch == "\x03":
            flush()
            pos += 1
            out.append((MAGIC_SENTENCE, pos))
            out.append((MAGIC_PARAGRAPH, pos))
        elif index_sp and ch in ".?!" and _sentence_boundary(text, i, cur):
            flush()
            pos += 1
            out.append((MAGIC_SENTENCE, pos))
        else:
            flush()
            if ch in phrase_boundary and (i + 1 == len(text) or not (_is_word_char(text[i + 1]) or _is_cjk(text[i + 1]))):
                pos += phrase_boundary_step
    flush()
    return out


def tree_to_node(t):
    kind = t[0]
    if kind == "kw":
        mods = t[4] if len(t) > 4 else {}
        n = M.kw(t[1], t[2], field_start=bool(mods.get("start")), field_end=bool(mods.get("end")))
        if len(t) > 3:
            n.field_mask = t[3]
        if mods.get("max_pos"):
            n.field_max_pos = mods["max_pos"]
        return n
    if kind in ("and", "or", "andnot", "maybe"):
        op = {"and": M.OP_AND, "or": M.OP_OR, "andnot": M.OP_ANDNOT, "maybe": M.OP_MAYBE}[kind]
        return M.Node(op, children=[tree_to_node(c) for c in t[1:]])
    if kind == "phrase":
        n = M.PHRASE([(w, p) for w, p in t[1]])
        if len(t) > 2:
            n.field_mask = t[2]
        return n
    if kind == "prox":
        return M.PROXIMITY([(w, p) for w, p in t[2]], t[1])
    if kind == "quorum":
        return M.Node(M.OP_QUORUM, words=[M.Keyword(w, p) for w, p in t[2]], oparg=t[1])
    if kind == "near":
        return M.Node(M.OP_NEAR, children=[tree_to_node(c) for c in t[2:]], oparg=t[1])
    if kind == "before":
        return M.Node(M.OP_BEFORE, children=[tree_to_node(c) for c in t[1:]])
    if kind in ("sentence", "paragraph"):
        return M.Node(M.OP_SENTENCE if kind == "sentence" else M.OP_PARAGRAPH, children=[tree_to_node(c) for c in t[1:]])
    if kind == "notnear":
        return M.Node(M.OP_NOTNEAR, children=[tree_to_node(t[2]), tree_to_node(t[3])], oparg=t[1])
    raise ValueError(kind)


RANKERS = {"proximity_bm25": M.RANK_PROXIMITY_BM25, "bm25": M.RANK_BM25, "none": M.RANK_NONE, "wordcount": M.RANK_WORDCOUNT,
           "proximity": M.RANK_PROXIMITY, "matchany": M.RANK_MATCHANY, "fieldmask": M.RANK_FIELDMASK, "sph04": M.RANK_SPH04}


def load_golden():
    with open(GOLDEN, encoding="utf-8") as f:
        return json.load(f)["cases"]


def case_tokens(case, text):
    """a golden case's document text -> [(keyword, pos)] with the case's indexing settings"""
    return tokenize(html_strip(text) if case.get("html_strip") else text, case.get("min_word_len", 1), case.get("stopwords", ()),
                    case.get("phrase_boundary", ""), case.get("phrase_boundary_step", 0), index_sp=bool(case.get("index_sp")))


def build_golden_index(case, prefix, dict_crc=False):
    docs = []
    for d in case["docs"]:
        docs.append({"id": d["id"], "fields": [case_tokens(case, t) for t in d["fields"]], "attrs": d.get("attrs", [])})
    M.build_index(prefix, case["fields"], docs, attr_names=case.get("attrs", ()), dict_crc=dict_crc)


def golden_query(case, q):
    # "sort": "id_asc" = SphinxQL `order by id asc`: documents are indexed in id order, so that is rowid ascending
    # "sort": "weight_asc" = sortmode expr "-@weight" (ties by id ascending, as every comparator of the reference ends)
    sort_keys = {"id_asc": [M.SortKey(M.KEYPART_ROWID, 0, False)], "weight_asc": [M.SortKey(M.KEYPART_WEIGHT, 0, False)]}.get(q.get("sort"))
    # "filters": [[attribute name, min, max]...] (attribute 0 of a v62 schema is the document id, the case's "attrs" follow)
    filters = [M.Filter(1 + case["attrs"].index(a), lo, hi) for a, lo, hi in q.get("filters", [])]
    return M.Query(tree_to_node(q["tree"]), ranker=RANKERS[q["ranker"]], field_weights=q.get("field_weights"), sort_keys=sort_keys,
                   filters=filters, max_matches=1000)


def assert_same_results(a, b, ctx=""):
    """bit-exact comparison of two ResultSet.get() dicts: ids, integer weights, order, total_found"""
    assert a["status"] == b["status"], (ctx, a["status"], b["status"])
    assert a["total_found"] == b["total_found"], (ctx, "total_found", a["total_found"], b["total_found"])
    assert a["rowid"] == b["rowid"], (ctx, "rowid/order", a["rowid"][:10], b["rowid"][:10])
    assert a["weight"] == b["weight"], (ctx, "weight", a["weight"][:10], b["weight"][:10])
    assert a["docid"] == b["docid"], (ctx, "docid")

"""ctypes mirror of include/mgpu.h.

Query trees are built from the same pieces the reference's parser produces (XQNode_t / XQKeyword_t,
src/sphinxquery.h:21-39, 134-280): keyword nodes carry one word; AND/OR/ANDNOT/MAYBE nodes carry
children; PHRASE/PROXIMITY nodes carry a word list.  `Query.pack()` flattens a tree into the C structs.
The same packed structs are accepted by the CPU oracle (oracle/oracle.cpp), which is how the parity
tests feed both sides identical inputs.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))

MGPU_OK = 0
MGPU_E_IO, MGPU_E_FORMAT, MGPU_E_UNSUPPORTED, MGPU_E_BAD_QUERY, MGPU_E_NO_DEVICE, MGPU_E_CUDA, MGPU_E_NOMEM = -1, -2, -3, -4, -5, -6, -7

OP_AND, OP_OR, OP_MAYBE, OP_NOT, OP_ANDNOT, OP_BEFORE, OP_PHRASE, OP_PROXIMITY, OP_QUORUM, OP_NEAR, OP_NOTNEAR, OP_SENTENCE, OP_PARAGRAPH = range(13)
RANK_PROXIMITY_BM25, RANK_BM25, RANK_NONE, RANK_WORDCOUNT, RANK_PROXIMITY, RANK_MATCHANY, RANK_FIELDMASK, RANK_SPH04 = range(8)
KEYPART_ROWID, KEYPART_WEIGHT, KEYPART_INT, KEYPART_FLOAT = range(4)
FILTER_RANGE, FILTER_VALUES = range(2)


class MgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mgpu error %d: %s" % (code, msg))
        self.code = code


class c_xqkeyword(C.Structure):
    _fields_ = [("word", C.c_char_p), ("atom_pos", C.c_int32), ("boost", C.c_float),
                ("field_start", C.c_uint8), ("field_end", C.c_uint8), ("excluded", C.c_uint8), ("expanded", C.c_uint8)]


class c_xqnode(C.Structure):
    _fields_ = [("op", C.c_int32), ("oparg", C.c_int32), ("first_child", C.c_int32), ("n_children", C.c_int32),
                ("first_word", C.c_int32), ("n_words", C.c_int32), ("field_mask", C.c_uint32), ("field_max_pos", C.c_int32),
                ("not_weighted", C.c_uint8), ("pad", C.c_uint8 * 3)]


class c_sortkey(C.Structure):
    _fields_ = [("kind", C.c_int32), ("attr", C.c_int32), ("desc", C.c_int32)]


class c_filter(C.Structure):
    _fields_ = [("kind", C.c_int32), ("attr", C.c_int32), ("min_value", C.c_int64), ("max_value", C.c_int64),
                ("values", C.POINTER(C.c_int64)), ("n_values", C.c_int32), ("exclude", C.c_int32)]


class c_query(C.Structure):
    _fields_ = [("nodes", C.POINTER(c_xqnode)), ("n_nodes", C.c_int32), ("root", C.c_int32),
                ("children", C.POINTER(C.c_int32)), ("n_children", C.c_int32),
                ("words", C.POINTER(c_xqkeyword)), ("n_words", C.c_int32),
                ("ranker", C.c_int32), ("field_weights", C.POINTER(C.c_int32)), ("n_field_weights", C.c_int32),
                ("sort_keys", C.POINTER(c_sortkey)), ("n_sort_keys", C.c_int32),
                ("filters", C.POINTER(c_filter)), ("n_filters", C.c_int32),
                ("max_matches", C.c_int32), ("index_weight", C.c_int32),
                ("plain_idf", C.c_uint8), ("unnormalized_tfidf", C.c_uint8), ("shard_of_global", C.c_uint8), ("pad", C.c_uint8 * 1),
                ("total_docs", C.c_int64), ("word_docs", C.POINTER(C.c_int64))]


class c_wordstat(C.Structure):
    _fields_ = [("docs", C.c_int64), ("hits", C.c_int64)]


class c_result(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_matches", C.c_int32), ("total_found", C.c_int64),
                ("rowid", C.POINTER(C.c_uint32)), ("weight", C.POINTER(C.c_int32)), ("docid", C.POINTER(C.c_int64)),
                ("sort_attr", C.POINTER(C.c_int64)), ("word_stats", C.POINTER(c_wordstat))]


class c_batch_stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("work_items", C.c_int64), ("algorithmic_bytes", C.c_int64),
                ("postings", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("eval_kernel_ms", C.c_float), ("merge_kernel_ms", C.c_float), ("hot_decode_ms", C.c_float), ("hot_terms", C.c_int32),
                ("class_ms", C.c_float * 7), ("class_queries", C.c_int32 * 7), ("class_bytes", C.c_int64 * 7),
                ("host_plan_ms", C.c_float), ("host_setup_ms", C.c_float), ("host_fetch_ms", C.c_float),
                ("host_wait_ms", C.c_float), ("host_total_ms", C.c_float),
                ("or_kernel", C.c_int32), ("hitlist_bytes", C.c_int64), ("attr_rows", C.c_int64)]


class c_sharded_stats(C.Structure):
    _fields_ = [("n_shards", C.c_int32), ("nccl", C.c_int32), ("host_total_ms", C.c_float), ("host_plan_ms", C.c_float),
                ("host_setup_ms", C.c_float), ("host_wait_ms", C.c_float), ("host_fetch_ms", C.c_float),
                ("max_eval_kernel_ms", C.c_float), ("max_hot_decode_ms", C.c_float), ("kernel_launches", C.c_int32),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("algorithmic_bytes", C.c_int64), ("postings", C.c_int64)]


class c_build_doc_input(C.Structure):
    _fields_ = [("n_docs", C.c_int32), ("n_fields", C.c_int32), ("field_names", C.POINTER(C.c_char_p)),
                ("n_attrs", C.c_int32), ("attr_names", C.POINTER(C.c_char_p)),
                ("docids", C.POINTER(C.c_int64)), ("attrs", C.POINTER(C.c_uint32)),
                ("n_keywords", C.c_int32), ("keywords", C.POINTER(C.c_char_p)),
                ("field_tok_offsets", C.POINTER(C.c_int64)), ("tok_keyword", C.POINTER(C.c_int32)), ("tok_pos", C.POINTER(C.c_int32)),
                ("skiplist_block", C.c_int32), ("hit_format_inline", C.c_int32), ("dict_crc", C.c_int32)]


class c_parser_settings(C.Structure):
    _fields_ = [("n_fields", C.c_int32), ("field_names", C.POINTER(C.c_char_p)), ("min_word_len", C.c_int32),
                ("n_stopwords", C.c_int32), ("stopwords", C.POINTER(C.c_char_p)), ("overshort_step", C.c_int32),
                ("stopword_step", C.c_int32), ("match_mode", C.c_int32), ("ngram_cjk", C.c_int32)]


class SynthParams(C.Structure):
    """mgpu_synth_params; defaults = the corpus of SURVEY.md 8(d)."""
    _fields_ = [("seed", C.c_uint64), ("first_doc", C.c_int64), ("n_docs", C.c_int64), ("vocab", C.c_int32),
                ("title_min", C.c_int32), ("title_max", C.c_int32), ("body_min", C.c_int32), ("body_max", C.c_int32),
                ("body_mu", C.c_float), ("body_sigma", C.c_float), ("threads", C.c_int32)]

    def __init__(self, n_docs, first_doc=0, seed=0x5EED0001, vocab=1 << 20, threads=0,
                 title_min=4, title_max=12, body_min=16, body_max=1024, body_mu=4.6, body_sigma=0.6):
        super().__init__(seed, first_doc, n_docs, vocab, title_min, title_max, body_min, body_max, body_mu, body_sigma, threads)


_lib = None


def load_library(path=None):
    """Loads libmgpu.so (building it first if the sources are newer). Raises if that is impossible."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    if path is None:
        from . import build as _build
        path = _build.LIB
        if _build.needs_build():
            _build.build()
    if not os.path.exists(path):
        raise MgpuError(MGPU_E_IO, "libmgpu.so is not built (run python -m manticoresearch_b200.build)")
    lib_ = C.CDLL(path)
    vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    sig = {
        "mgpu_abi_version": (C.c_int, []),
        "mgpu_index_open": (C.c_int, [C.c_char_p, C.c_int, u32, C.POINTER(vp)]),
        "mgpu_index_close": (C.c_int, [vp]),
        "mgpu_index_set_stream": (C.c_int, [vp, vp]),
        "mgpu_index_set_option": (C.c_int, [vp, C.c_char_p, C.c_int64]),
        "mgpu_last_error": (C.c_char_p, [vp]),
        "mgpu_index_total_docs": (i64, [vp]),
        "mgpu_index_num_fields": (i32, [vp]),
        "mgpu_index_field_index": (i32, [vp, C.c_char_p]),
        "mgpu_index_attr_index": (i32, [vp, C.c_char_p]),
        "mgpu_index_word_stats": (C.c_int, [vp, C.c_char_p, C.POINTER(i64), C.POINTER(i64)]),
        "mgpu_index_word_bytes": (C.c_int, [vp, C.c_char_p, C.POINTER(i64), C.POINTER(i64)]),
        "mgpu_search_batch": (C.c_int, [vp, C.POINTER(c_query), C.c_int, C.POINTER(c_result)]),
        "mgpu_batch_prepare": (C.c_int, [vp, C.POINTER(c_query), C.c_int, C.POINTER(vp)]),
        "mgpu_batch_run": (C.c_int, [vp]),
        "mgpu_batch_sync": (C.c_int, [vp]),
        "mgpu_batch_fetch": (C.c_int, [vp, C.POINTER(c_result)]),
        "mgpu_batch_free": (None, [vp]),
        "mgpu_batch_get_stats": (C.c_int, [vp, C.POINTER(c_batch_stats)]),
        "mgpu_index_last_search_stats": (C.c_int, [vp, C.POINTER(c_batch_stats)]),
        "mgpu_batch_export_keys": (C.c_int, [vp, vp, vp, vp, C.c_int]),
        "mgpu_sharded_open": (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
        "mgpu_sharded_close": (None, [vp]),
        "mgpu_sharded_search_batch": (C.c_int, [vp, C.POINTER(c_query), C.c_int, C.POINTER(c_result)]),
        "mgpu_sharded_set_option": (C.c_int, [vp, C.c_char_p, C.c_int64]),
        "mgpu_sharded_total_docs": (i64, [vp]),
        "mgpu_sharded_word_docs": (C.c_int, [vp, C.c_char_p, C.POINTER(i64)]),
        "mgpu_sharded_last_error": (C.c_char_p, [vp]),
        "mgpu_sharded_get_stats": (C.c_int, [vp, C.POINTER(c_sharded_stats)]),
        "mgpu_merge_shard_keys": (C.c_int, [C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
        "mgpu_unpack_key": (None, [C.POINTER(C.c_uint64), C.POINTER(u32), C.POINTER(i32), C.POINTER(C.c_uint64)]),
        "mgpu_decode_doclist": (C.c_int, [vp, C.c_char_p, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_uint64), i64, C.POINTER(i64)]),
        "mgpu_api_create": (C.c_int, [vp, C.c_char_p, C.POINTER(c_parser_settings), C.POINTER(vp)]),
        "mgpu_api_create_sharded": (C.c_int, [vp, C.POINTER(C.c_char_p), C.c_int, C.POINTER(c_parser_settings), C.POINTER(vp)]),
        "mgpu_sharded_word_stats": (C.c_int, [vp, C.c_char_p, C.POINTER(i64), C.POINTER(i64)]),
        "mgpu_api_handle": (C.c_int, [vp, C.c_char_p, C.c_size_t, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "mgpu_api_describe_last": (C.c_char_p, [vp]),
        "mgpu_api_free": (None, [vp]),
        "mgpu_index_field_name": (C.c_char_p, [vp, i32]),
        "mgpu_index_check": (C.c_int, [C.c_char_p, C.POINTER(i64), C.c_char_p, C.c_int]),
        "mgpu_parse_query": (C.c_int, [C.POINTER(c_parser_settings), C.c_char_p, C.POINTER(vp)]),
        "mgpu_parsed_fill": (C.c_int, [vp, C.POINTER(c_query)]),
        "mgpu_parsed_error": (C.c_char_p, [vp]),
        "mgpu_parsed_warning": (C.c_char_p, [vp]),
        "mgpu_parsed_explain": (C.c_char_p, [vp]),
        "mgpu_parsed_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib_, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib_
    return lib_


def lib():
    return load_library()


_wlib = None


def writer_lib():
    """libmgpu_writer.so (include/mgpu_writer.h): the v62 index writer + synthetic corpus; host-only, no CUDA"""
    global _wlib
    if _wlib is not None:
        return _wlib
    from . import build as _build
    if _build.needs_build():
        _build.build()
    if not os.path.exists(_build.WRITER_LIB):
        raise MgpuError(MGPU_E_IO, "libmgpu_writer.so is not built (run python -m manticoresearch_b200.build)")
    w = C.CDLL(_build.WRITER_LIB)
    i32, i64 = C.c_int32, C.c_int64
    sig = {
        "mgpu_writer_abi_version": (C.c_int, []),
        "mgpu_build_index": (C.c_int, [C.c_char_p, C.POINTER(c_build_doc_input), C.c_char_p, C.c_int]),
        "mgpu_build_synthetic": (C.c_int, [C.c_char_p, C.POINTER(SynthParams), C.c_char_p, C.c_int]),
        "mgpu_synth_field_len": (i32, [C.POINTER(SynthParams), i64, C.c_int]),
        "mgpu_synth_token": (i32, [C.POINTER(SynthParams), i64, C.c_int, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(w, name)
        fn.restype = res
        fn.argtypes = args
    _wlib = w
    return w


WRITER_EXPORTED_SYMBOLS = ["mgpu_writer_abi_version", "mgpu_build_index", "mgpu_build_synthetic", "mgpu_synth_field_len", "mgpu_synth_token"]


EXPORTED_SYMBOLS = [
    "mgpu_abi_version", "mgpu_index_open", "mgpu_index_close", "mgpu_index_set_stream", "mgpu_index_set_option", "mgpu_last_error", "mgpu_index_total_docs",
    "mgpu_index_num_fields", "mgpu_index_field_index", "mgpu_index_attr_index", "mgpu_index_word_stats",
    "mgpu_index_word_bytes", "mgpu_search_batch", "mgpu_batch_prepare", "mgpu_batch_run", "mgpu_batch_sync",
    "mgpu_batch_fetch", "mgpu_batch_free", "mgpu_batch_get_stats", "mgpu_index_last_search_stats", "mgpu_batch_export_keys",
    "mgpu_merge_shard_keys", "mgpu_unpack_key", "mgpu_decode_doclist",
    "mgpu_sharded_open", "mgpu_sharded_close", "mgpu_sharded_search_batch", "mgpu_sharded_set_option", "mgpu_sharded_total_docs",
    "mgpu_sharded_word_docs", "mgpu_sharded_last_error", "mgpu_sharded_get_stats",
    "mgpu_api_create", "mgpu_api_create_sharded", "mgpu_sharded_word_stats", "mgpu_api_handle", "mgpu_api_describe_last", "mgpu_api_free",
    "mgpu_index_field_name", "mgpu_index_check", "mgpu_parse_query", "mgpu_parsed_fill", "mgpu_parsed_error", "mgpu_parsed_warning", "mgpu_parsed_explain", "mgpu_parsed_free",
]

# ---------------------------------------------------------------------------------------------
# query trees
# ---------------------------------------------------------------------------------------------

ALL_FIELDS = 0xFFFFFFFF


class Keyword:
    """XQKeyword_t"""
    def __init__(self, word, atom_pos, boost=1.0, field_start=False, field_end=False, excluded=False, expanded=False):
        self.word, self.atom_pos, self.boost = word, atom_pos, boost
        self.field_start, self.field_end, self.excluded, self.expanded = field_start, field_end, excluded, expanded


class Node:
    """XQNode_t: either words (keyword / phrase / proximity) or children (AND/OR/...)."""
    def __init__(self, op, children=None, words=None, oparg=0, field_mask=ALL_FIELDS, field_max_pos=0, not_weighted=False):
        self.op, self.children, self.words = op, list(children or []), list(words or [])
        self.oparg, self.field_mask, self.field_max_pos, self.not_weighted = oparg, field_mask, field_max_pos, not_weighted

    def fields(self, mask):
        """apply a field limit (@title ...) to this node and everything below it"""
        self.field_mask = mask
        for c in self.children:
            c.fields(mask)
        return self

    def all_keywords(self):
        out = list(self.words)
        for c in self.children:
            out += c.all_keywords()
        return out


def kw(word, atom_pos, **kwargs):
    return Node(OP_AND, words=[Keyword(word, atom_pos, **kwargs)])


def AND(*children):
    return Node(OP_AND, children=children)


def OR(*children):
    return Node(OP_OR, children=children)


def ANDNOT(*children):
    return Node(OP_ANDNOT, children=children)


def MAYBE(*children):
    return Node(OP_MAYBE, children=children)


def PHRASE(words_with_pos, **kwargs):
    return Node(OP_PHRASE, words=[Keyword(w, p) for w, p in words_with_pos], **kwargs)


def PROXIMITY(words_with_pos, distance, **kwargs):
    return Node(OP_PROXIMITY, words=[Keyword(w, p) for w, p in words_with_pos], oparg=distance, **kwargs)


def NEAR(distance, *children):
    """a NEAR/n b ...: children are nodes (the CUDA path runs the two-keyword form)"""
    return Node(OP_NEAR, children=list(children), oparg=distance)


def BEFORE(*children):
    """a << b << c"""
    return Node(OP_BEFORE, children=list(children))


def NOTNEAR(distance, must, unwanted):
    """must NOTNEAR/n unwanted"""
    return Node(OP_NOTNEAR, children=[must, unwanted], oparg=distance)


def QUORUM(words_with_pos, threshold, **kwargs):
    """"a b c"/N with an ABSOLUTE threshold"""
    return Node(OP_QUORUM, words=[Keyword(w, p) for w, p in words_with_pos], oparg=threshold, **kwargs)


class SortKey:
    def __init__(self, kind, attr=0, desc=True):
        self.kind, self.attr, self.desc = kind, attr, desc


class Filter:
    def __init__(self, attr, min_value=None, max_value=None, values=None, exclude=False):
        self.attr, self.min_value, self.max_value, self.values, self.exclude = attr, min_value, max_value, values, exclude


MATCH_ALL, MATCH_ANY, MATCH_PHRASE, MATCH_BOOLEAN, MATCH_EXTENDED = range(5)


class ApiResponder:
    """mgpu_api_*: SphinxAPI `search` / `keywords` packets in, reply packets out. index: an Index (path_prefix = its prefix), a ShardedIndex
    (path_prefix = the shards' prefixes in order) or None (host-only: parse / describe / error replies)."""
    def __init__(self, index, path_prefix, min_word_len=None, stopwords=()):
        self._lib = lib()
        self._h = C.c_void_p()
        st = None
        if min_word_len is not None or stopwords:
            st = c_parser_settings()
            self._stops = (C.c_char_p * max(1, len(stopwords)))(*[w.encode("utf-8") for w in stopwords])
            st.min_word_len, st.n_stopwords, st.stopwords = min_word_len or 1, len(stopwords), self._stops
            st.overshort_step, st.stopword_step, st.ngram_cjk = 1, 1, 1
        pst = C.byref(st) if st is not None else None
        if isinstance(index, ShardedIndex):
            prefixes = list(path_prefix)
            arr = (C.c_char_p * len(prefixes))(*[p.encode() for p in prefixes])
            rc = self._lib.mgpu_api_create_sharded(index._h, arr, len(prefixes), pst, C.byref(self._h))
        else:
            rc = self._lib.mgpu_api_create(index._h if index is not None else None, path_prefix.encode(), pst, C.byref(self._h))
        if rc != MGPU_OK:
            raise MgpuError(rc, "mgpu_api_create")

    def handle(self, request):
        reply, n = C.c_void_p(), C.c_size_t()
        rc = self._lib.mgpu_api_handle(self._h, bytes(request), len(request), C.byref(reply), C.byref(n))
        if rc != MGPU_OK:
            raise MgpuError(rc, "mgpu_api_handle")
        return C.string_at(reply, n.value)

    def describe_last(self):
        return self._lib.mgpu_api_describe_last(self._h).decode("utf-8", "replace")

    def close(self):
        if self._h:
            self._lib.mgpu_api_free(self._h)
            self._h = C.c_void_p()


def check_index(path_prefix):
    """mgpu_index_check -> (number of failures, report text). Host only."""
    n = C.c_int64(0)
    buf = C.create_string_buffer(1 << 16)
    rc = lib().mgpu_index_check(path_prefix.encode(), C.byref(n), buf, len(buf))
    if rc != MGPU_OK:
        raise MgpuError(rc, "mgpu_index_check: " + buf.value.decode("utf-8", "replace").strip())
    return n.value, buf.value.decode("utf-8", "replace")


def explain_query(text, field_names, **kwargs):
    """mgpu_parsed_explain: the parsed tree in SHOW PLAN's text form"""
    return parse_query(text, field_names, _explain=True, **kwargs)


def parse_query(text, field_names, min_word_len=1, stopwords=(), match_mode=MATCH_EXTENDED, ngram_cjk=True, overshort_step=1, stopword_step=1, _explain=False):
    """mgpu_parse_query -> (root Node, ranker forced by a legacy match mode or None, warning). Raises MgpuError on a parse error.
    Host only: works without a GPU."""
    l = lib()
    st = c_parser_settings()
    names = (C.c_char_p * max(1, len(field_names)))(*[f.encode() for f in field_names])
    stops = (C.c_char_p * max(1, len(stopwords)))(*[w.encode("utf-8") for w in stopwords])
    st.n_fields, st.field_names, st.min_word_len = len(field_names), names, min_word_len
    st.n_stopwords, st.stopwords = len(stopwords), stops
    st.overshort_step, st.stopword_step, st.match_mode, st.ngram_cjk = overshort_step, stopword_step, match_mode, int(ngram_cjk)
    h = C.c_void_p()
    rc = l.mgpu_parse_query(C.byref(st), text.encode("utf-8"), C.byref(h))
    if not h:
        raise MgpuError(rc, "mgpu_parse_query: bad arguments")
    try:
        if rc != MGPU_OK:
            raise MgpuError(rc, l.mgpu_parsed_error(h).decode("utf-8", "replace"))
        if _explain:
            return l.mgpu_parsed_explain(h).decode("utf-8", "replace")
        q = c_query()
        q.ranker = -1
        rc = l.mgpu_parsed_fill(h, C.byref(q))
        if rc != MGPU_OK:
            raise MgpuError(rc, "mgpu_parsed_fill")

        def build(i):
            cn = q.nodes[i]
            n = Node(cn.op, oparg=cn.oparg, field_mask=cn.field_mask, field_max_pos=cn.field_max_pos)
            for k in range(cn.first_word, cn.first_word + cn.n_words):
                w = q.words[k]
                n.words.append(Keyword(w.word.decode("utf-8"), w.atom_pos, boost=w.boost, field_start=bool(w.field_start), field_end=bool(w.field_end), excluded=bool(w.excluded)))
            for k in range(cn.first_child, cn.first_child + cn.n_children):
                n.children.append(build(q.children[k]))
            return n

        root = build(q.root)
        return root, (q.ranker if q.ranker >= 0 else None), l.mgpu_parsed_warning(h).decode("utf-8", "replace")
    finally:
        l.mgpu_parsed_free(h)


class Query:
    """CSphQuery subset + the parsed tree."""
    def __init__(self, root, ranker=RANK_PROXIMITY_BM25, field_weights=None, sort_keys=None, filters=None,
                 max_matches=1000, index_weight=1, plain_idf=False, unnormalized_tfidf=False, total_docs=0, word_docs=None,
                 shard_of_global=False):
        self.root, self.ranker, self.field_weights = root, ranker, field_weights
        self.sort_keys, self.filters = sort_keys or [], filters or []
        self.max_matches, self.index_weight = max_matches, index_weight
        self.plain_idf, self.unnormalized_tfidf = plain_idf, unnormalized_tfidf
        self.shard_of_global = shard_of_global
        self.total_docs, self.word_docs = total_docs, word_docs
        self._keep = []

    def keywords(self):
        return self.root.all_keywords() if self.root is not None else []

    def pack(self, out=None):
        """fills (and returns) a c_query; keeps the backing arrays alive on self"""
        nodes, children, words = [], [], []

        def walk(n):
            idx = len(nodes)
            nodes.append(None)
            cn = c_xqnode()
            cn.op, cn.oparg = n.op, n.oparg
            cn.field_mask, cn.field_max_pos, cn.not_weighted = n.field_mask & 0xFFFFFFFF, n.field_max_pos, int(n.not_weighted)
            cn.first_word, cn.n_words = len(words), len(n.words)
            for k in n.words:
                ck = c_xqkeyword(k.word.encode("utf-8"), k.atom_pos, k.boost, int(k.field_start), int(k.field_end), int(k.excluded), int(k.expanded))
                words.append(ck)
            kids = [walk(c) for c in n.children]
            cn.first_child, cn.n_children = len(children), len(kids)
            children.extend(kids)
            nodes[idx] = cn
            return idx

        q = out if out is not None else c_query()
        root = walk(self.root) if self.root is not None else -1
        a_nodes = (c_xqnode * max(1, len(nodes)))(*nodes)
        a_children = (C.c_int32 * max(1, len(children)))(*children)
        a_words = (c_xqkeyword * max(1, len(words)))(*words)
        self._keep = [a_nodes, a_children, a_words]
        q.nodes, q.n_nodes, q.root = a_nodes, len(nodes), root
        q.children, q.n_children = a_children, len(children)
        q.words, q.n_words = a_words, len(words)
        q.ranker = self.ranker
        if self.field_weights is not None:
            a_w = (C.c_int32 * len(self.field_weights))(*self.field_weights)
            self._keep.append(a_w)
            q.field_weights, q.n_field_weights = a_w, len(self.field_weights)
        else:
            q.field_weights, q.n_field_weights = None, 0
        if self.sort_keys:
            a_s = (c_sortkey * len(self.sort_keys))(*[c_sortkey(s.kind, s.attr, int(s.desc)) for s in self.sort_keys])
            self._keep.append(a_s)
            q.sort_keys, q.n_sort_keys = a_s, len(self.sort_keys)
        else:
            q.sort_keys, q.n_sort_keys = None, 0
        if self.filters:
            fl = []
            for f in self.filters:
                cf = c_filter()
                cf.attr, cf.exclude = f.attr, int(f.exclude)
                if f.values is not None:
                    a_v = (C.c_int64 * len(f.values))(*f.values)
                    self._keep.append(a_v)
                    cf.kind, cf.values, cf.n_values = FILTER_VALUES, a_v, len(f.values)
                else:
                    cf.kind, cf.min_value, cf.max_value = FILTER_RANGE, f.min_value, f.max_value
                fl.append(cf)
            a_f = (c_filter * len(fl))(*fl)
            self._keep.append(a_f)
            q.filters, q.n_filters = a_f, len(fl)
        else:
            q.filters, q.n_filters = None, 0
        q.max_matches, q.index_weight = self.max_matches, self.index_weight
        q.plain_idf, q.unnormalized_tfidf = int(self.plain_idf), int(self.unnormalized_tfidf)
        q.shard_of_global = int(self.shard_of_global)
        q.total_docs = self.total_docs
        if self.word_docs is not None:
            a_d = (C.c_int64 * len(self.word_docs))(*self.word_docs)
            self._keep.append(a_d)
            q.word_docs = a_d
        else:
            q.word_docs = None
        return q


class ResultSet:
    """host buffers for a batch of mgpu_result + numpy-free accessors"""
    def __init__(self, queries):
        n = len(queries)
        self.results = (c_result * max(1, n))()
        self._keep = []
        for i, q in enumerate(queries):
            k = q.max_matches if q.max_matches > 0 else 1000
            nw = max(1, len(q.keywords()))
            bufs = ((C.c_uint32 * k)(), (C.c_int32 * k)(), (C.c_int64 * k)(), (C.c_int64 * k)(), (c_wordstat * nw)())
            self._keep.append(bufs)
            r = self.results[i]
            r.rowid, r.weight, r.docid, r.sort_attr, r.word_stats = bufs
        self.n = n

    def get(self, i):
        r = self.results[i]
        n = r.n_matches
        return {
            "status": r.status, "total_found": r.total_found,
            "rowid": list(r.rowid[:n]), "weight": list(r.weight[:n]), "docid": list(r.docid[:n]), "sort_attr": list(r.sort_attr[:n]),
        }

    def word_stats(self, i, nwords):
        r = self.results[i]
        return [(r.word_stats[w].docs, r.word_stats[w].hits) for w in range(nwords)]


def pack_queries(queries):
    arr = (c_query * max(1, len(queries)))()
    for i, q in enumerate(queries):
        q.pack(arr[i])
    return arr


class Index:
    """Handle of an index resident in one GPU's HBM (mgpu_index)."""
    def __init__(self, path_prefix, device=0, rowid_base=0):
        self._lib = lib()
        h = C.c_void_p()
        rc = self._lib.mgpu_index_open(path_prefix.encode(), device, rowid_base, C.byref(h))
        if rc != MGPU_OK:
            raise MgpuError(rc, (self._lib.mgpu_last_error(None) or b"").decode())
        self._h = h
        self.device = device

    def close(self):
        if self._h:
            rc = self._lib.mgpu_index_close(self._h)
            if rc != MGPU_OK:       # batches of this handle are still alive: the handle stays open
                raise MgpuError(rc, (self._lib.mgpu_last_error(self._h) or b"").decode())
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, rc):
        raise MgpuError(rc, (self._lib.mgpu_last_error(self._h) or b"").decode())

    def set_stream(self, cuda_stream_handle):
        """run on the caller's stream (e.g. torch.cuda.current_stream().cuda_stream); 0/None = private stream"""
        self._lib.mgpu_index_set_stream(self._h, C.c_void_p(cuda_stream_handle or 0))

    def set_option(self, name, value):
        """engine option of this handle (mgpu_index_set_option): tuning and the A/B switches of the launch classes"""
        rc = self._lib.mgpu_index_set_option(self._h, name.encode(), int(value))
        if rc != 0:
            self._err(rc)

    @property
    def total_docs(self):
        return self._lib.mgpu_index_total_docs(self._h)

    @property
    def num_fields(self):
        return self._lib.mgpu_index_num_fields(self._h)

    def field_index(self, name):
        return self._lib.mgpu_index_field_index(self._h, name.encode())

    def attr_index(self, name):
        return self._lib.mgpu_index_attr_index(self._h, name.encode())

    def word_stats(self, word):
        d, h = C.c_int64(), C.c_int64()
        if not self._lib.mgpu_index_word_stats(self._h, word.encode(), C.byref(d), C.byref(h)):
            return None
        return d.value, h.value

    def word_bytes(self, word):
        d, s = C.c_int64(), C.c_int64()
        if not self._lib.mgpu_index_word_bytes(self._h, word.encode(), C.byref(d), C.byref(s)):
            return None
        return d.value, s.value

    def search(self, queries):
        """mgpu_search_batch: host buffers in, host buffers out. Returns a ResultSet."""
        arr = pack_queries(queries)
        rs = ResultSet(queries)
        rc = self._lib.mgpu_search_batch(self._h, arr, len(queries), rs.results)
        if rc != MGPU_OK:
            self._err(rc)
        return rs

    def search_packed(self, packed, n, result_set):
        """mgpu_search_batch on already marshalled host buffers (pack_queries / ResultSet): the bare C-ABI call"""
        rc = self._lib.mgpu_search_batch(self._h, packed, n, result_set.results)
        if rc != MGPU_OK:
            self._err(rc)
        return result_set

    def last_search_stats(self):
        st = c_batch_stats()
        self._lib.mgpu_index_last_search_stats(self._h, C.byref(st))
        return {k: (list(getattr(st, k)) if k.startswith("class_") else getattr(st, k)) for k, _ in c_batch_stats._fields_}

    def prepare(self, queries, packed=None):
        return Batch(self, queries, packed)

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
        rc = self._lib.mgpu_decode_doclist(self._h, word.encode(),
                                           rowid.ctypes.data_as(C.POINTER(C.c_uint32)), hits.ctypes.data_as(C.POINTER(C.c_uint32)),
                                           fields.ctypes.data_as(C.POINTER(C.c_uint32)), pos.ctypes.data_as(C.POINTER(C.c_uint64)), n, C.byref(nout))
        if rc != MGPU_OK:
            self._err(rc)
        return rowid, hits, fields, pos

    def decode_doclist_timed(self, word):
        """decode only, nothing copied back (roofline probe of kernel K1)"""
        nout = C.c_int64()
        rc = self._lib.mgpu_decode_doclist(self._h, word.encode(), None, None, None, None, 0, C.byref(nout))
        if rc != MGPU_OK:
            self._err(rc)
        return nout.value


class ShardedIndex:
    """Rowid-range shards of one index behind one handle, one GPU per shard (mgpu_sharded): plan once, one host thread per
    shard, NCCL exchange of the K keys per query, results of the unsharded index (rowid = global rowid)."""
    def __init__(self, path_prefixes, devices):
        self._lib = lib()
        n = len(path_prefixes)
        arr = (C.c_char_p * n)(*[p.encode() for p in path_prefixes])
        dev = (C.c_int * n)(*devices)
        h = C.c_void_p()
        rc = self._lib.mgpu_sharded_open(arr, dev, n, C.byref(h))
        if rc != MGPU_OK:
            raise MgpuError(rc, (self._lib.mgpu_sharded_last_error(None) or b"").decode())
        self._h = h
        self.n_shards = n

    def close(self):
        if self._h:
            self._lib.mgpu_sharded_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, rc):
        raise MgpuError(rc, (self._lib.mgpu_sharded_last_error(self._h) or b"").decode())

    def set_option(self, name, value):
        rc = self._lib.mgpu_sharded_set_option(self._h, name.encode(), int(value))
        if rc != 0:
            self._err(rc)

    @property
    def total_docs(self):
        return self._lib.mgpu_sharded_total_docs(self._h)

    def word_docs(self, word):
        d = C.c_int64()
        return d.value if self._lib.mgpu_sharded_word_docs(self._h, word.encode("utf-8"), C.byref(d)) else 0

    def search(self, queries):
        arr = pack_queries(queries)
        rs = ResultSet(queries)
        return self.search_packed(arr, len(queries), rs)

    def search_packed(self, packed, n, result_set):
        """mgpu_sharded_search_batch on already marshalled host buffers: the bare C-ABI call"""
        rc = self._lib.mgpu_sharded_search_batch(self._h, packed, n, result_set.results)
        if rc != MGPU_OK:
            self._err(rc)
        return result_set

    def stats(self):
        st = c_sharded_stats()
        self._lib.mgpu_sharded_get_stats(self._h, C.byref(st))
        return {k: getattr(st, k) for k, _ in c_sharded_stats._fields_}


class Batch:
    """mgpu_batch: plan uploaded once; run() may be repeated (the device-resident timed region)."""
    def __init__(self, index, queries, packed=None):
        self.index, self.queries = index, queries
        self._lib = index._lib
        self._arr = packed if packed is not None else pack_queries(queries)     # `packed`: already marshalled host buffers
        h = C.c_void_p()
        rc = self._lib.mgpu_batch_prepare(index._h, self._arr, len(queries), C.byref(h))
        if rc != MGPU_OK:
            index._err(rc)
        self._h = h

    def run(self):
        rc = self._lib.mgpu_batch_run(self._h)
        if rc != MGPU_OK:
            self.index._err(rc)

    def sync(self):
        rc = self._lib.mgpu_batch_sync(self._h)
        if rc != MGPU_OK:
            self.index._err(rc)

    def fetch(self):
        rs = ResultSet(self.queries)
        rc = self._lib.mgpu_batch_fetch(self._h, rs.results)
        if rc != MGPU_OK:
            self.index._err(rc)
        return rs

    def stats(self):
        st = c_batch_stats()
        self._lib.mgpu_batch_get_stats(self._h, C.byref(st))
        return {k: (list(getattr(st, k)) if k.startswith("class_") else getattr(st, k)) for k, _ in c_batch_stats._fields_}

    def export_keys(self, dev_keys_ptr, dev_counts_ptr, dev_total_ptr, k):
        rc = self._lib.mgpu_batch_export_keys(self._h, dev_keys_ptr, dev_counts_ptr, dev_total_ptr, k)
        if rc != MGPU_OK:
            self.index._err(rc)

    def free(self):
        if self._h:
            self._lib.mgpu_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# index building (host only)
# ---------------------------------------------------------------------------------------------

def build_index(path_prefix, field_names, docs, attr_names=(), skiplist_block=32, hit_format_inline=True, dict_crc=False):
    """docs: list of dicts {"id": int, "fields": [[(keyword, pos), ...] per field], "attrs": [uint32...]}.

    Row order = list order (the reference assigns rowids in source order)."""
    kw_index, keywords = {}, []
    offsets, tok_kw, tok_pos = [0], [], []
    for d in docs:
        for f in range(len(field_names)):
            for (w, p) in d["fields"][f]:
                if w not in kw_index:
                    kw_index[w] = len(keywords)
                    keywords.append(w)
                tok_kw.append(kw_index[w])
                tok_pos.append(p)
            offsets.append(len(tok_kw))
    inp = c_build_doc_input()
    n = len(docs)
    a_fields = (C.c_char_p * len(field_names))(*[f.encode() for f in field_names])
    a_attrs = (C.c_char_p * max(1, len(attr_names)))(*[a.encode() for a in attr_names])
    a_ids = (C.c_int64 * max(1, n))(*[d["id"] for d in docs])
    flat_attrs = [v for d in docs for v in d.get("attrs", [])]
    a_attrv = (C.c_uint32 * max(1, len(flat_attrs)))(*flat_attrs)
    a_kw = (C.c_char_p * max(1, len(keywords)))(*[k.encode("utf-8") for k in keywords])
    a_off = (C.c_int64 * len(offsets))(*offsets)
    a_tk = (C.c_int32 * max(1, len(tok_kw)))(*tok_kw)
    a_tp = (C.c_int32 * max(1, len(tok_pos)))(*tok_pos)
    inp.n_docs, inp.n_fields, inp.field_names = n, len(field_names), a_fields
    inp.n_attrs, inp.attr_names = len(attr_names), a_attrs
    inp.docids, inp.attrs = a_ids, a_attrv
    inp.n_keywords, inp.keywords = len(keywords), a_kw
    inp.field_tok_offsets, inp.tok_keyword, inp.tok_pos = a_off, a_tk, a_tp
    inp.skiplist_block, inp.hit_format_inline = skiplist_block, int(hit_format_inline)
    inp.dict_crc = int(dict_crc)
    err = C.create_string_buffer(512)
    rc = writer_lib().mgpu_build_index(path_prefix.encode(), C.byref(inp), err, 512)
    if rc != MGPU_OK:
        raise MgpuError(rc, err.value.decode())


def build_synthetic(path_prefix, params):
    err = C.create_string_buffer(512)
    rc = writer_lib().mgpu_build_synthetic(path_prefix.encode(), C.byref(params), err, 512)
    if rc != MGPU_OK:
        raise MgpuError(rc, err.value.decode())


def synth_field_len(params, doc, field):
    return writer_lib().mgpu_synth_field_len(C.byref(params), doc, field)


def synth_token(params, doc, field, pos0):
    return writer_lib().mgpu_synth_token(C.byref(params), doc, field, pos0)


def synth_keyword(term):
    """keyword string of 0-based term id (0 = most frequent)"""
    return "t%07d" % (term + 1)

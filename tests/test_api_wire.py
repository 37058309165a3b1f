"""Wire responder: mgpu_api_handle answers the binary SphinxAPI `search` command.

The request packets are the bytes the reference's own client (api/sphinxapi.py) puts on the wire (tests/golden/api_requests.json,
made by tests/golden/make_api_fixtures.py). CPU tests: packets parse, are described, protocol errors are answered as searchd answers
them. GPU test: the replies carry exactly what the same queries give through mgpu_search_batch, plus the rows' attributes. When
/root/reference is present, the replies captured on the GPU (tests/golden/api_replies.json) are read back by the reference's client."""
import json
import os
import struct

import pytest

import helpers
import manticoresearch_b200.mgpu as M

GOLD = os.path.join(helpers.ROOT, "tests", "golden")
REQUESTS = {k: bytes.fromhex(v) for k, v in json.load(open(os.path.join(GOLD, "api_requests.json")))["requests"].items()}

SEARCHD_OK, SEARCHD_ERROR, SEARCHD_WARNING = 0, 1, 3
WORDS = ["hello", "world", "there", "extra", "filler"]


def corpus_docs():
    """60 documents, fields title / body, attributes group_id / stamp; deterministic"""
    docs = []
    for i in range(60):
        title = [(WORDS[(i + k) % 5], k + 1) for k in range(1 + i % 3)]
        body = [(WORDS[(i * 7 + k * 3) % 5], k + 1) for k in range(2 + i % 5)]
        docs.append({"id": 2 + i, "fields": [title, body], "attrs": [i % 4, 100 + (i * 13) % 90]})
    return docs


def build_corpus(prefix):
    M.build_index(prefix, ["title", "body"], corpus_docs(), attr_names=["group_id", "stamp"])


class Reader:
    def __init__(self, raw):
        self.raw, self.p = raw, 0

    def u32(self):
        v = struct.unpack_from(">L", self.raw, self.p)[0]
        self.p += 4
        return v

    def u64(self):
        v = struct.unpack_from(">Q", self.raw, self.p)[0]
        self.p += 8
        return v

    def string(self):
        n = self.u32()
        s = self.raw[self.p:self.p + n].decode("utf-8")
        self.p += n
        return s


def parse_reply(raw, n_queries):
    """the reply packet as the protocol defines it (SendResult, src/searchd.cpp:3398-3510) -> (status, [result dict])"""
    status, ver, length = struct.unpack_from(">2HL", raw, 0)
    assert length == len(raw) - 8
    r = Reader(raw[8:])
    if status != SEARCHD_OK:
        return status, r.string()
    assert ver == 0x121
    out = []
    for _ in range(n_queries):
        res = {"status": r.u32(), "error": "", "warning": ""}
        out.append(res)
        if res["status"] != SEARCHD_OK:
            msg = r.string()
            if res["status"] == SEARCHD_WARNING:
                res["warning"] = msg
            else:
                res["error"] = msg
                continue
        res["fields"] = [r.string() for _ in range(r.u32())]
        res["attrs"] = [(r.string(), r.u32()) for _ in range(r.u32())]
        count, id64 = r.u32(), r.u32()
        assert id64 == 1
        res["matches"] = []
        for _ in range(count):
            m = {"id": r.u64(), "weight": r.u32(), "attrs": {}}
            for name, typ in res["attrs"]:
                m["attrs"][name] = r.u64() if typ == 6 else r.u32()
            res["matches"].append(m)
        res["total"], res["total_found"], res["time_msec"] = r.u32(), r.u32(), r.u32()
        res["words"] = {}
        for _ in range(r.u32()):
            w = r.string()
            res["words"][w] = (r.u32(), r.u32())
    assert r.p == len(r.raw)
    return status, out


SEARCH_NAMES = ["default", "any_attr_desc", "extended_sort_filter_weights", "phrase_range_idrange", "multi", "positional_weights"]


def parse_keywords_reply(raw, stats, ver=0x100):
    status, rver, length = struct.unpack_from(">2HL", raw, 0)
    assert length == len(raw) - 8
    r = Reader(raw[8:])
    if status != SEARCHD_OK:
        return status, r.string()
    out = []
    for _ in range(r.u32()):
        e = {"tokenized": r.string(), "normalized": r.string()}
        if ver >= 0x101:
            e["qpos"] = r.u32()
        if stats:
            e["docs"], e["hits"] = r.u32(), r.u32()
        out.append(e)
    assert r.p == len(r.raw)
    return status, out


N_QUERIES = {"default": 1, "any_attr_desc": 1, "extended_sort_filter_weights": 1, "phrase_range_idrange": 1, "multi": 4, "positional_weights": 1}


def test_requests_of_the_reference_client_parse(tmp_path):
    prefix = str(tmp_path / "api")
    build_corpus(prefix)
    api = M.ApiResponder(None, prefix)
    try:
        expect = {
            "default": ["MATCH('hello world')", "FROM idx", "ORDER BY relevance", "LIMIT 0,20", "mode=extended2", "ranker=0", "max_matches=1000"],
            "any_attr_desc": ["MATCH('hello there')", "ORDER BY attr_desc(group_id)", "LIMIT 1,3", "mode=any", "max_matches=50", "/* a comment */"],
            "extended_sort_filter_weights": ["group_id IN (1,3)", "ORDER BY extended(@weight desc, group_id asc)", "ranker=1", "field_weights=(title=5)"],
            "phrase_range_idrange": ["stamp NOT BETWEEN 100 AND 160", "id BETWEEN 3 AND 40", "mode=phrase", "LIMIT 0,100"],
            "multi": ["MATCH('\"hello world\"~3 | extra')", "GROUP BY group_id", "MATCH('hello | (world')", "ORDER BY attr_asc(stamp)", "ranker=3"],
            "positional_weights": ["weights=(3,1)", "MATCH('world there')"],
        }
        for name in SEARCH_NAMES:
            req = REQUESTS[name]
            raw = api.handle(req)
            desc = api.describe_last()
            assert desc.count(";\n") == N_QUERIES[name]
            for frag in expect[name]:
                assert frag in desc, (name, frag, desc)
            status, results = parse_reply(raw, N_QUERIES[name])
            assert status == SEARCHD_OK
            errors = [r["error"] for r in results]
            assert all(r["status"] == SEARCHD_ERROR for r in results)
            if name == "multi":
                assert "group-by is not supported" in errors[1] and "query error: syntax error" in errors[2]
                assert "no index is attached" in errors[0] and "no index is attached" in errors[3]
            else:
                assert "no index is attached" in errors[0]
    finally:
        api.close()


def test_keywords_command(tmp_path):
    """SEARCHD_COMMAND_KEYWORDS (HandleCommandKeywords): the index tokenizer's view of a text; without an index only the no-stats form"""
    prefix = str(tmp_path / "api")
    build_corpus(prefix)
    api = M.ApiResponder(None, prefix)
    try:
        status, words = parse_keywords_reply(api.handle(REQUESTS["keywords_plain"]), False)
        assert status == SEARCHD_OK
        assert [w["tokenized"] for w in words] == ["hello", "world", "there", "extra", "nosuchword"] and all(w["normalized"] == w["tokenized"] for w in words)
        assert "CALL KEYWORDS('Hello, WORLD there-extra  nosuchword', 'idx', 0)" in api.describe_last()
        status, msg = parse_keywords_reply(api.handle(REQUESTS["keywords_stats"]), True)
        assert status == SEARCHD_ERROR and "no index" in msg
        # protocol 1.1 adds four flags to the request and the in-query position to every keyword
        body = REQUESTS["keywords_plain"][8:] + struct.pack(">4L", 0, 0, 0, 0)
        status, words = parse_keywords_reply(api.handle(struct.pack(">2HL", 3, 0x101, len(body)) + body), False, ver=0x101)
        assert [w["qpos"] for w in words] == [1, 2, 3, 4, 5]
        status, msg = parse_keywords_reply(api.handle(struct.pack(">2HL", 3, 0x102, len(body)) + body), False)
        assert status == SEARCHD_ERROR and "client version is higher" in msg
    finally:
        api.close()
    api = M.ApiResponder(None, prefix, min_word_len=3, stopwords=("there",))
    try:
        body = REQUESTS["keywords_plain"][8:] + struct.pack(">4L", 0, 0, 0, 0)
        _, words = parse_keywords_reply(api.handle(struct.pack(">2HL", 3, 0x101, len(body)) + body), False, ver=0x101)
        assert [(w["tokenized"], w["qpos"]) for w in words] == [("hello", 1), ("world", 2), ("extra", 4), ("nosuchword", 5)]
    finally:
        api.close()


def test_protocol_errors(tmp_path):
    prefix = str(tmp_path / "api")
    build_corpus(prefix)
    api = M.ApiResponder(None, prefix)
    try:
        good = REQUESTS["default"]

        def err(packet):
            status, msg = parse_reply(api.handle(packet), 0)
            assert status == SEARCHD_ERROR
            return msg

        assert "truncated" in err(good[:-3])
        assert "truncated" in err(good[:40])
        assert "truncated" in err(b"")
        assert "unknown command" in err(struct.pack(">2HL", 7, 0x120, len(good) - 8) + good[8:])
        assert "client version is higher" in err(struct.pack(">2HL", 0, 0x125, len(good) - 8) + good[8:])
        assert "major command version mismatch" in err(struct.pack(">2HL", 0, 0x220, len(good) - 8) + good[8:])
        assert "bad multi-query count" in err(good[:12] + struct.pack(">L", 0) + good[16:])
        assert "bad multi-query count" in err(good[:12] + struct.pack(">L", 1000) + good[16:])
        assert "truncated" in err(good[:12] + struct.pack(">L", 2) + good[16:])      # says two queries, carries one
        # every prefix of every packet is answered, never crashes
        for name in SEARCH_NAMES:
            req = REQUESTS[name]
            for cut in range(0, len(req), 7):
                status, _ = parse_reply(api.handle(struct.pack(">2HL", 0, 0x120, max(cut - 8, 0)) + req[8:cut]), 0)
                assert status == SEARCHD_ERROR
        for name in ("keywords_plain", "keywords_stats"):
            req = REQUESTS[name]
            for cut in range(8, len(req)):
                status, _ = parse_keywords_reply(api.handle(struct.pack(">2HL", 3, 0x100, cut - 8) + req[8:cut]), False)
                assert status == SEARCHD_ERROR
    finally:
        api.close()


def expected_queries(gid, stamp):
    """the mgpu queries the packets stand for, built by hand (not through the responder)"""
    A = lambda *ws: M.AND(*[M.kw(w, i + 1) for i, w in enumerate(ws)])
    return {
        "default": [(M.Query(A("hello", "world"), max_matches=1000), 0, 20)],
        "any_attr_desc": [(M.Query(M.OR(M.kw("hello", 1), M.kw("there", 2)), ranker=M.RANK_MATCHANY, max_matches=50,
                                   sort_keys=[M.SortKey(M.KEYPART_INT, gid, True), M.SortKey(M.KEYPART_WEIGHT, 0, True)]), 1, 3)],
        "extended_sort_filter_weights": [(M.Query(M.OR(M.kw("hello", 1), M.kw("world", 2), M.kw("there", 3)), ranker=M.RANK_BM25, max_matches=1000,
                                                  field_weights=[5, 1], filters=[M.Filter(gid, values=[1, 3])],
                                                  sort_keys=[M.SortKey(M.KEYPART_WEIGHT, 0, True), M.SortKey(M.KEYPART_INT, gid, False)]), 0, 20)],
        "phrase_range_idrange": [(M.Query(M.PHRASE([("hello", 1), ("world", 2)]), ranker=M.RANK_PROXIMITY, max_matches=100,
                                          filters=[M.Filter(stamp, 100, 160, exclude=True), M.Filter(0, 3, 40)]), 0, 100)],
        "multi": [(M.Query(M.OR(M.PROXIMITY([("hello", 1), ("world", 2)], 3), M.kw("extra", 4)), ranker=M.RANK_WORDCOUNT, max_matches=1000), 0, 20), None, None,
                  (M.Query(M.ANDNOT(M.kw("hello", 1).fields(1), M.kw("there", 2).fields(1)), ranker=M.RANK_WORDCOUNT, max_matches=1000,
                           sort_keys=[M.SortKey(M.KEYPART_INT, stamp, False), M.SortKey(M.KEYPART_WEIGHT, 0, True)]), 0, 20)],
        "positional_weights": [(M.Query(A("world", "there"), max_matches=1000, field_weights=[3, 1]), 0, 1000)],
    }


@pytest.mark.gpu
def test_replies_equal_direct_search(tmp_path):
    prefix = str(tmp_path / "api")
    build_corpus(prefix)
    docs = {d["id"]: d for d in corpus_docs()}
    gpu = M.Index(prefix, device=0)
    api = M.ApiResponder(gpu, prefix)
    replies = {}
    try:
        gid, stamp = gpu.attr_index("group_id"), gpu.attr_index("stamp")
        exp = expected_queries(gid, stamp)
        nonempty = 0
        raw = api.handle(REQUESTS["keywords_stats"])
        replies["keywords_stats"] = raw.hex()
        status, words = parse_keywords_reply(raw, True)
        assert status == SEARCHD_OK and [w["tokenized"] for w in words] == ["hello", "world", "hello", "zzz"]
        for w in words:
            assert (w["docs"], w["hits"]) == (gpu.word_stats(w["tokenized"]) or (0, 0))
        assert words[0]["docs"] > 0 and words[3]["docs"] == 0
        for name in SEARCH_NAMES:
            req = REQUESTS[name]
            raw = api.handle(req)
            replies[name] = raw.hex()
            status, results = parse_reply(raw, N_QUERIES[name])
            assert status == SEARCHD_OK
            for qi, res in enumerate(results):
                if exp[name][qi] is None:
                    assert res["status"] == SEARCHD_ERROR and res["error"]
                    continue
                query, offset, limit = exp[name][qi]
                d = gpu.search([query]).get(0)
                assert d["status"] == 0
                assert res["status"] == SEARCHD_OK, (name, qi, res["error"])
                assert res["fields"] == ["title", "body"] and res["attrs"] == [("group_id", 1), ("stamp", 1)]
                want = list(zip(d["docid"], d["weight"]))[offset:offset + limit]
                assert [(m["id"], m["weight"]) for m in res["matches"]] == want, (name, qi)
                for m in res["matches"]:
                    assert m["attrs"] == {"group_id": docs[m["id"]]["attrs"][0], "stamp": docs[m["id"]]["attrs"][1]}
                assert res["total"] == len(d["docid"]) and res["total_found"] == d["total_found"]
                kws = query.keywords()
                stats = gpu.search([query]).word_stats(0, len(kws))
                assert res["words"] == {k.word: tuple(s) for k, s in zip(kws, stats)}
                nonempty += len(want) > 0
        assert nonempty >= 6
        out_dir = os.path.join(helpers.ROOT, "gpurun_out")
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, "api_replies.json"), "w") as f:
                json.dump({"source": "mgpu_api_handle on the B200, tests/test_api_wire.py::test_replies_equal_direct_search", "replies": replies}, f, indent=1)
    finally:
        api.close()
        gpu.close()


@pytest.mark.skipif(not os.path.exists("/root/reference/api/sphinxapi.py") or not os.path.exists(os.path.join(GOLD, "api_replies.json")),
                    reason="needs the reference's client (this container) and the replies captured on the GPU")
def test_reference_client_reads_the_replies():
    """the reference's client sends its request into a fake socket and parses the reply packet the responder produced on the GPU"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_api_fixtures", os.path.join(GOLD, "make_api_fixtures.py"))
    F = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(F)
    S = F.load_client()
    replies = {k: bytes.fromhex(v) for k, v in json.load(open(os.path.join(GOLD, "api_replies.json")))["replies"].items()}
    docs = {d["id"]: d for d in corpus_docs()}
    sent, words, error = F.run_keywords(S, "keywords_stats", replies["keywords_stats"])
    assert sent == REQUESTS["keywords_stats"] and words is not None, error
    _, mine = parse_keywords_reply(replies["keywords_stats"], True)
    assert words == mine and [w["tokenized"] for w in words] == ["hello", "world", "hello", "zzz"]
    for name, raw in replies.items():
        if name.startswith("keywords"):
            continue
        sent, results, error = F.run_client(S, name, raw)
        assert sent == REQUESTS[name]                   # the fixture is what this client sends
        assert results is not None, error
        _, mine = parse_reply(raw, N_QUERIES[name])
        assert len(results) == len(mine)
        for theirs, ours in zip(results, mine):
            assert theirs["status"] == ours["status"]
            if ours["status"] == SEARCHD_ERROR:
                assert theirs["error"] == ours["error"]
                continue
            assert theirs["fields"] == ours["fields"] and [tuple(a) for a in theirs["attrs"]] == ours["attrs"]
            assert [(m["id"], m["weight"], m["attrs"]) for m in theirs["matches"]] == [(m["id"], m["weight"], m["attrs"]) for m in ours["matches"]]
            assert (theirs["total"], theirs["total_found"]) == (ours["total"], ours["total_found"])
            assert {w["word"]: (w["docs"], w["hits"]) for w in theirs["words"]} == ours["words"]
            for m in theirs["matches"]:
                assert m["attrs"]["group_id"] == docs[m["id"]]["attrs"][0]


@pytest.mark.gpu
def test_sharded_responder_answers_like_the_single_index_one(tmp_path):
    """mgpu_api_create_sharded: the same packets over two rowid-range shards (one mgpu_sharded_search_batch per packet): same matches,
    weights, attribute values (read from the shard that holds the row), totals and whole-index keyword statistics"""
    full = str(tmp_path / "full")
    build_corpus(full)
    docs = corpus_docs()
    prefixes = [str(tmp_path / "s0"), str(tmp_path / "s1")]
    M.build_index(prefixes[0], ["title", "body"], docs[:37], attr_names=["group_id", "stamp"])
    M.build_index(prefixes[1], ["title", "body"], docs[37:], attr_names=["group_id", "stamp"])
    gpu, sh = M.Index(full, device=0), M.ShardedIndex(prefixes, [0, 0])
    one, two = M.ApiResponder(gpu, full), M.ApiResponder(sh, prefixes)
    try:
        for name in SEARCH_NAMES:
            _, a = parse_reply(one.handle(REQUESTS[name]), N_QUERIES[name])
            _, b = parse_reply(two.handle(REQUESTS[name]), N_QUERIES[name])
            assert len(a) == len(b)
            for x, y in zip(a, b):
                x.pop("time_msec", None)
                y.pop("time_msec", None)
                assert x == y, name
            assert any(r["status"] == SEARCHD_OK and r["matches"] for r in b)
        _, wa = parse_keywords_reply(one.handle(REQUESTS["keywords_stats"]), True)
        _, wb = parse_keywords_reply(two.handle(REQUESTS["keywords_stats"]), True)
        assert wa == wb and wb[0]["docs"] > 0
    finally:
        one.close()
        two.close()
        sh.close()
        gpu.close()


def test_mutated_packets_never_crash(tmp_path):
    """byte-level fuzz of the daemon-facing entry point: every mutation of a valid packet is answered with a well-formed reply"""
    import random
    rng = random.Random(4242)
    prefix = str(tmp_path / "api")
    build_corpus(prefix)
    api = M.ApiResponder(None, prefix)
    try:
        names = list(REQUESTS)
        ok = err = 0
        for _ in range(4000):
            name = rng.choice(names)
            raw = bytearray(REQUESTS[name])
            for _ in range(rng.randint(1, 4)):
                how = rng.random()
                pos = rng.randrange(8, len(raw))
                if how < 0.5:
                    raw[pos] = rng.randrange(256)
                elif how < 0.7:
                    raw[pos:pos + 4] = struct.pack(">L", rng.choice([0, 1, 0xFFFFFFFF, 0x7FFFFFFF, 0x80000000, len(raw), 1 << 20]))
                elif how < 0.85:
                    del raw[pos:pos + rng.randint(1, 16)]
                else:
                    raw[pos:pos] = bytes(rng.randrange(256) for _ in range(rng.randint(1, 16)))
            if rng.random() < 0.8:
                raw[4:8] = struct.pack(">L", len(raw) - 8)      # keep the framing right so that the body parser is what gets exercised
            reply = api.handle(bytes(raw))
            status, ver, length = struct.unpack_from(">2HL", reply, 0)
            assert length == len(reply) - 8 and status in (SEARCHD_OK, SEARCHD_ERROR)
            ok += status == SEARCHD_OK
            err += status == SEARCHD_ERROR
        assert ok > 200 and err > 200
    finally:
        api.close()


# ---------------------------------------------------------------------------------------------
# the reference's own C client (api/libsphinxclient, protocol 1.30) over a loopback socket: oracle/_ref/refclient is compiled from the
# reference's sources by oracle/Makefile (this container) and travels to the GPU box as a binary
# ---------------------------------------------------------------------------------------------
REFCLIENT = os.path.join(helpers.ROOT, "oracle", "_ref", "refclient")


def run_refclient(api, scenario, n_connections=1):
    """serves n_connections on 127.0.0.1 the way searchd does (handshake, one command packet per connection -> mgpu_api_handle) while
    the reference client runs `scenario`; -> (its stdout lines, the request packets it sent)"""
    import socket
    import subprocess
    import threading
    srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    srv.bind(("127.0.0.1", 0))
    srv.listen(4)
    srv.settimeout(20)
    port = srv.getsockname()[1]
    packets = []

    def recv_all(conn, n):
        buf = b""
        while len(buf) < n:
            chunk = conn.recv(n - len(buf))
            if not chunk:
                break
            buf += chunk
        return buf

    def serve():
        for _ in range(n_connections):
            conn, _ = srv.accept()
            conn.settimeout(20)
            try:
                conn.sendall(struct.pack(">L", 1))          # searchd's protocol version first
                recv_all(conn, 4)                           # the client's
                head = recv_all(conn, 8)
                if len(head) < 8:
                    continue
                body = recv_all(conn, struct.unpack(">L", head[4:8])[0])
                packets.append(head + body)
                conn.sendall(api.handle(head + body))
            finally:
                conn.close()

    t = threading.Thread(target=serve, daemon=True)
    t.start()
    out = subprocess.run([REFCLIENT, str(port), scenario], capture_output=True, text=True, timeout=60)
    t.join(20)
    srv.close()
    assert out.returncode == 0, out.stderr
    return out.stdout.splitlines(), packets


@pytest.mark.skipif(not os.path.exists(REFCLIENT), reason="oracle/_ref/refclient is built where /root/reference is present")
def test_reference_c_client_talks_to_the_responder(tmp_path):
    """protocol 1.30 requests of the reference's C client parse (its packets differ from the Python client's 1.32 ones: no token-filter
    fields), and the responder's error replies are what that client reports"""
    prefix = str(tmp_path / "api")
    build_corpus(prefix)
    api = M.ApiResponder(None, prefix)
    try:
        lines, packets = run_refclient(api, "extended_sort_filter_weights")
        assert struct.unpack_from(">2H", packets[0], 0) == (0, 0x11E)
        desc = api.describe_last()
        for frag in ("MATCH('hello | world | there')", "group_id IN (1,3)", "ORDER BY extended(@weight desc, group_id asc)", "ranker=1", "field_weights=(title=5)"):
            assert frag in desc, (frag, desc)
        assert any("no index is attached" in l for l in lines)      # (sphinx_query() hands a failed single query back as the client's error)
        lines, packets = run_refclient(api, "multi")
        assert api.describe_last().count(";\n") == 4
        assert [l for l in lines if l.startswith("query ")] == ["query 0", "query 1", "query 2", "query 3"]
        assert any("group-by is not supported" in l for l in lines) and any("syntax error" in l for l in lines)
        lines, packets = run_refclient(api, "keywords_stats")
        assert struct.unpack_from(">2H", packets[0], 0) == (3, 0x100) and any("no index is attached" in l for l in lines)
    finally:
        api.close()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REFCLIENT), reason="oracle/_ref/refclient is built where /root/reference is present and travels with the snapshot")
def test_reference_c_client_end_to_end_on_gpu(tmp_path):
    """the UNMODIFIED reference client, compiled from its own sources, queries the B200 through the responder over a socket: what it
    prints equals a direct mgpu_search_batch of the same queries"""
    prefix = str(tmp_path / "api")
    build_corpus(prefix)
    docs = {d["id"]: d for d in corpus_docs()}
    gpu = M.Index(prefix, device=0)
    api = M.ApiResponder(gpu, prefix)
    try:
        gid, stamp = gpu.attr_index("group_id"), gpu.attr_index("stamp")
        exp = expected_queries(gid, stamp)
        checked = 0
        for name in ("default", "any_attr_desc", "extended_sort_filter_weights", "phrase_range_idrange", "multi"):
            lines, _ = run_refclient(api, name)
            blocks, cur = [], None
            for l in lines:
                if l.startswith("status "):
                    cur = {"status": int(l.split()[1]), "matches": [], "words": {}}
                    blocks.append(cur)
                elif l.startswith("match "):
                    f = l.split()
                    cur["matches"].append((int(f[1]), int(f[2]), int(f[3]), int(f[4])))
                elif l.startswith("total "):
                    f = l.split()
                    cur["total"], cur["total_found"] = int(f[1]), int(f[3])
                elif l.startswith("word "):
                    f = l.split()
                    cur["words"][f[1]] = (int(f[2]), int(f[3]))
            assert len(blocks) == N_QUERIES[name], lines
            for qi, b in enumerate(blocks):
                if exp[name][qi] is None:
                    assert b["status"] == SEARCHD_ERROR
                    continue
                query, offset, limit = exp[name][qi]
                d = gpu.search([query]).get(0)
                assert b["status"] == SEARCHD_OK, (name, qi, lines)
                want = [(i, w, docs[i]["attrs"][0], docs[i]["attrs"][1]) for i, w in list(zip(d["docid"], d["weight"]))[offset:offset + limit]]
                assert b["matches"] == want, (name, qi)
                assert (b["total"], b["total_found"]) == (len(d["docid"]), d["total_found"])
                checked += len(want) > 0
        assert checked >= 5
        lines, _ = run_refclient(api, "keywords_stats")
        kws = [l.split() for l in lines if l.startswith("keyword ")]
        assert [k[1] for k in kws] == ["hello", "world", "hello", "zzz"]
        assert (int(kws[0][3]), int(kws[0][4])) == gpu.word_stats("hello") and (int(kws[3][3]), int(kws[3][4])) == (0, 0)
    finally:
        api.close()
        gpu.close()


def _smoke_case():
    return next(c for c in helpers.load_golden() if c["name"] == "libsphinxclient_smoke")


def _parse_refclient(lines):
    blocks, cur, kws = [], None, []
    for l in lines:
        f = l.split()
        if l.startswith("status "):
            cur = {"status": int(f[1]), "matches": [], "words": {}}
            blocks.append(cur)
        elif l.startswith("match "):
            cur["matches"].append([int(f[1]), int(f[2])])
        elif l.startswith("total "):
            cur["total"], cur["total_found"] = int(f[1]), int(f[3])
        elif l.startswith("word "):
            cur["words"][f[1]] = [int(f[2]), int(f[3])]
        elif l.startswith("keyword "):
            kws.append((f[1], f[2], int(f[3]), int(f[4])))
    return blocks, kws


@pytest.mark.skipif(not os.path.exists(REFCLIENT), reason="oracle/_ref/refclient is built where /root/reference is present")
def test_c_client_smoke_requests_parse(tmp_path):
    prefix = str(tmp_path / "smoke")
    helpers.build_golden_index(_smoke_case(), prefix)
    api = M.ApiResponder(None, prefix)
    try:
        lines, packets = run_refclient(api, "smoke", n_connections=5)
        assert len(packets) == 5 and struct.unpack_from(">2H", packets[0], 0) == (3, 0x100)
        assert "group_id IN (1)" in api.describe_last() and "field_weights=(title=100,content=1)" in api.describe_last()
    finally:
        api.close()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REFCLIENT), reason="oracle/_ref/refclient is built where /root/reference is present and travels with the snapshot")
def test_c_client_smoke_test_against_its_reference_output(tmp_path):
    """api/libsphinxclient/smoke_ref.txt: what the C client's own test program printed against a real searchd. The same client, the same
    calls, the same data, against the responder on the GPU: keyword statistics of build_keywords, matches with weights, totals and
    per-query keyword statistics of the three queries and of the filtered one"""
    case = _smoke_case()
    prefix = str(tmp_path / "smoke")
    helpers.build_golden_index(case, prefix)
    gpu = M.Index(prefix, device=0)
    api = M.ApiResponder(gpu, prefix)
    try:
        lines, _ = run_refclient(api, "smoke", n_connections=5)
        blocks, kws = _parse_refclient(lines)
        assert kws == [("hello", "hello", 0, 0), ("test", "test", 3, 5), ("one", "one", 1, 2)]      # smoke_ref.txt, test_keywords
        assert len(blocks) == 4
        for b, q in zip(blocks, case["queries"]):
            assert b["status"] == SEARCHD_OK
            assert b["matches"] == q["expect"]["matches"], q["text"]
            assert b["total"] == len(q["expect"]["matches"]) and b["total_found"] == q["expect"]["total_found"]
            assert b["words"] == q["expect"]["words"], q["text"]
    finally:
        api.close()
        gpu.close()

#!/usr/bin/env python3
"""Generates tests/golden/api_requests.json: SEARCHD_COMMAND_SEARCH request packets built by the reference's OWN client
(/root/reference/api/sphinxapi.py, protocol 1.32), captured from a fake socket. This container only; the JSON is committed.

Also used by tests/test_api_wire.py (when /root/reference is present) to parse the responder's reply packets with the reference's
client: the same client that built the request reads the answer."""
import importlib.util
import json
import os
import sys
from struct import pack

HERE = os.path.dirname(os.path.abspath(__file__))
CLIENT = "/root/reference/api/sphinxapi.py"


def load_client():
    spec = importlib.util.spec_from_file_location("ref_sphinxapi", CLIENT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class FakeSocket:
    def __init__(self, reply):
        self.sent = bytearray()
        self.reply = bytearray(reply)

    def send(self, data):
        self.sent += data
        return len(data)

    def recv(self, n):
        out, self.reply = self.reply[:n], self.reply[n:]
        return bytes(out)

    def close(self):
        pass


ERROR_REPLY = pack(">2HL", 1, 0, 4 + 2) + pack(">L", 2) + b"no"


def scenarios(S):
    """name -> function configuring a fresh client and adding its queries"""
    def s_default(c):
        c.AddQuery("hello world", "idx")

    def s_any_attr_desc(c):
        c.SetMatchMode(S.SPH_MATCH_ANY)
        c.SetSortMode(S.SPH_SORT_ATTR_DESC, "group_id")
        c.SetLimits(1, 3, 50)
        c.AddQuery("hello there", "idx", "a comment")

    def s_extended_sort_filter_weights(c):
        c.SetMatchMode(S.SPH_MATCH_EXTENDED2)
        c.SetSortMode(S.SPH_SORT_EXTENDED, "@weight DESC, group_id ASC")
        c.SetFilter("group_id", [3, 1])
        c.SetFieldWeights({"title": 5})
        c.SetRankingMode(S.SPH_RANK_BM25)
        c.AddQuery("hello | world | there", "idx")

    def s_phrase_range_idrange(c):
        c.SetMatchMode(S.SPH_MATCH_PHRASE)
        c.SetFilterRange("stamp", 100, 160, exclude=1)
        c.SetIDRange(3, 40)
        c.SetLimits(0, 100, 100)
        c.AddQuery("hello world", "idx")

    def s_multi(c):
        c.SetMatchMode(S.SPH_MATCH_EXTENDED2)
        c.SetRankingMode(S.SPH_RANK_WORDCOUNT)
        c.AddQuery('"hello world"~3 | extra', "idx")
        c.SetGroupBy("group_id", S.SPH_GROUPBY_ATTR)
        c.AddQuery("hello", "idx")
        c.ResetGroupBy()
        c.AddQuery("hello | (world", "idx")
        c.SetSortMode(S.SPH_SORT_ATTR_ASC, "stamp")
        c.AddQuery("@title hello -there", "idx")

    def s_positional_weights(c):
        c.SetMatchMode(S.SPH_MATCH_EXTENDED2)
        c._weights = [3, 1]      # (this client has no SetWeights(); the legacy positional list is still on the wire)
        c.SetRankingMode(S.SPH_RANK_PROXIMITY_BM25)
        c.SetLimits(0, 1000, 1000)
        c.AddQuery("world there", "idx")

    return {"default": s_default, "any_attr_desc": s_any_attr_desc, "extended_sort_filter_weights": s_extended_sort_filter_weights,
            "phrase_range_idrange": s_phrase_range_idrange, "multi": s_multi, "positional_weights": s_positional_weights}


def run_client(S, name, reply):
    """-> (request bytes the client sent, what its RunQueries() parsed out of `reply`, its last error)"""
    c = S.SphinxClient()
    scenarios(S)[name](c)
    sock = FakeSocket(reply)
    c._Connect = lambda: sock
    res = c.RunQueries()
    return bytes(sock.sent), res, c.GetLastError()


KEYWORDS = {"keywords_plain": ("Hello, WORLD there-extra  nosuchword", "idx", 0), "keywords_stats": ("hello world hello zzz", "idx", 1)}


def run_keywords(S, name, reply):
    """-> (request bytes of BuildKeywords, what the client parsed out of `reply`, its last error)"""
    c = S.SphinxClient()
    sock = FakeSocket(reply)
    c._Connect = lambda: sock
    res = c.BuildKeywords(*KEYWORDS[name])
    return bytes(sock.sent), res, c.GetLastError()


def main():
    S = load_client()
    out = {"source": "requests built by ravelry/manticoresearch api/sphinxapi.py (VER_COMMAND_SEARCH 0x120)", "requests": {}}
    for name in scenarios(S):
        req, _, _ = run_client(S, name, ERROR_REPLY)
        out["requests"][name] = req.hex()
    for name in KEYWORDS:
        req, _, _ = run_keywords(S, name, ERROR_REPLY)
        out["requests"][name] = req.hex()
    with open(os.path.join(HERE, "api_requests.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote %d request packets" % len(out["requests"]))


if __name__ == "__main__":
    sys.exit(main())

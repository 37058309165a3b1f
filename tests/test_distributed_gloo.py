"""World-size-2 test of the distributed-local search logic on CPU (gloo): rowid-range shards, global IDF inputs through
all_reduce, 128-bit match keys with global rowids, all_gather + merge == the unsharded result.

The per-shard searcher here is the CPU oracle (this is a test of the host-side sharding / exchange logic of
manticoresearch_b200/distributed.py, which is the same code bench.py runs over NCCL with the CUDA searcher)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL_DOCS = 6000
K = 50


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helpers
    import manticoresearch_b200.mgpu as M
    from manticoresearch_b200 import distributed as D
    from manticoresearch_b200 import workload

    first, n = D.shard_range(TOTAL_DOCS, rank, world)
    prefix = os.path.join(tmp, "shard%d" % rank)
    M.build_synthetic(prefix, M.SynthParams(n, first_doc=first, vocab=1 << 14, threads=2))
    shard = helpers.OracleIndex(prefix)

    def make_queries():
        qs = workload.cfg2_queries(n=60, max_rank=3000, max_matches=K, with_andnot=0.2)
        qs += workload.cfg1_queries(n=20, max_matches=K)
        # multi-keyword ANDs over keywords of nearly equal df: a shard's local counts order them differently than the whole index
        # does; the order (and with it the fp32 TF*IDF summation) must follow the global df (mgpu_query::shard_of_global)
        mids = [M.synth_keyword(r) for r in range(30, 70)]
        for i in range(0, 36, 3):
            qs.append(M.Query(M.AND(*[M.kw(w, j + 1) for j, w in enumerate(mids[i:i + 4])]), ranker=M.RANK_BM25, field_weights=[3, 1], max_matches=K))
        return qs

    queries = make_queries()
    gdf = D.global_keyword_docs(lambda w: (shard.word_stats(w) or (0, 0))[0], queries, torch.device("cpu"))
    D.apply_global_idf(queries, TOTAL_DOCS, gdf)
    rs = shard.search(queries)

    nq = len(queries)
    keys = np.zeros((nq, K, 2), dtype=np.uint64)
    counts = torch.zeros((nq,), dtype=torch.int32)
    totals = torch.zeros((nq,), dtype=torch.int64)
    for qi in range(nq):
        r = rs.get(qi)
        assert r["status"] == 0
        counts[qi] = len(r["rowid"])
        totals[qi] = r["total_found"]
        for i, (row, w) in enumerate(zip(r["rowid"], r["weight"])):
            keys[qi, i] = D.pack_key(w, first + row)
    tkeys = torch.from_numpy(keys.view(np.int64))
    all_keys = [torch.zeros_like(tkeys) for _ in range(world)]
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_keys, tkeys)
    dist.all_gather(all_counts, counts)
    dist.all_reduce(totals)

    if rank == 0:
        full_prefix = os.path.join(tmp, "full")
        M.build_synthetic(full_prefix, M.SynthParams(TOTAL_DOCS, vocab=1 << 14, threads=2))
        full = helpers.OracleIndex(full_prefix)
        plain = make_queries()
        ref = full.search(plain)
        bad = []
        for qi in range(nq):
            per_shard = []
            for s in range(world):
                kk = all_keys[s].numpy().view(np.uint64)[qi]
                per_shard.append([(int(kk[i, 0]), int(kk[i, 1])) for i in range(int(all_counts[s][qi]))])
            merged = [D.unpack_key(hi, lo) for hi, lo in D.merge_keys_host(per_shard, K)]
            e = ref.get(qi)
            if [m[0] for m in merged] != e["rowid"] or [m[1] for m in merged] != e["weight"] or int(totals[qi]) != e["total_found"]:
                bad.append(qi)
        ret["bad"] = bad
        ret["checked"] = nq
        ret["nonempty"] = sum(1 for qi in range(nq) if ref.get(qi)["total_found"] > 0)
    dist.barrier()
    dist.destroy_process_group()


def test_two_shards_over_gloo_equal_unsharded(tmp_path):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), ret), nprocs=2, join=True)
    assert ret["checked"] == 92
    assert ret["nonempty"] >= 46
    assert ret["bad"] == [], "sharded != unsharded for queries %s" % ret["bad"]


def test_shard_ranges_cover_the_corpus():
    sys.path.insert(0, ROOT)
    from manticoresearch_b200 import distributed as D
    for total in (1, 7, 1000, 10_000_000):
        for world in (1, 2, 3, 8):
            pos = 0
            for r in range(world):
                first, n = D.shard_range(total, r, world)
                assert first == pos and n >= 0
                pos += n
            assert pos == total


def test_key_pack_orders_like_the_comparator():
    """MatchRelevanceLt_fn (src/sphinxsort.cpp:4534-4548): weight desc, then rowid asc -- incl. negative weights"""
    sys.path.insert(0, ROOT)
    from manticoresearch_b200 import distributed as D
    ms = [(5, 10), (5, 3), (-7, 1), (0, 2), (2147483647, 9), (-2147483648, 0), (5, 4000000000)]
    by_key = sorted(ms, key=lambda m: D.pack_key(*m), reverse=True)
    by_cmp = sorted(ms, key=lambda m: (-m[0], m[1]))
    assert by_key == by_cmp
    for m in ms:
        assert D.unpack_key(*D.pack_key(*m)) == (m[1], m[0])

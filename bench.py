#!/usr/bin/env python3
"""Benchmark of the full-text query hot path (BASELINE.json metric: queries/sec & compressed-postings GB/s).

A "step" = one pass of the hot path over one batch of synthetic queries. The bench line (no flags) is
  --workload cfg2 = BASELINE.json configs[1]: 10M-doc synthetic Zipfian index, 10k-query batch of 2-8 term AND/OR mixes,
                    SPH_RANK_BM25 with field weights (title=10, body=1), top-100.
The other configurations run at the sizes BASELINE.json states, each with its own parity sample, roofline and CPU baseline:
  --workload cfg1  configs[0]: 1M docs, 1 000 two-term AND queries, SPH_RANK_PROXIMITY_BM25, max_matches 1000 (top-20)
  --workload cfg3  configs[2]: 10M docs, 2 000 phrase / proximity queries, PROXIMITY_BM25 (LCS), top-1000
  --workload cfg4  configs[3]: 100M docs over 8 GPUs = 12.5M docs per GPU, the cfg2 mix + 10% ANDNOT, top-1000, local top-K + NCCL merge
  --workload cfg5  configs[4]: 100M docs, 500 stop-word ORs + filter gid BETWEEN 100 AND 299 + ORDER BY ts DESC, top-10000

--gpus N (under torchrun, one rank per GPU): the index is split into N contiguous rowid-range shards.
  value (device-timed, inputs resident): every rank evaluates the whole batch on its shard, K best keys per query are
        all-gathered over NCCL and merged on the GPU (global IDF inputs from the all-reduced per-shard dictionaries);
  e2e:  ONE call of the C ABI's mgpu_sharded_search_batch per step on rank 0 (C++: plan once, a host thread per GPU,
        ncclSend/ncclRecv + shard_merge_kernel, pinned result buffers), host buffers in and out; the other ranks wait.

  python bench.py --gpus 1 --steps 5 --warmup 3            # our arm
  python bench.py --impl reference --gpus 1 ...            # CPU arm: the oracle (the reference cannot be built here)
"""
import argparse
import hashlib
import json
import os
import shutil
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CORPUS = "synthetic Zipfian plain index (seed 0x5EED0001, V=2^20, hit_format=inline, skiplist 32)"
WORKLOADS = {
    "cfg1": {"docs": 1_000_000, "queries": 1000, "K": 1000, "scaling": "strong",
             "desc": "configs[0]: 1M-doc " + CORPUS + ", 1 000 two-term AND queries (rank bands [10,100] x [100,10000]), SPH_RANK_PROXIMITY_BM25, max_matches 1000 (top-20)"},
    "cfg2": {"docs": 10_000_000, "queries": 10_000, "K": 100, "scaling": "strong",
             "desc": "configs[1]: 10M-doc " + CORPUS + ", 10k-query batch of 2-8 term AND/OR/(a b)|(c d) mixes, SPH_RANK_BM25 field_weights=(title=10,body=1), top-100"},
    "cfg3": {"docs": 10_000_000, "queries": 2000, "K": 1000, "scaling": "strong",
             "desc": "configs[2]: 10M-doc " + CORPUS + " with hitlists, 2 000 phrase and proximity (\"a b c\"~5) queries, SPH_RANK_PROXIMITY_BM25 (LCS), top-1000"},
    "cfg4": {"docs": 12_500_000, "queries": 10_000, "K": 1000, "scaling": "weak",
             "desc": "configs[3]: " + CORPUS + " sharded by rowid range, 12.5M docs per GPU (100M over 8), 10k-query batch of 2-8 term AND/OR/(a b)|(c d) mixes with 10% ANDNOT, SPH_RANK_BM25 field_weights=(title=10,body=1), top-1000, local top-K + NCCL merge"},
    "cfg5": {"docs": 100_000_000, "queries": 500, "K": 10000, "scaling": "strong",
             "desc": "configs[4]: 100M-doc " + CORPUS + ", 500 OR queries of 3-6 stop words (ranks 1-50) + filter gid BETWEEN 100 AND 299 + ORDER BY ts DESC, SPH_RANK_BM25, top-10000"},
}
CLASS_NAMES = ["stream_kernel<512>", "eval_kernel<hits>", "and_kernel", "stream_kernel<256>", "and_kernel<hits>", "stream_kernel<512,or>", "stream_kernel<512,dnf>"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--docs", type=int, default=0, help="override the workload's document count (per GPU for cfg4)")
    ap.add_argument("--queries", type=int, default=0, help="override the workload's batch size")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity", type=int, default=48, help="queries of the full-size parity sample (0 = none)")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="engine option for mgpu_index_set_option (A/B switches, e.g. --opt or_bits=0 --opt stats=1)")
    ap.add_argument("--only", default="", choices=["", "and", "or", "mix"], help="analysis only (cfg2/cfg4): keep one query shape of the batch")
    return ap.parse_args()


def workload_config(args, world):
    """the `config` object: identical in both arms (the driver compares them)"""
    w = WORKLOADS[args.workload]
    per_gpu = args.docs or w["docs"]
    docs = per_gpu * world if w["scaling"] == "weak" else per_gpu
    nq = args.queries or w["queries"]
    return {"workload": w["desc"] + (" [ANALYSIS SUBSET: only %s queries]" % args.only if args.only else ""),
            "docs": docs, "queries_per_batch": nq, "max_matches": w["K"], "n_shards": world,
            "l2": "inputs >> L2 (GBs of postings per step against 126 MB)"}


def bench_dir():
    d = os.environ.get("MGPU_BENCH_DIR")
    if not d:
        d = "/tmp/mgpu_bench"
        try:
            st = shutil.disk_usage("/dev/shm")
            if st.free > 24 << 30:
                d = "/dev/shm/mgpu_bench"
        except Exception:
            pass
    os.makedirs(d, exist_ok=True)
    return d


def ensure_index(M, total_docs, shard, n_shards):
    """builds (or reuses) shard `shard` of `n_shards` of the seeded corpus; returns (prefix, first_doc, n_docs, build_seconds)"""
    first = total_docs * shard // n_shards
    n = total_docs * (shard + 1) // n_shards - first
    prefix = os.path.join(bench_dir(), "zipf_%d_%dof%d" % (total_docs, shard, n_shards))
    t0 = time.time()
    if not os.path.exists(prefix + ".ok"):
        M.build_synthetic(prefix, M.SynthParams(n, first_doc=first))
        open(prefix + ".ok", "w").write("ok")
    return prefix, first, n, time.time() - t0


def make_queries(M, workload, args, total_docs, attr_source):
    """the seeded query set of the workload (manticoresearch_b200/workload.py); attr_source: anything with attr_index(name)"""
    w = WORKLOADS[args.workload]
    nq, K = args.queries or w["queries"], w["K"]
    if args.workload == "cfg1":
        return workload.cfg1_queries(n=nq, max_matches=K)
    if args.workload == "cfg3":
        return workload.cfg3_queries(M.SynthParams(total_docs), n=nq, max_matches=K)
    if args.workload == "cfg5":
        return workload.cfg5_queries(attr_source, n=nq, max_matches=K)
    queries = workload.cfg2_queries(n=nq, max_matches=K, with_andnot=0.1 if args.workload == "cfg4" else 0.0)
    if args.only:
        def shape(q):
            r = q.root
            if r.op == M.OP_AND:
                return "and"
            return "mix" if any(c.children for c in r.children) else "or"
        queries = [q for q in queries if shape(q) == args.only]
    return queries


class SchemaOnly:
    """attribute indexes of the synthetic schema (id, gid, ts) for the arm that has no GPU index handle"""
    def attr_index(self, name):
        return {"id": 0, "gid": 1, "ts": 2}[name]


class ClockSampler:
    """samples SM clocks / throttle reasons with nvidia-smi during the timed region"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.samples, self.stop, self.th = gpu_index, [], False, None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernels_sha():
    """identifies the kernel sources an ncu capture belongs to"""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "manticoresearch_b200", "csrc", "cuda")
    for f in sorted(os.listdir(d)):
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic():
    """per-kernel DRAM bytes per launch from the committed ncu --set full capture of this command at N=1
    (profiles/kernel_traffic.json, written by scripts/ncu_traffic.py). The capture is stamped with the hash of the kernel
    sources it was taken on: after any kernel change the figures are void (null) until the capture is redone."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))
    except Exception:
        return {}, "no capture"
    if t.get("kernels_sha") != kernels_sha():
        return {}, "stale: captured on kernels %s, these are %s" % (t.get("kernels_sha"), kernels_sha())
    return t, "ncu --set full, kernels %s" % t.get("kernels_sha")


def oracle_index(prefix):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    return helpers.OracleIndex(prefix)


def cpu_oracle_qps(prefix, queries, seconds, threads):
    """oracle (CPU restatement of the reference path) on a bounded sample: one query per thread, striped"""
    idx = oracle_index(prefix)
    done = [0] * threads
    t_end = time.time() + seconds
    t0 = time.time()

    def work(t):
        i = t
        while i < len(queries) and time.time() < t_end:
            idx.search([queries[i]])
            done[t] += 1
            i += threads

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.time() - t0
    idx.close()
    n = sum(done)
    return n / dt if dt > 0 else 0.0, n, dt


def cpu_baseline(prefix, queries, seconds):
    """all host cores and one core, each on a bounded sample of the same seeded batch"""
    threads = os.cpu_count() or 1
    qps, n, dt = cpu_oracle_qps(prefix, queries, seconds * 0.75, threads)
    qps1, n1, dt1 = cpu_oracle_qps(prefix, queries[len(queries) // 2:] + queries[:len(queries) // 2], seconds * 0.25, 1)
    return {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": "%d queries of the same seeded batch, one query per thread, %d threads, %.1f s" % (n, threads, dt),
            "one_core": {"value": qps1, "unit": "queries/s", "sample": "%d queries, %.1f s" % (n1, dt1)},
            "per_core": qps / threads}


def _same(g, c):
    return (g["status"], g["total_found"], g["rowid"], g["weight"]) == (c["status"], c["total_found"], c["rowid"], c["weight"])


def parity_sample(prefix, queries, get_result, n, what):
    """full-size spot check outside the timed region: n queries of the batch re-run on the CPU oracle (the checker) over the
    UNSHARDED index, compared bit-exactly (rowids, weights, order, total_found) with what the CUDA path returned"""
    idx = oracle_index(prefix)
    pick = list(range(0, len(queries), max(1, len(queries) // n)))[:n]
    bad = []
    lock = threading.Lock()

    def work(t, nt):
        for j in range(t, len(pick), nt):
            qi = pick[j]
            c = idx.search([queries[qi]]).get(0)
            if not _same(get_result(qi), c):
                with lock:
                    bad.append(qi)

    nt = min(16, os.cpu_count() or 1)
    ths = [threading.Thread(target=work, args=(t, nt)) for t in range(nt)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    idx.close()
    return {"checked": len(pick), "mismatches": len(bad), "bad_queries": sorted(bad)[:8], "checker": what}


def parity_sample_sharded_oracle(prefixes, firsts, total_docs, queries, get_result, n, gdf_of):
    """the same check when the unsharded index is too big to build next to the shards (cfg4 at 100M docs): the oracle searches
    every shard with the statistics of the whole index (total docs, global df, keyword order by global df =
    mgpu_query.shard_of_global) and the per-shard results are merged on the host by (weight desc, global rowid asc);
    tests/test_distributed_gloo.py proves this equal to the unsharded oracle"""
    import copy
    shards = [oracle_index(p) for p in prefixes]
    pick = list(range(0, len(queries), max(1, len(queries) // n)))[:n]
    bad = []
    lock = threading.Lock()

    def work(t, nt):
        for j in range(t, len(pick), nt):
            qi = pick[j]
            q = copy.copy(queries[qi])
            q.total_docs = total_docs
            q.word_docs = [gdf_of(k.word) for k in q.keywords()]
            q.shard_of_global = True
            rows, total = [], 0
            for s, first in zip(shards, firsts):
                r = s.search([q]).get(0)
                total += r["total_found"]
                rows += [(-w, first + row) for row, w in zip(r["rowid"], r["weight"])]
            rows.sort()
            rows = rows[:q.max_matches]
            c = {"status": 0, "total_found": total, "rowid": [r[1] for r in rows], "weight": [-r[0] for r in rows]}
            if not _same(get_result(qi), c):
                with lock:
                    bad.append(qi)

    nt = min(16, os.cpu_count() or 1)
    ths = [threading.Thread(target=work, args=(t, nt)) for t in range(nt)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    for s in shards:
        s.close()
    return {"checked": len(pick), "mismatches": len(bad), "bad_queries": sorted(bad)[:8],
            "checker": "oracle/oracle.cpp per shard with global statistics, merged on the host (= the unsharded oracle, tests/test_distributed_gloo.py)"}


def run_reference(args):
    """the CPU arm: the oracle on the host cores (the reference searchd cannot be built here: no bison/flex/boost)"""
    import manticoresearch_b200.mgpu as M
    from manticoresearch_b200 import workload
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_config(args, args.gpus)
    prefix, _, _, _ = ensure_index(M, cfg["docs"], 0, 1)       # index files from libmgpu_writer.so (host-only); libmgpu.so is never loaded
    queries = make_queries(M, workload, args, cfg["docs"], SchemaOnly())
    threads = os.cpu_count() or 1
    qps_runs = []
    budget = max(5.0, min(args.cpu_seconds, 60.0))
    n_done = 0
    for _ in range(args.warmup and 1):
        cpu_oracle_qps(prefix, queries[:threads * 2], 5.0, threads)
    t_all = time.time()
    for s in range(max(1, args.steps)):
        qps, n, dt = cpu_oracle_qps(prefix, queries[s * 997 % max(1, len(queries) - threads * 8):], budget / max(1, args.steps), threads)
        qps_runs.append(qps)
        n_done += n
    dt_all = time.time() - t_all
    value = statistics.mean(qps_runs)
    line = {
        "impl": "reference", "metric": "queries/sec", "value": value, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dt_all / max(1, args.steps), "higher_is_better": True, "scaling": WORKLOADS[args.workload]["scaling"], "vs_baseline": None,
        "dtype": "u32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": "port", "per_core": value / threads,
                         "sample": "%d queries of the same seeded batch, one query per thread, %d threads, %.0f s" % (n_done, threads, dt_all)},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference searchd cannot be built in this image (no bison/flex/boost); this is oracle/oracle.cpp, the CPU restatement of its path",
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import manticoresearch_b200.mgpu as M
    from manticoresearch_b200 import workload
    from manticoresearch_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        # stdout carries rank 0's JSON line only: NCCL's own log (its version banner when NCCL_DEBUG is set) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")     # host-side barriers that keep the GPUs free (the e2e leg runs from rank 0)
    cfg = workload_config(args, world)
    wl = WORKLOADS[args.workload]
    total_docs, K = cfg["docs"], wl["K"]

    # ---- index shard of this rank
    prefix, first_doc, n_docs, build_s = ensure_index(M, total_docs, rank, world)
    t0 = time.time()
    index = M.Index(prefix, device=local_rank, rowid_base=first_doc)
    for kv in args.opt:
        name, _, value = kv.partition("=")
        index.set_option(name, int(value))
    load_s = time.time() - t0
    # one explicit stream for everything: the index's kernels, torch's NCCL collectives and the timing events
    # (torch's default stream has handle 0, which mgpu_index_set_stream reads as "use the index's private stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    index.set_stream(stream.cuda_stream)

    # ---- queries; global IDF inputs when sharded (CSphMultiQueryArgs::m_iTotalDocs / m_pLocalDocs)
    queries = make_queries(M, workload, args, total_docs, index)
    plain_queries = make_queries(M, workload, args, total_docs, index) if world > 1 else queries    # without the global statistics
    gdf = None
    if world > 1:
        gdf = D.global_keyword_docs(lambda w: (index.word_stats(w) or (0, 0))[0], queries, dev)
        D.apply_global_idf(queries, total_docs, gdf)

    nq = len(queries)
    merger = D.ShardMerger(nq, K, dev, local_rank, stream) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: plan uploaded once, K steps of (eval + merge [+ shard merge])
    batch = index.prepare(queries)
    for _ in range(args.warmup):
        batch.run()
        if world > 1:
            merger.merge(batch)
    barrier()
    eval_ms, merge_ms, hot_ms, class_ms = [], [], [], []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            batch.run()
            if world > 1:
                merger.merge(batch)     # asynchronous: export + all_gather + merge are stream-ordered, no host sync per step
            else:
                batch.sync()
                st = batch.stats()
                eval_ms.append(st["eval_kernel_ms"])
                merge_ms.append(st["merge_kernel_ms"])
                hot_ms.append(st["hot_decode_ms"])
                class_ms.append(st["class_ms"])
        ev1.record(stream)
        barrier()
    if world > 1:
        batch.sync()                    # kernel times of the last timed step
        st = batch.stats()
        eval_ms.append(st["eval_kernel_ms"])
        merge_ms.append(st["merge_kernel_ms"])
        hot_ms.append(st["hot_decode_ms"])
        class_ms.append(st["class_ms"])
    total_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    st = batch.stats()
    n_unsupported = 0
    fetched = None
    merged = None
    if world == 1:
        fetched = batch.fetch()
        n_unsupported = sum(1 for i in range(nq) if fetched.results[i].status != 0)
    elif rank == 0:
        merged = merger.fetch()

    # algorithmic bytes / postings of the whole job (all shards)
    agg = torch.tensor([st["algorithmic_bytes"], st["postings"], st["hitlist_bytes"], st["attr_rows"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(agg)
    job_bytes = float(agg[0].item()) + float(agg[2].item()) + 8.0 * float(agg[3].item())
    job_postings = float(agg[1].item())
    batch.free()

    # ---- end to end through the C ABI with host buffers: plan + H2D + kernels + D2H every step
    packed = M.pack_queries(plain_queries)      # the caller's host buffers: flattened XQNode trees in, result arrays out
    host_results = M.ResultSet(plain_queries)
    host_ms = None
    e2e_launches = None
    if world == 1:
        for _ in range(min(args.warmup, 2)):        # untimed: first-call costs of this path (pinned staging, pool growth)
            index.search_packed(packed, nq, host_results)
        torch.cuda.synchronize()
        e2e_t0 = time.perf_counter()
        for _ in range(args.steps):
            rs = index.search_packed(packed, nq, host_results)      # mgpu_search_batch
            assert rs.results[0].status in (0, M.MGPU_E_UNSUPPORTED)
        e2e_s = time.perf_counter() - e2e_t0
        b3 = index.prepare(queries)
        b3.run()
        b3.fetch()
        s3 = b3.stats()
        h2d, d2h = s3["h2d_bytes"], s3["d2h_bytes"]
        b3.free()
        ls = index.last_search_stats()      # of the last timed mgpu_search_batch call
        host_ms = {k: round(ls[k], 2) for k in ("host_total_ms", "host_plan_ms", "host_setup_ms", "host_wait_ms", "host_fetch_ms")}
        e2e_path = "mgpu_search_batch"
        index.close()
    else:
        # rank 0 drives all N GPUs through ONE C-ABI call per step; the other ranks release their GPU and wait on the host
        index.close()
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        e2e_s, h2d, d2h = 0.0, 0, 0
        if rank == 0:
            prefixes = [os.path.join(bench_dir(), "zipf_%d_%dof%d" % (total_docs, r, world)) for r in range(world)]
            sh = M.ShardedIndex(prefixes, list(range(world)))
            for kv in args.opt:
                name, _, value = kv.partition("=")
                sh.set_option(name, int(value))
            for _ in range(min(args.warmup, 2)):
                sh.search_packed(packed, nq, host_results)
            e2e_t0 = time.perf_counter()
            for _ in range(args.steps):
                sh.search_packed(packed, nq, host_results)          # mgpu_sharded_search_batch
            e2e_s = time.perf_counter() - e2e_t0
            ss = sh.stats()
            h2d, d2h = ss["h2d_bytes"], ss["d2h_bytes"]
            host_ms = {k: round(ss[k], 2) for k in ("host_total_ms", "host_plan_ms", "host_setup_ms", "host_wait_ms", "host_fetch_ms", "max_eval_kernel_ms", "max_hot_decode_ms")}
            host_ms["nccl"] = ss["nccl"]
            e2e_launches = ss["kernel_launches"]
            n_unsupported = sum(1 for i in range(nq) if host_results.results[i].status != 0)
            sh.close()
        dist.barrier(group=host_group)
        e2e_path = "mgpu_sharded_search_batch (rank 0: one process, a host thread per GPU)"

    if rank == 0:
        qps = nq / (ms_per_step / 1000.0)
        peak, peak_src = measured_peak()
        eval_avg_ms = statistics.mean(eval_ms)
        names = list(CLASS_NAMES)
        if st["or_kernel"] == 3:
            names[5] = "orbits_kernel"
        NC = len(names)
        cms = [statistics.mean(x[c] for x in class_ms) for c in range(NC)]
        # roofline of the DOMINANT kernel: the launch class with the largest share of the step; its own algorithmic bytes
        # (.spd + .spe extents of every keyword of its queries, + the hitlist bytes of the matched documents for the
        # hit-consuming classes, + 8 attribute bytes per row the filters / sort keys read) over its own CUDA-event duration
        cbytes = list(st["class_bytes"])
        hit_classes = [c for c in (1, 4) if cms[c] > 0]
        for c in hit_classes:
            cbytes[c] += st["hitlist_bytes"] * (cms[c] / sum(cms[x] for x in hit_classes))
        cbytes[6] += 8 * st["attr_rows"]
        dom = max(range(NC), key=lambda c: cms[c])
        achieved = cbytes[dom] / (cms[dom] / 1000.0) / 1e9
        step_bytes = st["algorithmic_bytes"] + st["hitlist_bytes"] + 8 * st["attr_rows"]
        all_achieved = step_bytes / (eval_avg_ms / 1000.0) / 1e9
        traffic, traffic_src = ncu_traffic()
        tkey = {"orbits_kernel": "orbits_kernel"}.get(names[dom], names[dom])
        touched = {k: v.get("dram_bytes_per_launch") for k, v in traffic.items() if isinstance(v, dict)}
        line = {
            "metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": cfg,
            "setup": {"index_build_s": round(build_s, 1), "index_load_s": round(load_s, 1), "unsupported_queries": n_unsupported,
                      "algorithmic_GB_per_step": round(job_bytes / 1e9, 1), "parallelism": "rowid-range shards x%d" % world},
            "postings_per_sec": job_postings / (ms_per_step / 1000.0),
            "compressed_GBps": job_bytes / (ms_per_step / 1000.0) / 1e9,
            "e2e": {"value": nq / (e2e_s / args.steps), "unit": "queries/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1000.0 * e2e_s / args.steps, "path": e2e_path, "host_ms": host_ms},
            "gpu_launches": int(args.steps * (st["kernel_launches"] * world + (world if world > 1 else 0))),
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": touched.get(tkey), "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": cbytes[dom], "kernel_ms": cms[dom],
                         "kernel_share_of_step": cms[dom] / ms_per_step,
                         "dram_frac": (touched[tkey] / (cms[dom] / 1000.0) / 1e9 / peak) if touched.get(tkey) else None,
                         "note": "achieved = SURVEY 8(d) no-skipping algorithmic bytes / kernel time; kernels that skip blocks or read the batch's "
                                 "decoded hot-term store move fewer bytes: bytes_touched (ncu dram bytes per launch) and dram_frac say how many",
                         "all_kernels": {"achieved": all_achieved, "frac": all_achieved / peak, "algorithmic_bytes": step_bytes, "ms": eval_avg_ms,
                                         "hitlist_bytes": st["hitlist_bytes"], "attr_bytes": 8 * st["attr_rows"]},
                         "class_ms": dict(zip(names, cms)), "class_queries": dict(zip(names, st["class_queries"])),
                         "class_GBps": {names[c]: (cbytes[c] / (cms[c] / 1000.0) / 1e9 if cms[c] > 0 else None) for c in range(NC)},
                         "bytes_touched": touched,
                         "merge_kernel_ms": statistics.mean(merge_ms), "hot_decode_ms": statistics.mean(hot_ms), "hot_terms": st["hot_terms"],
                         "frac_of_nominal_8TBs": achieved / 8000.0},
        }
        if e2e_launches is not None:
            line["e2e"]["gpu_launches_per_step"] = e2e_launches
        if args.parity > 0:
            full_prefix = prefix
            if world == 1:
                def get_dev(qi):
                    return fetched.get(qi)
                line["parity_sample"] = parity_sample(full_prefix, plain_queries, get_dev, args.parity,
                                                      "oracle/oracle.cpp at full size (%d docs)" % total_docs)
                line["parity_sample"]["path"] = "mgpu_batch_run + fetch (the device-timed leg)"
                line["parity_sample_e2e"] = parity_sample(full_prefix, plain_queries, lambda qi: host_results.get(qi), min(args.parity, 16),
                                                          "oracle/oracle.cpp at full size (%d docs)" % total_docs)
                line["parity_sample_e2e"]["path"] = e2e_path
            else:
                def get_nccl(qi):
                    m = merged[qi]
                    return {"status": 0, "total_found": m["total_found"], "rowid": [x[0] for x in m["matches"]], "weight": [x[1] for x in m["matches"]]}
                prefixes = [os.path.join(bench_dir(), "zipf_%d_%dof%d" % (total_docs, r, world)) for r in range(world)]
                firsts = [total_docs * r // world for r in range(world)]
                if total_docs <= 25_000_000:
                    full_prefix, _, _, _ = ensure_index(M, total_docs, 0, 1)     # the unsharded index, for the checker only
                    chk = "oracle/oracle.cpp on the UNSHARDED index (%d docs)" % total_docs
                    line["parity_sample"] = parity_sample(full_prefix, plain_queries, get_nccl, args.parity, chk)
                    line["parity_sample_e2e"] = parity_sample(full_prefix, plain_queries, lambda qi: host_results.get(qi), args.parity, chk)
                else:
                    line["parity_sample"] = parity_sample_sharded_oracle(prefixes, firsts, total_docs, plain_queries, get_nccl, args.parity, lambda w: gdf[w])
                    line["parity_sample_e2e"] = parity_sample_sharded_oracle(prefixes, firsts, total_docs, plain_queries, lambda qi: host_results.get(qi),
                                                                             args.parity, lambda w: gdf[w])
                line["parity_sample"]["path"] = "torch.distributed NCCL all_gather + shard_merge_kernel (the device-timed leg)"
                line["parity_sample_e2e"]["path"] = e2e_path
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(prefix, plain_queries, args.cpu_seconds)
        print(json.dumps(line))
    if world > 1:
        dist.barrier(group=host_group)
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        # the CPU arm builds the host-only libraries and the oracle; it never loads libmgpu.so
        from manticoresearch_b200 import build as b
        b.build()
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
        run_reference(args)
    else:
        from __graft_entry__ import build
        build()
        run_ours(args)


if __name__ == "__main__":
    main()

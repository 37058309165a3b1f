#!/usr/bin/env python3
"""Benchmark of the full-text query hot path (BASELINE.json metric: queries/sec & compressed-postings GB/s).

A "step" = one pass of the hot path over one batch of synthetic queries:
  workload = BASELINE.json configs[1]: 10M-doc synthetic Zipfian index, 10k-query batch of 2-8 term AND/OR
  mixes, SPH_RANK_BM25 with field weights (title=10, body=1), top-100.
  --gpus N: the SAME index is split into N contiguous rowid-range shards, one per GPU (strong scaling);
  every rank evaluates the whole batch on its shard, K best keys per query are all-gathered over NCCL
  and merged on the GPU (global IDF inputs come from the all-reduced per-shard dictionaries).

  python bench.py --gpus 1 --steps 5 --warmup 3            # our arm
  python bench.py --impl reference --gpus 1 ...            # CPU arm: the oracle (the reference cannot be built here)
"""
import argparse
import ctypes as C
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

WORKLOAD = ("configs[1]: 10M-doc synthetic Zipfian plain index (seed 0x5EED0001, V=2^20, hit_format=inline, skiplist 32), "
            "10k-query batch of 2-8 term AND/OR/(a b)|(c d) mixes, SPH_RANK_BM25 field_weights=(title=10,body=1), top-100")


WORKLOAD_CFG4 = ("configs[3]: synthetic Zipfian plain index sharded by rowid range (seed 0x5EED0001, V=2^20), 10k-query batch of 2-8 term "
                 "AND/OR/(a b)|(c d) mixes with 10% ANDNOT, SPH_RANK_BM25 field_weights=(title=10,body=1), top-1000, local top-K + NCCL merge")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs", type=int, default=int(os.environ.get("MGPU_BENCH_DOCS", 10_000_000)))
    ap.add_argument("--queries", type=int, default=int(os.environ.get("MGPU_BENCH_QUERIES", 10_000)))
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="engine option for mgpu_index_set_option (A/B switches, e.g. --opt or_bits=0 --opt stats=1)")
    ap.add_argument("--only", default="", choices=["", "and", "or", "mix"], help="analysis only: keep one query shape of the batch")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"],
                    help="cfg2 = the bench line (BASELINE.json configs[1]); cfg4 = configs[3] shape for extra runs: same mix + 10%% ANDNOT, top-1000 (use with --docs 100000000 --gpus 8)")
    return ap.parse_args()


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


def ncu_traffic(kernel):
    """dram bytes (read+write) per launch of `kernel` from the committed ncu --set full capture of this command at N=1
    (profiles/kernel_traffic.json, see profiles/r01_v21_summary.md)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json"))).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def cpu_oracle_qps(prefix, queries, seconds, threads):
    """oracle (CPU restatement of the reference path) on a bounded sample: one query per thread, striped"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    idx = helpers.OracleIndex(prefix)
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


def parity_sample(prefix, queries, fetched, n):
    """full-size spot check outside the timed region: n queries of the batch re-run on the CPU oracle (the checker), compared
    bit-exactly (rowids, weights, order, total_found) with what the CUDA path returned"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    idx = helpers.OracleIndex(prefix)
    pick = list(range(0, len(queries), max(1, len(queries) // n)))[:n]
    bad = []
    lock = threading.Lock()

    def work(t, nt):
        for j in range(t, len(pick), nt):
            qi = pick[j]
            c = idx.search([queries[qi]]).get(0)
            g = fetched.get(qi)
            if (g["status"], g["total_found"], g["rowid"], g["weight"]) != (c["status"], c["total_found"], c["rowid"], c["weight"]):
                with lock:
                    bad.append(qi)

    nt = min(16, os.cpu_count() or 1)
    ths = [threading.Thread(target=work, args=(t, nt)) for t in range(nt)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    idx.close()
    return {"checked": len(pick), "mismatches": len(bad), "bad_queries": bad[:8], "checker": "oracle/oracle.cpp at full size (10M docs)"}


def run_reference(args):
    import manticoresearch_b200.mgpu as M
    from manticoresearch_b200 import workload
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prefix, _, _, _ = ensure_index(M, args.docs, 0, 1)
    queries = workload.cfg2_queries(n=args.queries, max_matches=100)
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
        "ms_per_step": 1000.0 * dt_all / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "docs": args.docs, "queries_per_batch": args.queries, "l2": "inputs >> L2"},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": "port",
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

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries rank 0's JSON line only: NCCL's own log (its version banner when NCCL_DEBUG is set) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    # ---- index shard of this rank
    prefix, first_doc, n_docs, build_s = ensure_index(M, args.docs, rank, world)
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
    from manticoresearch_b200 import distributed as D
    K = 100
    queries = workload.cfg2_queries(n=args.queries, max_matches=K)
    if args.workload == "cfg4":
        K = 1000
        queries = workload.cfg2_queries(n=args.queries, max_matches=K, with_andnot=0.1)
    if args.only:
        def shape(q):
            r = q.root
            if r.op == M.OP_AND:
                return "and"
            return "mix" if any(c.children for c in r.children) else "or"
        queries = [q for q in queries if shape(q) == args.only]
    if world > 1:
        gdf = D.global_keyword_docs(lambda w: (index.word_stats(w) or (0, 0))[0], queries, dev)
        D.apply_global_idf(queries, args.docs, gdf)

    nq = len(queries)
    merger = D.ShardMerger(nq, K, dev, local_rank, stream) if world > 1 else None

    def merge_step(batch):
        """local top-K keys -> NCCL all-gather -> GPU merge; total_found via all-reduce"""
        merger.merge(batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: plan uploaded once, K steps of (eval + merge [+ shard merge])
    batch = index.prepare(queries)
    for _ in range(args.warmup):
        batch.run()
        if world > 1:
            merge_step(batch)
    barrier()
    eval_ms, merge_ms, hot_ms, class_ms = [], [], [], []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            batch.run()
            if world > 1:
                merge_step(batch)       # asynchronous: export + all_gather + merge are stream-ordered, no host sync per step
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
    if rank == 0:
        fetched = batch.fetch()
        n_unsupported = sum(1 for i in range(nq) if fetched.results[i].status != 0)

    # algorithmic bytes / postings of the whole job (all shards)
    agg = torch.tensor([st["algorithmic_bytes"], st["postings"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(agg)
    job_bytes, job_postings = float(agg[0].item()), float(agg[1].item())

    # ---- end to end through the C ABI with host buffers: plan + H2D + kernels + D2H every step
    packed = M.pack_queries(queries)            # the caller's host buffers: flattened XQNode trees in, result arrays out
    host_results = M.ResultSet(queries)
    if world == 1:
        for _ in range(min(args.warmup, 2)):        # untimed: first-call costs of this path (pinned staging, pool growth)
            index.search_packed(packed, nq, host_results)
    barrier()
    e2e_t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        if world == 1:
            rs = index.search_packed(packed, nq, host_results)          # mgpu_search_batch
            assert rs.results[0].status == 0
            b2 = None
        else:
            b2 = index.prepare(queries, packed)
            b2.run()
            merge_step(b2)
            host_keys = merger.out_keys.cpu()
            host_counts = merger.out_counts.cpu()
            host_totals = merger.totals.cpu()
            d2h = host_keys.numel() * 8 + host_counts.numel() * 4 + host_totals.numel() * 8
            h2d = b2.stats()["h2d_bytes"]
            b2.free()
    barrier()
    e2e_s = time.perf_counter() - e2e_t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    if world == 1:
        b3 = index.prepare(queries)
        b3.run()
        b3.fetch()
        s3 = b3.stats()
        h2d, d2h = s3["h2d_bytes"], s3["d2h_bytes"]
        b3.free()
        ls = index.last_search_stats()      # of the last timed mgpu_search_batch call
        host_ms = {k: round(ls[k], 2) for k in ("host_total_ms", "host_plan_ms", "host_setup_ms", "host_wait_ms", "host_fetch_ms")}

    if rank == 0:
        qps = nq / (ms_per_step / 1000.0)
        peak, peak_src = measured_peak()
        eval_avg_ms = statistics.mean(eval_ms)
        # roofline of the DOMINANT kernel: the launch class with the largest share of the step; its own algorithmic bytes
        # (.spd + .spe extents of every keyword of its queries) over its own CUDA-event duration
        names = ["stream_kernel<512>", "eval_kernel<hits>", "and_kernel", "stream_kernel<256>", "and_kernel<hits>", "stream_kernel<512,or>", "stream_kernel<512,dnf>"]
        NC = len(names)
        cms = [statistics.mean(x[c] for x in class_ms) for c in range(NC)]
        dom = max(range(NC), key=lambda c: cms[c])
        achieved = st["class_bytes"][dom] / (cms[dom] / 1000.0) / 1e9
        all_achieved = st["algorithmic_bytes"] / (eval_avg_ms / 1000.0) / 1e9
        line = {
            "metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": (WORKLOAD if args.workload == "cfg2" else WORKLOAD_CFG4) + (" [ANALYSIS SUBSET: only %s queries]" % args.only if args.only else ""), "docs": args.docs, "queries_per_batch": nq, "parallelism": "rowid-range shards x%d" % world,
                       "l2": "inputs >> L2 (%.1f GB algorithmic bytes per step)" % (job_bytes / 1e9),
                       "index_build_s": round(build_s, 1), "index_load_s": round(load_s, 1), "unsupported_queries": n_unsupported},
            "postings_per_sec": job_postings / (ms_per_step / 1000.0),
            "compressed_GBps": job_bytes / (ms_per_step / 1000.0) / 1e9,
            "e2e": {"value": nq / (e2e_s / args.steps), "unit": "queries/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1000.0 * e2e_s / args.steps, "host_ms": host_ms if world == 1 else None},
            "gpu_launches": int(args.steps * (st["kernel_launches"] + (1 if world > 1 else 0))),
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(names[dom]), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": st["class_bytes"][dom], "kernel_ms": cms[dom],
                         "kernel_share_of_step": cms[dom] / ms_per_step,
                         "all_kernels": {"achieved": all_achieved, "frac": all_achieved / peak, "algorithmic_bytes": st["algorithmic_bytes"], "ms": eval_avg_ms},
                         "class_ms": dict(zip(names, cms)), "class_queries": dict(zip(names, st["class_queries"])),
                         "class_GBps": {names[c]: (st["class_bytes"][c] / (cms[c] / 1000.0) / 1e9 if cms[c] > 0 else None) for c in range(NC)},
                         "merge_kernel_ms": statistics.mean(merge_ms), "hot_decode_ms": statistics.mean(hot_ms), "hot_terms": st["hot_terms"],
                         "frac_of_nominal_8TBs": achieved / 8000.0},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["parity_sample"] = parity_sample(prefix, queries, fetched, 48)
            threads = os.cpu_count() or 1
            cqps, cn, cdt = cpu_oracle_qps(prefix, queries, args.cpu_seconds, threads)
            line["cpu_baseline"] = {"value": cqps, "unit": "queries/s", "cores": threads, "kind": "port",
                                    "sample": "%d queries of the same seeded batch, one query per thread, %d threads, %.1f s" % (cn, threads, cdt)}
        print(json.dumps(line))
    batch.free()
    index.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    from __graft_entry__ import build
    build()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

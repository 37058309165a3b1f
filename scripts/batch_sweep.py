"""Batch-size sweep of the reference-facing call (mgpu_search_batch, host buffers in and out) on the cfg2 workload:
end-to-end queries/s and p50 / p99 latency per call for batches of 1, 16, 256, 1k and 10k queries.

The reference calls the seam with one query at a time; batches are where the GPU path earns its keep.
Every repetition takes a different slice of the seeded 10k-query batch, so no call repeats the previous one's keywords.
The decoded hot-term store is rebuilt inside every call (never kept across calls).

    python scripts/batch_sweep.py [docs] > profiles/r02_batch_sweep.json
"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import manticoresearch_b200.mgpu as M  # noqa: E402
from manticoresearch_b200 import workload  # noqa: E402


def main():
    docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    prefix, _, _, _ = bench.ensure_index(M, docs, 0, 1)
    os.sync()       # the index files were just written: let their write-back finish before latencies are taken (it stalls host threads for 100s of ms)
    idx = M.Index(prefix, device=0)
    allq = workload.cfg2_queries(n=10_000, max_matches=100)
    out = {"workload": bench.WORKLOADS["cfg2"]["desc"], "docs": docs, "call": "mgpu_search_batch (host buffers, plan + H2D + kernels + D2H per call)", "sizes": []}
    for nq, reps in ((1, 300), (16, 150), (256, 40), (1000, 12), (10_000, 5)):
        lat = []
        slices = []
        for r in range(reps + 2):
            o = (r * 997 * max(1, nq // 3 + 1)) % max(1, len(allq) - nq + 1)
            qs = allq[o:o + nq]
            slices.append((M.pack_queries(qs), M.ResultSet(qs), qs))
        import gc
        gc.collect()
        gc.disable()    # a generation-2 collection over the harness's own query objects costs 100s of ms and would land inside a timed call
        for r, (packed, rs, qs) in enumerate(slices):
            t0 = time.perf_counter()
            idx.search_packed(packed, nq, rs)
            dt = time.perf_counter() - t0
            if r >= 2:      # two untimed calls: pinned staging, pool growth
                lat.append(dt * 1000.0)
        gc.enable()
        st = idx.last_search_stats()
        lat.sort()
        out["sizes"].append({
            "queries_per_call": nq, "calls": len(lat), "queries_per_s": nq * len(lat) / (sum(lat) / 1000.0),
            "latency_ms": {"p50": statistics.median(lat), "p99": lat[min(len(lat) - 1, int(0.99 * len(lat)))], "min": lat[0], "max": lat[-1]},
            "last_call": {k: round(st[k], 3) for k in ("host_plan_ms", "host_setup_ms", "host_wait_ms", "host_fetch_ms", "eval_kernel_ms", "hot_decode_ms")},
            "hot_terms_last_call": st["hot_terms"], "class_queries_last_call": st["class_queries"],
        })
        print("nq=%5d  %9.1f q/s  p50 %8.3f ms  p99 %8.3f ms" % (nq, out["sizes"][-1]["queries_per_s"], out["sizes"][-1]["latency_ms"]["p50"],
                                                                 out["sizes"][-1]["latency_ms"]["p99"]), file=sys.stderr, flush=True)
    idx.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

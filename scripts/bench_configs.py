"""Throughput of the other BASELINE.json configs' query shapes on the bench index (analysis; the bench line is cfg2).
usage: python scripts/bench_configs.py [docs] [n_queries]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import manticoresearch_b200.mgpu as M
from manticoresearch_b200 import workload

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
prefix, first, n, _ = bench.ensure_index(M, docs, 0, 1)
idx = M.Index(prefix, device=0)
params = M.SynthParams(docs)
sets = {
    "cfg1 two-term AND, PROXIMITY_BM25, K=1000": workload.cfg1_queries(n=nq),
    "cfg3 phrase/proximity, PROXIMITY_BM25, K=1000": workload.cfg3_queries(params, n=nq),
    "cfg5 stopword OR + filter + ORDER BY ts DESC, K=10000": workload.cfg5_queries(idx, n=max(10, nq // 10)),
    "cfg2 mix BM25, K=100": workload.cfg2_queries(n=nq),
}
for name, qs in sets.items():
    packed, rs = M.pack_queries(qs), M.ResultSet(qs)
    idx.search_packed(packed, len(qs), rs)      # warm-up
    t0 = time.perf_counter()
    idx.search_packed(packed, len(qs), rs)
    dt = time.perf_counter() - t0
    st = idx.last_search_stats()
    bad = sum(1 for i in range(len(qs)) if rs.results[i].status != 0)
    print("%-55s %6d queries  %9.1f q/s e2e  kernels %8.1f ms  classes(ms) %s  unsupported %d" % (
        name, len(qs), len(qs) / dt, st["eval_kernel_ms"], [round(x, 1) for x in st["class_ms"]], bad), flush=True)
    if "--cpu" in sys.argv:
        import helpers
        cpu = helpers.OracleIndex(prefix)
        m = min(len(qs), 16)
        t0 = time.perf_counter()
        cpu.search(qs[:m])
        print("    oracle, 1 thread: %.2f q/s" % (m / (time.perf_counter() - t0)), flush=True)
        cpu.close()
idx.close()

"""small reproducer: cfg3-shaped phrase/proximity queries on a 50k-doc corpus, GPU vs oracle"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
import manticoresearch_b200.mgpu as M
from manticoresearch_b200 import workload

tmp = tempfile.mkdtemp()
prefix = os.path.join(tmp, "s")
params = M.SynthParams(int(os.environ.get("DOCS", 50000)))
M.build_synthetic(prefix, params)
gpu = M.Index(prefix, device=0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
qs = workload.cfg3_queries(params, n=n)
g = gpu.search(qs)
print("ran", len(qs), [g.get(i)["total_found"] for i in range(min(n, 10))])
if "--check" in sys.argv:
    cpu = helpers.OracleIndex(prefix)
    c = cpu.search(qs)
    for i in range(len(qs)):
        helpers.assert_same_results(g.get(i), c.get(i), ctx="q%d" % i)
    print("parity ok")

import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import manticoresearch_b200.mgpu as M
from manticoresearch_b200 import workload, distributed as D
total, world = 10_000_000, 8
tmp = tempfile.mkdtemp()
prefixes = []
t0 = time.time()
for r in range(world):
    first, n = D.shard_range(total, r, world)
    p = os.path.join(tmp, "s%d" % r); prefixes.append(p)
    M.build_synthetic(p, M.SynthParams(n, first_doc=first))
print("built", time.time() - t0, flush=True)
sh = M.ShardedIndex(prefixes, [0] * world)
qs = workload.cfg2_queries(n=10000)
if len(sys.argv) > 1:
    sh.set_option("timing", 1)
for i in range(4):
    t = time.time(); r = sh.search(qs); dt = time.time() - t
    print("call %d %.1f ms" % (i, dt * 1e3), sh.stats(), flush=True)
sh.close()

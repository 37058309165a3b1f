"""Design study for launch class 5 (pure OR programs, BM25, top-K): which fraction of the rows survives each candidate bound
at the final K-th-best threshold?  CPU only (oracle doclist decode + numpy); not part of the product path.

    python scripts/or_bound_study.py [docs] [queries]
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import manticoresearch_b200.mgpu as M  # noqa: E402
from manticoresearch_b200 import workload  # noqa: E402


def main():
    ndocs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    K = 100
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, "s")
        M.build_synthetic(prefix, M.SynthParams(ndocs))
        idx = helpers.OracleIndex(prefix)
        queries = [q for q in workload.cfg2_queries(n=2000) if q.root.op == M.OP_OR and all(not c.children for c in q.root.children)][:nq]
        cache = {}
        tot = {}
        for q in queries:
            words = [w.word for w in q.keywords()]
            n = len(words)
            score = np.zeros(ndocs, np.float64)
            present = np.zeros(ndocs, bool)
            F = np.zeros(ndocs, np.uint8)
            ub_pres = np.zeros(ndocs, np.float64)       # sum over present of max(idf,0)
            ub_neg = np.zeros(ndocs, np.float64)        # ... + 0.4545*idf for present negative keywords
            ub_h2 = np.zeros(ndocs, np.float64)         # 3 levels: hits 1 / 2-3 / >=4
            ub_cls = np.zeros(ndocs, np.float64)        # 4-bit tf class
            idfs = []
            lits = []   # (weight, bool array): positive-weighted literals of the 3-level bound; total = sum of weights + const
            lits1 = []  # the same for the presence-only bound with negative keywords
            const = const1 = 0.0
            pos_kw = []     # (idf, presence) of positive keywords
            negpen1 = np.zeros(ndocs, np.float64)   # penalty of present negative keywords, one level
            negpen3 = np.zeros(ndocs, np.float64)   # ... three levels
            for w in words:
                if w not in cache:
                    cache[w] = idx.decode_doclist(w)
                rowid, hits, fields = cache[w][0], cache[w][1], cache[w][2]
                df = len(rowid)
                if not df:
                    continue
                idf = np.log((ndocs - df + 1) / df) / (2 * np.log(1 + ndocs)) / n
                idfs.append((idf, df))
                tf = hits / (hits + 1.2)
                score[rowid] += tf * idf
                present[rowid] = True
                F[rowid] |= (fields & 3).astype(np.uint8)
                cls = np.ceil(15 * tf) / 15
                Pm = np.zeros(ndocs, bool); Pm[rowid] = True
                H2 = np.zeros(ndocs, bool); H2[rowid[hits >= 2]] = True
                H4 = np.zeros(ndocs, bool); H4[rowid[hits >= 4]] = True
                if idf > 0:
                    pos_kw.append((idf, Pm))
                else:
                    negpen1[rowid] += 0.4545 * idf
                    negpen3[rowid] += idf * np.where(hits >= 4, 0.769, np.where(hits >= 2, 0.625, 0.4545))
                if idf > 0:
                    lits += [(0.4546 * idf, Pm), ((0.7143 - 0.4546) * idf, H2), ((1 - 0.7143) * idf, H4)]
                    lits1 += [(idf, Pm)]
                else:
                    a, b, d = 0.4545 * idf, 0.625 * idf, 0.769 * idf
                    const += d
                    lits += [(b - d, ~H4), (a - b, ~H2), (-a, ~Pm)]
                    const1 += a
                    lits1 += [(-a, ~Pm)]
                if idf > 0:
                    ub_pres[rowid] += idf
                    ub_neg[rowid] += idf
                    ub_h2[rowid] += idf * np.where(hits >= 4, 1.0, np.where(hits >= 2, 0.7143, 0.4546))
                    ub_cls[rowid] += idf * cls
                else:
                    ub_neg[rowid] += idf * 0.4545
                    ub_h2[rowid] += idf * np.where(hits >= 4, 0.769, np.where(hits >= 2, 0.625, 0.4545))
                    ub_cls[rowid] += idf * (np.ceil(15 * tf) - 1) / 15
            rank = np.array([0, 10, 1, 11])[F] * 1000.0
            weight = rank + np.floor((score + 0.5) * 1000)
            weight[~present] = -1
            if present.sum() <= K:
                continue
            thr = np.partition(weight, -K)[-K]
            res = {}
            for name, ub in (("presence", ub_pres), ("presence+neg", ub_neg), ("3 levels", ub_h2), ("tf class", ub_cls)):
                bound = rank + (ub + 0.5) * 1000 + 1
                res[name] = float((present & (bound >= thr)).sum()) / ndocs
            for name, ll, cc in (("req-AND 3lvl", lits, const), ("req-AND pres", lits1, const1)):
                total = cc + sum(w for w, _ in ll)
                slack = rank + (total + 0.5) * 1000 + 1 - thr       # per row, weight units
                ok = present & (slack >= 0)
                for w, arr in ll:
                    ok &= arr | (w * 1000 <= slack)
                res[name] = float(ok.sum()) / ndocs
            pos_kw.sort(key=lambda t: -t[0])
            suffix = np.cumsum([t[0] for t in pos_kw][::-1])[::-1] if pos_kw else []
            for name, pen in (("essOR neg1", negpen1), ("essOR neg3", negpen3)):
                need = (thr - 1 - rank) / 1000.0 - 0.5 - pen      # tf*idf the positive keywords must bring
                ok = need <= 0
                for i, (idf, Pm) in enumerate(pos_kw):
                    ok |= Pm & (suffix[i] >= need)
                res[name] = float((ok & present).sum()) / ndocs
            res["present"] = float(present.sum()) / ndocs
            res["exact"] = float((weight >= thr).sum()) / ndocs
            for k, v in res.items():
                tot.setdefault(k, []).append(v)
            print("%-60s idf/df %s  %s" % (" ".join(words), ["%.4f/%.3f" % (a, b / ndocs) for a, b in idfs],
                                           {k: round(v, 5) for k, v in res.items()}))
        print("MEAN", {k: round(float(np.mean(v)), 5) for k, v in tot.items()})
        print("MAX ", {k: round(float(np.max(v)), 5) for k, v in tot.items()})
        idx.close()


if __name__ == "__main__":
    main()

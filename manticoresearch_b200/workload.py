"""Seeded synthetic query sets over the synthetic Zipfian corpus.

Pure host-side input generation for the tests and bench.py: builds XQNode-shaped trees (mgpu.Node).
"""
import bisect
import math
import random

from . import mgpu as M


class ZipfRanks:
    """ranks Zipf(s=1)-sampled from [lo, hi] (1-based term ranks)"""
    def __init__(self, lo, hi):
        self.lo = lo
        self.cum = []
        acc = 0.0
        for r in range(lo, hi + 1):
            acc += 1.0 / r
            self.cum.append(acc)

    def sample(self, rng):
        x = rng.random() * self.cum[-1]
        return self.lo + bisect.bisect_left(self.cum, x)


def _kw(rank, pos):
    return M.kw(M.synth_keyword(rank - 1), pos)


def distinct_ranks(rng, sampler, n):
    seen = []
    while len(seen) < n:
        r = sampler.sample(rng)
        if r not in seen:
            seen.append(r)
    return seen


def cfg1_queries(n=1000, seed=0x5EED0002, max_matches=1000, ranker=M.RANK_PROXIMITY_BM25):
    """two-term AND, terms from rank bands [10,100] x [100,10000]"""
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        a = rng.randint(10, 100)
        b = int(math.exp(rng.uniform(math.log(100), math.log(10000))))
        out.append(M.Query(M.AND(_kw(a, 1), _kw(b, 2)), ranker=ranker, max_matches=max_matches))
    return out


def cfg2_queries(n=10000, seed=0x5EED0002, max_rank=100000, max_matches=100, with_andnot=0.0):
    """2-8 terms; 50% pure AND, 30% pure OR, 20% (a b)|(c d) mixes; SPH_RANK_BM25, field_weights=(title=10, body=1)"""
    rng = random.Random(seed)
    sampler = ZipfRanks(1, max_rank)
    out = []
    for _ in range(n):
        x = rng.random()
        nterms = rng.randint(2, 8)
        ranks = distinct_ranks(rng, sampler, nterms)
        leaves = [_kw(r, i + 1) for i, r in enumerate(ranks)]
        if with_andnot and x < with_andnot and nterms >= 3:
            root = M.ANDNOT(M.AND(*leaves[:-1]) if nterms > 2 else leaves[0], leaves[-1])
        elif x < 0.5 + with_andnot * 0.5:
            root = M.AND(*leaves)
        elif x < 0.8:
            root = M.OR(*leaves)
        else:
            if nterms < 4:
                ranks = distinct_ranks(rng, sampler, 4)
                leaves = [_kw(r, i + 1) for i, r in enumerate(ranks)]
            h = len(leaves) // 2
            root = M.OR(M.AND(*leaves[:h]), M.AND(*leaves[h:]))
        out.append(M.Query(root, ranker=M.RANK_BM25, field_weights=[10, 1], max_matches=max_matches))
    return out


def cfg3_queries(params, n=2000, seed=0x5EED0002, max_matches=1000):
    """50% 2-3 word phrases sampled from adjacent tokens of random docs, 50% "a b c"~5 with words co-occurring within 8 positions"""
    rng = random.Random(seed)
    out = []
    ndocs = params.n_docs
    while len(out) < n:
        doc = params.first_doc + rng.randrange(ndocs)
        blen = M.synth_field_len(params, doc, 1)
        if blen < 12:
            continue
        if len(out) % 2 == 0:
            k = rng.randint(2, 3)
            p0 = rng.randrange(blen - k)
            terms = [M.synth_token(params, doc, 1, p0 + i) for i in range(k)]
            if len(set(terms)) != k:
                continue
            node = M.PHRASE([(M.synth_keyword(t), i + 1) for i, t in enumerate(terms)])
        else:
            p0 = rng.randrange(blen - 8)
            offs = sorted(rng.sample(range(8), 3))
            terms = [M.synth_token(params, doc, 1, p0 + o) for o in offs]
            if len(set(terms)) != 3:
                continue
            node = M.PROXIMITY([(M.synth_keyword(t), i + 1) for i, t in enumerate(terms)], 5)
        out.append(M.Query(node, ranker=M.RANK_PROXIMITY_BM25, max_matches=max_matches))
    return out


def cfg5_queries(index, n=500, seed=0x5EED0002, max_matches=10000):
    """3-6 term OR from ranks [1,50] + gid BETWEEN 100 AND 299 + ORDER BY ts DESC"""
    rng = random.Random(seed)
    gid, ts = index.attr_index("gid"), index.attr_index("ts")
    out = []
    for _ in range(n):
        ranks = rng.sample(range(1, 51), rng.randint(3, 6))
        root = M.OR(*[_kw(r, i + 1) for i, r in enumerate(ranks)])
        out.append(M.Query(root, ranker=M.RANK_BM25, max_matches=max_matches,
                           sort_keys=[M.SortKey(M.KEYPART_INT, ts, True)], filters=[M.Filter(gid, 100, 299)]))
    return out


def random_boolean_queries(n, seed, max_rank=20000, nfields=2, rankers=(M.RANK_BM25, M.RANK_NONE), max_matches=50):
    """parity fuzz set: random trees over AND/OR/ANDNOT/MAYBE with field limits, weights, boosts, missing words"""
    rng = random.Random(seed)
    sampler = ZipfRanks(1, max_rank)
    out = []
    for qi in range(n):
        pos = [0]

        def leaf():
            pos[0] += 1
            r = sampler.sample(rng)
            word = M.synth_keyword(r - 1) if rng.random() > 0.03 else "zzmissing%d" % r
            node = M.kw(word, pos[0], boost=rng.choice([1.0, 1.0, 1.0, 2.0, 0.5]))
            u = rng.random()
            if u < 0.15:
                node.field_mask = 1
            elif u < 0.3:
                node.field_mask = 2
            return node

        def tree(depth):
            u = rng.random()
            if depth >= 2 or u < 0.25:
                return leaf()
            k = rng.randint(2, 4)
            if u < 0.55:
                return M.AND(*[tree(depth + 1) if rng.random() < 0.3 else leaf() for _ in range(k)])
            if u < 0.85:
                return M.OR(*[tree(depth + 1) if rng.random() < 0.3 else leaf() for _ in range(k)])
            if u < 0.95:
                return M.ANDNOT(tree(depth + 1), leaf() if rng.random() < 0.7 else tree(depth + 1))
            return M.MAYBE(tree(depth + 1), leaf())

        root = tree(0)
        fw = [rng.choice([1, 1, 2, 10, 0, -3]) for _ in range(nfields)] if rng.random() < 0.5 else None
        out.append(M.Query(root, ranker=rng.choice(rankers), field_weights=fw, max_matches=rng.choice([1, 7, max_matches, 1000]),
                           index_weight=rng.choice([1, 1, 1, 3])))
    return out


def random_hit_queries(params, n, seed, max_matches=200, with_hitops=False):
    """parity fuzz set of the hit stage: phrases / proximities sampled from real docs (some of stop words), combined with
    AND / OR / ANDNOT / MAYBE and plain keywords, ranked by every ranker; some queries repeat a keyword (dupes path)"""
    rng = random.Random(seed)
    sampler = ZipfRanks(1, 20000)
    out = []
    while len(out) < n:
        pos = [0]

        def leaf(word=None):
            pos[0] += 1
            v = rng.random()      # position filters: ^word, word$, ^word$, @field[N] (ExtTermPos_T)
            node = M.kw(word or M.synth_keyword(sampler.sample(rng) - 1), pos[0], field_start=(v < 0.08 or 0.16 <= v < 0.18), field_end=(0.08 <= v < 0.18))
            if 0.18 <= v < 0.26:
                node.field_max_pos = rng.choice([1, 3, 10, 50])
            u = rng.random()
            if u < 0.1:
                node.field_mask = 1
            elif u < 0.2:
                node.field_mask = 2
            return node

        def nway():
            while True:
                doc = params.first_doc + rng.randrange(params.n_docs)
                field = 1 if rng.random() < 0.8 else 0
                flen = M.synth_field_len(params, doc, field)
                if flen >= 10:
                    break
            if rng.random() < 0.5:
                k = rng.randint(2, 4)
                p0 = rng.randrange(flen - k)
                terms = [M.synth_token(params, doc, field, p0 + i) for i in range(k)]
                words = []
                for t in terms:
                    pos[0] += 1
                    words.append((M.synth_keyword(t), pos[0]))
                if rng.random() < 0.15:
                    pos[0] += 1                     # a gap in atom positions (a stop word dropped from the phrase)
                    words[-1] = (words[-1][0], pos[0])
                node = M.PHRASE(words)
            else:
                k = rng.randint(2, 4)
                span = min(flen - 1, 9)
                p0 = rng.randrange(flen - span)
                offs = sorted(rng.sample(range(span), k))
                terms = [M.synth_token(params, doc, field, p0 + o) for o in offs]
                if rng.random() < 0.3:
                    rng.shuffle(terms)
                words = []
                for t in terms:
                    pos[0] += 1
                    words.append((M.synth_keyword(t), pos[0]))
                node = M.PROXIMITY(words, rng.choice([1, 2, 5, 8]))
            if rng.random() < 0.15:
                node.field_mask = 1 << field
            return node

        def hitop():
            """NEAR / BEFORE / NOTNEAR / quorum over keywords that really sit close together in some document"""
            while True:
                doc = params.first_doc + rng.randrange(params.n_docs)
                field = 1 if rng.random() < 0.8 else 0
                flen = M.synth_field_len(params, doc, field)
                if flen >= 10:
                    break
            span = min(flen - 1, 12)
            p0 = rng.randrange(flen - span)
            k = rng.randint(2, 4)
            offs = sorted(rng.sample(range(span), k))
            terms = [M.synth_token(params, doc, field, p0 + o) for o in offs]
            v = rng.random()
            if v < 0.3:
                a, b = rng.sample(terms, 2) if rng.random() < 0.85 else (terms[0], terms[0])
                return M.NEAR(rng.choice([1, 2, 4, 9]), leaf(M.synth_keyword(a)), leaf(M.synth_keyword(b)))
            if v < 0.55:
                if rng.random() < 0.3:
                    rng.shuffle(terms)
                if rng.random() < 0.15:
                    terms[-1] = terms[0]
                return M.BEFORE(*[leaf(M.synth_keyword(t)) for t in terms])
            if v < 0.8:
                other = terms[1] if rng.random() < 0.7 else sampler.sample(rng) - 1
                return M.NOTNEAR(rng.choice([1, 2, 5, 12]), leaf(M.synth_keyword(terms[0])), leaf(M.synth_keyword(other)))
            words = list(terms) + [sampler.sample(rng) - 1 for _ in range(rng.randint(0, 2))]
            if rng.random() < 0.3:
                words.insert(rng.randrange(len(words) + 1), words[0])    # a repeated keyword: counts up to its hits
            ws = []
            for t in words:
                pos[0] += 1
                ws.append((M.synth_keyword(t), pos[0]))
            return M.QUORUM(ws, rng.randint(2, max(2, len(ws) - 1)))

        u = rng.random()
        if with_hitops and u < 0.3:
            h = hitop()
            v = rng.random()
            root = h if v < 0.6 else M.OR(h, leaf()) if v < 0.75 else M.AND(h, leaf()) if v < 0.9 else M.ANDNOT(h, leaf())
        elif u < 0.25:
            root = nway()
        elif u < 0.4:
            root = M.OR(nway(), leaf())
        elif u < 0.5:
            root = M.AND(nway(), leaf())
        elif u < 0.58:
            root = M.OR(nway(), nway())
        elif u < 0.64:
            root = M.ANDNOT(nway(), leaf())
        elif u < 0.70:
            root = M.MAYBE(leaf(), nway())
        elif u < 0.80:
            root = M.AND(*[leaf() for _ in range(rng.randint(2, 4))])
        elif u < 0.88:
            root = M.OR(M.AND(leaf(), leaf()), leaf())
        elif u < 0.94:
            w = M.synth_keyword(sampler.sample(rng) - 1)    # the same keyword twice: HasQwordDupes
            root = M.AND(leaf(w), leaf(), leaf(w)) if rng.random() < 0.5 else M.OR(leaf(w), M.AND(leaf(w), leaf()))
        else:
            root = M.OR(M.ANDNOT(leaf(), leaf()), M.MAYBE(leaf(), leaf()))
        ranker = rng.choice([M.RANK_PROXIMITY_BM25] * 5 + [M.RANK_WORDCOUNT] * 2 + [M.RANK_BM25, M.RANK_NONE]
                            + [M.RANK_SPH04] * 3 + [M.RANK_PROXIMITY] * 2 + [M.RANK_MATCHANY] * 2 + [M.RANK_FIELDMASK])
        fw = [rng.choice([1, 2, 10, 0, -3]) for _ in range(2)] if rng.random() < 0.5 else None
        out.append(M.Query(root, ranker=ranker, field_weights=fw, max_matches=rng.choice([5, max_matches, 1000]),
                           index_weight=rng.choice([1, 1, 2])))
    return out

"""Distributed-local search over contiguous rowid-range shards, one shard per GPU / process.

Replaces the reference's local-shard fan-out + merge (RunLocalSearches / MergeAllMatches / KillPlainDupes,
src/searchd.cpp:5596-5814, 4653-4738, 3910-3952) for disjoint rowid ranges:

  * every rank searches the whole query batch on its own shard (mgpu_batch_run);
  * bit-exactness against the UNSHARDED reference needs global IDF inputs: N = sum of shard docs and
    df(t) = sum of shard df(t), all-reduced once per batch and handed to the planner through
    mgpu_query.total_docs / word_docs (the reference's CSphMultiQueryArgs::m_iTotalDocs / m_pLocalDocs,
    src/sphinx.h:2914-2916 -> src/sphinxsearch.cpp:4298-4315);
  * ties break on the GLOBAL rowid = shard base + local rowid, which mgpu packs into the 128-bit match key;
  * the only data-path exchange is K packed keys per query per rank: all_gather -> shard_merge_kernel;
    total_found via all_reduce(sum).

torch.distributed is plumbing only (NCCL on GPUs; gloo in the CPU tests of the host-side logic).
"""
import torch
import torch.distributed as dist

from . import mgpu as M


def shard_range(total_docs, rank, world):
    """docs [first, first+n) of the corpus belong to shard `rank` of `world` (contiguous rowid ranges)"""
    first = total_docs * rank // world
    return first, total_docs * (rank + 1) // world - first


def global_keyword_docs(local_docs_of, queries, device, group=None):
    """all-reduces per-keyword document counts over the shards. local_docs_of(word) -> df on this shard (0 if absent).
    Returns {word: global df}."""
    words = sorted({k.word for q in queries for k in q.keywords()})
    df = torch.tensor([int(local_docs_of(w)) for w in words], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(df, group=group)
    return dict(zip(words, df.tolist()))


def apply_global_idf(queries, total_docs, gdf):
    """hands the global N and df(t) to every query of the batch (CSphMultiQueryArgs::m_iTotalDocs / m_pLocalDocs)"""
    for q in queries:
        q.total_docs = total_docs
        q.word_docs = [gdf[k.word] for k in q.keywords()]
        q.shard_of_global = True     # keyword order (and the fp32 summation order) of the unsharded index


def pack_key(weight, global_rowid, sort_hi=None):
    """host restatement of the device key layout (device_types.h Key128_t) for relevance order:
    hi = (weight ^ 0x80000000) << 32, lo = ~global_rowid << 32 | weight; bigger key = better match"""
    w = weight & 0xFFFFFFFF
    hi = ((w ^ 0x80000000) << 32) if sort_hi is None else sort_hi
    lo = (((~global_rowid) & 0xFFFFFFFF) << 32) | w
    return hi, lo


def merge_keys_host(per_shard_keys, k):
    """MergeAllMatches for disjoint shards on the host: concatenate, order by key descending, keep k.
    per_shard_keys: list (one per shard) of lists of (hi, lo). Used by the CPU tests; the product path is ShardMerger."""
    allk = [key for shard in per_shard_keys for key in shard]
    allk.sort(reverse=True)
    return allk[:k]


def unpack_key(hi, lo):
    """-> (global_rowid, weight)"""
    rowid = (~(lo >> 32)) & 0xFFFFFFFF
    w = lo & 0xFFFFFFFF
    return rowid, w - (1 << 32) if w & 0x80000000 else w


class ShardMerger:
    """device buffers + the NCCL exchange of one rank; merge(batch) leaves the global top-K keys on every rank.
    Everything is asynchronous and ordered on `stream`, which must be the stream the index runs on
    (Index.set_stream(stream.cuda_stream)) and torch's current stream (NCCL collectives)."""

    def __init__(self, n_queries, k, device, local_rank, stream):
        self.nq, self.k, self.device, self.local_rank, self.stream = n_queries, k, device, local_rank, stream
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.keys = torch.zeros((n_queries, k, 2), dtype=torch.int64, device=device)
        self.counts = torch.zeros((n_queries,), dtype=torch.int32, device=device)
        self.totals = torch.zeros((n_queries,), dtype=torch.int64, device=device)
        self.all_keys = torch.zeros((self.world, n_queries, k, 2), dtype=torch.int64, device=device)
        self.all_counts = torch.zeros((self.world, n_queries), dtype=torch.int32, device=device)
        self.out_keys = torch.zeros((n_queries, k, 2), dtype=torch.int64, device=device)
        self.out_counts = torch.zeros((n_queries,), dtype=torch.int32, device=device)
        self._lib = M.lib()

    def merge(self, batch):
        """local top-K keys -> all_gather over NVLink -> shard_merge_kernel; total_found via all_reduce"""
        batch.export_keys(self.keys.data_ptr(), self.counts.data_ptr(), self.totals.data_ptr(), self.k)
        if self.world > 1:
            dist.all_gather_into_tensor(self.all_keys, self.keys)
            dist.all_gather_into_tensor(self.all_counts, self.counts)
            dist.all_reduce(self.totals)
        else:
            self.all_keys.copy_(self.keys.unsqueeze(0))
            self.all_counts.copy_(self.counts.unsqueeze(0))
        rc = self._lib.mgpu_merge_shard_keys(self.local_rank, self.all_keys.data_ptr(), self.all_counts.data_ptr(), self.world, self.nq, self.k,
                                             self.out_keys.data_ptr(), self.out_counts.data_ptr(), self.stream.cuda_stream)
        if rc != M.MGPU_OK:
            raise M.MgpuError(rc, "mgpu_merge_shard_keys failed")

    def fetch(self):
        """-> per query list of (global_rowid, weight), best first, plus total_found"""
        keys = self.out_keys.cpu().numpy().astype("uint64")
        counts = self.out_counts.cpu().tolist()
        totals = self.totals.cpu().tolist()
        out = []
        for qi in range(self.nq):
            rows = [unpack_key(int(keys[qi, i, 0]), int(keys[qi, i, 1])) for i in range(counts[qi])]
            out.append({"matches": rows, "total_found": totals[qi]})
        return out

"""Multi-GPU decomposition of the mapping path: one process per GPU,
no collective inside the mapping path.

Two ways to shard, both result-preserving:

* **query sharding, replicated index** -- queries are independent (Mapper._query_draft only reads
  the index, pyx:1052-1055): every rank builds (or receives) the whole index and maps its own
  share of the queries.  `partition_queries` balances the shares by fragment count (the unit of
  L1/L2 work), not by query count.
* **reference sharding** -- whole reference GENOMES are dealt to ranks.  Candidates need seeds in
  one contig (computeMap.hpp:328), both computeCGI filters are keyed inside one reference genome
  (computeCoreIdentity.hpp:218-250) and the min-fraction filter uses that genome's own length
  (pyx:1124-1126), so every rank produces FINAL hit rows for its genomes -- upstream FastANI does
  the same per thread (splitReferenceGenomes / correctRefGenomeIds, computeCoreIdentity.hpp:
  454-484).  `gather_hits` moves the per-query rows (16 bytes each) to every rank with two small
  collectives (counts, then payload) and `merge_hits` restores the reference's ordering: identity
  descending, stable in ascending global genome id (pyx:1135).

The collectives go through ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU
tests); torch is imported lazily and is plumbing only.
"""
import heapq

import numpy as np

HIT_DT = np.dtype([("ref_genome", "<i4"), ("matches", "<i4"), ("fragments", "<i4"), ("identity", "<f4")])


def partition_queries(fragment_counts, world_size):
    """Greedy longest-processing-time assignment of queries to ranks by fragment count.

    Returns a list of `world_size` index lists (each ascending).  Deterministic: ties go to the
    lower query index / lower rank.
    """
    if world_size < 1:
        raise ValueError("world_size must be positive")
    order = sorted(range(len(fragment_counts)), key=lambda i: (-int(fragment_counts[i]), i))
    heap = [(0, r) for r in range(world_size)]
    shares = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        shares[r].append(i)
        heapq.heappush(heap, (load + int(fragment_counts[i]), r))
    return [sorted(s) for s in shares]


def reference_shards(genome_lengths, world_size):
    """Contiguous blocks of whole reference genomes with about equal total length.

    Returns `world_size + 1` offsets: rank r owns genomes [off[r], off[r + 1]); the local genome
    id of a hit plus off[r] is its global id (the analogue of correctRefGenomeIds,
    computeCoreIdentity.hpp:477-484).
    """
    if world_size < 1:
        raise ValueError("world_size must be positive")
    n = len(genome_lengths)
    total = float(sum(int(x) for x in genome_lengths))
    offsets, acc, g = [0], 0.0, 0
    for r in range(1, world_size):
        target = total * r / world_size
        while g < n and acc + int(genome_lengths[g]) / 2.0 <= target:
            acc += int(genome_lengths[g])
            g += 1
        offsets.append(g)
    offsets.append(n)
    return offsets


def hits_to_rows(hits, name_to_id):
    """Hit objects of one query -> HIT_DT rows with local genome ids (column by column: a structured
    assignment per hit costs 2 us, which is a millisecond per query against a 500-genome shard)."""
    n = len(hits)
    rows = np.zeros(n, dtype=HIT_DT)
    if n:
        rows["ref_genome"] = np.fromiter((name_to_id[h.name] for h in hits), dtype=np.int32, count=n)
        rows["matches"] = np.fromiter((h.matches for h in hits), dtype=np.int32, count=n)
        rows["fragments"] = np.fromiter((h.fragments for h in hits), dtype=np.int32, count=n)
        rows["identity"] = np.fromiter((h.identity for h in hits), dtype=np.float32, count=n)
    return rows


def merge_hits(rows_per_rank, offsets):
    """Rows of ONE query from every rank (local ids) -> one array with global ids in the
    reference's order: identity descending, stable w.r.t. ascending genome id (pyx:1135; the
    per-rank rows arrive sorted that way and the merge keeps it)."""
    parts = []
    for r, rows in enumerate(rows_per_rank):
        rows = np.asarray(rows, dtype=HIT_DT).copy()
        rows["ref_genome"] += int(offsets[r])
        parts.append(rows)
    allrows = np.concatenate(parts) if parts else np.zeros(0, dtype=HIT_DT)
    order = np.lexsort((allrows["ref_genome"], -allrows["identity"].astype(np.float64)))
    return allrows[order]


def gather_hits(rows_per_query, group=None, device=None, cap=None):
    """All-gather the hit rows of a list of queries.

    `rows_per_query`: list (same length on every rank) of HIT_DT arrays with LOCAL genome ids.
    Returns `out[q][r]` = rows of query q from rank r.

    `cap`: an upper bound, known on every rank, of the rows one rank can hold for the whole list
    (queries x genomes of the largest shard).  With it the exchange is ONE collective of a fixed-width
    block per rank -- the per-query counts followed by the 16-byte rows -- and one device-to-host copy;
    without it (or when the block would pass 4 MiB) two collectives: the counts, then a payload padded
    to the largest total.  A few KB per query either way: latency-bound on NVLink, nothing to fuse with
    the mapping kernels.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    nq = len(rows_per_query)
    dev = device if device is not None else torch.device("cpu")
    mine = np.concatenate([np.asarray(r, dtype=HIT_DT) for r in rows_per_query]) if nq else np.zeros(0, dtype=HIT_DT)
    out = [[None] * world for _ in range(nq)]

    if cap is not None and len(mine) > cap:
        raise ValueError("gather_hits: %d rows on this rank exceed cap=%d (cap must bound every rank)" % (len(mine), cap))
    if cap is not None and (nq + 4 * cap) * 4 <= (4 << 20):         # (the same decision on every rank)
        width = nq + 4 * int(cap)
        block = np.zeros(width, dtype=np.int32)
        block[:nq] = [len(r) for r in rows_per_query]
        block[nq:nq + 4 * len(mine)] = mine.view(np.int32)
        send = torch.from_numpy(block).to(dev)
        recv = torch.empty(world * width, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
        host = recv.cpu().numpy().reshape(world, width)
        for r in range(world):
            cnt = host[r, :nq]
            rows = host[r, nq:nq + 4 * int(cnt.sum())].copy().view(HIT_DT)
            pos = 0
            for q in range(nq):
                out[q][r] = rows[pos:pos + int(cnt[q])]
                pos += int(cnt[q])
        return out

    counts = torch.tensor([len(r) for r in rows_per_query], dtype=torch.int64, device=dev)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    totals = [int(c.sum().item()) for c in all_counts]
    width = max(max(totals), 1)
    flat = np.zeros(width, dtype=HIT_DT)
    flat[:len(mine)] = mine
    payload = torch.from_numpy(flat.view(np.int32).reshape(width, 4).copy()).to(dev)
    all_payload = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(all_payload, payload, group=group)
    for r in range(world):
        rows = all_payload[r].cpu().numpy().reshape(-1).view(HIT_DT)
        cnt = all_counts[r].cpu().numpy()
        pos = 0
        for q in range(nq):
            out[q][r] = rows[pos:pos + int(cnt[q])].copy()
            pos += int(cnt[q])
    return out


def query_reference_sharded(mapper, queries, offsets, group=None, device=None, drafts=False):
    """Map every query against this rank's reference shard and return, on every rank, the merged
    global hit rows per query.  `mapper.names` must be the LOCAL genome ids 0..n_local-1 (or any
    names whose position in `mapper.names` is the local id)."""
    name_to_id = {n: i for i, n in enumerate(mapper.names)}
    # one call for the whole list: light queries share passes of the pipeline and the next pass is staged while the
    # current one is mapped (fa_query_batch)
    queries = list(queries)
    items = [list(q) for q in queries] if drafts else queries
    local = [hits_to_rows(hits, name_to_id) for hits in mapper.query_many(items)]
    shard = max(int(offsets[r + 1]) - int(offsets[r]) for r in range(len(offsets) - 1))
    gathered = gather_hits(local, group=group, device=device, cap=len(local) * shard)
    return [merge_hits(per_rank, offsets) for per_rank in gathered]

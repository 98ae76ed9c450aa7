"""Multi-GPU decomposition of the mapping path: one process per GPU,
no collective inside the mapping kernels.

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
  454-484).  The exchange lives in the library (`fa_gather_hits` / `fa_query_batch_sharded`,
  csrc/fa_comm.cu): the query sketches are made once across the ranks (each sketches 1/world of
  the fragments of a group of queries; one ncclAllGather per group, a group ahead of the mapping),
  and the per-query rows (16 bytes each) of all ranks travel in one small ncclAllGather over
  NVLink and are merged into the reference's ordering, identity descending, stable in ascending
  global genome id (pyx:1135).  `merge_hits` is the same merge in numpy, kept
  as the specification the tests compare the library against.

No torch here: the NCCL communicator belongs to the library (`Communicator`), and `connect` hands
its 128-byte id from rank 0 to the other ranks over a plain TCP socket (MASTER_ADDR / MASTER_PORT
of the launcher, port + 1) -- or pass an `exchange` callable to use a channel you already have.
"""
import heapq
import os
import socket
import struct
import time

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


def connect(world_size=None, rank=None, device=None, exchange=None, timeout=120.0):
    """Create this rank's `Communicator` (collective over all ranks).

    `world_size`, `rank`, `device` default to the launcher's WORLD_SIZE / RANK / LOCAL_RANK.  Rank 0
    draws the NCCL unique id; `exchange(id_or_None) -> id` broadcasts it (rank 0 passes the id, the
    others None) -- by default a one-shot TCP hand-out on MASTER_ADDR:(MASTER_PORT + 1)
    (FA_COMM_PORT overrides the port)."""
    from ._fastani import Communicator
    world_size = int(os.environ.get("WORLD_SIZE", "1")) if world_size is None else int(world_size)
    rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
    device = int(os.environ.get("LOCAL_RANK", str(rank))) if device is None else int(device)
    uid = Communicator.unique_id() if rank == 0 else None
    if world_size > 1:
        uid = (exchange or _tcp_broadcast(world_size, rank, timeout))(uid)
    return Communicator(uid, world_size, rank, device)


def _tcp_broadcast(world_size, rank, timeout):
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(os.environ.get("FA_COMM_PORT", str(int(os.environ.get("MASTER_PORT", "29500")) + 1)))

    def exchange(uid):
        if rank == 0:
            with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as srv:
                srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                srv.bind((addr, port))
                srv.listen(world_size)
                srv.settimeout(timeout)
                for _ in range(world_size - 1):
                    conn, _peer = srv.accept()
                    with conn:
                        conn.sendall(struct.pack("<I", len(uid)) + uid)
            return uid
        deadline = time.monotonic() + timeout
        while True:
            try:
                with socket.create_connection((addr, port), timeout=5.0) as conn:
                    buf = b""
                    while len(buf) < 4 or len(buf) < 4 + struct.unpack("<I", buf[:4])[0]:
                        chunk = conn.recv(4096)
                        if not chunk:
                            raise ConnectionError("rank 0 closed the connection early")
                        buf += chunk
                    return buf[4:4 + struct.unpack("<I", buf[:4])[0]]
            except (ConnectionRefusedError, ConnectionError, socket.timeout, OSError):
                if time.monotonic() > deadline:
                    raise
                time.sleep(0.05)

    return exchange


def query_reference_sharded(mapper, queries, offsets, comm):
    """Map every query against this rank's reference shard and return, on every rank, the merged
    global hit rows per query (numpy structured arrays, HIT_DT).  Collective; one library call:
    no `Hit` objects, no host language between the mapping and the gather."""
    return mapper.query_many(list(queries), rows=True, comm=comm, genome_offsets=[int(o) for o in offsets])

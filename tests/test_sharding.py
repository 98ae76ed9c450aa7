"""The multi-GPU decomposition (pyfastani_b200/sharding.py, csrc/fa_comm.cu).  On CPU: the
partitions, the merge specification, the product's own TCP rendezvous (three processes) and bench.py's
torch.distributed plumbing on a world-size-2 gloo group.  On the GPU: reference sharding reproduces the single-index
hits exactly, and the library's NCCL gather equals the numpy merge (two ranks when two GPUs are visible)."""
import os
import socket

import numpy as np
import pytest

from pyfastani_b200 import sharding
from pyfastani_b200.sharding import HIT_DT


def test_partition_queries_balances_fragments():
    counts = [1666, 10, 900, 901, 5, 1500, 700, 300]
    shares = sharding.partition_queries(counts, 3)
    assert sorted(i for s in shares for i in s) == list(range(len(counts)))
    loads = [sum(counts[i] for i in s) for s in shares]
    assert max(loads) - min(loads) <= max(counts)              # LPT bound
    assert max(loads) <= 4 / 3 * sum(counts) / 3 + max(counts) / 3 + 1
    assert sharding.partition_queries(counts, 3) == shares       # deterministic
    assert sharding.partition_queries([], 2) == [[], []]
    assert sharding.partition_queries([5], 4) == [[0], [], [], []]
    with pytest.raises(ValueError):
        sharding.partition_queries(counts, 0)


def test_reference_shards_are_contiguous_whole_genomes():
    lengths = [5_000_000] * 10 + [1_000_000] * 10
    off = sharding.reference_shards(lengths, 4)
    assert off[0] == 0 and off[-1] == 20 and off == sorted(off) and len(off) == 5
    per = [sum(lengths[off[r]:off[r + 1]]) for r in range(4)]
    assert max(per) - min(per) <= 5_000_000
    assert sharding.reference_shards(lengths, 1) == [0, 20]
    assert sharding.reference_shards([], 3) == [0, 0, 0, 0]
    assert sharding.reference_shards([7], 3)[-1] == 1


def _rows(lst):
    a = np.zeros(len(lst), dtype=HIT_DT)
    for i, t in enumerate(lst):
        a[i] = t
    return a


def test_merge_hits_restores_reference_order():
    # rank 0 owns global genomes 0..2, rank 1 owns 3..5; equal identities keep ascending genome id
    r0 = _rows([(2, 10, 20, 99.5), (0, 9, 20, 97.25)])
    r1 = _rows([(1, 11, 20, 99.5), (0, 8, 20, 98.0), (2, 7, 20, 97.25)])
    m = sharding.merge_hits([r0, r1], [0, 3, 6])
    assert m["ref_genome"].tolist() == [2, 4, 3, 0, 5]
    assert m["identity"].tolist() == [99.5, 99.5, 98.0, 97.25, 97.25]
    assert m["matches"].tolist() == [10, 11, 8, 9, 7]
    assert len(sharding.merge_hits([_rows([]), _rows([])], [0, 1, 2])) == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_local_rows(rank):
    """Three queries; rank r reports hits on its own (local) genomes only, ragged and empty cases."""
    rng = np.random.default_rng(100 + rank)
    local = []
    for q in range(3):
        n = [2, 0, 3][q] if rank == 0 else [1, 0, 0][q]
        rows = np.zeros(n, dtype=HIT_DT)
        rows["ref_genome"] = np.arange(n)
        rows["matches"] = rng.integers(1, 100, n)
        rows["fragments"] = 100 + q
        rows["identity"] = np.sort(rng.uniform(80, 100, n).astype(np.float32))[::-1]
        local.append(rows)
    return local


def _rendezvous_worker(rank, world, port, out_dir):
    """The product's own rendezvous (no torch): rank 0 hands a 128-byte id to the others over TCP."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    uid = bytes(range(128)) if rank == 0 else None
    got = sharding._tcp_broadcast(world, rank, 30.0)(uid)
    open(os.path.join(out_dir, "id%d.bin" % rank), "wb").write(got)


def test_tcp_rendezvous_world3(tmp_path):
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_rendezvous_worker, args=(r, 3, port, str(tmp_path))) for r in (2, 1, 0)]   # rank 0 last: the others retry
    for p in procs:
        p.start()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in range(3):
        assert open(tmp_path / ("id%d.bin" % r), "rb").read() == bytes(range(128))


def _gloo_worker(rank, world, port, out_dir):
    """World-size-2 gloo group on CPU: the N > 1 host logic of bench.py (its partition of the query list, the max over
    ranks of the step time, the hit total) with torch.distributed as the plumbing, as on the GPU box."""
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        counts = [1666, 1666, 1500, 900, 1666, 400, 1666, 1200, 1666]
        mine = sharding.partition_queries(counts, world)[rank]
        t = bench.max_over_ranks(dist, torch, torch.device("cpu"), 1.0 + rank, world)
        n = bench.sum_over_ranks(dist, torch, torch.device("cpu"), len(mine), world)
        # the unique id reaches every rank through the group as well (what bench.py hands to sharding.connect)
        uid = bench.broadcast_bytes(dist, torch, torch.device("cpu"), bytes(range(128)) if rank == 0 else None, 128, world)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), mine=np.array(mine), t=t, n=n, uid=np.frombuffer(uid, np.uint8))
    finally:
        dist.destroy_process_group()


def test_bench_plumbing_gloo_world2(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "rank0.npz")
    b = np.load(tmp_path / "rank1.npz")
    assert sorted(a["mine"].tolist() + b["mine"].tolist()) == list(range(9))
    assert float(a["t"]) == float(b["t"]) == 2.0 and int(a["n"]) == int(b["n"]) == 9
    assert a["uid"].tobytes() == b["uid"].tobytes() == bytes(range(128))


def _nccl_worker(rank, world, port, out_dir):
    """One rank of the library's own exchange: fa_gather_hits over NCCL (no torch), against the numpy merge."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    comm = sharding.connect(world, rank, rank)
    local = _fake_local_rows(rank)
    merged = comm.gather_hits(local, [0, 3, 4])
    # a batch with more rows than travel with the counts (2048): the second collective
    big = np.zeros(3000 + 500 * rank, dtype=HIT_DT)
    big["ref_genome"] = np.arange(len(big)) % 3000
    big["identity"] = np.linspace(99, 80, len(big)).astype(np.float32)
    merged_big = comm.gather_hits([big[:10], big[10:]], [0, 3000, 6500])
    info = comm.info
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), *merged, big0=merged_big[0], big1=merged_big[1], mybig=big,
             collectives=info["collectives"], **{"local%d" % q: local[q] for q in range(3)})


@pytest.mark.gpu
def test_gather_hits_nccl_world2(tmp_path):
    import multiprocessing as mp
    import pyfastani_b200 as pf

    if pf.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL refuses two ranks on one device); run with gpurun --gpus 2")
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    a = np.load(tmp_path / "rank0.npz")
    b = np.load(tmp_path / "rank1.npz")
    for q in range(3):
        ma, mb = a["arr_%d" % q], b["arr_%d" % q]
        assert np.array_equal(ma, mb)                                   # every rank holds the same merged rows
        expect = sharding.merge_hits([a["local%d" % q], b["local%d" % q]], [0, 3, 4])
        assert np.array_equal(ma, expect)
        assert len(ma) == [3, 0, 3][q]
        assert np.all(np.diff(ma["identity"]) <= 0)
    for q, sl in enumerate((slice(0, 10), slice(10, None))):
        expect = sharding.merge_hits([a["mybig"][sl], b["mybig"][sl]], [0, 3000, 6500])
        assert np.array_equal(a["big%d" % q], expect) and np.array_equal(b["big%d" % q], expect)
    assert int(a["collectives"]) == 3                                   # one for the small batch, two for the large one


@pytest.mark.gpu
def test_reference_sharding_equals_single_index():
    """Mapping against two reference shards (whole genomes each) and merging gives exactly the
    hits of the single index: same genomes, matches, fragments, identity bits, same order."""
    import pyfastani_b200 as pf
    import synth

    query, refs, _ = synth.one_to_many(77, 9, 120_000, lo=0.85, hi=0.99)
    rng = np.random.default_rng(5)
    drafts = [synth.fragment(rng, r, 3, min_end=500) for r in refs]
    lengths = [sum(len(c) for c in d) for d in drafts]
    full = pf.Sketch()
    for i, d in enumerate(drafts):
        full.add_draft(i, d)
    mapper = full.index()
    want = sharding.hits_to_rows(mapper.query_genome(query), {i: i for i in range(len(drafts))})
    for world in (2, 3):
        off = sharding.reference_shards(lengths, world)
        per_rank = []
        for r in range(world):
            sk = pf.Sketch()
            for i in range(off[r], off[r + 1]):
                sk.add_draft(i - off[r], drafts[i])
            m = sk.index()
            per_rank.append(sharding.hits_to_rows(m.query_genome(query), {i: i for i in range(off[r + 1] - off[r])}))
        got = sharding.merge_hits(per_rank, off)
        assert np.array_equal(got, want)
    assert len(want) == 9


@pytest.mark.gpu
def test_query_reference_sharded_world1():
    """`query_reference_sharded` end to end on one rank (an NCCL communicator of one): the list of queries goes through
    ONE library call -- fa_query_batch, the all-gather, the merge -- and comes back as the rows of plain queries."""
    import pyfastani_b200 as pf
    import synth

    query, refs, _ = synth.one_to_many(78, 7, 90_000, lo=0.85, hi=0.99)
    sk = pf.Sketch()
    for i, r in enumerate(refs):
        sk.add_genome(i, r)
    mapper = sk.index()
    queries = [query, refs[2], synth.revcomp(refs[5]), b"ACGT" * 10]
    ident = {i: i for i in range(len(refs))}
    want = [sharding.hits_to_rows(mapper.query_genome(q), ident) for q in queries]
    comm = sharding.connect(1, 0, 0)
    assert comm.info["nccl_version"] >= 21800
    got = sharding.query_reference_sharded(mapper, queries, [0, len(refs)], comm)
    rows = mapper.query_many(queries, rows=True)                    # the same rows without the communicator
    assert len(got) == len(queries)
    for a, b, c in zip(got, want, rows):
        assert np.array_equal(a, b) and np.array_equal(c, b)
    assert len(want[0]) == 7 and len(want[3]) == 0
    assert comm.info["collectives"] == 1


def _sharded_worker(rank, world, port, out_dir, exchange):
    """One rank of the reference-sharded layout end to end: its shard of the genomes, all queries through
    fa_query_batch_sharded -- sketches made once across the ranks (sketch_exchange) unless switched off."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if not exchange:
        os.environ["FA_NO_SKETCH_EXCHANGE"] = "1"
    if exchange == "tight":                        # slots of 96 hashes: most sketches do not fit, their passes sketch at home
        os.environ["FA_EXCHANGE_STRIDE"] = "96"
    import pyfastani_b200 as pf
    import synth

    drafts, queries, lengths = _sharded_inputs(synth)
    off = sharding.reference_shards(lengths, world)
    comm = sharding.connect(world, rank, rank)
    sk = pf.Sketch(device=rank)
    for i in range(off[rank], off[rank + 1]):
        sk.add_draft(i - off[rank], drafts[i])
    mapper = sk.index()
    got = sharding.query_reference_sharded(mapper, queries, off, comm)
    info = dict(mapper.last_query_info)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), *got, collectives=comm.info["collectives"], ms_sketch=info["ms_sketch"])


def _sharded_inputs(synth):
    query, refs, _ = synth.one_to_many(91, 10, 150_000, lo=0.85, hi=0.99)
    rng = np.random.default_rng(6)
    drafts = [synth.fragment(rng, r, 3, min_end=500) for r in refs]
    lengths = [sum(len(c) for c in d) for d in drafts]
    # whole genomes, a draft in seven contigs (short ones among them), a packed one, a str, one without a fragment
    import pyfastani_b200 as pf
    q_draft = synth.fragment(rng, refs[3], 7, min_end=100) + [b"ACGT", b"ACGTTGCA" * 300]
    queries = [query, q_draft, pf.PackedSequence.pack(refs[7]), synth.revcomp(refs[5]).decode("ascii"), b"ACGT" * 10,
               refs[1][:40_000], refs[9]]
    return drafts, queries, lengths


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", [True, False, "tight"])
def test_reference_sharded_queries_world2(tmp_path, exchange):
    """Two ranks, genomes sharded, every query mapped by both: merged rows equal those of one index over all genomes --
    with the query sketches made once across the ranks and all-gathered (the default) and with every rank sketching
    every query."""
    import multiprocessing as mp
    import pyfastani_b200 as pf
    import synth

    if pf.device_count() < 2:
        pytest.skip("needs two GPUs; run with gpurun --gpus 2")
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, str(tmp_path), exchange)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    drafts, queries, _ = _sharded_inputs(synth)
    full = pf.Sketch()
    for i, d in enumerate(drafts):
        full.add_draft(i, d)
    want = full.index().query_many([q if isinstance(q, list) else [q] for q in queries], rows=True)
    a, b = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for q in range(len(queries)):
        assert np.array_equal(a["arr_%d" % q], b["arr_%d" % q])
        assert np.array_equal(a["arr_%d" % q], want[q]), q
    assert len(want[0]) == 10 and len(want[4]) == 0
    # one all-gather of hit rows; with the exchange, one more per group of queries (seven light queries: one group)
    assert int(a["collectives"]) == (2 if exchange else 1)

"""Host-side logic of the multi-GPU decomposition (pyfastani_b200/sharding.py): CPU tests with a
world_size-2 gloo group, and one GPU test that reference sharding reproduces the single-index
hits exactly."""
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


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # three queries; rank r reports hits on its own (local) genomes only, ragged and empty cases
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
        gathered = sharding.gather_hits(local)                       # two collectives: counts, padded payload
        merged = [sharding.merge_hits(per_rank, [0, 3, 4]) for per_rank in gathered]
        with pytest.raises(ValueError):
            sharding.gather_hits(local, cap=-1)                      # a bound below the rows held: refused before any collective
        one = sharding.gather_hits(local, cap=3 * 3)                 # one fixed-width collective (3 queries x 3 genomes)
        for q in range(3):
            for r in range(world):
                assert np.array_equal(one[q][r], gathered[q][r]), (q, r)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), *merged, **{"local%d" % q: local[q] for q in range(3)})
    finally:
        dist.destroy_process_group()


def test_gather_hits_gloo_world2(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "rank0.npz")
    b = np.load(tmp_path / "rank1.npz")
    for q in range(3):
        ma, mb = a["arr_%d" % q], b["arr_%d" % q]
        assert np.array_equal(ma, mb)                                   # every rank holds the same merged rows
        expect = sharding.merge_hits([a["local%d" % q], b["local%d" % q]], [0, 3, 4])
        assert np.array_equal(ma, expect)
        assert len(ma) == [3, 0, 3][q]
        assert np.all(np.diff(ma["identity"]) <= 0)


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
def test_query_reference_sharded_world1(tmp_path):
    """`query_reference_sharded` end to end on one rank (a gloo group of one): the list of queries goes through one
    `query_many` call, the rows through the one-collective gather, and come back as the rows of plain queries."""
    import torch.distributed as dist
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
    dist.init_process_group("gloo", init_method="file://" + str(tmp_path / "rdv"), rank=0, world_size=1)
    try:
        got = sharding.query_reference_sharded(mapper, queries, [0, len(refs)])
    finally:
        dist.destroy_process_group()
    assert len(got) == len(queries)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    assert len(want[0]) == 7 and len(want[3]) == 0

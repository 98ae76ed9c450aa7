"""The reference's own tests (src/pyfastani/tests/test_ani.py, test_sketch.py) re-run against
pyfastani_b200, with the same adapters (str, bytes, numpy view) and the same expected values."""
import pickle
import warnings

import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pf():
    import pyfastani_b200
    return pyfastani_b200


@pytest.fixture(scope="module")
def genomes():
    return {n: golden_io.genome(n) for n in ("ecoli", "shigella")}


ADAPTERS = {
    "str": lambda b: b.decode("ascii"),
    "bytes": lambda b: b,
    "numpy": lambda b: np.frombuffer(b, dtype=np.uint8),
    "bytearray": lambda b: bytearray(b),
    "memoryview": lambda b: memoryview(b),
}


@pytest.mark.parametrize("adapter", list(ADAPTERS))
def test_fastani_example(pf, genomes, adapter):          # test_ani.py:29-51
    get = ADAPTERS[adapter]
    sketch = pf.Sketch()
    sketch.add_draft("Escherichia_coli_str_K12_MG1655", [get(c) for c in genomes["ecoli"]])
    mapper = sketch.index()
    hits = mapper.query_draft(map(get, genomes["shigella"]))
    assert len(hits) == 1
    assert hits[0].name == "Escherichia_coli_str_K12_MG1655"
    assert hits[0].matches == 1303
    assert hits[0].fragments == 1608
    assert round(abs(hits[0].identity - 97.7507), 4) == 0


def test_fastani_example_result_files(pf, genomes):
    """The two result files of the reference CLI (outputCGI / outputPhylip) from `query_many` on the reference's own example:
    the line FastANI's README shows, and the 2 x 2 lower-triangular matrix."""
    from pyfastani_b200 import output
    sketch = pf.Sketch()
    sketch.add_draft("ecoli.fna", genomes["ecoli"])
    sketch.add_draft("shigella.fna", genomes["shigella"])
    mapper = sketch.index()
    results = mapper.query_many([genomes["shigella"]])
    lines = list(output.tabular_lines(["shigella.fna"], results))
    assert lines[0].startswith("shigella.fna\tshigella.fna\t100\t") and lines[0].endswith("\t1608")
    assert lines[1] == "shigella.fna\tecoli.fna\t97.7507\t1303\t1608"
    both = mapper.query_many([genomes["ecoli"], genomes["shigella"]])
    m = list(output.matrix_lines(["ecoli.fna", "shigella.fna"], ["ecoli.fna", "shigella.fna"], both))
    ab = [h for h in both[0] if h.name == "shigella.fna"][0].identity
    ba = [h for h in both[1] if h.name == "ecoli.fna"][0].identity
    assert m == ["2", "ecoli.fna", "shigella.fna\t%f" % float((np.float32(ba) + np.float32(ab)) / np.float32(2))]


def test_escherichia_minimizers(pf, genomes):           # test_ani.py:54-71
    sketch = pf.Sketch()
    assert sketch.window_size == 24
    sketch.add_draft("Escherichia_coli_str_K12_MG1655", genomes["ecoli"])
    assert len(sketch.minimizers) == 371301
    mapper = sketch.index()
    assert len(mapper.minimizers) == 371301
    assert len(mapper.lookup_index) == 361568
    hits = mapper.query_draft(genomes["ecoli"])
    assert len(hits) == 1
    assert (hits[0].matches, hits[0].fragments, hits[0].identity) == (1547, 1547, 100.0)


def test_shigella_minimizers(pf, genomes):              # test_ani.py:74-91
    sketch = pf.Sketch()
    sketch.add_draft("Shigella_flexneri_2a_01", genomes["shigella"])
    assert len(sketch.minimizers) == 386387
    first = [(m.hash, m.sequence_id, m.window_position) for m in (sketch.minimizers[i] for i in range(4))]
    assert first == [(21161528, 0, 0), (25007321, 0, 18), (159674326, 0, 24), (262432603, 0, 25)]
    assert sketch.minimizers[-1] == sketch.minimizers[386386]
    with pytest.raises(IndexError):
        sketch.minimizers[386387]
    mapper = sketch.index()
    assert len(mapper.lookup_index) == 347908
    hits = mapper.query_draft(genomes["shigella"])
    assert (hits[0].matches, hits[0].fragments, hits[0].identity) == (1600, 1608, 100.0)


def test_config1(pf, genomes):
    sketch = pf.Sketch()
    sketch.add_draft("shigella", genomes["shigella"])
    mapper = sketch.index()
    hits = mapper.query_genome(genomes["ecoli"][0])
    assert hits == [pf.Hit("shigella", float.fromhex("0x1.86a7fap+6"), 1322, 1547)]
    info = mapper.last_query_info
    assert (info["fragments"], info["candidates"], info["mappings"]) == (1547, 5038, 4101)
    assert mapper.query_genome(genomes["ecoli"][0], threads=4) == hits
    with pytest.raises(ValueError):
        mapper.query_genome(genomes["ecoli"][0], threads=-1)


def test_sketch_pickling(pf, genomes):                  # test_ani.py:136-153
    sketch = pf.Sketch()
    sketch.add_draft("Escherichia_coli_str_K12_MG1655", genomes["ecoli"])
    sketch = pickle.loads(pickle.dumps(sketch))
    assert sketch.names == ["Escherichia_coli_str_K12_MG1655"]
    assert len(sketch.minimizers) == 371301
    mapper = sketch.index()
    hits = mapper.query_draft(genomes["shigella"])
    assert (hits[0].name, hits[0].matches, hits[0].fragments) == ("Escherichia_coli_str_K12_MG1655", 1303, 1608)
    assert round(abs(hits[0].identity - 97.7507), 4) == 0


def test_mapper_pickling(pf, genomes):                  # test_ani.py:156-173
    sketch = pf.Sketch()
    sketch.add_draft("Escherichia_coli_str_K12_MG1655", genomes["ecoli"])
    mapper = pickle.loads(pickle.dumps(sketch.index()))
    assert len(mapper.lookup_index) == 361568
    hits = mapper.query_draft(genomes["shigella"])
    assert (hits[0].name, hits[0].matches, hits[0].fragments) == ("Escherichia_coli_str_K12_MG1655", 1303, 1608)
    assert round(abs(hits[0].identity - 97.7507), 4) == 0


def test_reinit(pf):                                    # test_sketch.py:25-35
    sketch = pf.Sketch(fragment_length=100)
    sketch.add_genome("test", "ATGC" * 100)
    assert sketch.names == ["test"] and sketch.fragment_length == 100
    sketch.__init__(fragment_length=200)
    assert sketch.names == [] and sketch.fragment_length == 200


def test_add_draft_warnings(pf):                        # test_sketch.py:37-52
    sketch = pf.Sketch()
    with warnings.catch_warnings(record=True) as catch:
        warnings.simplefilter("always")
        sketch.add_draft("short_seq", ["ATGC" * 1000, "ATGC"])
        assert len(catch) == 1
    assert sketch.names == ["short_seq"]


def test_pickle_small(pf):                              # test_sketch.py:54-61
    sketch = pf.Sketch()
    sketch.add_genome("short_seq", "ATGC" * 1000)
    sketch2 = pickle.loads(pickle.dumps(sketch))
    assert sketch2.names == ["short_seq"]
    assert len(sketch2.minimizers) == len(sketch.minimizers)


def test_query_warnings_and_views(pf):
    rng = np.random.default_rng(3)
    g = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 60_000))
    sketch = pf.Sketch()
    sketch.add_genome("g", g)
    sketch.clear()
    assert sketch.names == [] and len(sketch.minimizers) == 0
    sketch.add_genome("g", g)
    assert sketch.occurences_threshold == 2147483647
    mapper = sketch.index()
    assert sketch.names == [] and len(sketch.minimizers) == 0     # ownership moved, pyx:795-804
    with warnings.catch_warnings(record=True) as catch:
        warnings.simplefilter("always")
        hits = mapper.query_draft([g, b"ACGTACGTAC"])
        assert len(catch) == 1 and len(hits) == 1
    assert mapper.query_draft([g[:2999]]) == []
    li = mapper.lookup_index
    key = next(iter(li))
    assert key in li and (key + 1 in li) in (True, False)
    pos = li[key]
    assert all(isinstance(p, pf.Position) for p in pos)
    with pytest.raises(KeyError):
        li[0xFFFFFFFF]
    k2, p2 = next(li.items())
    assert k2 == key and p2 == pos
    dev = pf.DeviceSequence.from_host(g)
    assert mapper.query_genome(dev) == mapper.query_genome(g)


def test_lookup_index_mutation(pf):
    """MinimizerIndex.__setitem__ / __delitem__ (pyx:1480-1507) on `Mapper.lookup_index`: the reference edits the hash
    table L1 seeding reads, here the CSR table in device memory is rebuilt around the entry.  Erase, restore, move a
    list to a new hash, an empty list, bad positions; the seed count of a query follows the table, and with every
    key erased nothing maps."""
    import synth
    q, refs, _ = synth.one_to_many(21, 3, 40_000, lo=0.90, hi=0.99)
    sketch = pf.Sketch()
    for i, r in enumerate(refs):
        sketch.add_genome("ref%d" % i, r)
    mapper = sketch.index()
    base = mapper.query_genome(q)
    seeds0 = mapper.last_query_info["seeds"]
    assert len(base) == 3 and seeds0 > 0
    li = mapper.lookup_index
    keys = list(li)
    n0 = len(li)
    assert keys == sorted(keys) and n0 == len(keys)
    # a hash of the query's own sketch that occurs in the references: its positions are seeds of the query
    qs = pf.Sketch()
    qs.add_genome("q", q[:3_000])
    h = next(m.hash for m in qs.minimizers if m.hash in li)
    pos = li[h]
    assert pos and all(isinstance(p, pf.Position) for p in pos)
    del li[h]
    assert h not in li and len(li) == n0 - 1
    with pytest.raises(KeyError):
        li[h]
    with pytest.raises(KeyError):
        del li[h]
    mapper.query_genome(q)
    assert mapper.last_query_info["seeds"] < seeds0                    # those seeds are gone
    li[h] = pos
    assert h in li and li[h] == pos and len(li) == n0 and list(li) == keys
    assert mapper.query_genome(q) == base and mapper.last_query_info["seeds"] == seeds0
    li[h] = pos[:1]                                                    # a shorter list, then the original again
    assert li[h] == pos[:1]
    li[h] = list(reversed(pos))
    assert li[h] == list(reversed(pos))                                # (the order given is kept, as in the reference)
    assert mapper.query_genome(q) == base and mapper.last_query_info["seeds"] == seeds0
    # a hash that was not a key: inserted in hash order
    h2 = next(x for x in range(1, 100_000) if x not in li)
    li[h2] = pos
    assert h2 in li and li[h2] == list(pos) and len(li) == n0 + 1 and list(li) == sorted(keys + [h2])
    assert mapper.query_genome(q) == base                              # (h2 is not in the query's sketch)
    del li[h2]
    li[h] = []
    assert h in li and li[h] == [] and len(li) == n0
    li[h] = pos
    with pytest.raises(ValueError):
        li[h] = [pf.Position(0, 1_000_000_000)]
    with pytest.raises(ValueError):
        li[h] = [pos[0], pos[0]]
    with pytest.raises(TypeError):
        li[h] = [(0, 1)]
    assert li[h] == pos and mapper.query_genome(q) == base
    # a standalone table is a plain host-side mapping, as in the reference (and pickles as one)
    own = pf.MinimizerIndex()
    own[7] = pos
    own[3] = []
    assert len(own) == 2 and 7 in own and own[7] == pos and dict(own.items()) == {7: pos, 3: []}
    back = pickle.loads(pickle.dumps(own))
    assert dict(back.items()) == {7: pos, 3: []}
    del own[7]
    assert 7 not in own
    # every key erased: no seeds, no hits
    for k in keys:
        del li[k]
    assert len(li) == 0 and list(li) == []
    assert mapper.query_genome(q) == [] and mapper.last_query_info["seeds"] == 0


def test_sketch_files(pf, tmp_path):
    """The on-disk sketch (`Sketch.save` / `Sketch.load`, `Mapper.save` / `Mapper.load`): parameters, names, lengths and
    the minimizer columns survive the round trip, the reloaded mapper answers like the original, and nothing is
    sketched again."""
    import synth
    q, refs, _ = synth.one_to_many(31, 4, 70_000, lo=0.86, hi=0.99)
    sketch = pf.Sketch(fragment_length=2_000, percentage_identity=82.0)
    for i, r in enumerate(refs[:3]):
        sketch.add_genome("ref%d" % i, r)
    sketch.add_draft("draft", [refs[3][:30_000], refs[3][30_000:31_000], refs[3][31_000:]])
    h, s, w = sketch.minimizers.arrays()
    st = sketch.minimizers.__getstate__()
    assert (h.tolist(), s.tolist(), w.tolist()) == (st["hashes"], st["ids"], st["offsets"]) and len(h) == st["length"]
    path = tmp_path / "refs.sketch.npz"
    sketch.save(path)
    back = pf.Sketch.load(path)
    assert back.names == sketch.names and back.fragment_length == 2_000 and back.percentage_identity == 82.0
    assert back.window_size == sketch.window_size and back.k == sketch.k
    for a, b in zip(back.minimizers.arrays(), (h, s, w)):
        assert np.array_equal(a, b)
    assert back.__getstate__() == sketch.__getstate__()
    back.add_genome("late", refs[0][5_000:60_000])                       # the counter of sequence ids goes on where it was
    sketch.add_genome("late", refs[0][5_000:60_000])
    assert back.__getstate__() == sketch.__getstate__()
    mapper = sketch.index()
    want = mapper.query_genome(q)
    assert len(want) == 5
    assert back.index().query_genome(q) == want
    mpath = tmp_path / "refs.mapper.npz"
    mapper.save(mpath)
    again = pf.Mapper.load(mpath)
    assert again.names == mapper.names and len(again.lookup_index) == len(mapper.lookup_index)
    assert again.query_genome(q) == want and again.query_draft([q[:40_000], q[40_000:]]) == mapper.query_draft([q[:40_000], q[40_000:]])
    with pytest.raises(ValueError):
        np.savez(tmp_path / "other.npz", header=np.frombuffer(b'{"format": "x"}', np.uint8))
        pf.Sketch.load(tmp_path / "other.npz")
    odd = pf.Sketch()
    odd.add_genome(("a", "tuple"), refs[0])                               # names that are not JSON: pickle is the way
    odd.save(tmp_path / "odd.npz")                                        # (a tuple becomes a list: still JSON)
    odd2 = pf.Sketch()
    odd2.add_genome(object(), refs[0])
    with pytest.raises(TypeError):
        odd2.save(tmp_path / "odd2.npz")


def test_query_many_equals_single_queries(pf):
    """`Mapper.query_many` / `fa_query_batch`: one call, results identical to query_genome / query_draft
    item by item (genomes, drafts, a device-resident sequence, a query without hits, an empty batch)."""
    import synth
    q, refs, _ = synth.one_to_many(99, 5, 90_000, lo=0.85, hi=0.99)
    sketch = pf.Sketch()
    for i, r in enumerate(refs):
        sketch.add_genome(i, r)
    mapper = sketch.index()
    rng = np.random.default_rng(5)
    unrelated = synth.to_bytes(synth.random_codes(rng, 30_000))
    draft = synth.fragment(rng, q, 5, min_end=500)
    queries = [q, draft, unrelated, pf.DeviceSequence.from_host(refs[2]), tuple(draft[:2]), refs[4].decode()]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        single = [mapper.query_draft(x) if isinstance(x, (list, tuple)) else mapper.query_genome(x) for x in queries]
        many = mapper.query_many(queries)
    assert many == single and single[2] == [] and len(single[0]) == 5
    assert mapper.last_query_info["queries"] == len(queries)
    assert mapper.query_many([]) == []
    with pytest.raises(ValueError):
        mapper.query_many(queries, threads=-1)


def test_device_resident_drafts(pf):
    """Drafts whose contigs already live in HBM (`DeviceSequence`): more than a handful go through one gather launch
    instead of one copy each -- odd lengths, a contig shorter than a fragment, a host contig in between -- as
    references (`add_draft`) and as queries (`query_draft`, `query_many`); results equal the host-bytes path."""
    import synth
    q, refs, _ = synth.one_to_many(123, 4, 150_000, lo=0.86, hi=0.99)
    rng = np.random.default_rng(8)
    drafts = [synth.fragment(rng, r, 11, min_end=300) for r in refs]
    qd = synth.fragment(rng, q, 13, min_end=300) + [b"ACGTTGCA" * 20]
    host, devs = pf.Sketch(), pf.Sketch()
    for i, d in enumerate(drafts):
        host.add_draft(i, d)
        devs.add_draft(i, [pf.DeviceSequence.from_host(c) if j != 3 else c for j, c in enumerate(d)])
    assert devs.minimizers.__getstate__() == host.minimizers.__getstate__()
    mh, md = host.index(), devs.index()
    qdev = [pf.DeviceSequence.from_host(c) if j != 5 else c for j, c in enumerate(qd)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = mh.query_draft(qd)
        assert md.query_draft(qdev) == want and len(want) == 4
        assert md.query_many([qdev, qd, qdev[:6], qdev]) == [want, want, mh.query_draft(qd[:6]), want]


def test_protein_mode(pf):                                   # test_ani.py:96-115
    """Sketch(protein=True): alphabet 20, window 1; the reference's MIBiG cluster test and pickling."""
    gold = golden_io.protein_golden()[0]["bgc"]
    bgc = {n: [c.decode() for c in golden_io.genome(n)] for n in ("BGC0001425", "BGC0001427", "BGC0001428")}
    sketch = pf.Sketch(protein=True, fragment_length=100)
    assert sketch.protein and sketch.window_size == 1
    sketch.add_draft("BGC0001425", bgc["BGC0001425"])
    sketch.add_draft("BGC0001427", bgc["BGC0001425"])
    mapper = pickle.loads(pickle.dumps(sketch)).index()
    assert mapper.protein and mapper.window_size == 1
    hits = mapper.query_draft(bgc["BGC0001428"])
    assert len(hits) == 2
    assert [(h.name, h.matches, h.fragments) for h in hits] == [("BGC0001425", 130, 176), ("BGC0001427", 130, 176)]
    assert [np.float32(h.identity) for h in hits] == [golden_io.f32(r[1]) for r in gold["as_in_test_ani"]]

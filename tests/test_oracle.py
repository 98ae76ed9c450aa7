"""CPU tests: pin the oracle (oracle/fastani_oracle.c) against golden vectors made by the
real pyfastani, the reference's own known answers, and the reference headers compiled in
place (oracle/_ref) where that library is present."""
import hashlib

import numpy as np
import pytest

import cases
import golden_io
from oracle.oracle import Oracle, available

PORT = Oracle("port")
HAVE_REF = "reference" in available()


def _contig_minimizers(orc, contigs, params):
    sk = orc.sketch(**params)
    sk.add_draft("g", contigs)
    return sk


@pytest.mark.parametrize("idx", range(len(cases.minimizer_cases())))
def test_minimizers_match_pyfastani(idx):
    case = cases.minimizer_cases()[idx]
    man, h, s, w = golden_io.minimizer_golden()[idx]
    assert man["name"] == case["name"]
    sk = _contig_minimizers(PORT, case["contigs"], case["params"])
    assert sk.params.window == man["window"]
    gh, gs, gw = sk.minimizers()
    assert len(gh) == man["n"]
    assert np.array_equal(gh, h) and np.array_equal(gs, s) and np.array_equal(gw, w)
    assert sk.warnings == man["warnings"]


def _check_hits(hits, names, rows):
    assert len(hits) == len(rows)
    for h, (name, ident, matches, frags) in zip(hits, rows):
        assert names[h["ref_genome"]] == name
        assert h["matches"] == matches and h["fragments"] == frags
        assert h["identity"] == golden_io.f32(ident), (float(h["identity"]).hex(), ident)


@pytest.mark.parametrize("name", [c["name"] for c in cases.query_cases()])
def test_queries_match_pyfastani(name):
    case = next(c for c in cases.query_cases() if c["name"] == name)
    gold = golden_io.query_golden()[name]
    sk = PORT.sketch(**case["params"])
    for rname, contigs in case["refs"]:
        sk.add_draft(rname, contigs)
    sk.index()
    assert sk.params.window == gold["window"]
    assert len(sk.minimizers()[0]) == gold["minimizers"]
    assert sk.unique() == gold["unique"]
    assert sk.warnings == gold["ref_warnings"]
    for q, res in zip(case["queries"], gold["results"]):
        hits, info = sk.query_draft(q)
        _check_hits(hits, sk.names, res["hits"])
        assert info["short_contigs"] == res["warnings"]


def test_config1_known_answers():
    """BASELINE config 1 and the reference's own known answers (test_ani.py:47-91)."""
    gold = golden_io.config1_golden()
    genomes = {n: golden_io.genome(n) for n in ("ecoli", "shigella")}
    assert [len(c) for c in genomes["ecoli"]] == [4641652]
    assert [len(c) for c in genomes["shigella"]] == [4607202, 221618]
    for rname in ("shigella", "ecoli"):
        sk = PORT.sketch()
        sk.add_draft(rname, genomes[rname])
        sk.index()
        h, s, w = sk.minimizers()
        for qname in ("ecoli", "shigella"):
            g = gold["%s_vs_%s" % (qname, rname)]
            assert len(h) == g["minimizers"] and sk.unique() == g["unique"]
            assert hashlib.sha256(h.tobytes() + s.tobytes() + w.tobytes()).hexdigest() == g["sha256"]
            if qname != rname:      # the self-queries are covered on the GPU side; keep the CPU suite short
                hits, _ = sk.query_draft(genomes[qname])
                _check_hits(hits, sk.names, g["hits"])
    assert gold["shigella_vs_ecoli"]["hits"][0][2:] == [1303, 1608]          # test_ani.py:49-50
    assert abs(float.fromhex(gold["shigella_vs_ecoli"]["hits"][0][1]) - 97.7507) < 5e-5   # test_ani.py:51
    assert (gold["ecoli_vs_ecoli"]["minimizers"], gold["ecoli_vs_ecoli"]["unique"]) == (371301, 361568)
    assert (gold["shigella_vs_shigella"]["minimizers"], gold["shigella_vs_shigella"]["unique"]) == (386387, 347908)


def test_murmur_known_vectors():
    # MurmurHash3_x64_128 seed 42, low 32 bits of h1; first Shigella minimizer
    gold = golden_io.config1_golden()["ecoli_vs_shigella"]["first"]
    sf = golden_io.genome("shigella")[0]
    h, s, w = PORT.minimizers(sf[:200])
    assert [int(h[0]), int(s[0]), int(w[0])] == gold[0]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (reference absent)")
class TestAgainstReferenceHeaders:
    def setup_class(cls):
        cls.ref = Oracle("reference")

    def test_window(self):
        for kw in ({}, {"k": 12}, {"fragment_length": 1000}, {"percentage_identity": 95.0}, {"p_value": 1e-6},
                   {"reference_size": 5_000_000_000}, {"fragment_length": 500, "k": 11}):
            assert PORT.recommended_window(**kw) == self.ref.recommended_window(**kw), kw

    def test_hash(self):
        rng = np.random.default_rng(0)
        for n in list(range(1, 50)) + [64, 100, 2048]:
            b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            assert PORT.hash(b) == self.ref.hash(b), n

    def test_stat_table_exhaustive_small(self):
        """identity / pass flag / minHits for every (s, x), s <= 320: the Boost-free binomial
        quantile must reproduce the reference's Boost path bit for bit."""
        for k, pid in ((16, 80.0), (12, 90.0)):
            for s in list(range(1, 321)):
                assert PORT.minimum_hits(s, k, pid) == self.ref.minimum_hits(s, k, pid), (s, k, pid)
                for x in range(0, s + 1):
                    assert PORT.l2_stat(x, s, k, pid) == self.ref.l2_stat(x, s, k, pid), (x, s, k, pid)

    def test_stat_table_sampled_large(self):
        rng = np.random.default_rng(1)
        for s in (400, 777, 1024, 2000, 2962, 5000):
            assert PORT.minimum_hits(s) == self.ref.minimum_hits(s)
            xs = set(rng.integers(0, s + 1, 60).tolist()) | set(range(0, 90)) | {s, s - 1}
            for x in xs:
                assert PORT.l2_stat(x, s) == self.ref.l2_stat(x, s), (x, s)

    def test_intermediates_boundary_case(self):
        """L1 candidates and L2 mappings, field by field, on the draft-heavy boundary case."""
        import synth
        refs, query = synth.boundary_case(99)
        out = {}
        for orc in (PORT, self.ref):
            sk = orc.sketch()
            for name, contigs in refs:
                sk.add_draft(name, contigs)
            sk.index()
            hits, info = sk.query_draft(query, dump=True)
            m = np.sort(info["mappings"], order=["frag", "seq", "ref_start"])
            out[orc.kind] = (sk.minimizers(), info["candidates"], m, hits, info["stats"])
        a, b = out["port"], out["reference"]
        for x, y in zip(a[0], b[0]):
            assert np.array_equal(x, y)
        assert np.array_equal(a[1], b[1]) and len(a[1]) > 100
        assert np.array_equal(a[2], b[2]) and len(a[2]) > 100
        assert np.array_equal(a[3], b[3])
        assert a[4]["seeds"] == b[4]["seeds"]


# ---- protein mode (pyx:225-309, 548-550) -------------------------------------------------------
@pytest.mark.parametrize("idx", range(len(cases.protein_minimizer_cases())))
def test_protein_minimizers_match_pyfastani(idx):
    case = cases.protein_minimizer_cases()[idx]
    gold, arr = golden_io.protein_golden()
    man = gold["minimizers"][idx]
    assert man["name"] == case["name"] and man["window"] == 1
    sk = _contig_minimizers(PORT, case["contigs"], case["params"])
    assert sk.params.window == 1 and sk.params.alphabet == 20
    gh, gs, gw = sk.minimizers()
    assert len(gh) == man["n"]
    assert np.array_equal(gh, arr["h%d" % idx]) and np.array_equal(gs, arr["s%d" % idx]) and np.array_equal(gw, arr["w%d" % idx])
    assert sk.warnings == man["warnings"]


@pytest.mark.parametrize("idx", range(len(cases.protein_query_cases())))
def test_protein_queries_match_pyfastani(idx):
    case = cases.protein_query_cases()[idx]
    gold = golden_io.protein_golden()[0]["queries"][idx]
    sk = PORT.sketch(**case["params"])
    for rname, contigs in case["refs"]:
        sk.add_draft(rname, contigs)
    sk.index()
    assert len(sk.minimizers()[0]) == gold["minimizers"] and sk.unique() == gold["unique"]
    for q, res in zip(case["queries"], gold["results"]):
        hits, info = sk.query_draft(q)
        _check_hits(hits, sk.names, res["hits"])
        assert info["short_contigs"] == res["warnings"]


def test_protein_bgc_known_answer():
    """The reference's own protein test (test_ani.py:96-115): 130 / 176 for both names."""
    gold = golden_io.protein_golden()[0]["bgc"]
    bgc = {n: golden_io.genome(n) for n in ("BGC0001425", "BGC0001427", "BGC0001428")}
    sk = PORT.sketch(protein=True, fragment_length=100)
    sk.add_draft("BGC0001425", bgc["BGC0001425"])
    sk.add_draft("BGC0001427", bgc["BGC0001425"])
    sk.index()
    assert len(sk.minimizers()[0]) == gold["minimizers"] and sk.unique() == gold["unique"]
    hits, _ = sk.query_draft(bgc["BGC0001428"])
    _check_hits(hits, sk.names, gold["as_in_test_ani"])
    assert [(int(h["matches"]), int(h["fragments"])) for h in hits] == [(130, 176), (130, 176)]
    sk2 = PORT.sketch(protein=True, fragment_length=100)
    sk2.add_draft("BGC0001425", bgc["BGC0001425"])
    sk2.add_draft("BGC0001427", bgc["BGC0001427"])
    sk2.index()
    _check_hits(sk2.query_draft(bgc["BGC0001428"])[0], sk2.names, gold["distinct_refs"])
    _check_hits(sk2.query_draft(bgc["BGC0001427"])[0], sk2.names, gold["self"])

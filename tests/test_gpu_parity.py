"""GPU parity tests: the CUDA path, called through the C ABI, against (1) golden vectors made by
the real pyfastani, (2) the CPU oracle on the same seeded inputs, intermediates included.

Bar (BASELINE.json north_star): minimizer hashes/positions, candidate regions, matches and
fragments bit-exact; identity within 1e-4 absolute -- these tests ask for bit-exact identity too.
"""
import hashlib

import numpy as np
import pytest

import capi
import cases
import golden_io
import synth
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu

IDENTITY_TOL = 1e-4      # the north-star tolerance; asserted bit-exact first, this is the fallback bar


def _port():
    return Oracle("port")


@pytest.mark.parametrize("batched", [True, False])
@pytest.mark.parametrize("idx", range(len(cases.minimizer_cases())))
def test_minimizers_match_pyfastani(idx, batched):
    case = cases.minimizer_cases()[idx]
    man, h, s, w = golden_io.minimizer_golden()[idx]
    sk = capi.Sketch(batched=batched, **case["params"])
    sk.add_draft("g", case["contigs"])
    gh, gs, gw = sk.minimizers()
    assert len(gh) == man["n"], case["name"]
    assert np.array_equal(gh, h) and np.array_equal(gs, s) and np.array_equal(gw, w), case["name"]
    assert sk.warnings == man["warnings"]


def _check_hits(hits, names, rows):
    assert len(hits) == len(rows), (hits, rows)
    for h, (name, ident, matches, frags) in zip(hits, rows):
        assert names[h["ref_genome"]] == name
        assert h["matches"] == matches and h["fragments"] == frags
        assert abs(float(h["identity"]) - float(golden_io.f32(ident))) <= IDENTITY_TOL
        assert h["identity"] == golden_io.f32(ident), (float(h["identity"]).hex(), ident)


def _l1_parts(ix, l1, mean):
    """The large class of the on-chip L1 kernel whole ("chip-large"), cut into parts at genome boundaries (a fixed number,
    or what the library picks), or in parts too small for half of the fragments, which fall back to the whole shape."""
    if l1.startswith("parts"):
        ix.set_l1_small_cap(0)
    # the warp-per-fragment shape (fragments of at most 256 hits) stays on where the mode does not ask for a CTA shape
    ix.set_l1_tiny_cap(-1 if l1 in ("chip", "sort", "mixed") else 0)
    ix.set_l1_parts(*{"chip-large": (0,), "parts-3": (3,), "parts-8": (8,), "parts-overflow": (4, max(mean // 4, 0)), "parts-auto": (-1,)}.get(l1, (0,)))


@pytest.mark.parametrize("l1", ["chip", "chip-large", "shapes", "sort", "mixed", "parts-3", "parts-overflow"])
@pytest.mark.parametrize("name", [c["name"] for c in cases.query_cases()])
def test_queries_match_pyfastani_and_oracle(name, l1):
    """`l1` selects the L1 path: all fragments through the on-chip kernel (the default), all through
    the device-wide radix sort, or split by seed count."""
    case = next(c for c in cases.query_cases() if c["name"] == name)
    gold = golden_io.query_golden()[name]
    sk = capi.Sketch(**case["params"])
    osk = _port().sketch(**case["params"])
    for rname, contigs in case["refs"]:
        sk.add_draft(rname, contigs)
        osk.add_draft(rname, contigs)
    assert sk.counts()[0] == gold["minimizers"]
    assert sk.warnings == gold["ref_warnings"]
    ix = sk.index()
    osk.index()
    assert ix.params().window == gold["window"]
    assert ix.counts()[:2] == (gold["minimizers"], gold["unique"])
    for a, b in zip(ix.minimizers(), osk.minimizers()):
        assert np.array_equal(a, b)
    for q, res in zip(case["queries"], gold["results"]):
        ohits, oinfo = osk.query_draft(q, dump=True)
        st = oinfo["stats"]
        mean = st["seeds"] // max(st["fragments"], 1)
        ix.set_l1_seed_cap({"sort": 0, "mixed": mean - 1}.get(l1, -1))
        ix.set_l1_small_cap({"chip-large": 0, "shapes": mean}.get(l1, -1))
        _l1_parts(ix, l1, mean)
        hits, out = ix.query_draft(q, dump=True)
        if l1.startswith("parts") and len(case["refs"]) >= 2 and st["seeds"]:
            want = min(int(l1[6:]) if l1[6:].isdigit() else 4, len(case["refs"]))
            # (k = 12 at 90 %: minHits of the largest sketch the index supports exceeds the look-ahead of the 256-thread
            # shape, so neither the small shape nor the parts are used)
            assert out["info"]["l1_parts"] == (0 if name == "k12_pid90" else want)
        if l1 in ("chip", "chip-large", "shapes") or l1.startswith("parts"):
            assert out["info"]["l1_sorted_fragments"] == 0
            assert out["info"]["l1_tiny_fragments"] <= (st["fragments"] if l1 == "chip" else 0)
            if l1 == "chip-large":
                assert out["info"]["l1_small_fragments"] == 0
        elif st["seeds"]:
            assert out["info"]["l1_sorted_fragments"] > 0
        # L1 candidate regions, bit-exact
        assert np.array_equal(out["candidates"], oinfo["candidates"])
        # L2 mappings, every field (the oracle emits them in the same fragment/candidate order)
        assert np.array_equal(out["mappings"], oinfo["mappings"])
        info = out["info"]
        assert (info["fragments"], info["seeds"], info["candidates"], info["mappings"], info["sketch_sum"]) == \
               (st["fragments"], st["seeds"], st["candidates"], st["mappings"], st["sketch_sum"])
        assert info["scanned"] == st["scanned"]
        _check_hits(hits, ix.names, res["hits"])
        assert out["short_contigs"] == res["warnings"]
        assert info["kernel_launches"] > 0 or st["fragments"] == 0


@pytest.mark.parametrize("l1", ["chip", "chip-large", "shapes", "sort", "mixed", "small-72k", "small-110k", "parts-3", "parts-8",
                                "parts-auto", "parts-overflow"])
def test_l1_many_references(l1):
    """120 related references: every fragment has several thousand seed hits (more than one 4096-hit
    tile of the on-chip L1 kernel) spread over nine 2^16-minimizer chunks of the index, plus a
    reverse-complemented and a shuffled-contig query.  Candidates, mappings and hits vs the oracle."""
    q, refs, _ = synth.one_to_many(808, 120, 60_000, lo=0.85, hi=0.995)
    sk, osk = capi.Sketch(), _port().sketch()
    for i, r in enumerate(refs):
        contigs = [r] if i % 3 else [r[:25_000], r[25_000:25_900], r[25_900:]]
        sk.add_draft(i, contigs)
        osk.add_draft(i, contigs)
    ix = sk.index()
    osk.index()
    assert ix.counts()[0] > 8 * 65536
    for query in ([q], [synth.revcomp(q)], [q[30_000:], q[:30_000]]):
        ohits, oinfo = osk.query_draft(query, dump=True)
        st = oinfo["stats"]
        assert st["seeds"] // st["fragments"] > 4096
        mean = st["seeds"] // st["fragments"]
        ix.set_l1_seed_cap({"sort": 0, "mixed": mean - 1}.get(l1, -1))
        ix.set_l1_small_cap({"chip-large": 0, "shapes": mean}.get(l1, -1))
        # the small shape with three / two CTAs per SM (72 KB / 110 KB each): what an index of a few thousand genomes
        # selects because its chunk histogram leaves no room for the hits in 54 KB
        ix.set_l1_small_shape({"small-72k": 1, "small-110k": 2}.get(l1, -1))
        _l1_parts(ix, l1, mean)
        hits, out = ix.query_draft(query, dump=True)
        assert (out["info"]["l1_sorted_fragments"] == 0) == (l1 not in ("sort", "mixed"))
        if l1.startswith("parts"):
            assert out["info"]["l1_small_fragments"] == 0
            assert out["info"]["l1_parts"] == {"parts-3": 3, "parts-8": 8, "parts-overflow": 4}.get(l1, out["info"]["l1_parts"]) >= 2
        if l1 in ("chip", "small-72k", "small-110k"):          # several 1024-hit tiles per fragment in the small shape
            assert out["info"]["l1_small_fragments"] == st["fragments"]
        if l1 == "chip-large":
            assert out["info"]["l1_small_fragments"] == 0
        if l1 == "shapes":
            assert 0 < out["info"]["l1_small_fragments"] < st["fragments"]
        assert np.array_equal(out["candidates"], oinfo["candidates"])
        assert np.array_equal(out["mappings"], oinfo["mappings"])
        assert np.array_equal(hits, ohits)


def test_l1_warp_per_fragment_shape():
    """Fragments with a few dozen hits each (two references at 90 % / 88 %, one unrelated, a draft among them) all take
    the warp-per-fragment shape of the L1 kernel; with it switched off the CTA shapes give the same regions.  Both
    against the oracle, for a whole query, its reverse complement and a query cut into contigs."""
    rng = np.random.default_rng(31)
    base = synth.random_codes(rng, 150_000)
    q = synth.to_bytes(base)
    refs = [[synth.to_bytes(synth.mutate_codes(rng, base, 0.90))],
            synth.fragment(rng, synth.to_bytes(synth.mutate_codes(rng, base, 0.88)), 6, min_end=400),
            [synth.to_bytes(synth.random_codes(rng, 80_000))]]
    sk, osk = capi.Sketch(), _port().sketch()
    for i, r in enumerate(refs):
        sk.add_draft(i, r)
        osk.add_draft(i, r)
    ix = sk.index()
    osk.index()
    for query in ([q], [synth.revcomp(q)], synth.fragment(rng, q, 5, min_end=3100)):
        ohits, oinfo = osk.query_draft(query, dump=True)
        mean = oinfo["stats"]["seeds"] // oinfo["stats"]["fragments"]
        for cap in (-1, 0, mean):                                # all fragments, none, about half of them
            ix.set_l1_tiny_cap(cap)
            hits, out = ix.query_draft(query, dump=True)
            tiny, frags = out["info"]["l1_tiny_fragments"], out["info"]["fragments"]
            assert tiny == frags if cap == -1 else tiny == 0 if cap == 0 else 0 < tiny < frags
            assert np.array_equal(out["candidates"], oinfo["candidates"])
            assert np.array_equal(out["mappings"], oinfo["mappings"])
            assert np.array_equal(hits, ohits)
    ix.set_l1_tiny_cap(-1)


def test_config1_known_answers():
    """BASELINE config 1 (E. coli query vs Shigella draft reference) and the reference's own
    known answers (test_ani.py:47-91), both directions plus the self queries."""
    gold = golden_io.config1_golden()
    genomes = {n: golden_io.genome(n) for n in ("ecoli", "shigella")}
    for rname in ("shigella", "ecoli"):
        sk = capi.Sketch()
        sk.add_draft(rname, genomes[rname])
        h, s, w = sk.minimizers()
        ix = sk.index()
        for qname in ("ecoli", "shigella"):
            g = gold["%s_vs_%s" % (qname, rname)]
            assert ix.counts()[:2] == (g["minimizers"], g["unique"])
            assert hashlib.sha256(h.tobytes() + s.tobytes() + w.tobytes()).hexdigest() == g["sha256"]
            hits, _ = ix.query_draft(genomes[qname])
            _check_hits(hits, ix.names, g["hits"])


def test_config1_intermediates_vs_oracle():
    genomes = {n: golden_io.genome(n) for n in ("ecoli", "shigella")}
    sk = capi.Sketch(); sk.add_draft("shigella", genomes["shigella"]); ix = sk.index()
    osk = _port().sketch(); osk.add_draft("shigella", genomes["shigella"]); osk.index()
    hits, out = ix.query_draft(genomes["ecoli"], dump=True)
    ohits, oinfo = osk.query_draft(genomes["ecoli"], dump=True)
    assert len(oinfo["candidates"]) == 5038 and len(oinfo["mappings"]) == 4101
    assert np.array_equal(out["candidates"], oinfo["candidates"])
    assert np.array_equal(out["mappings"], oinfo["mappings"])
    assert np.array_equal(hits, ohits)
    assert float(hits[0]["identity"]).hex() == "0x1.86a7fa0000000p+6"


def test_lookup_index_views():
    q, refs, _ = synth.one_to_many(5, 3, 60_000)
    sk = capi.Sketch()
    for i, r in enumerate(refs):
        sk.add_genome(i, r)
    ix = sk.index()
    h, s, w = ix.minimizers()
    keys = ix.keys()
    assert np.array_equal(keys, np.unique(h))
    for key in list(keys[:20]) + list(keys[-20:]):
        ps, pw = ix.lookup(int(key))
        sel = h == key
        assert np.array_equal(ps, s[sel]) and np.array_equal(pw, w[sel])     # insertion order, winSketch.hpp:180-185
    ps, pw = ix.lookup(int(keys[0]) + 1 if keys[0] + 1 not in keys else 0xFFFFFFFE)
    assert len(ps) == 0


def test_full_size_properties():
    """Size-independent properties at a larger scale than the oracle is run on: self-ANI of
    unrelated genomes hits only itself; reverse complement and contig permutation of the query
    leave matches/identity of a single-contig reference unchanged."""
    rng = np.random.default_rng(31)
    genomes = [synth.to_bytes(synth.random_codes(rng, 1_000_000)) for _ in range(6)]
    sk = capi.Sketch()
    for i, g in enumerate(genomes):
        sk.add_genome(i, g)
    ix = sk.index()
    n_min, n_uniq, n_contigs, n_genomes = ix.counts()
    assert n_contigs == 6 and n_genomes == 6
    assert abs(n_min / 6e6 - 0.08) < 0.002                 # density 2/(w+1), SURVEY.md section 8
    for i, g in enumerate(genomes):
        hits, out = ix.query_genome(g)
        assert len(hits) == 1 and hits[0]["ref_genome"] == i
        assert hits[0]["fragments"] == 333 and hits[0]["matches"] >= 331
        assert hits[0]["identity"] > 99.9
        rc_hits, _ = ix.query_genome(synth.revcomp(g[:999_000]))
        assert len(rc_hits) == 1 and rc_hits[0]["ref_genome"] == i and rc_hits[0]["matches"] >= 330


def test_low_complexity_fragments_take_the_exact_fallback():
    """Fragments that are mostly poly-A have tiny sketches, so one bucket of the L2 state holds
    more than 127 reference-only hashes: the fast kernel hands those candidates to the exact
    fallback kernel.  Results must still equal the oracle's, field by field."""
    rng = np.random.default_rng(77)
    ref = synth.to_bytes(synth.random_codes(rng, 40_000))
    frags = []
    for i in range(8):
        off = 2_000 + 4_000 * i
        frags.append(ref[off:off + 40 + 15 * i] + b"A" * (3000 - 40 - 15 * i))
    query = b"".join(frags) + ref[1000:7000]
    sk = capi.Sketch(); sk.add_genome("r", ref); sk.add_genome("r2", synth.revcomp(ref)); ix = sk.index()
    osk = _port().sketch(); osk.add_genome("r", ref); osk.add_genome("r2", synth.revcomp(ref)); osk.index()
    hits, out = ix.query_genome(query, dump=True)
    ohits, oinfo = osk.query_genome(query, dump=True)
    assert out["info"]["l2_fallback"] > 0
    assert len(oinfo["candidates"]) > 8
    assert np.array_equal(out["candidates"], oinfo["candidates"])
    assert np.array_equal(out["mappings"], oinfo["mappings"])
    assert np.array_equal(hits, ohits)


def test_l2_early_stop_keeps_first_and_last_optimum():
    """The slide kernel stops once the sketch matches still to come cannot reach the best window.
    Cases built to catch a stop that comes too soon: tandem copies of the query fragments inside one
    candidate region (the last optimum lies in the second copy), a diverged copy behind an exact
    one, and an exact copy behind a diverged one.  Mappings field by field against the oracle,
    and the stop must really have fired (fewer events replayed than the lists hold)."""
    rng = np.random.default_rng(4242)
    base = synth.random_codes(rng, 30_000)
    q = synth.to_bytes(base)
    filler = lambda n: synth.to_bytes(synth.random_codes(rng, n))
    mut = lambda ident: synth.to_bytes(synth.mutate_codes(rng, base, ident))
    refs = [
        filler(5_000) + q + filler(700) + q + filler(5_000),                   # tandem exact copies, 700 bp apart
        filler(3_000) + q[:9_000] + filler(1_200) + mut(0.93)[:9_000] + filler(3_000),
        filler(3_000) + mut(0.90)[3_000:15_000] + filler(900) + q[3_000:15_000] + filler(8_000),
        mut(0.97), mut(0.85),
    ]
    sk, osk = capi.Sketch(), _port().sketch()
    for i, r in enumerate(refs):
        sk.add_genome(i, r)
        osk.add_genome(i, r)
    ix = sk.index()
    osk.index()
    for query in (q, synth.revcomp(q), q[1_234:28_000]):
        hits, out = ix.query_genome(query, dump=True)
        ohits, oinfo = osk.query_genome(query, dump=True)
        assert np.array_equal(out["candidates"], oinfo["candidates"])
        assert np.array_equal(out["mappings"], oinfo["mappings"])
        assert np.array_equal(hits, ohits)
        assert 0 < out["info"]["events_replayed"] < out["info"]["events"]


def test_l2_both_directions_from_the_middle():
    """The slide starts at the window in the middle of a candidate region and replays its events to the right, takes the
    start state again and replays them to the left, each side until the sketch matches still to come cannot reach the
    best window.  Regions built
    so that the optimum -- or the first / last of several equal optima -- lies left of, right of and far from that start
    window: graded copies in both orders, three tandem copies, partial copies at the region's edges, copies on the
    reverse strand, and a reference made of many short random pieces of the query (the seeds' centre is anywhere)."""
    rng = np.random.default_rng(777)
    base = synth.random_codes(rng, 24_000)
    q = synth.to_bytes(base)
    filler = lambda n: synth.to_bytes(synth.random_codes(rng, n))
    mut = lambda ident: synth.to_bytes(synth.mutate_codes(rng, base, ident))
    pieces = b"".join(q[a:a + int(rng.integers(400, 2_500))] + filler(int(rng.integers(50, 900)))
                      for a in rng.integers(0, 21_000, size=40))
    refs = [
        filler(4_000) + q + filler(300) + mut(0.95) + filler(300) + mut(0.90) + filler(4_000),      # best copy leftmost
        filler(4_000) + mut(0.90) + filler(300) + mut(0.95) + filler(300) + q + filler(4_000),      # best copy rightmost
        filler(2_000) + q + filler(500) + q + filler(500) + q + filler(2_000),                      # three equal optima
        q[:7_000] + filler(6_000) + q[17_000:],                                                     # copies at both edges
        filler(3_000) + synth.revcomp(mut(0.96)) + filler(200) + synth.revcomp(q) + filler(3_000),
        filler(1_000) + mut(0.88)[2_000:20_000] + filler(2_800) + mut(0.99)[2_000:20_000] + filler(1_000),
        pieces,
    ]
    sk, osk = capi.Sketch(), _port().sketch()
    for i, r in enumerate(refs):
        sk.add_genome(i, r)
        osk.add_genome(i, r)
    ix = sk.index()
    osk.index()
    for query in (q, synth.revcomp(q), q[777:23_000], mut(0.97)):
        hits, out = ix.query_genome(query, dump=True)
        ohits, oinfo = osk.query_genome(query, dump=True)
        assert np.array_equal(out["candidates"], oinfo["candidates"])
        assert np.array_equal(out["mappings"], oinfo["mappings"])
        assert np.array_equal(hits, ohits)
        # (the few regions of the `pieces` reference that span more than EV_RMAX = 1024 minimizers go to the exact kernel)
        assert out["info"]["l2_fallback"] * 10 < len(out["candidates"]) and 0 < out["info"]["events_replayed"]


@pytest.mark.parametrize("seed", range(6))
def test_l2_random_regions_against_the_oracle(seed):
    """Fuzz of the event path: references stitched from random slices of mutated copies of the query (both strands,
    random gaps), so candidate regions hold several loci of unequal quality in random order; every mapping field
    against the oracle."""
    rng = np.random.default_rng(9000 + seed)
    base = synth.random_codes(rng, 18_000)
    q = synth.to_bytes(base)
    refs = []
    for _ in range(6):
        parts = []
        for _ in range(int(rng.integers(3, 9))):
            copy = synth.to_bytes(synth.mutate_codes(rng, base, float(rng.uniform(0.82, 1.0))))
            a = int(rng.integers(0, 12_000))
            piece = copy[a:a + int(rng.integers(1_500, 6_000))]
            parts.append(synth.revcomp(piece) if rng.random() < 0.3 else piece)
            parts.append(synth.to_bytes(synth.random_codes(rng, int(rng.integers(0, 3_000)))))
        refs.append(b"".join(parts))
    sk, osk = capi.Sketch(), _port().sketch()
    for i, r in enumerate(refs):
        sk.add_genome(i, r)
        osk.add_genome(i, r)
    ix = sk.index()
    osk.index()
    hits, out = ix.query_genome(q, dump=True)
    ohits, oinfo = osk.query_genome(q, dump=True)
    assert np.array_equal(out["candidates"], oinfo["candidates"])
    assert np.array_equal(out["mappings"], oinfo["mappings"])
    assert np.array_equal(hits, ohits)


def test_l1_irregular_blocks_next_to_candidate_loci():
    """The L1 kernel decides "closer than a fragment" from index distances wherever every step between neighbouring
    minimizers is at most one window, and gathers positions only in blocks of 1024 minimizers the index marked
    otherwise: contig ends and runs without minimizers (every k-mer of an (AT)n run is its own reverse complement and
    is skipped).  Loci cut by such runs and by contig ends a few hundred bases apart, against the oracle."""
    rng = np.random.default_rng(4242)
    base = synth.random_codes(rng, 40_000)
    q = synth.to_bytes(base)
    mut = lambda ident: synth.to_bytes(synth.mutate_codes(rng, base, ident))
    at = lambda n: b"AT" * n
    filler = lambda n: synth.to_bytes(synth.random_codes(rng, n))
    g1, g2, g3 = mut(0.97), mut(0.93), mut(0.99)
    refs = [
        [g1[:6_000] + at(400) + g1[6_000:9_000] + at(60) + g1[9_000:20_000] + at(1_700) + g1[20_000:]],
        [g2[:7_500], g2[7_500:7_900], at(300) + g2[7_900:16_000] + at(300), g2[16_000:16_050], g2[16_050:]],
        [filler(1_000) + g3[10_000:13_100] + at(1_450) + g3[13_100:30_000] + filler(500), at(2_000), g3[:10_000]],
    ]
    sk, osk = capi.Sketch(), _port().sketch()
    for i, r in enumerate(refs):
        sk.add_draft(i, r)
        osk.add_draft(i, r)
    ix = sk.index()
    osk.index()
    for query in ([q], [synth.revcomp(q)], [q[:11_000] + at(700) + q[11_000:]]):
        for cap in (-1, 0):                                     # the small and the large shape of the kernel
            ix.set_l1_small_cap(cap)
            hits, out = ix.query_draft(query, dump=True)
            ohits, oinfo = osk.query_draft(query, dump=True)
            assert out["info"]["l1_sorted_fragments"] == 0
            assert np.array_equal(out["candidates"], oinfo["candidates"])
            assert np.array_equal(out["mappings"], oinfo["mappings"])
            assert np.array_equal(hits, ohits)
    ix.set_l1_small_cap(-1)


def test_config2_shaped_drafts_vs_reference():
    """BASELINE configs[2] at its own shape, few genomes: six fragmented assemblies of 2.0-2.6 Mbp in 200-320 contigs
    (two species 88 % apart, strains at 96-99.7 %, half of the contigs reverse-complemented, order permuted; contig
    ends down to 200 bp), all-vs-all through add_draft / query_draft and one fa_query_batch call: every hit row against
    the CPU reference (the compiled reference when present, else the C port)."""
    from oracle.oracle import Oracle, available
    rng = np.random.default_rng(2024)
    root = synth.random_codes(rng, 2_000_000)
    species = [root, synth.mutate_codes(rng, root, 0.88)]
    drafts = []
    for sp in species:
        for ident in (0.997, 0.98, 0.96):
            codes = synth.mutate_codes(rng, sp, ident)
            extra = synth.random_codes(rng, int(rng.integers(0, 600_000)))        # strain-specific DNA: 2.0-2.6 Mbp
            g = synth.to_bytes(np.concatenate([codes, extra]))
            drafts.append(synth.fragment(rng, g, int(rng.integers(200, 321)), min_end=200))
    assert all(200 <= len(d) <= 320 and sum(map(len, d)) >= 2_000_000 for d in drafts)
    kind = "reference" if "reference" in available() else "port"
    sk, osk = capi.Sketch(), Oracle(kind).sketch()
    for i, d in enumerate(drafts):
        sk.add_draft(i, d)
        osk.add_draft(i, d)
    ix = sk.index()
    osk.index()
    batch, _ = ix.query_batch(drafts)
    for i, d in enumerate(drafts):
        ohits, _ = osk.query_draft(d, threads=8) if kind == "reference" else osk.query_draft(d)
        hits = ix.query_draft(d)[0]
        assert np.array_equal(hits, ohits) and np.array_equal(batch[i], ohits)
        assert len(hits) == 6 and hits[0]["ref_genome"] == i


def test_concurrent_queries_on_one_index():
    """Mapper.query_* may be called from several threads at once (pyx:1158-1161: the reference maps under `nogil` with
    its own state per call).  Here the calls of one index share a workspace and are serialised inside the library:
    four threads, three rounds over six queries each, every result equal to the single-threaded one."""
    import threading
    q, refs, _ = synth.one_to_many(77, 12, 120_000, lo=0.85, hi=0.99)
    rng = np.random.default_rng(5)
    queries = [q, synth.revcomp(q), q[10_000:90_000], refs[3], refs[7][5_000:], synth.to_bytes(synth.random_codes(rng, 50_000))]
    sk = capi.Sketch()
    for i, r in enumerate(refs):
        sk.add_genome(i, r)
    ix = sk.index()
    want = [ix.query_genome(x)[0] for x in queries]
    got, errors = {}, []

    def work(t):
        try:
            for rep in range(3):
                for j in np.random.default_rng(t).permutation(len(queries)):
                    got[(t, rep, int(j))] = ix.query_genome(queries[int(j)])[0]
        except Exception as e:                                   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    assert len(got) == 4 * 3 * len(queries)
    for (t, rep, j), h in got.items():
        assert np.array_equal(h, want[j]), (t, rep, j)


def test_indexes_release_their_memory():
    """every buffer of an index, its workspace included, is released by fa_index_free -- twenty indexes built,
    queried and dropped leave the device where it was."""
    import gc
    q, refs, _ = synth.one_to_many(5, 8, 150_000, lo=0.85, hi=0.99)

    def cycle():
        sk = capi.Sketch()
        for i, r in enumerate(refs):
            sk.add_genome(i, r)
        ix = sk.index()
        assert len(ix.query_genome(q)[0]) == 8
        del ix, sk
        gc.collect()

    cycle()
    free0 = capi.mem_info()[0]
    for _ in range(20):
        cycle()
    assert abs(int(capi.mem_info()[0]) - int(free0)) < (64 << 20), (free0, capi.mem_info())


def test_query_batch_equals_single_queries():
    """fa_query_batch: CSR hit rows of many queries in one call == fa_query per query; counters add up."""
    q, refs, _ = synth.one_to_many(515, 6, 80_000, lo=0.84, hi=0.99)
    sk = capi.Sketch()
    for i, r in enumerate(refs):
        sk.add_genome(i, r)
    ix = sk.index()
    rng = np.random.default_rng(6)
    queries = [[q], synth.fragment(rng, q, 4, min_end=500), [refs[3]], [q[:2000], b"ACGTACGT"], [], [synth.revcomp(refs[5])]]
    singles, frags, cands = [], 0, 0
    for qq in queries:
        h, out = ix.query_draft(qq)
        singles.append(h)
        frags += out["info"]["fragments"]; cands += out["info"]["candidates"]
    batch, info = ix.query_batch(queries)
    assert len(batch) == len(queries)
    for a, b in zip(batch, singles):
        assert np.array_equal(a, b)
    assert info["fragments"] == frags and info["candidates"] == cands and info["short_contigs"] == 1
    assert ix.query_batch([])[0] == []


# ---- protein mode (pyx:225-309, 548-550) ----------------------------------------
@pytest.mark.parametrize("idx", range(len(cases.protein_minimizer_cases())))
def test_protein_minimizers_match_pyfastani(idx):
    case = cases.protein_minimizer_cases()[idx]
    gold, arr = golden_io.protein_golden()
    man = gold["minimizers"][idx]
    sk = capi.Sketch(**case["params"])
    sk.add_draft("g", case["contigs"])
    gh, gs, gw = sk.minimizers()
    assert len(gh) == man["n"]
    assert np.array_equal(gh, arr["h%d" % idx]) and np.array_equal(gs, arr["s%d" % idx]) and np.array_equal(gw, arr["w%d" % idx])
    assert sk.warnings == man["warnings"]


@pytest.mark.parametrize("idx", range(len(cases.protein_query_cases())))
def test_protein_queries_match_pyfastani_and_oracle(idx):
    case = cases.protein_query_cases()[idx]
    gold = golden_io.protein_golden()[0]["queries"][idx]
    sk, osk = capi.Sketch(**case["params"]), _port().sketch(**case["params"])
    for rname, contigs in case["refs"]:
        sk.add_draft(rname, contigs)
        osk.add_draft(rname, contigs)
    ix = sk.index()
    osk.index()
    assert ix.counts()[0] == gold["minimizers"] and ix.counts()[1] == gold["unique"]
    for q, res in zip(case["queries"], gold["results"]):
        hits, out = ix.query_draft(q, dump=True)
        ohits, oinfo = osk.query_draft(q, dump=True)
        assert np.array_equal(out["candidates"], oinfo["candidates"])
        assert np.array_equal(out["mappings"], oinfo["mappings"])
        assert np.array_equal(hits, ohits)
        assert [[ix.names[h["ref_genome"]], int(h["matches"]), int(h["fragments"])] for h in hits] == [[r[0], r[2], r[3]] for r in res["hits"]]
        assert [h["identity"] for h in hits] == [golden_io.f32(r[1]) for r in res["hits"]]
        assert out["short_contigs"] == res["warnings"]


def test_protein_bgc_known_answer():
    """The reference's own protein test (test_ani.py:96-115): 130 / 176 under both names."""
    gold = golden_io.protein_golden()[0]["bgc"]
    bgc = {n: golden_io.genome(n) for n in ("BGC0001425", "BGC0001427", "BGC0001428")}
    sk = capi.Sketch(protein=True, fragment_length=100)
    sk.add_draft("BGC0001425", bgc["BGC0001425"])
    sk.add_draft("BGC0001427", bgc["BGC0001425"])
    ix = sk.index()
    assert ix.counts()[0] == gold["minimizers"] and ix.counts()[1] == gold["unique"]
    hits, _ = ix.query_draft(bgc["BGC0001428"])
    assert [(ix.names[h["ref_genome"]], int(h["matches"]), int(h["fragments"])) for h in hits] == \
           [("BGC0001425", 130, 176), ("BGC0001427", 130, 176)]
    assert [h["identity"] for h in hits] == [golden_io.f32(r[1]) for r in gold["as_in_test_ani"]]


def test_draft_all_vs_all_clusters():
    """BASELINE configs 3 / 4 in the small: two 'species' (85 % apart) of five 'strains' each (95-99.9 %),
    every genome cut into 40 contigs, half of them reverse-complemented, order permuted; all ten as
    references (add_draft) and as queries (query_draft, one fa_query_batch call).  Candidates,
    mappings and hits of every query against the oracle; the batch equals the single queries."""
    rng = np.random.default_rng(303)
    root = synth.random_codes(rng, 180_000)
    species = [root, synth.mutate_codes(rng, root, 0.85)]
    drafts = []
    for sp in species:
        for ident in (0.999, 0.99, 0.98, 0.965, 0.95):
            g = synth.to_bytes(synth.mutate_codes(rng, sp, ident))
            drafts.append(synth.fragment(rng, g, 40, min_end=300))
    sk, osk = capi.Sketch(), _port().sketch()
    for i, d in enumerate(drafts):
        sk.add_draft(i, d)
        osk.add_draft(i, d)
    ix = sk.index()
    osk.index()
    batch, _ = ix.query_batch(drafts)
    for i, d in enumerate(drafts):
        hits, out = ix.query_draft(d, dump=True)
        ohits, oinfo = osk.query_draft(d, dump=True)
        assert np.array_equal(out["candidates"], oinfo["candidates"])
        assert np.array_equal(out["mappings"], oinfo["mappings"])
        assert np.array_equal(hits, ohits) and np.array_equal(batch[i], ohits)
        assert len(hits) == 10 and hits[0]["ref_genome"] == i and hits[0]["identity"] > 99.9
        same = {int(h["ref_genome"]) // 5 for h in hits[:5]}
        assert same == {i // 5}                              # the own species ranks first


def test_many_to_many_tree_vs_reference():
    """BASELINE config 4 in the small: 24 genomes of 0.4-0.8 Mbp on a two-level tree (80-100 % ANI inside a clade, unrelated across clades),
    a third of them as drafts, all-vs-all = 576 genome pairs in one fa_query_batch call.  Every hit
    row (reference id, matches, fragments, identity bits, order) against the CPU reference
    (the compiled reference when present, else the C port)."""
    from oracle.oracle import available
    rng = np.random.default_rng(44)
    genomes = []
    for clade in range(4):
        anc = synth.random_codes(rng, int(rng.integers(400_000, 800_000)))
        for sp in range(2):
            spc = synth.mutate_codes(rng, anc, float(rng.uniform(0.91, 0.97)))
            for strain in range(3):
                g = synth.to_bytes(synth.mutate_codes(rng, spc, float(rng.uniform(0.95, 0.999))))
                genomes.append(synth.fragment(rng, g, int(rng.integers(20, 60)), min_end=200) if len(genomes) % 3 == 0 else [g])
    kind = "reference" if "reference" in available() else "port"
    sk, osk = capi.Sketch(), Oracle(kind).sketch()
    for i, g in enumerate(genomes):
        sk.add_draft(i, g)
        osk.add_draft(i, g)
    ix = sk.index()
    osk.index()
    batch, info = ix.query_batch(genomes)
    kw = {"threads": 0} if kind == "reference" else {}
    pairs = 0
    for i, g in enumerate(genomes):
        ohits, _ = osk.query_draft(g, **kw)
        assert np.array_equal(batch[i], ohits), i
        assert ohits[0]["ref_genome"] == i
        pairs += len(ohits)
    assert pairs >= 24 * 5 and info["l2_fallback"] == 0


@pytest.mark.parametrize("params", [dict(fragment_length=40), dict(fragment_length=30, k=12), dict(fragment_length=24, k=16),
                                    dict(fragment_length=64, k=20, percentage_identity=70.0)])
def test_tiny_fragments_vs_oracle(params):
    """Very short fragments: the L2 window (fragment_length - (w - 1) - (k - 1) positions) shrinks to a few
    positions, down to the sizes where the event path of L2 hands over to the exact kernel.  Against the oracle."""
    rng = np.random.default_rng(9)
    base = synth.random_codes(rng, 6_000)
    refs = [synth.to_bytes(base), synth.to_bytes(synth.mutate_codes(rng, base, 0.97)), synth.to_bytes(synth.random_codes(rng, 5_000))]
    query = synth.to_bytes(synth.mutate_codes(rng, base, 0.99))[500:4_500]
    sk, osk = capi.Sketch(**params), _port().sketch(**params)
    for i, r in enumerate(refs):
        sk.add_genome(i, r)
        osk.add_genome(i, r)
    ix = sk.index()
    osk.index()
    hits, out = ix.query_genome(query, dump=True)
    ohits, oinfo = osk.query_genome(query, dump=True)
    assert np.array_equal(out["candidates"], oinfo["candidates"])
    assert np.array_equal(out["mappings"], oinfo["mappings"])
    assert np.array_equal(hits, ohits)

"""Case definitions shared by the golden-vector generator (tests/golden/make_golden.py,
run against the real pyfastani) and the parity tests (run against the oracle and the
CUDA path).  Everything is regenerated from seeds, so only outputs are stored."""
import numpy as np

import synth


def _rand(seed, n):
    return synth.to_bytes(synth.random_codes(np.random.default_rng(seed), n))


def minimizer_cases():
    """Edge cases of byte normalisation, hashing and winnowing."""
    base = _rand(1, 5000)
    cases = []

    def add(name, contigs, **params):
        cases.append({"name": name, "contigs": contigs, "params": params})

    add("random5000", [base])
    add("lowercase", [base.lower()])
    add("str", [base.decode()])
    add("mixedcase", [bytes(b | 0x20 if i % 3 == 0 else b for i, b in enumerate(base))])
    ucs2 = base[:1000].decode() + "ĀŁ" + base[1000:3000].decode() + "€" + base[3000:].decode()
    add("ucs2", [ucs2])
    ucs4 = base[:700].decode() + "\U0001F600" + base[700:2500].decode()
    add("ucs4", [ucs4])
    nrun = bytearray(base); nrun[1200:1700] = b"N" * 500
    add("n_run", [bytes(nrun)])
    add("all_n", [b"N" * 100])
    add("poly_a_start", [b"A" * 100])
    add("palindrome_at", [b"AT" * 50])
    add("prefix_then_poly_a", [base[:300] + b"A" * 200 + base[300:600]])
    add("poly_a_long_start", [b"A" * 5000 + base[:500]])
    add("two_tandem", [base[:500] * 6])
    iupac = bytearray(base[:4200])
    for i, c in enumerate(b"RYKMBDHVSWNUrykmbdhvswnu"):
        iupac[50 + 97 * i] = c
        iupac[4100 + 3 * i] = c
    add("iupac", [bytes(iupac)])
    # non-letters: SSE2 `& ~0x20` on full 16-byte chunks vs toupper on the tail of each 2048 block
    odd = bytearray(base[:5000])          # blocks: 2048, 2048, 904 = 56*16 + 8 tail bytes
    for i, c in enumerate(b"0123456789-*.{|}~`@[]^_ \t\n\x0b\x1b\x00\x7f"):
        odd[100 + 61 * i] = c             # chunk region of block 0
        odd[2048 + 37 * i] = c            # chunk region of block 1
    for i, c in enumerate(b"1-{.z9a~"):
        odd[4992 + i] = c                 # toupper tail of the last block
    for i, c in enumerate(bytes([0x80, 0xC1, 0xE1, 0xF4, 0xFF, 0xA0, 0x9b, 0xd4])):
        odd[3000 + 53 * i] = c
        odd[4984 + i] = c
    add("odd_bytes", [bytes(odd)])
    add("odd_bytes_tail_only", [bytes(odd[4096:])])
    for n in (16, 23, 24, 38, 39, 40, 41, 2047, 2048, 2049, 2063, 2064, 4096, 4111):
        add("len%d" % n, [_rand(100 + n, n)])
    add("draft_short_ids", [_rand(7, 3000), b"ACGT", _rand(8, 30), b"", _rand(9, 2500), b"ACGTACGTACGTACGTACGTACGTACG"])
    add("k8", [base], k=8)
    add("k11_frag1000", [base], k=11, fragment_length=1000)
    add("k5", [base[:2000]], k=5)
    add("k17", [base], k=17)
    add("k21", [base], k=21)
    add("k32", [base], k=32)
    add("k33_odd", [bytes(odd)], k=33)
    add("frag500", [base], fragment_length=500)
    add("frag10000_pid95", [base], fragment_length=10000, percentage_identity=95.0)
    add("numpy_view", [np.frombuffer(base, dtype=np.uint8)])
    add("bytearray", [bytearray(base)])
    return cases


def query_cases():
    """Whole-path cases: refs, queries, params.  Queries are lists of contigs."""
    cases = []

    # config-2 in miniature
    q, refs, idents = synth.one_to_many(12345, 8, 200_000)
    cases.append({"name": "one_to_many_8x200k", "params": {},
                  "refs": [("ref%02d" % i, [r]) for i, r in enumerate(refs)],
                  "queries": [[q], [q.lower()], [q.decode()], [q[:2999]], [q[:3000]], [q[:2999], b"ACGT"],
                              [q[100_000:130_000], b"ACGTACGTAC", q[:50_000]]]})

    # config-3/4 in miniature: drafts, strand flips, all-vs-all
    drafts = synth.clustered_drafts(2026, 3, 3)
    cases.append({"name": "drafts_3x3", "params": {}, "refs": drafts, "queries": [c for _, c in drafts]})

    # SURVEY B.12 boundary case
    refs, query = synth.boundary_case(99)
    cases.append({"name": "boundary99", "params": {}, "refs": refs, "queries": [query]})

    # embedded genome + minimum-fraction behaviour
    g = _rand(5, 60_000)
    junk = _rand(6, 300_000)
    cases.append({"name": "embedded", "params": {},
                  "refs": [("g", [g]), ("junk", [_rand(11, 80_000)])],
                  "queries": [[g], [junk[:150_000] + g + junk[150_000:]], [g[:30_000]], [synth.revcomp(g)],
                              [g[:10_000] + b"N" * 500 + g[10_500:]]]})

    # non-default parameters
    q, refs, _ = synth.one_to_many(77, 5, 120_000, lo=0.85, hi=0.99)
    for name, params in (("frag1000", {"fragment_length": 1000}),
                         ("k12_pid90", {"k": 12, "percentage_identity": 90.0}),
                         ("minfrac05", {"minimum_fraction": 0.5}),
                         ("frag5000_k14", {"fragment_length": 5000, "k": 14})):
        cases.append({"name": name, "params": params,
                      "refs": [("r%d" % i, [r]) for i, r in enumerate(refs)], "queries": [[q], [refs[2]]]})

    # repeats inside one reference + duplicated contigs (ties, duplicate hashes in a window)
    rng = np.random.default_rng(4242)
    unit = synth.random_codes(rng, 3500)
    rep = np.concatenate([unit, synth.random_codes(rng, 800), unit, unit, synth.random_codes(rng, 20_000),
                          synth.mutate_codes(rng, unit, 0.97), synth.random_codes(rng, 10_000)])
    repb = synth.to_bytes(rep)
    cases.append({"name": "repeats", "params": {},
                  "refs": [("rep", [repb, repb[:20_000], synth.revcomp(repb[5_000:30_000])]),
                           ("rep2", [synth.to_bytes(synth.mutate_codes(rng, rep, 0.93))])],
                  "queries": [[repb], [synth.to_bytes(unit) * 3], [repb[1000:]]]})
    return cases


# ---- protein mode (reference: pyx:225-309, 548-550, test_ani.py:96-115) ------------------
AA = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)


def _rand_aa(rng, n):
    return AA[rng.integers(0, 20, size=n)].tobytes()


def _mutate_aa(rng, seq, identity):
    a = np.frombuffer(seq, dtype=np.uint8).copy()
    hit = rng.random(a.size) < (1.0 - identity)
    a[hit] = AA[rng.integers(0, 20, size=int(hit.sum()))]
    return a.tobytes()


def protein_minimizer_cases():
    rng = np.random.default_rng(2020)
    cases = []

    def add(name, contigs, **params):
        cases.append({"name": name, "contigs": contigs, "params": dict(protein=True, **params)})

    p1 = _rand_aa(rng, 5000)
    add("prot_default", [p1], fragment_length=100)
    add("prot_k5", [p1[:3000]], k=5, fragment_length=100)
    add("prot_lower_mixed", [p1[:2500].lower(), p1[2500:4133]], fragment_length=100)       # blocks of 2048, 16-byte chunks + tails
    add("prot_str", [p1[:2100].decode(), p1[:700].decode().lower()], fragment_length=100)
    add("prot_ucs2", [p1[:1500].decode() + "Δ" + p1[1500:2300].decode()], fragment_length=100)
    add("prot_x_runs", [p1[:300] + b"X" * 40 + p1[300:600] + b"*" + p1[600:900], b"M" * 64, b"MK" * 40], fragment_length=100)
    add("prot_short", [p1[:15], p1[:16], p1[:17], b""], fragment_length=100)
    add("prot_non_letters", [p1[:100] + b"-.*12" + p1[100:2050] + b"[]{}|" + p1[2050:2063]], fragment_length=100)
    return cases


def protein_query_cases():
    rng = np.random.default_rng(2021)
    fam = [_rand_aa(rng, int(n)) for n in rng.integers(150, 1200, size=40)]
    close = [_mutate_aa(rng, p, 0.93) for p in fam]
    far = [_mutate_aa(rng, p, 0.80) for p in fam]
    other = [_rand_aa(rng, int(n)) for n in rng.integers(150, 1200, size=40)]
    shuffled = [close[i] for i in rng.permutation(len(close))]
    cases = [{"name": "prot_families", "params": dict(protein=True, fragment_length=100),
              "refs": [("fam", fam), ("close", shuffled), ("far", far), ("other", other)],
              "queries": [fam, close, other, fam[:5], [fam[0][:99]]]},
             {"name": "prot_k7_frag60", "params": dict(protein=True, fragment_length=60, k=7, percentage_identity=70.0),
              "refs": [("fam", fam[:20]), ("far", far[:20])],
              "queries": [close[:20], far[5:15]]}]
    return cases

"""Deterministic synthetic genomes for the parity tests, golden vectors and bench.py.

Generators follow SURVEY.md section 8(d): uniform i.i.d. ACGT, substitution-only
mutation x' = (x + U{1,2,3}) mod 4 at per-base rate 1 - identity, numpy
``default_rng(seed)`` (PCG64, stable across numpy versions).
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = np.arange(256)
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[a] = b


def random_codes(rng, n):
    return rng.integers(0, 4, size=n, dtype=np.uint8)


def mutate_codes(rng, codes, identity):
    out = codes.copy()
    hit = rng.random(codes.size) < (1.0 - identity)
    out[hit] = (out[hit] + rng.integers(1, 4, size=int(hit.sum()), dtype=np.uint8)) & 3
    return out


def to_bytes(codes):
    return ACGT[codes].tobytes()


def revcomp(seq):
    a = np.frombuffer(seq, dtype=np.uint8)
    return _COMP[a[::-1]].tobytes()


def fragment(rng, seq, n_contigs, flip=True, permute=True, min_end=1000):
    """Cut `seq` at uniformly random points into n_contigs pieces."""
    n = len(seq)
    if n_contigs <= 1 or n < 2 * min_end + n_contigs:
        return [seq]
    cuts = np.sort(rng.choice(np.arange(min_end, n - min_end), size=n_contigs - 1, replace=False))
    edges = [0] + [int(c) for c in cuts] + [n]
    parts = [seq[edges[i]:edges[i + 1]] for i in range(n_contigs)]
    if flip:
        parts = [revcomp(p) if rng.random() < 0.5 else p for p in parts]
    if permute:
        parts = [parts[i] for i in rng.permutation(len(parts))]
    return parts


def one_to_many(seed, n_refs, length, lo=0.80, hi=0.99):
    """Config 2: base genome as query; n_refs independent mutations of it at linspace(lo, hi)."""
    rng = np.random.default_rng(seed)
    base = random_codes(rng, length)
    idents = np.linspace(lo, hi, n_refs)
    refs = [to_bytes(mutate_codes(rng, base, float(i))) for i in idents]
    return to_bytes(base), refs, idents


def ref_stream(seed, n_refs, length, lo=0.80, hi=0.99):
    """Config 2 as a generator (constant memory): yields (query) first, then each reference."""
    rng = np.random.default_rng(seed)
    base = random_codes(rng, length)
    yield to_bytes(base)
    for i in np.linspace(lo, hi, n_refs):
        yield to_bytes(mutate_codes(rng, base, float(i)))


def clustered_drafts(seed, n_species, strains, length_range=(60_000, 120_000), contigs_range=(5, 20),
                     species_identity=(0.80, 0.90), strain_identity=(0.95, 0.999), min_end=300):
    """Configs 3/4 in miniature: a root per genus, species at 0.80-0.90 from it, strains at
    0.95-0.999 from their species; each genome cut into contigs with random strand flips
    and permuted order.  Returns a list of (name, [contigs])."""
    rng = np.random.default_rng(seed)
    root = random_codes(rng, int(rng.integers(*length_range)))
    out = []
    for sp in range(n_species):
        anc = mutate_codes(rng, root, float(rng.uniform(*species_identity)))
        for st in range(strains):
            g = to_bytes(mutate_codes(rng, anc, float(rng.uniform(*strain_identity))))
            nc = int(rng.integers(contigs_range[0], contigs_range[1] + 1))
            out.append(("sp%d_st%d" % (sp, st), fragment(rng, g, nc, min_end=min_end)))
    return out


def boundary_case(seed=99):
    """repeats, tiny contigs, strand flips, an N-run and a poly-A run."""
    rng = np.random.default_rng(seed)
    root = random_codes(rng, 400_000)
    rep = random_codes(rng, 1500)
    for p in sorted(rng.integers(10_000, 390_000, size=5)):
        root = np.concatenate([root[:p], rep, root[p:]])
    refs = []
    for ident, nc in ((0.97, 120), (0.88, 200), (0.82, 60)):
        g = to_bytes(mutate_codes(rng, root, ident))
        n = len(g)
        cuts = np.sort(rng.choice(np.arange(1, n), size=nc - 1, replace=False))
        # force a few tiny contigs (2 .. 30 bp)
        cuts[1] = cuts[0] + 2
        cuts[5] = cuts[4] + 19
        cuts[9] = cuts[8] + 30
        cuts = np.sort(cuts)
        edges = [0] + [int(c) for c in cuts] + [n]
        parts = [g[edges[i]:edges[i + 1]] for i in range(nc) if edges[i + 1] > edges[i]]
        parts = [revcomp(p) if rng.random() < 0.5 else p for p in parts]
        parts = [parts[i] for i in rng.permutation(len(parts))]
        refs.append(("ref%.2f" % ident, parts))
    q = bytearray(to_bytes(mutate_codes(rng, root, 0.995)))
    q[50_000:50_040] = b"N" * 40
    q[120_000:120_060] = b"A" * 60
    query = fragment(rng, bytes(q), 90, min_end=500)
    return refs, query

"""Synthetic genome collections for bench_many.py (BASELINE configs[2]-[4]) that are reproducible ANYWHERE: every base is
a pure function of (seed, node, position) through a 64-bit integer hash, evaluated with torch on the GPU when the
collection is generated for a run and with numpy on the CPU when a sample of it is regenerated to check the run's hit
rows against the CPU reference (`bench_many.py --verify`) -- a box without a GPU can verify a run of eight.

Mutation model of SURVEY.md 8(d): substitutions only, x' = (x + U{1,2,3}) mod 4 at per-base rate 1 - identity.

Two trees:

* ``genus``    GTDB-like: independent genus roots (3-6 Mbp) -> species at 0.80-0.90 of the root -> strains at 0.95-0.999 of
               their species.  Pairs inside a genus span 64-100 % identity, pairs of different genera are unrelated
               sequence (FastANI reports nothing for them), which is what an all-vs-all over a reference database looks
               like: most of the pairs cost only their chance seed hits.
* ``related``  BASELINE configs[3] as written (every pair at 75-100 %): ONE 6 Mbp root; families = a 3-6 Mbp circular window
               of it at 0.92-0.94; genera at 0.95-0.97 of the family; species at 0.975-0.99; strains at 0.99-0.999.  Two
               genomes of different families share the overlap of their windows at about 0.75-0.78.
"""
import numpy as np

M64 = (1 << 64) - 1
C1, C2, GOLD = 0xBF58476D1CE4E5B9, 0x94D049BB133111EB, 0x9E3779B97F4A7C15


def _s64(x):
    """The 64-bit pattern `x` as a signed Python int (what torch int64 scalars accept)."""
    x &= M64
    return x - (1 << 64) if x >> 63 else x


def node_key(seed, *path):
    """A well-mixed 64-bit key for a node of the tree (pure Python, both back ends use the same number)."""
    z = (seed * GOLD + 0x1234567) & M64
    for p in path:
        z = (z ^ (int(p) + 0x632BE59BD9B4E019)) & M64
        z = ((z ^ (z >> 30)) * C1) & M64
        z = ((z ^ (z >> 27)) * C2) & M64
        z ^= z >> 31
    return z


class NumpyBackend:
    name = "numpy"

    def arange(self, n):
        return np.arange(n, dtype=np.uint64)

    def mix(self, idx, key):
        with np.errstate(over="ignore"):
            z = idx + np.uint64(key)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(C1)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(C2)
            return z ^ (z >> np.uint64(31))

    def field(self, h, shift, bits):
        return ((h >> np.uint64(shift)) & np.uint64((1 << bits) - 1)).astype(np.int64)

    def to_codes(self, x):
        return x.astype(np.uint8)

    def where_add(self, codes, hit, shift):
        return ((codes.astype(np.int64) + np.where(hit, shift, 0)) & 3).astype(np.uint8)

    def roll_window(self, codes, start, length):
        n = codes.shape[0]
        idx = (np.arange(length, dtype=np.int64) + start) % n
        return codes[idx]


class TorchBackend:
    name = "torch"

    def __init__(self, torch, device):
        self.t, self.dev = torch, device

    def arange(self, n):
        return self.t.arange(n, dtype=self.t.int64, device=self.dev)

    def _lsr(self, z, s):                                   # logical shift right of an int64 tensor
        return (z >> s) & ((1 << (64 - s)) - 1)

    def mix(self, idx, key):
        z = idx + _s64(key)
        z = (z ^ self._lsr(z, 30)) * _s64(C1)
        z = (z ^ self._lsr(z, 27)) * _s64(C2)
        return z ^ self._lsr(z, 31)

    def field(self, h, shift, bits):
        return (h >> shift) & ((1 << bits) - 1)

    def to_codes(self, x):
        return x.to(self.t.uint8)

    def where_add(self, codes, hit, shift):
        return ((codes.to(self.t.int64) + hit.to(self.t.int64) * shift) & 3).to(self.t.uint8)

    def roll_window(self, codes, start, length):
        n = codes.shape[0]
        idx = (self.t.arange(length, dtype=self.t.int64, device=self.dev) + start) % n
        return codes[idx]


def random_codes(be, key, n):
    return be.to_codes(be.field(be.mix(be.arange(n), key), 33, 2))


def mutate(be, codes, identity, key):
    """Substitute each base with probability 1 - identity (24-bit threshold), by 1, 2 or 3 modulo 4."""
    h = be.mix(be.arange(codes.shape[0]), key)
    thr = int(round((1.0 - float(identity)) * (1 << 24)))
    hit = be.field(h, 40, 24) < thr
    shift = be.field(h, 8, 20) % 3 + 1
    return be.where_add(codes, hit, shift)


class Collection:
    """G genomes of a tree; `codes(be, i)` is genome i as 2-bit codes on the back end `be` (ancestors are cached per back
    end, so generating the genomes in order costs one mutation pass each)."""

    def __init__(self, tree, genomes, seed, species=8, genus=4, scale=1):
        if tree not in ("genus", "related"):
            raise ValueError("tree must be 'genus' or 'related'")
        self.tree, self.G, self.seed = tree, int(genomes), int(seed)
        lo, hi = 3_000_000 // scale, 6_000_000 // scale            # (scale > 1: small genomes for the tests)
        rng = np.random.default_rng(seed)
        self.plan = []                       # per genome: list of (node path, kind, parameters) from the root down
        if tree == "genus":
            per_genus = genus * species
            for g in range((self.G + per_genus - 1) // per_genus):
                length = int(rng.integers(lo, hi + 1))
                for sp in range(genus):
                    sp_id = float(rng.uniform(0.80, 0.90))
                    for st in range(species):
                        st_id = float(rng.uniform(0.95, 0.999))
                        if len(self.plan) < self.G:
                            self.plan.append([(("r", g), "root", length), (("r", g, sp), "mut", sp_id), (("r", g, sp, st), "mut", st_id)])
        else:
            strains, species_n, genera = 4, 4, 5
            per_family = strains * species_n * genera
            for f in range((self.G + per_family - 1) // per_family):
                start, length = int(rng.integers(0, hi)), int(rng.integers(lo, hi + 1))
                f_id = float(rng.uniform(0.92, 0.94))
                for ge in range(genera):
                    g_id = float(rng.uniform(0.95, 0.97))
                    for sp in range(species_n):
                        s_id = float(rng.uniform(0.975, 0.99))
                        for st in range(strains):
                            t_id = float(rng.uniform(0.99, 0.999))
                            if len(self.plan) < self.G:
                                self.plan.append([(("R",), "root", hi), (("R", f, "w"), "win", (start, length)), (("R", f), "mut", f_id),
                                                  (("R", f, ge), "mut", g_id), (("R", f, ge, sp), "mut", s_id),
                                                  (("R", f, ge, sp, st), "mut", t_id)])
        self._cache = {}

    def length(self, i):
        n = 0
        for _path, kind, par in self.plan[i]:
            n = par if kind == "root" else (par[1] if kind == "win" else n)
        return n

    def lengths(self):
        return [self.length(i) for i in range(self.G)]

    def _key(self, path):
        return node_key(self.seed, *[p if isinstance(p, int) else sum(ord(c) for c in p) + 1000 for p in path])

    def codes(self, be, i):
        cache = self._cache.setdefault(be.name, {})
        steps = self.plan[i]
        keep = {}
        cur = None
        for depth, (path, kind, par) in enumerate(steps):
            if path in cache:
                cur = cache[path]
            elif kind == "root":
                cur = random_codes(be, self._key(path), par)
            elif kind == "win":
                cur = be.roll_window(cur, par[0], par[1])
            else:
                cur = mutate(be, cur, par, self._key(path))
            if depth < len(steps) - 1:
                keep[path] = cur
        self._cache[be.name] = keep          # only the ancestors of the genome just made stay cached
        return cur

    def expected_identity(self, i, j):
        """Product of the edge identities on the path between genomes i and j (None when they share no root)."""
        a, b = self.plan[i], self.plan[j]
        if a[0][0] != b[0][0]:
            return None
        d = 0
        while d < min(len(a), len(b)) and a[d][0] == b[d][0]:
            d += 1
        ident = 1.0
        for steps in (a[d:], b[d:]):
            for _p, kind, par in steps:
                if kind == "mut":
                    ident *= par
        return ident


ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.frombuffer(b"TGCA", dtype=np.uint8)


def draft_plan(seed, i, length):
    """Contig boundaries, order and strand flips of genome i as a draft of 200-500 contigs (configs[2])."""
    rng = np.random.default_rng(node_key(seed, 77, i) & 0xFFFFFFFF)
    n = int(rng.integers(200, 501))
    n = max(2, min(n, length // 4000))
    cuts = np.sort(rng.choice(np.arange(1000, length - 1000), size=n - 1, replace=False))
    bounds = np.concatenate([[0], cuts, [length]]).astype(np.int64)
    order = rng.permutation(n)
    flips = rng.random(n) < 0.5
    return [(int(bounds[k]), int(bounds[k + 1]), bool(flips[k])) for k in order]


def contigs_numpy(codes, plan):
    """ASCII contigs (bytes) of one genome from its codes: one contig (plan None) or the draft's pieces."""
    if plan is None:
        return [ACGT[codes].tobytes()]
    out = []
    for a, b, flip in plan:
        part = codes[a:b]
        out.append(COMP[part[::-1]].tobytes() if flip else ACGT[part].tobytes())
    return out


def contigs_torch(torch, codes, plan, lut, comp):
    seq = None
    if plan is None:
        return [lut[codes.long()]]
    out = []
    for a, b, flip in plan:
        part = codes[a:b].long()
        out.append(comp[part].flip(0).contiguous() if flip else lut[part].contiguous())
    del seq
    return out

"""SURVEY.md 8(f)-2, the front of the path: 2-bit packed staging (fa_packed / fa_pack_2bit) and FASTA text parsed on the
device (fa_fasta_parse, the reference's Parser of src/pyfastani/_fasta.pyx:41-103).

CPU tests pin the host packer; the GPU tests ask that packed input and device-parsed FASTA give bit-identical minimizers
and hits to the plain bytes (which the other parity tests pin against the oracle and the golden vectors).
"""
import hashlib

import numpy as np
import pytest

import capi
import golden_io
import synth


def _noisy(seed, n):
    """ACGT with everything the packer has to keep as runs: N stretches, IUPAC codes, lower case, a non-letter."""
    rng = np.random.default_rng(seed)
    want, n = n, max(n, 500)
    a = np.frombuffer(synth.to_bytes(synth.random_codes(rng, n)), dtype=np.uint8).copy()
    for _ in range(6):
        p = int(rng.integers(0, n - 400))
        a[p:p + int(rng.integers(1, 300))] = ord("N")
    for ch in b"RYKMSWBDHVn-":
        for p in rng.integers(0, n, size=3):
            a[int(p)] = ch
    low = rng.random(n) < 0.1
    a[low] = np.where(a[low] < 0x60, a[low] | 0x20, a[low])
    a[:3] = np.frombuffer(b"NNa", dtype=np.uint8)
    a[-2:] = np.frombuffer(b"nN", dtype=np.uint8)
    return a.tobytes()[:want]


def _fold(b):
    """What unpacking gives back: acgt come back as capitals, every other byte as it was."""
    t = bytes.maketrans(b"acgt", b"ACGT")
    return bytes(b).translate(t)


# ---- CPU: the host packer -------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 17, 1000, 65537])
def test_pack_round_trip(n):
    data = _noisy(n, n)
    p = capi.Packed(data)
    assert p.length == n and p.bits.size == (n + 3) // 4
    assert p.unpack() == _fold(data)
    pos, ln = p.run_pos[:p.n_runs].astype(np.int64), p.run_len[:p.n_runs].astype(np.int64)
    assert np.all(pos[1:] >= pos[:-1] + ln[:-1])                   # ascending, disjoint
    kept = np.frombuffer(data, dtype=np.uint8)
    not_acgt = ~np.isin(kept, np.frombuffer(b"ACGTacgt", dtype=np.uint8))
    assert int(ln.sum()) == int(not_acgt.sum())


def test_pack_runs_split_by_value_and_count():
    p = capi.Packed(b"ACNNNNRRNAC" + b"NY" * 40)
    assert p.n_runs == 3 + 80
    assert list(p.run_pos[:3]) == [2, 6, 8] and list(p.run_len[:3]) == [4, 2, 1]
    assert bytes(p.run_byte[:3]) == b"NRN"


def test_python_packed_sequence():
    import pickle

    import pyfastani_b200 as pf
    data = _noisy(5, 20_000)
    p = pf.PackedSequence.pack(data)
    assert len(p) == len(data) and p.unpack() == _fold(data)
    assert p.nbytes < len(data) // 3
    assert pickle.loads(pickle.dumps(p)).unpack() == _fold(data)
    assert pf.PackedSequence.pack("acgtNn").unpack() == b"ACGTNn"
    with pytest.raises(UnicodeEncodeError):
        pf.PackedSequence.pack("ACGTŁ")
    with pytest.raises(ValueError):
        pf.PackedSequence(100, np.zeros(3, np.uint8), [], [], [])


# ---- GPU: packed input == plain bytes -------------------------------------------------------------------------------

@pytest.mark.gpu
def test_packed_references_give_identical_minimizers():
    contigs = [_noisy(11, 70_001), _noisy(12, 16), _noisy(13, 5), _noisy(14, 33_333), b"ACGT" * 5000]
    a, b = capi.Sketch(), capi.Sketch()
    a.add_draft("g", contigs)
    b.add_draft("g", [capi.Packed(c) for c in contigs])
    for x, y in zip(a.minimizers(), b.minimizers()):
        assert np.array_equal(x, y)
    assert a.warnings == b.warnings and a.meta() == b.meta()
    # mixed in one call: bytes, packed, str
    c = capi.Sketch()
    c.add_draft("g", [contigs[0], capi.Packed(contigs[1]), contigs[2].decode("latin-1"), capi.Packed(contigs[3]), contigs[4]])
    for x, y in zip(a.minimizers(), c.minimizers()):
        assert np.array_equal(x, y)


@pytest.mark.gpu
def test_packed_queries_give_identical_hits():
    query, refs, _ = synth.one_to_many(77, 5, 120_000, lo=0.85, hi=0.99)
    q = np.frombuffer(query, dtype=np.uint8).copy()
    q[40_000:40_700] = ord("N"); q[90_001] = ord("R")
    query = q.tobytes()
    sk = capi.Sketch()
    for i, r in enumerate(refs):
        sk.add_draft("r%d" % i, synth.fragment(np.random.default_rng(i), r, 3, min_end=500))
    ix = sk.index()
    plain, po = ix.query_genome(query, dump=True)
    packed, ko = ix.query_genome(capi.Packed(query), dump=True)
    assert np.array_equal(plain, packed) and len(plain) == 5
    assert np.array_equal(po["candidates"], ko["candidates"]) and np.array_equal(po["mappings"], ko["mappings"])
    assert ko["info"]["h2d_bytes"] < po["info"]["h2d_bytes"] / 2            # a quarter of the bases, plus tables
    # a draft query of packed contigs, and the batch entry staged ahead by its helper thread
    parts = synth.fragment(np.random.default_rng(5), query, 7, min_end=3500)
    d_plain, _ = ix.query_draft(parts)
    d_packed, _ = ix.query_draft([capi.Packed(p) for p in parts])
    assert np.array_equal(d_plain, d_packed)
    queries = [[query], parts, [refs[0]], [refs[1][:50_000]]]
    rows_plain, _ = ix.query_batch(queries)
    rows_packed, _ = ix.query_batch([[capi.Packed(c) for c in qq] for qq in queries])
    assert len(rows_plain) == len(rows_packed) == 4
    for x, y in zip(rows_plain, rows_packed):
        assert np.array_equal(x, y)


@pytest.mark.gpu
def test_packed_through_the_python_api():
    import pyfastani_b200 as pf
    query, refs, _ = synth.one_to_many(78, 4, 60_000, lo=0.85, hi=0.99)
    s1, s2 = pf.Sketch(), pf.Sketch()
    for i, r in enumerate(refs):
        s1.add_genome("r%d" % i, r)
        s2.add_genome("r%d" % i, pf.PackedSequence.pack(r))
    m1, m2 = s1.index(), s2.index()
    h1 = m1.query_genome(query)
    h2 = m2.query_genome(pf.PackedSequence.pack(query))
    assert h1 == h2 and len(h1) == 4
    many = m2.query_many([[pf.PackedSequence.pack(query)], [query]])
    assert many[0] == h1 and many[1] == h1


# ---- GPU: FASTA text parsed on the device ---------------------------------------------------------------------------

def _fasta_text(records, width=70, crlf=False, last_newline=True):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for name, seq in records:
        out.append(b">" + name + nl)
        for i in range(0, len(seq), width):
            out.append(seq[i:i + width] + nl)
    text = b"".join(out)
    return text if last_newline else text[:-1]


def _host_parse(text):
    """The reference parser's rule (_fasta.pyx:71-103) in a few lines of Python: split at '\\n'; a line that starts with
    '>' opens a record; other lines are appended upper-cased."""
    if not text.startswith(b">"):
        return []
    recs = []
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    for ln in lines:
        if ln.startswith(b">"):
            recs.append([ln[1:], bytearray()])
        else:
            recs[-1][1] += ln.upper()
    return [(bytes(a), bytes(b)) for a, b in recs]


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["plain", "no-final-newline", "crlf", "empty-lines", "not-fasta", "gt-inside"])
def test_device_fasta_matches_the_reference_parser(variant):
    rng = np.random.default_rng(3)
    recs = [(b"contig_%d some description" % i, _noisy(20 + i, int(rng.integers(600, 9000)))) for i in range(9)]
    recs.insert(4, (b"empty", b""))
    text = _fasta_text(recs, crlf=variant == "crlf", last_newline=variant != "no-final-newline")
    if variant == "empty-lines":
        text = text.replace(b"\n>contig_3", b"\n\n\n>contig_3") + b"\n\n"
    if variant == "not-fasta":
        text = b"# comment\n" + text
    if variant == "gt-inside":
        text = text.replace(b"contig_2 some", b"contig_2 > some").replace(b"NN", b"N>", 1)
    fa = capi.Fasta(text)
    want = _host_parse(text)
    assert fa.ids == [a.decode("latin-1") for a, _ in want]
    assert [fa.download(i) for i in range(len(want))] == [b for _, b in want]
    assert fa.n_bases == sum(len(b) for _, b in want)
    if variant == "not-fasta":
        assert fa.ids == []


@pytest.mark.gpu
def test_device_fasta_feeds_the_path():
    """The vendored E. coli / Shigella genomes as FASTA text: minimizers of the device-parsed records carry the golden
    sha256 (tests/golden/config1.json), the query gives the reference's known answer."""
    gold = golden_io.config1_golden()
    texts = {}
    for name in ("ecoli", "shigella"):
        contigs = golden_io.genome(name)
        texts[name] = _fasta_text([(b"%s_%d" % (name.encode(), i), c) for i, c in enumerate(contigs)], width=80)
    fa_ref = capi.Fasta(texts["shigella"])
    sk, plain = capi.Sketch(), capi.Sketch()
    sk.add_draft("shigella", fa_ref.contigs)
    plain.add_draft("shigella", golden_io.genome("shigella"))
    for x, y in zip(sk.minimizers(), plain.minimizers()):
        assert np.array_equal(x, y)
    ix, ix_plain = sk.index(), plain.index()
    fa_q = capi.Fasta(texts["ecoli"])
    hits, _ = ix.query_draft(fa_q.contigs)
    want, _ = ix_plain.query_draft(golden_io.genome("ecoli"))
    assert np.array_equal(hits, want) and len(hits) == 1
    assert (int(hits[0]["matches"]), int(hits[0]["fragments"])) == (1322, 1547)


@pytest.mark.gpu
def test_device_fasta_python_api(tmp_path):
    import pyfastani_b200 as pf
    query, refs, _ = synth.one_to_many(79, 3, 50_000, lo=0.9, hi=0.99)
    path = tmp_path / "refs.fna"
    parts = [synth.fragment(np.random.default_rng(i), r, 4, flip=False, permute=False, min_end=500) for i, r in enumerate(refs)]
    s1, s2 = pf.Sketch(), pf.Sketch()
    for g, ps in enumerate(parts):
        path.write_bytes(_fasta_text([(b"g%d_c%d" % (g, i), p) for i, p in enumerate(ps)], width=60))
        fa = pf.DeviceFasta(str(path))
        assert fa.ids == ["g%d_c%d" % (g, i) for i in range(len(ps))] and fa.bases == sum(map(len, ps))
        s1.add_draft("g%d" % g, fa.sequences)
        s2.add_draft("g%d" % g, ps)
    assert s1.minimizers.arrays()[0].tobytes() == s2.minimizers.arrays()[0].tobytes()
    m1, m2 = s1.index(), s2.index()
    qfa = pf.DeviceFasta(_fasta_text([(b"q", query)]))
    assert m1.query_draft(qfa.sequences) == m2.query_genome(query)
    assert len(pf.DeviceFasta(b"no header\nACGT\n")) == 0

"""Generate tests/golden/protein.json|npz from the REAL pyfastani (protein mode: pyx:225-309, 548-550).

Same recipe as make_golden.py (build the reference into /tmp/o/site, then
    PYTHONPATH=/tmp/o/site python tests/golden/make_golden_protein.py).
The three MIBiG clusters of the reference's own protein test (test_ani.py:96-115, known answer
130 / 176 twice) are stored gzip-compressed under tests/golden/data/ because /root/reference does
not exist on the GPU box.
"""
import gzip
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import pyfastani  # noqa: E402  (the reference, from PYTHONPATH)

import cases  # noqa: E402
from make_golden import f32hex, hits_to_rows, minimizer_triples  # noqa: E402

REF_DATA = "/root/reference/src/pyfastani/tests/data"


def fasta(path):
    seqs, cur = [], []
    for line in open(path):
        if line.startswith(">"):
            if cur:
                seqs.append("".join(cur).encode())
                cur = []
        else:
            cur.append(line.strip())
    if cur:
        seqs.append("".join(cur).encode())
    return seqs


def main():
    assert "/root/repo" not in os.path.abspath(pyfastani.__file__), "must import the REFERENCE pyfastani"
    arrays, out = {}, {"minimizers": [], "queries": []}
    for i, case in enumerate(cases.protein_minimizer_cases()):
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            sk = pyfastani.Sketch(**case["params"])
            sk.add_draft("g", case["contigs"])
            nwarn = len([x for x in w if "short contig" in str(x.message)])
        h, s, p, n = minimizer_triples(sk.minimizers)
        arrays["h%d" % i], arrays["s%d" % i], arrays["w%d" % i] = h, s, p
        out["minimizers"].append({"name": case["name"], "window": sk.window_size, "n": n, "warnings": nwarn})
        print(case["name"], n, nwarn)
    for case in cases.protein_query_cases():
        sk = pyfastani.Sketch(**case["params"])
        for name, contigs in case["refs"]:
            sk.add_draft(name, contigs)
        n_min = len(sk.minimizers)
        mapper = sk.index()
        res = []
        for q in case["queries"]:
            with warnings.catch_warnings(record=True) as w:
                warnings.simplefilter("always")
                res.append({"hits": hits_to_rows(mapper.query_draft(q, threads=1)), "warnings": len(w)})
        out["queries"].append({"name": case["name"], "window": mapper.window_size, "minimizers": n_min,
                               "unique": len(mapper.lookup_index), "results": res})
        print(case["name"], n_min, [r["hits"] for r in res][:3])
    # the reference's own protein test, test_ani.py:96-115 (it adds cluster 1425 under both names)
    bgc = {n: fasta(os.path.join(REF_DATA, n + ".faa")) for n in ("BGC0001425", "BGC0001427", "BGC0001428")}
    os.makedirs(os.path.join(HERE, "data"), exist_ok=True)
    for name, contigs in bgc.items():
        with gzip.open(os.path.join(HERE, "data", name + ".seq.gz"), "wb", compresslevel=9) as f:
            f.write(b"\n".join(contigs) + b"\n")
    sk = pyfastani.Sketch(protein=True, fragment_length=100)
    sk.add_draft("BGC0001425", bgc["BGC0001425"])
    sk.add_draft("BGC0001427", bgc["BGC0001425"])
    n_min = len(sk.minimizers)
    mapper = sk.index()
    hits = mapper.query_draft(bgc["BGC0001428"], threads=1)
    assert [(h.name, h.matches, h.fragments) for h in hits] == [("BGC0001425", 130, 176), ("BGC0001427", 130, 176)]
    sk2 = pyfastani.Sketch(protein=True, fragment_length=100)
    sk2.add_draft("BGC0001425", bgc["BGC0001425"])
    sk2.add_draft("BGC0001427", bgc["BGC0001427"])
    m2 = sk2.index()
    out["bgc"] = {"as_in_test_ani": hits_to_rows(hits), "minimizers": n_min, "unique": len(mapper.lookup_index),
                  "distinct_refs": hits_to_rows(m2.query_draft(bgc["BGC0001428"], threads=1)),
                  "self": hits_to_rows(m2.query_draft(bgc["BGC0001427"], threads=1))}
    print(out["bgc"])
    np.savez_compressed(os.path.join(HERE, "protein.npz"), **arrays)
    json.dump(out, open(os.path.join(HERE, "protein.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

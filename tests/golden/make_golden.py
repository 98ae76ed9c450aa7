"""Generate tests/golden/*.json|npz from the REAL pyfastani (the reference itself).

Run in the build container only (the reference cannot travel to the GPU box):

    # one-off: build the reference into a scratch site dir
    cp -r /root/reference /tmp/o/pyfastani && chmod -R u+w /tmp/o
    (cd /tmp/o/pyfastani && pip install --no-build-isolation --no-deps --no-index \
        --find-links /opt/wheelhouse --target /tmp/o/site .)
    PYTHONPATH=/tmp/o/site python tests/golden/make_golden.py

Inputs are either literal (stored in the fixture) or regenerated from seeds by
tests/synth.py + tests/cases.py, so fixtures stay small.  The two genomes of
BASELINE config 1 are stored gzip-compressed under tests/golden/data/ because
/root/reference does not exist on the GPU box.
"""
import gzip
import hashlib
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
import synth  # noqa: E402

REF_DATA = "/root/reference/vendor/FastANI/data"


def f32hex(x):
    return float(np.float32(x)).hex()


def hits_to_rows(hits):
    return [[h.name, f32hex(h.identity), int(h.matches), int(h.fragments)] for h in hits]


def minimizer_triples(view):
    n = len(view)
    st = view.__getstate__()
    return (np.array(st["hashes"], dtype=np.uint32), np.array(st["ids"], dtype=np.int32),
            np.array(st["offsets"], dtype=np.int32), n)


def main():
    assert "/root/repo" not in os.path.abspath(pyfastani.__file__), "must import the REFERENCE pyfastani"
    os.makedirs(os.path.join(HERE, "data"), exist_ok=True)

    # ---- A. minimizer cases -------------------------------------------------
    arrays, manifest = {}, []
    for i, case in enumerate(cases.minimizer_cases()):
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            sk = pyfastani.Sketch(**case["params"])
            sk.add_draft("g", case["contigs"])
            nwarn = len([x for x in w if "short contig" in str(x.message)])
        h, s, p, n = minimizer_triples(sk.minimizers)
        arrays["h%d" % i], arrays["s%d" % i], arrays["w%d" % i] = h, s, p
        manifest.append({"name": case["name"], "window": sk.window_size, "n": n, "warnings": nwarn})
    np.savez_compressed(os.path.join(HERE, "minimizers.npz"), **arrays)
    json.dump(manifest, open(os.path.join(HERE, "minimizers.json"), "w"), indent=1)
    print("minimizer cases:", len(manifest))

    # ---- B. query cases -----------------------------------------------------
    out = []
    for case in cases.query_cases():
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            sk = pyfastani.Sketch(**case["params"])
            for name, contigs in case["refs"]:
                sk.add_draft(name, contigs)
            n_min = len(sk.minimizers)
            mapper = sk.index()
            n_uniq = len(mapper.lookup_index)
            ref_warn = len(w)
        res = []
        for q in case["queries"]:
            with warnings.catch_warnings(record=True) as w:
                warnings.simplefilter("always")
                hits = mapper.query_draft(q, threads=1)
                res.append({"hits": hits_to_rows(hits), "warnings": len(w)})
        out.append({"name": case["name"], "window": mapper.window_size, "minimizers": n_min, "unique": n_uniq,
                    "ref_warnings": ref_warn, "results": res})
        print(case["name"], n_min, n_uniq, [len(r["hits"]) for r in res][:8])
    json.dump(out, open(os.path.join(HERE, "queries.json"), "w"), indent=1)

    # ---- C. BASELINE config 1 (+ the reference's own known answers) ---------
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

    genomes = {"ecoli": fasta(os.path.join(REF_DATA, "Escherichia_coli_str_K12_MG1655.fna")),
               "shigella": fasta(os.path.join(REF_DATA, "Shigella_flexneri_2a_01.fna"))}
    for name, contigs in genomes.items():
        with gzip.open(os.path.join(HERE, "data", name + ".seq.gz"), "wb", compresslevel=9) as f:
            f.write(b"\n".join(contigs) + b"\n")
    c1 = {}
    for rname, qname in (("shigella", "ecoli"), ("ecoli", "shigella"), ("ecoli", "ecoli"), ("shigella", "shigella")):
        sk = pyfastani.Sketch()
        sk.add_draft(rname, genomes[rname])
        h, s, p, n = minimizer_triples(sk.minimizers)
        mapper = sk.index()
        hits = mapper.query_draft(genomes[qname], threads=1)
        c1["%s_vs_%s" % (qname, rname)] = {
            "hits": hits_to_rows(hits), "minimizers": n, "unique": len(mapper.lookup_index),
            "sha256": hashlib.sha256(h.tobytes() + s.tobytes() + p.tobytes()).hexdigest(),
            "first": [[int(h[i]), int(s[i]), int(p[i])] for i in range(4)],
        }
        print(qname, "vs", rname, c1["%s_vs_%s" % (qname, rname)]["hits"])
    json.dump(c1, open(os.path.join(HERE, "config1.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

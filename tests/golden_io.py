"""Loaders for tests/golden (fixtures written by tests/golden/make_golden.py from the real pyfastani)."""
import gzip
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def minimizer_golden():
    man = json.load(open(os.path.join(GOLD, "minimizers.json")))
    arr = np.load(os.path.join(GOLD, "minimizers.npz"))
    return [(m, arr["h%d" % i], arr["s%d" % i], arr["w%d" % i]) for i, m in enumerate(man)]


def query_golden():
    return {c["name"]: c for c in json.load(open(os.path.join(GOLD, "queries.json")))}


def config1_golden():
    return json.load(open(os.path.join(GOLD, "config1.json")))


def genome(name):
    """Contigs of 'ecoli' / 'shigella' (vendor/FastANI/data in the reference) or the proteins of 'BGC000142x'
    (src/pyfastani/tests/data), as bytes."""
    with gzip.open(os.path.join(GOLD, "data", name + ".seq.gz"), "rb") as f:
        return f.read().split(b"\n")[:-1]


def protein_golden():
    """(manifest dict, npz arrays) of protein mode: minimizer cases, query cases, the reference's BGC test."""
    return json.load(open(os.path.join(GOLD, "protein.json"))), np.load(os.path.join(GOLD, "protein.npz"))


def f32(hexstr):
    return np.float32(float.fromhex(hexstr))

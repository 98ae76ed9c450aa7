"""FastANI's result files (outputCGI / outputPhylip, FA/cgi/include/computeCoreIdentity.hpp:303-445) written from
`Mapper.query_many` results: host-side formatting only, no device needed."""
import io

import numpy as np

import pyfastani_b200 as pf
from pyfastani_b200 import output


def _hits(*rows):
    return [pf.Hit(name, identity, matches, fragments) for name, identity, matches, fragments in rows]


def test_tabular_matches_the_reference_format():
    # the reference's own known answer (README / test_ani.py:47-51): identity printed with six significant digits
    res = [_hits(("shigella.fna", np.float32(97.7507), 1303, 1608)),
           _hits(("b.fna", np.float32(100.0), 1547, 1547), ("a.fna", np.float32(80.12345678), 12, 1547)),
           []]
    text = output.tabular_text(["ecoli.fna", "q2.fna", "q3.fna"], res)
    assert text == ("ecoli.fna\tshigella.fna\t97.7507\t1303\t1608\n"
                    "q2.fna\tb.fna\t100\t1547\t1547\n"
                    "q2.fna\ta.fna\t80.1235\t12\t1547\n")
    # structured rows (query_many(..., rows=True)) with the reference names
    dt = np.dtype([("ref_genome", "<i4"), ("matches", "<i4"), ("fragments", "<i4"), ("identity", "<f4")])
    rows = np.array([(1, 1547, 1547, 100.0), (0, 12, 1547, 80.12345678)], dtype=dt)
    assert output.tabular_text(["q2.fna"], [rows], ["a.fna", "b.fna"]) == "q2.fna\tb.fna\t100\t1547\t1547\nq2.fna\ta.fna\t80.1235\t12\t1547\n"


def test_matrix_matches_the_reference_format():
    # three genomes all-vs-all: self pairs ignored, both directions averaged (in float), missing pairs NA
    names = ["g0", "g1", "g2"]
    res = [_hits(("g0", 100.0, 10, 10), ("g1", np.float32(95.5), 9, 10)),
           _hits(("g1", 100.0, 10, 10), ("g0", np.float32(96.5), 9, 10)),
           _hits(("g2", 100.0, 10, 10), ("g1", np.float32(81.25), 3, 10))]
    buf = io.StringIO()
    output.write_matrix(buf, names, names, res)
    assert buf.getvalue() == "3\ng0\ng1\t96.000000\ng2\tNA\t81.250000\n"
    # queries that are not references come first in the numbering (computeCoreIdentity.hpp:361-377)
    lines = list(output.matrix_lines(["q"], ["r1", "r2"], [_hits(("r2", np.float32(88.8), 5, 9))]))
    assert lines == ["3", "q", "r1\tNA", "r2\t%f\tNA" % float(np.float32(88.8))]

"""FastANI's result files over `Mapper.query_many` -- a thin host-side layer.

The reference writes them from its `CGI_Results` vector (FA/cgi/include/computeCoreIdentity.hpp):

* `outputCGI` (:303-340): one line per reported pair, ``query<TAB>reference<TAB>identity<TAB>matches<TAB>fragments``,
  ordered by query and, inside a query, by identity descending (`CGI_Results::operator<`, cgid_types.hpp:76-79);
  the identity goes through ``ostream << float`` (six significant digits);
* `outputPhylip` (:349-445): the lower-triangular matrix ``<name>.matrix``: genomes numbered in order of first
  appearance (queries, then references), a pair computed in both directions is averaged, a pair without a reported
  mapping is ``NA``, values printed with ``std::to_string(float)`` (six decimals).

The minimum-fraction filter both functions apply (:327, :405) is the one `Mapper.query_*` has already applied to its
hits (pyx:1121-1135), so the rows here are the hits as returned.
"""
import io

import numpy as np

__all__ = ["tabular_lines", "write_tabular", "matrix_lines", "write_matrix"]


def _rows(hits):
    """(reference name, identity, matches, fragments) of one query's result: a list of `Hit`, or the structured rows
    of ``query_many(..., rows=True)`` together with the reference names."""
    for h in hits:
        yield h.name, float(np.float32(h.identity)), int(h.matches), int(h.fragments)


def _named(results, reference_names):
    """Per query, an iterable of objects with .name/.identity/.matches/.fragments.  `results` holds lists of `Hit`, or
    structured arrays with a ``ref_genome`` column (then `reference_names` resolves it)."""
    class _Row:
        __slots__ = ("name", "identity", "matches", "fragments")

        def __init__(self, name, identity, matches, fragments):
            self.name, self.identity, self.matches, self.fragments = name, identity, matches, fragments

    out = []
    for hits in results:
        if isinstance(hits, np.ndarray):
            if reference_names is None:
                raise ValueError("structured hit rows need the reference names")
            hits = [_Row(reference_names[int(r["ref_genome"])], r["identity"], r["matches"], r["fragments"]) for r in hits]
        out.append(hits)
    return out


def _g(x):
    """``ostream << float``: general format, six significant digits."""
    return "%g" % x


def tabular_lines(query_names, results, reference_names=None):
    """The lines of FastANI's tab-delimited output (outputCGI, computeCoreIdentity.hpp:303-340)."""
    query_names = list(query_names)
    results = _named(results, reference_names)
    if len(query_names) != len(results):
        raise ValueError("one result list per query name")
    for q, hits in zip(query_names, results):
        # (a query's hits arrive identity-descending; the stable sort keeps the library's order among equals)
        for name, identity, matches, fragments in sorted(_rows(hits), key=lambda r: -r[1]):
            yield "%s\t%s\t%s\t%d\t%d" % (q, name, _g(identity), matches, fragments)


def _open(out):
    if isinstance(out, (str, bytes)) or hasattr(out, "__fspath__"):
        return open(out, "w"), True
    return out, False


def write_tabular(out, query_names, results, reference_names=None):
    """Write the tab-delimited result file to a path or a text file object; returns the number of lines."""
    f, close = _open(out)
    try:
        n = 0
        for line in tabular_lines(query_names, results, reference_names):
            f.write(line + "\n")
            n += 1
        return n
    finally:
        if close:
            f.close()


def matrix_lines(query_names, reference_names, results):
    """The lines of FastANI's ``--matrix`` output (outputPhylip, computeCoreIdentity.hpp:349-445)."""
    query_names, reference_names = list(query_names), list(reference_names)
    results = _named(results, reference_names)
    if len(query_names) != len(results):
        raise ValueError("one result list per query name")
    index, order = {}, []
    for name in query_names + reference_names:
        if name not in index:
            index[name] = len(order)
            order.append(name)
    n = len(order)
    cell = {}
    for q, hits in zip(query_names, results):
        qi = index[q]
        for name, identity, _, _ in _rows(hits):
            ri = index[name]
            if qi == ri:
                continue
            key = (qi, ri) if qi > ri else (ri, qi)
            # float arithmetic, as the reference's std::vector<float> matrix
            if key in cell and cell[key] > 0:
                cell[key] = float((np.float32(cell[key]) + np.float32(identity)) / np.float32(2))
            else:
                cell[key] = float(np.float32(identity))
    yield "%d" % n
    for i in range(n):
        vals = [("%f" % cell[(i, j)]) if cell.get((i, j), 0.0) > 0.0 else "NA" for j in range(i)]
        yield "\t".join([str(order[i])] + vals)


def write_matrix(out, query_names, reference_names, results):
    """Write the lower-triangular matrix to a path (the reference appends ``.matrix`` to its output name itself; the
    path is used as given) or a text file object."""
    f, close = _open(out)
    try:
        for line in matrix_lines(query_names, reference_names, results):
            f.write(line + "\n")
    finally:
        if close:
            f.close()


def tabular_text(query_names, results, reference_names=None):
    buf = io.StringIO()
    write_tabular(buf, query_names, results, reference_names)
    return buf.getvalue()

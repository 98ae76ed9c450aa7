"""CPU tests of the boundary: the C-ABI library loads and exports every symbol the header
declares; host-side logic that needs no device (statistics, argument validation, value
classes) behaves like the reference; and computing calls fail loudly without a GPU."""
import ctypes
import pickle
import re
import subprocess
import warnings

import pytest

import capi
import golden_io
from oracle.oracle import Oracle, available


def _declared():
    src = open(capi.HEADER).read()
    return sorted(set(re.findall(r"^FA_API\s+[\w\s\*]+?\b(fa_\w+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 30
    lib = ctypes.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (fa_\w+)", out)))
    assert exported == names          # nothing undeclared leaks out either


def test_abi_struct_sizes():
    assert ctypes.sizeof(capi.Params) == 40
    assert ctypes.sizeof(capi.Contig) == 24
    assert ctypes.sizeof(capi.Hit) == 16


def test_recommended_window_matches_reference_tests():
    assert capi.recommended_window() == 24                        # test_ani.py:60, :80
    for man in golden_io.query_golden().values():
        pass
    import cases
    for case, man in zip(cases.minimizer_cases(), [m for m, *_ in golden_io.minimizer_golden()]):
        assert capi.recommended_window(**case["params"]) == man["window"], case["name"]


def test_statistics_match_oracle():
    port = Oracle("port")
    for k, pid in ((16, 80.0), (12, 90.0), (16, 95.0)):
        for s in list(range(1, 200, 7)) + [237, 238, 239, 250, 300, 1000, 2962]:
            assert capi.stat_minimum_hits(s, k, pid) == port.minimum_hits(s, k, pid)
            for x in sorted({0, 1, 2, 3, 4, 5, s // 50, s // 10, s // 2, s - 1, s}):
                if 0 <= x <= s:
                    assert capi.stat_l2(x, s, k, pid) == port.l2_stat(x, s, k, pid), (x, s, k, pid)


@pytest.mark.skipif("reference" not in available(), reason="oracle/_ref not built")
def test_filter_is_monotone_in_shared_count():
    """fa_stat.cpp tabulates the CI filter as `shared >= min_shared[s]`; check the underlying
    pass flag really is monotone, against the reference's own Boost path."""
    ref = Oracle("reference")
    for k, pid in ((16, 80.0), (16, 95.0), (12, 90.0)):
        for s in (1, 2, 3, 10, 57, 100, 238, 400, 1000):
            flags = [ref.l2_stat(x, s, k, pid)[0] for x in range(s + 1)]
            first = flags.index(True) if True in flags else s + 1
            assert all(flags[first:]) and not any(flags[:first]), (s, k, pid)
            assert [capi.stat_l2(x, s, k, pid)[0] for x in range(s + 1)] == flags


def test_stat_table_rows_match_the_walk_down_functions():
    """The device kernels read the TABLE (csrc/fa_stat.cpp stat_table: bisection + neighbourhood check), not the walk-down
    functions the two tests above exercise: every row, for the default and two other parameter sets, must equal
    max(1, estimateMinimumHitsRelaxed(s)) and the first shared count from which the filter of computeMap.hpp:380 passes."""
    lib = capi.lib()
    for k, pid, s_max, dense in ((16, 80.0, 2962, 320), (16, 95.0, 600, 200), (12, 90.0, 500, 200), (5, 70.0, 120, 120)):
        for s in list(range(1, dense + 1)) + list(range(dense + 1, s_max + 1, 97)) + [s_max]:
            mh, ms, irr = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
            capi.check(lib.fa_stat_table_row(s, s_max, k, ctypes.c_float(pid), ctypes.byref(mh), ctypes.byref(ms), ctypes.byref(irr)))
            assert irr.value == 0
            assert mh.value == max(1, capi.stat_minimum_hits(s, k, pid)), (s, k, pid)
            xs = range(s + 1) if s <= 150 else sorted({x for x in (0, 1, ms.value - 2, ms.value - 1, ms.value, ms.value + 1, s // 2, s) if 0 <= x <= s})
            for x in xs:
                assert capi.stat_l2(x, s, k, pid)[0] == (x >= ms.value), (x, s, k, pid)


def test_python_value_classes_and_errors():
    import pyfastani_b200 as pf
    assert pf.MAX_KMER_SIZE == 2048
    h = pf.Hit("g", 97.5, 3, 4)
    assert (h.name, h.matches, h.fragments, h.identity) == ("g", 3, 4, 97.5)
    assert pickle.loads(pickle.dumps(h)) == h
    assert repr(h) == "Hit(name='g', identity=97.5, matches=3, fragments=4)"
    m = pf.MinimizerInfo(21161528, 0, 18)
    assert pickle.loads(pickle.dumps(m)) == m and m.window_position == 18
    p = pf.Position(1, 2)
    assert pickle.loads(pickle.dumps(p)) == p
    # constructor validation, src/pyfastani/tests/test_sketch.py:12-23
    with pytest.raises(TypeError):
        pf.Sketch(k="1")
    with pytest.raises(TypeError):
        pf.Sketch(fragment_length="1")
    with pytest.raises(TypeError):
        pf.Sketch(minimum_fraction="0.5")
    with pytest.raises(OverflowError):
        pf.Sketch(k=2**32)
    with pytest.raises(ValueError):
        pf.Sketch(k=0)
    with pytest.raises(ValueError):
        pf.Sketch(p_value=-1.0)
    with pytest.raises(ValueError):
        pf.Sketch(percentage_identity=-1.0)
    with pytest.raises(ValueError):
        pf.Sketch(percentage_identity=200.0)
    with pytest.raises(BufferError):
        pf.Sketch(k=4096)
    with pytest.raises(TypeError):
        pf.Mapper()


def test_no_cpu_fallback():
    """Without a device the product must fail loudly, not compute on the host."""
    n = ctypes.c_int32(-1)
    rc = capi.lib().fa_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    import pyfastani_b200 as pf
    with pytest.raises(pf.CudaError):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pf.Sketch()
    with pytest.raises(capi.FaError) as e:
        capi.Sketch()
    assert e.value.code == 2


def test_product_does_not_touch_the_oracle():
    import os
    root = os.path.join(capi.ROOT, "pyfastani_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".pyx", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f == "__never__", os.path.join(dirpath, f)

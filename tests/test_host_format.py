"""Text rendering of results: Float64 values print as Rust's f64::to_string()
(shortest round-trip, positional, no exponent -- src/data/mod.rs:49-65), NoValue as "."."""
import ctypes
import math
import os
import random
import struct
from decimal import Decimal

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fmt():
    import __graft_entry__ as g

    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "granges_b200", "libgranges_host.so"))
    lib.grh_format_f64.restype = ctypes.c_int
    lib.grh_format_f64.argtypes = [ctypes.c_double, ctypes.c_char_p]
    buf = ctypes.create_string_buffer(512)

    def f(v):
        n = lib.grh_format_f64(v, buf)
        return buf.raw[:n].decode()

    return f


def rust_display(v):
    """What Rust's `format!("{}", v)` prints for an f64: repr's shortest digits, expanded positionally."""
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    s = format(Decimal(repr(v)), "f")  # repr = shortest round-trip digits; Decimal expands the exponent exactly
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    if s in ("", "-"):
        s += "0"
    return s


KNOWN = [(5.0, "5"), (0.1 + 0.2, "0.30000000000000004"), (1e-7, "0.0000001"), (1e21, "1000000000000000000000"), (-0.0, "-0"),
         (0.0, "0"), (2.5, "2.5"), (4.1, "4.1"), (1 / 3, "0.3333333333333333"), (123456789.125, "123456789.125"),
         (float("inf"), "inf"), (float("-inf"), "-inf"), (float("nan"), "NaN"), (5e-324, "0." + "0" * 323 + "5"),
         (1.7976931348623157e308, "17976931348623157" + "0" * 292)]


def test_known_values(fmt):
    for v, want in KNOWN:
        assert fmt(v) == want
        assert rust_display(v) == want  # the helper the CLI tests rely on


def test_random_bit_patterns_round_trip(fmt):
    rng = random.Random(3)
    for _ in range(20000):
        v = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
        s = fmt(v)
        assert s == rust_display(v)
        if not math.isnan(v):
            assert float(s) == v and "e" not in s.lower().replace("inf", "")

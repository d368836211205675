"""GPU: device self-tests of arithmetic shortcuts that must be exact."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def test_div10_shortcut_is_the_ieee_division_for_every_float(hpe):
    bad = C.c_ulonglong(123)
    assert hpe.lib().hpe_selftest_div10(C.byref(bad)) == 0
    assert bad.value == 0

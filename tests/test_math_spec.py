"""The transcendental specification (oracle/orc_math.h vs hijiki_b200/csrc/device/hjk_math.cuh).

GLSL leaves sin/cos/tan/exp/atan/asin precision to the driver; both sides implement one fixed
polynomial specification.  Here: the oracle's copy stays within a few ulp of libm on the
ranges the path uses, and the product's copy (compiled for the host) agrees bit for bit.
"""
import numpy as np
import pytest

import _libs

RNG = np.random.default_rng(20261017)


def _ulp_err(got, ref64):
    ref32 = ref64.astype(np.float32)
    spacing = np.maximum(np.spacing(np.abs(ref32)).astype(np.float64), 1e-45)
    return np.abs(got.astype(np.float64) - ref64) / spacing


def _inputs():
    two_pi = RNG.random(400_000).astype(np.float32) * np.float32(6.2831855)
    neg = -(RNG.random(400_000).astype(np.float32) * np.float32(80.0))
    tan_in = RNG.random(200_000).astype(np.float32) * np.float32(1.5)
    unit = (RNG.random(200_000).astype(np.float32) * 2 - 1).astype(np.float32)
    y = RNG.standard_normal(200_000).astype(np.float32)
    x = RNG.standard_normal(200_000).astype(np.float32)
    return two_pi, neg, tan_in, unit, y, x


def test_oracle_math_close_to_libm(oracle):
    two_pi, neg, tan_in, unit, y, x = _inputs()
    s = _libs.math_eval(oracle.orc_math_eval, 0, two_pi)
    c = _libs.math_eval(oracle.orc_math_eval, 1, two_pi)
    assert np.abs(s - np.sin(two_pi.astype(np.float64))).max() < 1.2e-7
    assert np.abs(c - np.cos(two_pi.astype(np.float64))).max() < 1.2e-7
    big = np.abs(np.sin(two_pi.astype(np.float64))) > 1e-3
    assert _ulp_err(s, np.sin(two_pi.astype(np.float64)))[big].max() <= 2.0
    e = _libs.math_eval(oracle.orc_math_eval, 3, neg)
    assert _ulp_err(e, np.exp(neg.astype(np.float64))).max() <= 1.5
    assert _libs.math_eval(oracle.orc_math_eval, 3, np.zeros(1, np.float32))[0] == 1.0
    assert _libs.math_eval(oracle.orc_math_eval, 3, np.array([-200.0], np.float32))[0] == 0.0
    t = _libs.math_eval(oracle.orc_math_eval, 2, tan_in)
    assert _ulp_err(t, np.tan(tan_in.astype(np.float64))).max() <= 4.0
    a = _libs.math_eval(oracle.orc_math_eval, 5, unit)
    assert _ulp_err(a, np.arcsin(unit.astype(np.float64))).max() <= 4.0
    # the FMA-specified exp(-e) of the reconstruction weight: faithful (< 1 ulp) on its whole range
    pos = np.concatenate([-neg, RNG.random(400_000).astype(np.float32) * np.float32(8.0),
                          np.float32(10.0) ** RNG.uniform(-8, 1.94, 200_000).astype(np.float32)])
    eb = _libs.math_eval(oracle.orc_math_eval, 6, pos)
    assert _ulp_err(eb, np.exp(-pos.astype(np.float64))).max() < 1.0
    edge = np.array([0.0, -0.0, 87.3365402, 87.34, 88.0, 1e30, np.inf], np.float32)
    assert np.array_equal(_libs.math_eval(oracle.orc_math_eval, 6, edge)[[0, 1, 3, 4, 5, 6]],
                          np.array([1, 1, 0, 0, 0, 0], np.float32))
    assert _libs.math_eval(oracle.orc_math_eval, 6, edge)[2] > 0
    assert np.isnan(_libs.math_eval(oracle.orc_math_eval, 6, np.array([np.nan], np.float32))[0])
    at = _libs.math_eval(oracle.orc_math_eval, 4, y, x)
    assert _ulp_err(at, np.arctan2(y.astype(np.float64), x.astype(np.float64))).max() <= 4.0


@pytest.mark.parametrize("fn", range(7))
def test_product_math_matches_oracle_bitwise(oracle, hosttest, fn):
    two_pi, neg, tan_in, unit, y, x = _inputs()
    a = [two_pi, two_pi, tan_in, neg, y, unit, -neg][fn]
    special = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-30, -1e-30, 88.0, -87.0, -87.4, 0.5,
                        -0.5, 3.1415927, 6.2831855, 1e7, 87.3365402, 87.34, 1e-9, 5.85645], dtype=np.float32)
    a = np.concatenate([a, special])
    b = np.concatenate([x, special[::-1]]) if fn == 4 else None
    o = _libs.math_eval(oracle.orc_math_eval, fn, a, b)
    p = _libs.math_eval(hosttest.ht_math_eval, fn, a, b)
    same = (o.view(np.uint32) == p.view(np.uint32)) | (np.isnan(o) & np.isnan(p))
    assert same.all()

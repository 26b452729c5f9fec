"""Host-side geometry (waterlily_b200/body.py) against the reference's AutoBody pins
(test/maintests.jl:189-205,226-229) and the time-derivative helper of accelerate! (src/Flow.jl:71-72)."""
import math

import numpy as np
import torch

from waterlily_b200 import AutoBody
from waterlily_b200.sim import _ddt


def norm2(x):
    return torch.sqrt((x * x).sum(-1))


def _measure(body, x, t, **kw):
    d, n, V = body.measure(torch.tensor([x], dtype=torch.float64), t, **kw)
    return float(d[0]), n[0].numpy(), V[0].numpy()


def _close(got, want):
    return np.isclose(got[0], want[0]) and np.allclose(got[1], want[1]) and np.allclose(got[2], want[2])


S2, SH = math.sqrt(2.0), math.sqrt(0.5)
body1 = AutoBody(lambda x, t: norm2(x) - 2 - t)
body2 = AutoBody(lambda x, t: norm2(x) - 2, lambda x, t: x + t ** 2)


def test_autodiff_pins():
    # maintests.jl:192-197
    assert _close(_measure(body1, [S2, S2], 0.0), (0, [SH, SH], [0, 0]))
    assert _close(_measure(body1, [2.0, 0.0, 0.0], 1.0), (-1.0, [1, 0, 0], [0, 0, 0]))
    assert _close(_measure(body2, [S2, S2], 0.0), (0, [SH, SH], [0, 0]))
    assert _close(_measure(body2, [1.0, -1.0, -1.0], 1.0), (0.0, [1, 0, 0], [-2, -2, -2]))


def test_boolean_pins():
    # maintests.jl:200-202: union, ∪ alias, and the difference a - b = a ∩ -b
    assert _close(_measure(body1 + body2, [-S2, -S2], 1.0), (-S2, [-SH, -SH], [-2, -2]))
    assert _close(_measure(body1 | body2, [-S2, -S2], 1.0), (-S2, [-SH, -SH], [-2, -2]))
    assert _close(_measure(body1 - body2, [-S2, -S2], 1.0), (S2, [SH, SH], [-2, -2]))


def test_csg_measure_matches_sdf():
    # measure().d and sdf() of every CSG operator must agree, on two overlapping circles and on the
    # nested annuli of maintests.jl:208-213
    a = AutoBody(lambda x, t: norm2(x - torch.tensor([0.3, 0.0], dtype=x.dtype)) - 1.0)
    b = AutoBody(lambda x, t: norm2(x + torch.tensor([0.4, 0.1], dtype=x.dtype)) - 0.8)
    rng = np.random.default_rng(0)
    pts = torch.from_numpy(rng.uniform(-2, 2, size=(400, 2)))
    for body in (a + b, a - b, a & b, b - a):
        d, n, _ = body.measure(pts, 0.0)
        assert torch.allclose(d, body.sdf(pts, 0.0))
        assert torch.allclose(norm2(n), torch.ones_like(d))
    circles = [AutoBody(lambda x, t, r=r: norm2(x) - r) for r in (1.0, 0.75, 0.5, 0.25)]
    body = circles[0] - circles[1] + circles[2] - circles[3]
    d, n, _ = body.measure(pts, 1.0)
    r = norm2(pts)
    want = torch.minimum(torch.maximum(r - 1.0, 0.75 - r), torch.maximum(r - 0.5, 0.25 - r))   # annulus ∪ annulus
    assert torch.allclose(d, want)
    assert torch.allclose(d, body.sdf(pts, 1.0))


def test_fast_measure():
    # maintests.jl:227-228
    full = _measure(body1, [3.0, 4.0], 0.0)
    assert _close(_measure(body1, [3.0, 4.0], 0.0, fastd2=9), full)
    got = _measure(body1, [3.0, 4.0], 0.0, fastd2=8)
    assert np.isclose(got[0], 3.0) and np.all(got[1] == 0) and np.all(got[2] == 0)


def test_ddt_of_host_closures():
    # accelerate! adds dU/dt (ForwardDiff.derivative in the reference, src/Flow.jl:71-72)
    assert abs(_ddt(lambda i, t: 0.5 * 3.0 * t * t, 1, 2.0) - 6.0) < 1e-9                 # polynomial ramp
    assert abs(_ddt(lambda i, t: math.sin(t), 1, 0.3) - math.cos(0.3)) < 1e-9             # smooth, pure Python
    assert abs(_ddt(lambda i, t: torch.sin(torch.as_tensor(t)), 1, 0.3) - math.cos(0.3)) < 1e-12   # torch closure: exact AD
    assert abs(_ddt(lambda i, t: max(t, 0.0), 1, 0.0) - 1.0) < 1e-9                       # ramp that starts at t: forward branch
    f = lambda i, t: t ** 3
    f.ddt = lambda i, t: 3 * t * t
    assert _ddt(f, 1, 2.0) == 12.0                                                          # user-supplied derivative

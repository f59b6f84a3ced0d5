"""Direct parity of the entry points that the stepper tests only reach through whole calls:

* ``gomelt_shift_window_f32`` (the device part of moveEverything cF:2439-2443 / 2460-2464) against the oracle's
  interpolatePoints composition ``T1on3 + (Tp2on3 + Tp3)`` and against the general interpolation kernel, bit for bit;
* ``gomelt_accum_single_step_f32`` (gm:339-357 + melting_temp cF:3696-3712) against the driver's NumPy sequence;
* ``gomelt_patch_copy_f32`` (box transfers of the distributed drop-in) against NumPy slicing, incl. its argument checks;
* ``gomelt_clamp_min_f32``.
"""
import numpy as np
import pytest

from oracle import computeFunctions as cF
from oracle.util import make_level

pytestmark = pytest.mark.gpu
F32 = np.float32


def _dev(a):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def _field(lv, seed, lo, hi):
    rng = np.random.default_rng(seed)
    x, y, z = lv["node_coords"]
    X, Y, Z = x[None, None, :], y[None, :, None], z[:, None, None]
    f = lo + (hi - lo) * np.exp(-0.5 * ((X - x.mean()) ** 2 + (Y - y.mean()) ** 2)) * np.exp(1.2 * (Z - z[-1]))
    return (f + rng.standard_normal(f.shape)).astype(F32).reshape(-1)


@pytest.mark.parametrize("with_mid", [True, False])
def test_shift_window_matches_the_oracle_composition(gm, with_mid):
    import torch

    ops = gm.ops
    L1 = make_level((50, 20, 30), ((0.0, 10.0), (0.0, 4.0), (-4.0, 2.0)))
    L2 = make_level((100, 100, 10), ((0.4, 4.4), (0.0, 4.0), (-0.4, 0.0)))
    old = make_level((100, 100, 10), ((1.4, 3.4), (1.0, 3.0), (-0.2, 0.0)))
    new = make_level((100, 100, 10), ((1.4 + 3 * 0.04, 3.4 + 3 * 0.04), (1.0 - 2 * 0.04, 3.0 - 2 * 0.04), (-0.2, 0.0)))
    T1, Tp2, Tp3 = _field(L1, 1, 300.0, 900.0), _field(L2, 2, -20.0, 35.0), _field(old, 3, -10.0, 15.0)
    tgt = new["node_coords"]
    want_tp = cF.interpolatePoints(old, Tp3, tgt)
    t1 = cF.interpolatePoints(L1, T1, tgt)
    want_T = (t1 + ((cF.interpolatePoints(L2, Tp2, tgt) + want_tp) if with_mid else want_tp)).astype(F32)
    c = lambda lv: [_dev(a) for a in lv["node_coords"]]
    Tp_new, T_new = torch.empty(new["nn"], device="cuda"), torch.empty(new["nn"], device="cuda")
    ops.shift_window(c(L1), _dev(T1), c(old), _dev(Tp3), c(new), Tp_new, T_new,
                     mid_coords=c(L2) if with_mid else None, Tp_mid=_dev(Tp2) if with_mid else None)
    torch.cuda.synchronize()
    got_tp, got_T = Tp_new.cpu().numpy(), T_new.cpu().numpy()
    assert np.array_equal(got_tp == 0, want_tp == 0)   # same support (nodes outside the old window get exactly 0)
    assert np.abs(got_tp - want_tp).max() <= 2e-6 * np.abs(want_tp).max()
    assert np.abs(got_T - want_T).max() <= 2e-6 * np.abs(want_T).max()
    # ... and bit for bit what the general interpolation kernel gives for the same three interpolations
    a, b, m = (torch.empty(new["nn"], device="cuda") for _ in range(3))
    ops.interp(c(old), _dev(Tp3), c(new), a)
    ops.interp(c(L1), _dev(T1), c(new), b)
    rest = a
    if with_mid:
        ops.interp(c(L2), _dev(Tp2), c(new), m)
        rest = m + a
    assert torch.equal(Tp_new, a) and torch.equal(T_new, b + rest)


def test_accum_single_step_matches_the_driver_sequence(gm):
    import torch

    rng = np.random.default_rng(5)
    big = (41, 23, 9)
    nx, ny, nz = 12, 7, 4
    iv = [np.arange(5, 5 + nx, dtype=np.int32), np.arange(9, 9 + ny, dtype=np.int32), np.arange(2, 2 + nz, dtype=np.int32)]
    idx = (iv[0][None, None, :] + iv[1][None, :, None] * big[0] + iv[2][:, None, None] * big[0] * big[1]).reshape(-1)
    nn0 = big[0] * big[1] * big[2]
    acc = (rng.random(nn0) * 1e-3).astype(F32)
    mx = (rng.random(nn0) * 1e-3).astype(F32)
    T3 = (1500.0 + 200.0 * rng.random(nx * ny * nz)).astype(F32)
    reset_mask = rng.random(nx * ny * nz) > 0.6
    dt, Tliq = F32(1e-5), F32(1609.0)
    # gm:339-357: reset = accum[idx] * (all_reset > 0); max = max(reset, max[idx]); accum[idx] -= reset; melting_temp
    reset = acc[idx] * reset_mask.astype(F32)
    want_mx = mx.copy()
    want_mx[idx] = np.maximum(reset, mx[idx])
    want_acc = acc.copy()
    want_acc[idx] = (acc[idx] + (-reset)).astype(F32)
    want_acc[idx] = (want_acc[idx] + (T3 > Tliq).astype(F32) * dt).astype(F32)
    dacc, dmx = _dev(acc), _dev(mx)
    gm.ops.accum_single_step(_dev(T3), _dev(reset_mask), float(dt), float(Tliq), dacc, dmx, [_dev(v) for v in iv], big[0], big[1])
    torch.cuda.synchronize()
    assert np.array_equal(dmx.cpu().numpy(), want_mx) and np.array_equal(dacc.cpu().numpy(), want_acc)


def test_patch_copy_boxes(gm):
    import torch

    ops = gm.ops
    rng = np.random.default_rng(7)
    sd, dd = (37, 11, 9), (20, 8, 30)
    src = rng.standard_normal(sd[2] * sd[1] * sd[0]).astype(F32)
    dst = rng.standard_normal(dd[2] * dd[1] * dd[0]).astype(F32)
    for slo, dlo, n in (((3, 2, 1), (0, 0, 0), (17, 6, 8)), ((0, 0, 0), (0, 1, 21), (20, 7, 9)), ((36, 10, 8), (19, 7, 29), (1, 1, 1)),
                        ((5, 5, 5), (5, 5, 5), (7, 0, 3))):
        want = dst.reshape(dd[2], dd[1], dd[0]).copy()
        want[dlo[2]:dlo[2] + n[2], dlo[1]:dlo[1] + n[1], dlo[0]:dlo[0] + n[0]] = \
            src.reshape(sd[2], sd[1], sd[0])[slo[2]:slo[2] + n[2], slo[1]:slo[1] + n[1], slo[0]:slo[0] + n[0]]
        d = _dev(dst)
        ops.patch_copy(_dev(src), sd, slo, d, dd, dlo, n)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy(), want.reshape(-1)), (slo, dlo, n)
    # pack into / unpack from a contiguous staging buffer (what a box transfer does on either side of a send / recv)
    n, lo = (9, 4, 5), (11, 3, 2)
    buf = torch.empty(n[0] * n[1] * n[2], device="cuda")
    ops.patch_copy(_dev(src), sd, lo, buf, n, (0, 0, 0), n)
    want = src.reshape(sd[2], sd[1], sd[0])[lo[2]:lo[2] + n[2], lo[1]:lo[1] + n[1], lo[0]:lo[0] + n[0]].reshape(-1)
    assert np.array_equal(buf.cpu().numpy(), want)
    with pytest.raises(gm.GomeltError):   # a box that leaves the destination
        ops.patch_copy(_dev(src), sd, (0, 0, 0), _dev(dst), dd, (15, 0, 0), (10, 2, 2))


def test_clamp_min(gm):
    import torch

    x = np.linspace(-5.0, 5.0, 100001, dtype=F32)
    d = _dev(x)
    gm.ops.clamp_min(d, 1.25)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), np.maximum(x, F32(1.25)))

"""CPU: the math of the warp-level candidate test (block_candidate in csrc/common.cuh), restated in numpy
fp32, never rejects an instance that contributes to some pixel of the block — over random and extreme
(needle-thin, huge, tiny-opacity, centre inside/outside) splats. The CUDA function itself is covered by the
GPU parity tests, which require bit-identical n_contrib / images against the reference rasterizer."""
import numpy as np

f32 = np.float32


def block_candidate(cx, cy, A, B, C, o, bx0, bx1, by0, by1):
    cx, cy, A, B, C, o = map(f32, (cx, cy, A, B, C, o))
    bx0, bx1, by0, by1 = map(f32, (bx0, bx1, by0, by1))
    chk = cx + cy + A + B + C + o
    if chk != chk:
        return True
    if not (A > 0 and C > 0 and A * C - B * B > 0):
        return True
    if o <= 0:
        return False
    tau = max(f32(np.log(f32(255.0) * o)), f32(0)) + f32(2e-3)
    dxn = min(max(cx, bx0), bx1) - cx
    dyn = min(max(cy, by0), by1) - cy
    Dx = max(abs(bx0 - cx), abs(bx1 - cx))
    Dy = max(abs(by0 - cy), abs(by1 - cy))
    E = f32(1e-6) * (A + C + f32(2) * abs(B)) * (Dx * Dx + Dy * Dy)
    qmin = f32(0)
    if dxn != 0 or dyn != 0:
        qmin = f32(np.inf)
        if dxn != 0:
            dy = min(max(-B * dxn / C, by0 - cy), by1 - cy)
            qmin = f32(0.5) * (A * dxn * dxn + C * dy * dy) + B * dxn * dy
        if dyn != 0:
            dx = min(max(-B * dyn / A, bx0 - cx), bx1 - cx)
            qmin = min(qmin, f32(0.5) * (A * dx * dx + C * dyn * dyn) + B * dx * dyn)
    return not (qmin > tau + E)


def contributes(cx, cy, A, B, C, o, bx0, by0):
    """Any pixel of the 8x4 block with power <= 0 and min(0.99, o*exp(power)) >= 1/255, evaluated in fp32 with
    the rasterizer's operation order (forward.cu:356-363)."""
    px = (np.arange(8, dtype=f32) + f32(bx0))[None, :]
    py = (np.arange(4, dtype=f32) + f32(by0))[:, None]
    dx = f32(cx) - px
    dy = f32(cy) - py
    inner = dx * (f32(A) * dx) + (f32(C) * dy) * dy
    power = inner * f32(-0.5) - (f32(B) * dx) * dy
    alpha = np.minimum(f32(0.99), f32(o) * np.exp(power.astype(f32)))
    return bool(np.any((power <= 0) & (alpha >= f32(1.0 / 255.0))))


def test_never_rejects_a_contributing_instance():
    rng = np.random.default_rng(0)
    rejected = hits = 0
    for it in range(60000):
        kind = it % 4
        s1 = 10 ** rng.uniform(-0.3, 2.5)          # sigma of the major axis, px
        s2 = s1 * 10 ** rng.uniform(-3, 0) if kind else s1
        th = rng.uniform(0, np.pi)
        c, s = np.cos(th), np.sin(th)
        cov = np.array([[c * c * s1 * s1 + s * s * s2 * s2, c * s * (s1 * s1 - s2 * s2)],
                        [c * s * (s1 * s1 - s2 * s2), s * s * s1 * s1 + c * c * s2 * s2]]) + 0.3 * np.eye(2)
        q = np.linalg.inv(cov)
        A, B, C = f32(q[0, 0]), f32(q[0, 1]), f32(q[1, 1])
        o = f32(10 ** rng.uniform(-2.6, 0)) if kind != 3 else f32(rng.uniform(0.0035, 0.0045))
        bx0, by0 = float(rng.integers(0, 1900)), float(rng.integers(0, 1070))
        # centres concentrated near the cutoff contour so both outcomes occur
        r = np.sqrt(2 * max(np.log(255 * float(o)), 0.01)) * rng.uniform(0.6, 1.4)
        ang = rng.uniform(0, 2 * np.pi)
        L = np.linalg.cholesky(cov)
        off = L @ (r * np.array([np.cos(ang), np.sin(ang)]))
        cx = f32(bx0 + rng.uniform(0, 7) + off[0])
        cy = f32(by0 + rng.uniform(0, 3) + off[1])
        cand = block_candidate(cx, cy, A, B, C, o, bx0, bx0 + 7, by0, by0 + 3)
        hit = contributes(cx, cy, A, B, C, o, bx0, by0)
        hits += hit
        rejected += (not cand)
        assert cand or not hit, (cx, cy, A, B, C, o, bx0, by0)
    # the test must be doing work in both directions
    assert hits > 5000 and rejected > 5000


def test_degenerate_inputs_are_kept():
    nan = float("nan")
    assert block_candidate(nan, 1, 1, 0, 1, 0.5, 0, 7, 0, 3)
    assert block_candidate(1e6, 1e6, 1, 0, 1, nan, 0, 7, 0, 3)
    assert block_candidate(1e6, 1e6, 1, 2, 1, 0.5, 0, 7, 0, 3)        # not positive definite
    assert block_candidate(1e6, 1e6, -1, 0, 1, 0.5, 0, 7, 0, 3)
    assert not block_candidate(3, 2, 1, 0, 1, 0.0, 0, 7, 0, 3)       # alpha == 0 everywhere
    assert not block_candidate(3, 2, 1, 0, 1, -0.5, 0, 7, 0, 3)
    assert block_candidate(3, 2, 1, 0, 1, 1.0 / 255.0, 0, 7, 0, 3)   # exactly at the threshold, centre on a pixel

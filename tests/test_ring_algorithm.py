"""CPU: the ring backward's ALGORITHM (csrc/blend_ring.cuh) restated in numpy / float64 against the
reference's per-pixel backward loop (backward.cu:546-674, colour-only, one channel):

* the schedule - lane l of a 16-lane ring holds element 16 m + l of the concatenated contributor lists of its blocks
  during steps [16 m + l, 16 (m + 1) + l); the pixel slot at lane l in step s is (s - l) mod 16; an element flagged
  first-of-block re-initialises every slot that reaches it - visits every (instance, pixel) pair of every block
  exactly once, each pixel meeting its block's instances in list order back to front;
* the per-instance moment sums (S0 .. S5, S7) and the eager accum_rec recursion give the reference's gradients.

The CUDA kernel is compared with the reference CUDA on the GPU; this pins the algebra and the index arithmetic."""
import numpy as np

LANES = 16


def forward_block(inst, px, py, bg):
    """Reference forward over one 4x4 block (forward.cu:331-396 semantics): per pixel final T, n_contrib (1-based position
    of the last blended instance), colour."""
    T = np.ones(16)
    C = np.zeros(16)
    ncontrib = np.zeros(16, int)
    done = np.zeros(16, bool)
    for j, (x, y, A, B, Cc, o, col) in enumerate(inst):
        dx, dy = x - px, y - py
        power = -0.5 * (A * dx * dx + Cc * dy * dy) - B * dx * dy
        alpha = np.minimum(0.99, o * np.exp(power))
        act = ~done & ~(power > 0) & ~(alpha < 1.0 / 255.0)
        test_T = T * (1 - alpha)
        stop = act & (test_T < 1e-4)
        done |= stop
        blend = act & ~stop
        C = np.where(blend, C + col * alpha * T, C)
        T = np.where(blend, test_T, T)
        ncontrib = np.where(blend, j + 1, ncontrib)
    return T, ncontrib, C + T * bg


def reference_backward(inst, px, py, T_final, ncontrib, dL, bg, ddelx, ddely):
    K = len(inst)
    g = np.zeros((K, 7))   # dmean.x, dmean.y, dconic.xx, .xy, .yy, dopacity, dcolour
    for p in range(16):
        T = T_final[p]
        accum_rec = last_alpha = last_color = 0.0
        for j in range(ncontrib[p] - 1, -1, -1):
            x, y, A, B, Cc, o, col = inst[j]
            dx, dy = x - px[p], y - py[p]
            power = -0.5 * (A * dx * dx + Cc * dy * dy) - B * dx * dy
            if power > 0:
                continue
            G = np.exp(power)
            alpha = min(0.99, o * G)
            if alpha < 1.0 / 255.0:
                continue
            T = T / (1 - alpha)
            accum_rec = last_alpha * last_color + (1 - last_alpha) * accum_rec
            last_color = col
            dL_dalpha = (col - accum_rec) * dL[p]
            g[j, 6] += alpha * T * dL[p]
            dL_dalpha *= T
            last_alpha = alpha
            dL_dalpha += (-T_final[p] / (1 - alpha)) * (bg * dL[p])
            dL_dG = o * dL_dalpha
            gdx, gdy = G * dx, G * dy
            g[j, 0] += dL_dG * (-gdx * A - gdy * B) * ddelx
            g[j, 1] += dL_dG * (-gdy * Cc - gdx * B) * ddely
            g[j, 2] += -0.5 * gdx * dx * dL_dG
            g[j, 3] += -0.5 * gdx * dy * dL_dG
            g[j, 4] += -0.5 * gdy * dy * dL_dG
            g[j, 5] += G * dL_dalpha
    return g


def ring_backward(blocks, bg, ddelx, ddely):
    """blocks: list of (inst, px, py, T_final, ncontrib, dL). Returns one gradient array per block."""
    # the concatenated sequence, every block back to front; element = (block, instance index, first-of-block)
    seq = []
    for b, (inst, *_rest) in enumerate(blocks):
        for n, j in enumerate(range(len(inst) - 1, -1, -1)):
            seq.append((b, j, n == 0))
    grads = [np.zeros((len(blk[0]), 7)) for blk in blocks]
    sums = {}                                   # element index -> S0..S5, S7
    slot = [None] * LANES                       # per lane: the pixel state currently there
    visited = set()
    nsteps = len(seq) + LANES
    for s in range(nsteps):
        slot = [slot[(l - 1) % LANES] for l in range(LANES)]          # the shuffle: state moves lane -> lane + 1
        for l in range(LANES):
            if s < l:
                continue
            k = LANES * ((s - l) // LANES) + l                         # the element lane l holds in this step
            if k >= len(seq):
                continue
            b, j, first = seq[k]
            inst, px, py, T_final, ncontrib, dL = blocks[b]
            p = (s - l) % LANES
            if first:                                                  # slot becomes pixel p of block b
                slot[l] = dict(b=b, p=p, T=T_final[p], Rp=0.0)
            st = slot[l]
            assert st is not None and st["b"] == b and st["p"] == p, (s, l, k)
            assert (b, j, p) not in visited
            visited.add((b, j, p))
            x, y, A, B, Cc, o, col = inst[j]
            dx, dy = x - px[p], y - py[p]
            power = -0.5 * (A * dx * dx + Cc * dy * dy) - B * dx * dy
            G = np.exp(power)
            alpha = min(0.99, o * G)
            if not (j < ncontrib[p]) or power > 0 or alpha < 1.0 / 255.0:
                continue
            S = sums.setdefault(k, np.zeros(7))
            oma = 1 - alpha
            st["T"] = st["T"] / oma
            w = alpha * st["T"]
            dL_dalpha = (col - st["Rp"]) * dL[p]
            st["Rp"] = col * alpha + oma * st["Rp"]                    # accum_rec as the next contributor will see it
            dL_dalpha *= st["T"]
            dL_dalpha += (-T_final[p] / oma) * (bg * dL[p])
            h = G * dL_dalpha
            q = o * h
            S += (h, q * dx, q * dy, q * dx * dx, q * dx * dy, q * dy * dy, w * dL[p])
    for k, S in sums.items():
        b, j, _ = seq[k]
        _, _, A, B, Cc, _, _ = blocks[b][0][j]
        grads[b][j] = (-ddelx * (A * S[1] + B * S[2]), -ddely * (Cc * S[2] + B * S[1]), -0.5 * S[3], -0.5 * S[4],
                       -0.5 * S[5], S[0], S[6])
    for b, (inst, *_r) in enumerate(blocks):
        assert sum(1 for v in visited if v[0] == b) == len(inst) * 16   # every (instance, pixel) pair exactly once
    return grads


def make_block(rng, K, ox, oy, opacity_hi):
    px = ox + np.arange(16) % 4 + 0.0
    py = oy + np.arange(16) // 4 + 0.0
    inst = []
    for _ in range(K):
        a, c = rng.uniform(0.05, 1.5, 2)
        b = rng.uniform(-0.9, 0.9) * np.sqrt(a * c)
        inst.append((ox + rng.uniform(-3, 7), oy + rng.uniform(-3, 7), a, b, c, rng.uniform(0.005, opacity_hi),
                     rng.uniform(0, 1)))
    return inst, px, py


def test_ring_schedule_and_moments_match_the_per_pixel_loop():
    rng = np.random.default_rng(0)
    for bg in (0.0, 0.35):
        blocks = []
        # lists shorter than, equal to and longer than a ring epoch; opaque enough to stop pixels early in one block
        for K, hi in ((5, 0.6), (16, 0.9), (41, 0.3), (1, 0.9), (70, 0.99)):
            inst, px, py = make_block(rng, K, rng.integers(0, 50) * 4.0, rng.integers(0, 50) * 4.0, hi)
            T_final, ncontrib, _ = forward_block(inst, px, py, bg)
            blocks.append((inst, px, py, T_final, ncontrib, rng.normal(size=16)))
        got = ring_backward(blocks, bg, 0.5 * 800, 0.5 * 600)
        for (inst, px, py, T_final, ncontrib, dL), g in zip(blocks, got):
            ref = reference_backward(inst, px, py, T_final, ncontrib, dL, bg, 0.5 * 800, 0.5 * 600)
            scale = np.abs(ref).max(axis=0) + 1e-300
            assert (np.abs(g - ref) / scale).max() < 1e-11

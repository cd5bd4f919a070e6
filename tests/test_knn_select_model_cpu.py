"""CPU model of the k-NN selection kernel's LOGIC (csrc/knn_select.cuh) in numpy: block of 4^3 cells, region of 8^3, quadrant
sub-regions of 6 x 6 columns, float32 histogram pass, band pass, exact float64 classification against the band edges, count
correction, rank inside the exact list, acceptance against the nearest face of the swept sub-region.  Every query the model
accepts must equal the brute-force K-th distance bit for bit; the others are the ones the kernel hands to the lock-step
kernel.  (The CUDA kernel itself is tested against scipy in tests/test_gpu_knn.py; this test pins the algorithm on the CPU.)"""
import numpy as np

NB, DELTA, LIST, BS, R = 32, np.float32(4e-3 / 32), 16, 4, 2


def kth_brute(pos, q, k, box):
    d = pos - q
    d = d + np.where(d < -0.5 * box, box, np.where(d > 0.5 * box, -box, 0.0))          # scipy's wrap
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    return np.partition(d2, k - 1)[k - 1], d2


def stage(u, sc):
    v = ((u - 4.0) * sc).astype(np.float32)
    w = (v[:, 0] * v[:, 0]).astype(np.float32)
    w = (v[:, 1] * v[:, 1] + w).astype(np.float32)
    return v, (v[:, 2] * v[:, 2] + w).astype(np.float32)


def test_selection_logic_matches_brute_force():
    rng = np.random.default_rng(11)
    G, box, k = 16, 1.0, 24
    n = int(1.9 * G ** 3)
    pos = rng.uniform(0, box, (n, 3))
    cs = box / G
    cell = np.minimum((pos / cs).astype(np.int64), G - 1)
    dref = (R + 0.5) * cs
    sc = cs / dref
    accepted = fallback = 0
    for b in [(0, 0, 0), (1, 2, 3), (3, 3, 3), (2, 0, 1)]:                              # blocks incl. the wrapping corners
        b0 = np.array(b) * BS
        rlo = b0 - R
        rel = (cell - rlo) % G                                                          # region-relative cell of every particle
        in_region = np.all(rel < BS + 2 * R, axis=1)
        cand = np.nonzero(in_region)[0]
        u = pos[cand] / cs - rlo
        u = np.where(u < 0, u + G, np.where(u >= G, u - G, u))
        vc, wc = stage(u, sc)
        for qx in range(2):
            for qy in range(2):
                s0 = b0 + np.array([2 * qx, 2 * qy, 0])
                colsel = (rel[cand, 0] >= 2 * qx) & (rel[cand, 0] < 2 * qx + 6) & (rel[cand, 1] >= 2 * qy) & (rel[cand, 1] < 2 * qy + 6)
                sub = np.nonzero(colsel)[0]                                             # the six runs of the staged list
                in_quad = np.all(cell[cand] >= s0, axis=1) & (cell[cand, 0] <= s0[0] + 1) & (cell[cand, 1] <= s0[1] + 1) & (cell[cand, 2] <= b0[2] + 3)
                for qi in np.nonzero(in_quad)[0][:40]:
                    q = pos[cand[qi]]
                    m2 = (np.float32(-2.0) * vc[qi]).astype(np.float32)
                    q2 = wc[qi]
                    t = (m2[2] * vc[sub, 2] + wc[sub]).astype(np.float32)
                    t = (m2[1] * vc[sub, 1] + t).astype(np.float32)
                    dot = (m2[0] * vc[sub, 0] + t).astype(np.float32)
                    s = np.clip((dot + q2).astype(np.float32), np.float32(0), np.float32(1))
                    bins = np.minimum((s * np.float32(NB)).astype(np.int64), NB)         # floor(32 s), 32 when s = 1
                    hist = np.bincount(bins, minlength=NB + 1)[:NB]
                    cum = np.cumsum(hist)
                    truth, d2_all = kth_brute(pos, q, k, box)
                    frac = q - cell[cand[qi]] * cs
                    below = cell[cand[qi]] - s0 + R
                    above = np.array([s0[0] + 1, s0[1] + 1, b0[2] + 3]) + R - cell[cand[qi]]
                    dmin = min((frac + below * cs).min(), ((cs - frac) + above * cs).min())
                    safe2 = (dmin - 1e-9 * cs) ** 2
                    ok = cum[-1] >= k
                    ans = np.inf
                    if ok:
                        bstar = int(np.argmax(cum >= k))
                        cnt = int(cum[bstar] - hist[bstar])
                        elo = np.float32(bstar) * np.float32(1.0 / NB)
                        midq = ((np.float32(bstar) + np.float32(0.5)) * np.float32(1.0 / NB) - q2).astype(np.float32)
                        hw = np.float32(0.5 / NB) + DELTA
                        band = np.nonzero(np.abs((dot - midq).astype(np.float32)) <= hw)[0]
                        TL, TH = bstar / NB * dref * dref, (bstar + 1) / NB * dref * dref
                        exact = []
                        ok = len(band) <= LIST
                        for j in band:
                            if s[j] < elo:
                                cnt -= 1
                            d2 = d2_all[cand[sub[j]]]
                            if d2 < TL:
                                cnt += 1
                            elif d2 <= TH:
                                exact.append(d2)
                        want = k - cnt
                        ok = ok and 1 <= want <= len(exact)
                        if ok:
                            ans = np.sort(exact)[want - 1]
                            ok = ans < safe2
                    if ok:
                        accepted += 1
                        assert ans == truth, (b, qx, qy, ans, truth)
                    else:
                        fallback += 1
    assert accepted > 300 and accepted > 3 * fallback, (accepted, fallback)

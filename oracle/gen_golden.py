#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running THE REFERENCE'S OWN projection code (oracle/_ref, built by
oracle/build_ref.py from /root/reference with non-arithmetic patches only) on seeded inputs.

The reference cannot travel to the GPU box, so its outputs are committed as small fixtures together
with this script (TEST INFRASTRUCTURE).  Each file holds the inputs and the reference map, so tests do
not depend on regenerating the inputs bit-for-bit.

  python oracle/gen_golden.py            # the fast cases (about a minute)
  python oracle/gen_golden.py --full     # adds the 64^3 -> 512^2 known-answer summary (about 3 minutes)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from astro_sph_tools_b200 import synthetic  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def wendland_c2_2d(r, h):
    """Wendland C2, 2-D normalisation, H = 2h: a plain Python callable for kernel_func."""
    r = np.asarray(r); h = np.asarray(h)
    H = 2.0 * h
    u = r / H
    t = 1.0 - u
    t = t * t
    t = t * t
    return np.where(u < 1.0, 7.0 / (np.pi * H * H) * t * (1.0 + 4.0 * u), 0.0)


def main():
    os.makedirs(GOLD, exist_ok=True)
    mod, Axes = oracle.reference_module()
    create_image, qsk = mod.create_image, mod.quartic_spline_kernel
    axes = {0: Axes.X, 1: Axes.Y, 2: Axes.Z}

    def run(name, pos, h, prop, npix, axis, bounds, chunk=32, kernel="cubic_spline_3d", extra=None):
        t0 = time.time()
        kf = qsk if kernel == "cubic_spline_3d" else wendland_c2_2d
        img = create_image(pos, h, prop, (npix, npix), chunk, axes[axis], *bounds, kernel_func=kf)
        dt = time.time() - t0
        d = dict(pos=pos, h=h, prop=prop, npix=npix, axis=axis, bounds=np.array(bounds, dtype=np.float64),
                 kernel=kernel, ref_map=img, ref_seconds=dt)
        if extra:
            d.update(extra)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **d)
        print(f"{name}: N={len(h)} {npix}^2 {kernel} axis={axis}  {dt:.2f}s  sum={img.sum():.15g} max={img.max():.15g}")
        return img

    # 1. single particle known answer (SURVEY App. C sanity list)
    run("single_particle", np.array([[3.0, 7.0, 5.0]]), np.array([0.6]), np.array([2.0]), 10, 2, (0.0, 10.0, 0.0, 10.0), chunk=5)
    for ax in (0, 1):
        run(f"single_particle_axis{ax}", np.array([[3.0, 7.0, 5.0]]), np.array([0.6]), np.array([2.0]), 10, ax,
            (0.0, 10.0, 0.0, 10.0), chunk=5)

    # 2. S1 lattice, SPH-realistic h = d_48 (periodic k-NN), default kernel
    s = synthetic.s1(16, k=48)
    run("s1_n16_p128_cubic_z", s["pos"], s["h"], s["mass"], 128, 2, (0.0, 1.0, 0.0, 1.0))
    # 3. config-1 shape at reduced size: periodic ghosts + Wendland C2 surface density through the reference
    P, H, M = synthetic.add_periodic_ghosts(s["pos"], s["h"], s["mass"], 1.0, cols=(0, 1))
    run("s1_n16_p128_wc2_periodic", P, H, M, 128, 2, (0.0, 1.0, 0.0, 1.0), kernel="wendland_c2_2d",
        extra=dict(base_pos=s["pos"], base_h=s["h"], base_prop=s["mass"], box=1.0))
    # 4. random cloud, window smaller than the cloud, signed weights, other axes
    rng = np.random.default_rng(777)
    pos = rng.uniform(0, 10, (3000, 3)); h = rng.uniform(0.0, 1.2, 3000); prop = rng.normal(size=3000)
    for ax in (0, 1, 2):
        run(f"cloud_axis{ax}", pos, h, prop, 64, ax, (2.0, 9.0, 1.0, 8.0), chunk=50)
    # 5. sub-pixel smoothing lengths (most particles touch 0-4 pixels) + a few huge ones
    h2 = np.concatenate([rng.uniform(0.0, 0.12, 2990), rng.uniform(4.0, 30.0, 10)])
    run("cloud_tiny_and_huge_h", pos, h2, np.abs(prop), 96, 2, (0.0, 10.0, 0.0, 10.0))
    # 6. temperature-weighted pair (config-2 shape at reduced size): two calls, mass and mass*T
    s = synthetic.s1(12, k=32, with_temperature=True)
    run("s1_n12_p96_mass", s["pos"], s["h"], s["mass"], 96, 2, (0.0, 1.0, 0.0, 1.0))
    run("s1_n12_p96_massT", s["pos"], s["h"], s["mass"] * s["T"], 96, 2, (0.0, 1.0, 0.0, 1.0))
    # 7. SURVEY 8(c) known answer: n=32, 128^2, h=d_48
    s = synthetic.s1(32, k=48)
    img = run("s1_n32_p128_cubic_z", s["pos"], s["h"], s["mass"], 128, 2, (0.0, 1.0, 0.0, 1.0))
    A = (1.0 / 128) ** 2
    print("   SURVEY says sum*A=9.352326287134296 max=10.307522073191322 img[0,0]=2.467664335010954 img[64,42]=9.879513044112905")
    print(f"   here        sum*A={img.sum() * A!r} max={img.max()!r} img[0,0]={img[0, 0]!r} img[64,42]={img[64, 42]!r}")

    if "--full" in sys.argv:
        # config 1 at full size; only a summary + a strided sample of the map is stored (2 MB map otherwise)
        s = synthetic.s1(64, k=48)
        t0 = time.time()
        img = create_image(s["pos"], s["h"], s["mass"], (512, 512), 32, axes[2], 0.0, 1.0, 0.0, 1.0, kernel_func=qsk)
        dt = time.time() - t0
        A = (1.0 / 512) ** 2
        np.savez_compressed(os.path.join(GOLD, "s1_n64_p512_cubic_z_summary.npz"), sumA=img.sum() * A, max=img.max(),
                            p00=img[0, 0], p256_170=img[256, 170], sample=img[::8, ::8].copy(), ref_seconds=dt)
        print(f"s1_n64_p512 (config 1, reference CPU path, 1 core): {dt:.1f}s sum*A={img.sum() * A!r} max={img.max()!r} "
              f"img[0,0]={img[0, 0]!r} img[256,170]={img[256, 170]!r}")
        print("   SURVEY says sum*A=19.31872920058476 max=20.378954886050302 img[0,0]=5.020691341041981 img[256,170]=19.93688569782197")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Build the reference's own projection path into oracle/_ref/ (TEST INFRASTRUCTURE, never shipped).

What this does: the reference's hot path
(`/root/reference/src/astro_sph_tools/tools/projections/`) does not build or run as shipped because of
five non-arithmetic defects (D1-D5).  This script re-packages those files from
where they lie under /root/reference into the git-ignored directory oracle/_ref/pkg, applies the
non-arithmetic patches, and cythonises the two .pyx files.  No arithmetic line is altered.

  D1  _projector.py:11 imports `._pixel_calculation` but the file is `_pixel_calculations.pyx`
  D2  _pixel_calculations.pyx:15-16 declares `cdef double[:]` memoryviews and then does numpy
      arithmetic on them (Cython 3 compiler crash) -> wrap the three inputs with np.asarray
  D3  _projector.py:63 passes bytes to a `char` parameter -> take element [0]
  D4  pyproject.toml:62 include_numpy=false although both .pyx `cimport numpy` -> add the include path
  D5  QuasarCode (external, not installed) is only used for Console.print_debug -> no-op stub

Nothing under oracle/_ref is tracked by git.  It is *not* gpurun-ignored, so the built tree travels to
the GPU box where `bench.py --impl reference` and the parity tests may import it.
Run:  python oracle/build_ref.py        (needs /root/reference; a no-op message otherwise)
"""
import os, shutil, subprocess, sys, sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AST_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "src", "astro_sph_tools")
OUT = os.path.join(HERE, "_ref", "pkg")


def build(force: bool = False) -> bool:
    if not os.path.isdir(SRC):
        print(f"[oracle/_ref] {SRC} not present: keeping any prebuilt oracle/_ref as is")
        return os.path.isdir(OUT)
    proj = os.path.join(OUT, "astro_sph_tools", "tools", "projections")
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    done = all(os.path.exists(os.path.join(proj, m + suffix)) for m in ("_kernels", "_pixel_calculation"))
    if done and not force:
        return True
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(proj)
    os.makedirs(os.path.join(OUT, "QuasarCode"))

    def cp(rel_src, rel_dst):
        dst = os.path.join(OUT, "astro_sph_tools", rel_dst)
        shutil.copyfile(os.path.join(SRC, rel_src), dst)
        os.chmod(dst, 0o644)

    cp("_CoordinateAxes.py", "_CoordinateAxes.py")
    for f in ("__init__.py", "_projector.py", "_kernels.pyx"):
        cp(os.path.join("tools", "projections", f), os.path.join("tools", "projections", f))
    # empty package inits: the real ones pull io/ -> h5py/unyt/QuasarCode which are not installed
    open(os.path.join(OUT, "astro_sph_tools", "__init__.py"), "w").close()
    open(os.path.join(OUT, "astro_sph_tools", "tools", "__init__.py"), "w").close()

    # D1 + D2
    s = open(os.path.join(SRC, "tools", "projections", "_pixel_calculations.pyx")).read()
    old = "    cdef double[:] dx, dy\n    cdef double[:] r2, r, weights\n"
    assert old in s, "reference source changed: D2 patch anchor not found"
    s = s.replace(old, "    positions_np = np.asarray(positions); smoothing_lengths_np = np.asarray(smoothing_lengths); "
                       "particle_properties_np = np.asarray(particle_properties)\n")
    s = (s.replace("positions[:,", "positions_np[:,")
          .replace("(2.0 * smoothing_lengths)", "(2.0 * smoothing_lengths_np)")
          .replace("smoothing_lengths[mask]", "smoothing_lengths_np[mask]")
          .replace("particle_properties[mask]", "particle_properties_np[mask]"))
    open(os.path.join(proj, "_pixel_calculation.pyx"), "w").write(s)
    # D3
    p = os.path.join(proj, "_projector.py")
    t = open(p).read()
    assert "str(projection_axis).encode()," in t, "reference source changed: D3 patch anchor not found"
    open(p, "w").write(t.replace("str(projection_axis).encode(),", "str(projection_axis).encode()[0],"))
    # D5
    open(os.path.join(OUT, "QuasarCode", "__init__.py"), "w").write(
        "class Console:\n    @staticmethod\n    def print_debug(*a, **k):\n        pass\n")
    # D4 + cythonize
    import numpy
    env = dict(os.environ)
    env["CFLAGS"] = (env.get("CFLAGS", "") + " -I" + numpy.get_include() + " -O2 -w").strip()
    subprocess.check_call([sys.executable, "-m", "cython", "--version"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call(["cythonize", "-i", "-3", "-q", "_kernels.pyx", "_pixel_calculation.pyx"], cwd=proj, env=env)
    shutil.rmtree(os.path.join(proj, "build"), ignore_errors=True)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    if ok:
        sys.path.insert(0, OUT)
        from astro_sph_tools.tools.projections import create_image, quartic_spline_kernel  # noqa: F401
        print("[oracle/_ref] built and importable:", OUT)
    sys.exit(0 if ok else 1)

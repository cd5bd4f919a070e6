"""Importable alias of the package directory ``astro-sph-tools_b200/``.

The task fixes the package directory name to ``astro-sph-tools_b200`` (hyphens), which Python cannot
import directly.  This stub makes ``import astro_sph_tools_b200`` resolve every submodule inside that
directory (``astro_sph_tools_b200.tools.projections`` -> ``astro-sph-tools_b200/tools/projections``).
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "astro-sph-tools_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f

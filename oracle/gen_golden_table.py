"""TEST INFRASTRUCTURE: generates tests/golden/tables/ion_table_golden.npz by running the REFERENCE's own IonisationTableBase
(/root/reference/src/astro_sph_tools/data_structures/_IonisationTable.py) on a small synthetic table.  The class file is
executed from where it lies; its one package-relative import (`.._Interface`, an abstract-class guard with no arithmetic)
is satisfied by a stub module.  Run in the build container only: python oracle/gen_golden_table.py"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from table_util import gas_state, synthetic_table  # noqa: E402

SRC = "/root/reference/src/astro_sph_tools/data_structures/_IonisationTable.py"


def load_reference_class():
    pkg = types.ModuleType("refpkg"); pkg.__path__ = []
    sub = types.ModuleType("refpkg.data_structures"); sub.__path__ = []
    iface = types.ModuleType("refpkg._Interface")

    class Interface:          # stub of astro_sph_tools/_Interface.py (abstract-class bookkeeping only)
        pass

    iface.Interface = Interface
    iface.ensure_not_interface = lambda cls, base: None
    sys.modules.update({"refpkg": pkg, "refpkg.data_structures": sub, "refpkg._Interface": iface})
    spec = importlib.util.spec_from_file_location("refpkg.data_structures._IonisationTable", SRC)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod.IonisationTableBase


if __name__ == "__main__":
    Base = load_reference_class()
    table, axes = synthetic_table(seed=21, shape=(13, 17, 7), uniform=False)
    t = Base(table, *axes, redshift_input_index=2)
    x = gas_state(5, 600, axes)
    x2 = np.ascontiguousarray(gas_state(6, 400, axes)[:, :2])
    z = 0.5 * (axes[2][2] + axes[2][3])
    out = os.path.join(ROOT, "tests", "golden", "tables", "ion_table_golden.npz")
    np.savez_compressed(out, table=table, axis0=axes[0], axis1=axes[1], axis2=axes[2], x=x, call=t(x), x2=x2, redshift=z,
                        at_redshift=t.evaluate_at_redshift(x2, z))
    print("wrote", out, os.path.getsize(out), "bytes")

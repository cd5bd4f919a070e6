"""numpy restatement of the reference's single-process ArrayReorder (tools/_ArrayReorder.py:988-1038 create, :937-961 call),
shared by the CPU pin test (against golden results produced by the reference's own class, oracle/gen_golden_reorder.py)
and the -m gpu parity tests.  Test infrastructure only."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reorder")


def reference_reorder(source_order, target_order, source_data, default_value, source_filter=None, target_filter=None):
    """the reference's arithmetic: argsort + np.isin membership filters + order conversion (lines 988-1032), then
    output[dest_filter] = source_data[source_filter][order_conversion_indexes] (line 959)"""
    sos, tos = source_order.argsort(), target_order.argsort()
    sus, tus = sos.argsort(), tos.argsort()
    t_search = target_order[tos] if target_filter is None else target_order[tos][target_filter[tos]]
    s_search = source_order[sos] if source_filter is None else source_order[sos][source_filter[sos]]
    fwd = np.isin(source_order[sos], t_search)[sus]
    bwd = np.isin(target_order[tos], s_search)[tus]
    if source_filter is not None: fwd &= source_filter
    if target_filter is not None: bwd &= target_filter
    sos2, tos2 = source_order[fwd].argsort(), target_order[bwd].argsort()
    conv = sos2[tos2.argsort()]
    if default_value is None:
        out = np.empty((len(target_order),) + source_data.shape[1:], dtype=source_data.dtype)
    else:
        out = np.full((len(target_order),) + source_data.shape[1:], default_value, dtype=source_data.dtype)
    out[bwd] = source_data[fwd][conv]
    return out, fwd, bwd


def golden_reorder_cases():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_reorder(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: g[k] for k in g.files}
    d.setdefault("source_order_filter", None); d.setdefault("target_order_filter", None)
    d["default_value"] = d["default_value"][()] if "default_value" in d else None
    return d

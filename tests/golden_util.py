"""Load a golden fixture (tests/golden/*.npz, produced by the unmodified reference) and bind it to a context exactly
the way a host walking `MappedWorkspace::maps` would: one device array per distinct bound container, DoF sets in
registration order, one potential per reference potential with its fetch table."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, name, data=None):
        """`data`: an already packed dump (tests/golden/make_golden.py:pack) instead of a committed fixture."""
        self.z = data if data is not None else np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.meta = json.loads(bytes(self.z["meta_json"]))
        self.name = name

    def __getitem__(self, k):
        return self.z[k]

    def potentials(self, only_nonempty=True):
        for i, p in enumerate(self.meta["potentials"]):
            if only_nonempty and p["n_elements"] == 0:
                continue
            yield i, p


def bind(ctx, g, kernel_names, rename=None, skip=(), override=None, split_fetch=None):
    """Returns {reference potential index: context potential handle}.
    override: {array index: replacement data}; split_fetch: (potential name, first_symbol) whose binding gets its own
    copy of the array (same values, different device buffer -- exercises non-canonical fetch tables)."""
    rename = rename or {}
    override = override or {}
    arrays = {}
    for i, a in enumerate(g.meta["arrays"]):
        arrays[i] = ctx.array(f"a{i}", a["stride"], override.get(i, g[f"array{i}"]))
    ids = [a["id"] for a in g.meta["arrays"]]
    for did in g.meta["dof_array_ids"]:
        ctx.dof_add(arrays[ids.index(did)])
    handles = {}
    for i, p in g.potentials():
        name = rename.get(p["name"], p["name"])
        if (name not in kernel_names and not p.get("user_ops")) or p["name"] in skip:
            continue
        fetch = [(arrays[m["array"]], m["conn_idx"], m["first_symbol"], m["stride"]) for m in p["maps"]]
        if p.get("user_ops"):
            # a user potential: the operation sequences the reference produced for it go through the NVRTC back-end
            h = ctx.potential_user(name, p["conn_stride"], fetch, p["n_in"], g[f"pot{i}_block_slots"],
                                   (g[f"pot{i}_ops_p"], g[f"pot{i}_opsc_p"]), (g[f"pot{i}_ops_pgh"], g[f"pot{i}_opsc_pgh"]))
            ctx.set_connectivity(h, g[f"pot{i}_conn"][g[f"pot{i}_active"].astype(bool)])
            handles[i] = h
            continue
        if split_fetch and split_fetch[0] == p["name"]:
            for k, m in enumerate(p["maps"]):
                if m["first_symbol"] == split_fetch[1]:
                    dup = ctx.array(f"dup{i}_{k}", m["stride"], override.get(m["array"], g[f"array{m['array']}"]))
                    fetch[k] = (dup, m["conn_idx"], m["first_symbol"], m["stride"])
        h = ctx.potential(name, p["conn_stride"], fetch)
        conn = g[f"pot{i}_conn"]
        active = g[f"pot{i}_active"].astype(bool)
        ctx.set_connectivity(h, conn[active])
        handles[i] = h
    return handles

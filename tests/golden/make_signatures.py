"""Dump the call signatures (parameter names, kinds, defaults) of every function and method the reference defines in
src/{model,fields,joint_prediction,point_prediction,sim,stat_tools}.py into tests/golden/signatures.json.

    python tests/golden/make_signatures.py        (build container only: needs /root/reference)
"""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_loader  # noqa: E402

MODULES = ("model", "fields", "joint_prediction", "point_prediction", "sim", "stat_tools")


def shape(f):
    return [[p.name, p.kind.name, None if p.default is inspect._empty else repr(p.default)]
            for p in inspect.signature(f).parameters.values()]


def collect(ns) -> dict:
    out = {}
    for mod in MODULES:
        m = getattr(ns, mod) if not isinstance(ns, dict) else ns[mod]
        for name, obj in vars(m).items():
            if name.startswith("__") or getattr(obj, "__module__", None) != m.__name__:
                continue
            if inspect.isfunction(obj):
                out[f"{mod}.{name}"] = shape(obj)
            elif inspect.isclass(obj):
                for mname, meth in vars(obj).items():
                    if isinstance(meth, (staticmethod, classmethod)):
                        meth = meth.__func__
                    if inspect.isfunction(meth):
                        out[f"{mod}.{name}.{mname}"] = shape(meth)
    return out


if __name__ == "__main__":
    sigs = collect(ref_loader.load())
    with open(os.path.join(HERE, "signatures.json"), "w") as f:
        json.dump(sigs, f, indent=0, sort_keys=True)
    print(len(sigs), "signatures written")

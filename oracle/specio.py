"""TEST INFRASTRUCTURE — (de)serialise a rollout spec dict to a flat .npz (no pickle).

Nested dicts use '/'-joined keys; lists/tuples use integer components; scalars, strings,
bools and None are stored in one JSON blob under '__meta__'."""
from __future__ import annotations

import json

import numpy as np


def flatten(obj, prefix="", arrays=None, meta=None):
    arrays = {} if arrays is None else arrays
    meta = {} if meta is None else meta
    if isinstance(obj, dict):
        meta[prefix + "@"] = "dict"
        for k, v in obj.items():
            flatten(v, f"{prefix}{str(k).replace('/', '|')}/", arrays, meta)
    elif isinstance(obj, (list, tuple)):
        meta[prefix + "@"] = f"list:{len(obj)}"
        for i, v in enumerate(obj):
            flatten(v, f"{prefix}{i}/", arrays, meta)
    elif isinstance(obj, np.ndarray):
        arrays[prefix.rstrip("/")] = obj
    elif isinstance(obj, (np.floating, np.integer)):
        meta[prefix.rstrip("/")] = obj.item()
    else:
        if isinstance(obj, float) and not np.isfinite(obj):
            obj = {"__float__": repr(obj)}
        meta[prefix.rstrip("/")] = obj
    return arrays, meta


def unflatten(arrays, meta):
    def build(prefix):
        tag = meta.get(prefix + "@")
        if tag == "dict":
            keys = []
            for src in (arrays, meta):
                for k in src:
                    if k.startswith(prefix) and k != prefix + "@":
                        head = k[len(prefix):].split("/")[0].rstrip("@")
                        if head and head not in keys:
                            keys.append(head)
            return {k.replace("|", "/"): build(f"{prefix}{k}/") for k in keys}
        if tag is not None and tag.startswith("list:"):
            n = int(tag.split(":")[1])
            return [build(f"{prefix}{i}/") for i in range(n)]
        key = prefix.rstrip("/")
        if key in arrays:
            return np.asarray(arrays[key])
        v = meta[key]
        if isinstance(v, dict) and "__float__" in v:
            return float(v["__float__"])
        return v

    return build("")


def save(path, obj):
    arrays, meta = flatten(obj)
    np.savez_compressed(path, __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)


def load(path):
    z = np.load(path, allow_pickle=False)
    meta = json.loads(bytes(z["__meta__"]).decode())
    arrays = {k: z[k] for k in z.files if k != "__meta__"}
    out = unflatten(arrays, meta)
    _tuples(out)
    return out


def _tuples(spec):
    """hidden layers are lists of (w, b) pairs."""
    def fix(d):
        if isinstance(d, dict):
            for k, v in d.items():
                if k == "hidden" and isinstance(v, list):
                    d[k] = [tuple(p) for p in v]
                else:
                    fix(v)
        elif isinstance(d, list):
            for v in d:
                fix(v)
    fix(spec)

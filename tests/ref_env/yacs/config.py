"""Minimal stand-in for ``yacs.config.CfgNode`` (yacs is not installed in this image).

Covers what the reference's lib/config/default.py and tools/zero_shot.py use: attribute access on a dict,
``CN(new_allowed=True)``, ``defrost`` / ``freeze``, ``merge_from_file`` (YAML), ``merge_from_list``
(KEY.PATH value pairs, values parsed as Python literals), ``dump`` and ``clone``.  Unknown keys raise unless the
node (or an ancestor) was created with ``new_allowed=True`` - the behaviour the reference's CUSTOM / TEST nodes
rely on."""
from __future__ import annotations

import ast
import copy

import yaml


class CfgNode(dict):
    IMMUTABLE = "__immutable__"
    NEW_ALLOWED = "__new_allowed__"

    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        self.__dict__[CfgNode.IMMUTABLE] = False
        self.__dict__[CfgNode.NEW_ALLOWED] = new_allowed
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v, new_allowed=new_allowed) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__[CfgNode.IMMUTABLE]:
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        self[name] = value

    def is_frozen(self):
        return self.__dict__[CfgNode.IMMUTABLE]

    def is_new_allowed(self):
        return self.__dict__[CfgNode.NEW_ALLOWED]

    def _set_immutable(self, flag):
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_immutable(flag)

    def freeze(self):
        self._set_immutable(True)

    def defrost(self):
        self._set_immutable(False)

    def clone(self):
        return copy.deepcopy(self)

    def _merge(self, other, path):
        for k, v in other.items():
            full = ".".join(path + [k])
            if k not in self:
                if not self.is_new_allowed():
                    raise KeyError(f"Non-existent config key: {full}")
                self[k] = CfgNode(v, new_allowed=True) if isinstance(v, dict) else v
                continue
            if isinstance(self[k], CfgNode):
                if not isinstance(v, dict):
                    raise ValueError(f"{full}: expected a mapping")
                self[k]._merge(v, path + [k])
            else:
                self[k] = _coerce(v, self[k], full)

    def merge_from_file(self, cfg_filename):
        with open(cfg_filename, "r") as f:
            self._merge(yaml.safe_load(f) or {}, [])

    def merge_from_other_cfg(self, other):
        self._merge(other, [])

    def merge_from_list(self, cfg_list):
        cfg_list = list(cfg_list or [])
        if len(cfg_list) % 2:
            raise ValueError(f"Override list has odd length: {cfg_list}; it must be a list of pairs")
        for full, raw in zip(cfg_list[0::2], cfg_list[1::2]):
            *parents, leaf = full.split(".")
            node = self
            for p in parents:
                if p not in node:
                    raise KeyError(f"Non-existent key: {full}")
                node = node[p]
            try:
                value = ast.literal_eval(raw) if isinstance(raw, str) else raw
            except (ValueError, SyntaxError):
                value = raw
            if leaf in node:
                value = _coerce(value, node[leaf], full)
            elif not node.is_new_allowed():
                raise KeyError(f"Non-existent key: {full}")
            node[leaf] = value

    def _plain(self):
        return {k: (v._plain() if isinstance(v, CfgNode) else v) for k, v in self.items()}

    def dump(self, **kwargs):
        return yaml.safe_dump(self._plain(), **kwargs)

    def __str__(self):
        return self.dump()

    def __repr__(self):
        return f"CfgNode({dict.__repr__(self)})"

    def __deepcopy__(self, memo):
        out = CfgNode(new_allowed=self.is_new_allowed())
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        out.__dict__[CfgNode.IMMUTABLE] = self.is_frozen()
        return out


def _coerce(value, current, key):
    """yacs' type check: the replacement must have the type of the default (with the tuple/list and
    int->float conversions yacs allows)."""
    if current is None or value is None or type(value) is type(current):
        return value
    if isinstance(current, (list, tuple)) and isinstance(value, (list, tuple)):
        return type(current)(value)
    if isinstance(current, float) and isinstance(value, int):
        return float(value)
    if isinstance(current, str) and not isinstance(value, str):
        return str(value)
    raise ValueError(f"Type mismatch ({type(current)} vs. {type(value)}) for config key: {key}")

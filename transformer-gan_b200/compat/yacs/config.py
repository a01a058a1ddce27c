"""``CfgNode``: attribute-style nested config with ``merge_from_file`` / ``merge_from_list`` / ``freeze`` / ``defrost`` /
``clone`` / ``dump`` (the calls made at train.py:147,1340-1343, generate.py:114-127,369-370, batch_generate.py:26-67
and utils/config_helper.py).  Semantics follow yacs: merging only accepts keys that already exist, values must keep
their type (int -> float and tuple <-> list are allowed), a frozen node rejects assignment."""
import copy

import yaml

_VALID = (tuple, list, str, int, float, bool, type(None))


class CfgNode(dict):
    IMMUTABLE = "__immutable__"

    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        init_dict = {} if init_dict is None else init_dict
        key_list = [] if key_list is None else key_list
        super().__init__()
        self.__dict__[CfgNode.IMMUTABLE] = False
        for k, v in init_dict.items():
            self[k] = CfgNode(v, key_list + [k]) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    # attribute access -------------------------------------------------------------------------------------
    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        if name in self.__dict__:
            raise AttributeError(f"Invalid attempt to modify internal CfgNode state: {name}")
        if not isinstance(value, _VALID + (dict,)):
            raise AssertionError(f"Invalid type {type(value)} for key {name}")
        self[name] = value

    # immutability -----------------------------------------------------------------------------------------
    def is_frozen(self):
        return self.__dict__[CfgNode.IMMUTABLE]

    def _immutable(self, flag):
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._immutable(flag)

    def freeze(self):
        self._immutable(True)

    def defrost(self):
        self._immutable(False)

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        out.__dict__[CfgNode.IMMUTABLE] = self.is_frozen()
        return out

    # merging ------------------------------------------------------------------------------------------------
    @staticmethod
    def _coerce(new, old, key):
        if old is None or new is None or type(new) is type(old):
            return new
        if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
            return float(new)
        if isinstance(old, tuple) and isinstance(new, list):
            return tuple(new)
        if isinstance(old, list) and isinstance(new, tuple):
            return list(new)
        raise ValueError(f"Type mismatch ({type(old)} vs. {type(new)}) with values ({old} vs. {new}) for config key: {key}")

    def _merge(self, other, path):
        for k, v in other.items():
            full = ".".join(path + [k])
            if k not in self:
                raise KeyError(f"Non-existent config key: {full}")
            if isinstance(v, dict):
                if not isinstance(self[k], CfgNode):
                    raise ValueError(f"config key {full} is not a node")
                self[k]._merge(v, path + [k])
            else:
                dict.__setitem__(self, k, self._coerce(copy.deepcopy(v), self[k], full))

    def merge_from_other_cfg(self, other):
        # like yacs, merging goes through item assignment and is NOT blocked by freeze() (train.py:147 merges the
        # experiment file into the frozen defaults)
        self._merge(other, [])

    def merge_from_file(self, cfg_filename):
        with open(cfg_filename, "r") as f:
            loaded = yaml.safe_load(f) or {}
        self.merge_from_other_cfg(loaded)

    def merge_from_list(self, cfg_list):
        if len(cfg_list) % 2:
            raise AssertionError(f"Override list has odd length: {cfg_list}; it must be a list of pairs")
        for full, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node, keys = self, full.split(".")
            for k in keys[:-1]:
                if k not in node:
                    raise KeyError(f"Non-existent config key: {full}")
                node = node[k]
            if keys[-1] not in node:
                raise KeyError(f"Non-existent config key: {full}")
            if isinstance(v, str):
                try:
                    v = yaml.safe_load(v)
                except yaml.YAMLError:
                    pass
            dict.__setitem__(node, keys[-1], self._coerce(v, node[keys[-1]], full))

    # output -------------------------------------------------------------------------------------------------
    def _plain(self):
        return {k: (v._plain() if isinstance(v, CfgNode) else v) for k, v in self.items()}

    def dump(self, **kwargs):
        return yaml.safe_dump(self._plain(), **kwargs)

    def __str__(self):
        return self.dump(default_flow_style=False)

    def __repr__(self):
        return f"CfgNode({dict.__repr__(self)})"

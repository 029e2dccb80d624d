"""Minimal stand-in for `yacs.config.CfgNode` (test infrastructure only).

The reference's FlowFormer config (`Module/Network/FlowFormer/configs/submission.py`)
only uses `CN()`, attribute assignment, item access and `.clone()`.
"""
import copy


class CfgNode(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

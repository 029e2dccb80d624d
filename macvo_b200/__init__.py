"""Import alias: the product package directory is `mac-vo_b200/` (not a valid Python identifier).

`import macvo_b200` executes `mac-vo_b200/__init__.py` and resolves sub-modules from that directory.
"""
import os as _os

_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mac-vo_b200")
__path__ = [_dir]
__file__ = _os.path.join(_dir, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))

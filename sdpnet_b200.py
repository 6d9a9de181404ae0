"""Import shim: the package directory is `sdp-net_b200/` (not a Python identifier), so
`import sdpnet_b200` loads it from there under this name."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "sdp-net_b200")
_spec = _u.spec_from_file_location("sdpnet_b200", _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["sdpnet_b200"] = _mod
_spec.loader.exec_module(_mod)

"""Import shim: the product package directory is named `simplemoc-kernel_b200/` (not a
valid Python identifier), so `import smk_b200` loads it through importlib."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)

_pkg = importlib.import_module("simplemoc-kernel_b200")
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
package = _pkg

"""Import alias: the package directory is `pacbot-rl_b200/` (not a valid identifier), so
`import pacbot_rl_b200` loads it through importlib and aliases the module tree."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("pacbot-rl_b200")
for _name, _mod in list(sys.modules.items()):
    if _name == "pacbot-rl_b200" or _name.startswith("pacbot-rl_b200."):
        sys.modules["pacbot_rl_b200" + _name[len("pacbot-rl_b200"):]] = _mod
sys.modules[__name__] = _pkg

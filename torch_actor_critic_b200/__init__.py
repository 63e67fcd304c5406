"""Importable alias of the product package directory `torch-actor-critic_b200/` (a hyphen cannot appear in a
Python module name).  `import torch_actor_critic_b200 as tac` resolves every submodule from that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "torch-actor-critic_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _os, _f

"""Import shim: the package directory is named ``go-snark-study_b200`` (not a
valid Python identifier), so ``import gosnark_b200`` maps onto it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "go-snark-study_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))

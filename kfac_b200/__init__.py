"""Importable alias of the `kfac-pytorch_b200/` package directory.

`import kfac_b200.preconditioner` resolves to
`kfac-pytorch_b200/preconditioner.py` (the hyphenated directory name required
by the project layout cannot be imported directly).
"""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                              'kfac-pytorch_b200'))

from kfac_b200 import enums  # noqa: E402,F401
from kfac_b200.enums import (AllreduceMethod, AssignmentStrategy,  # noqa: E402,F401
                             ComputeMethod, DistributedStrategy)

__version__ = '0.1.0'

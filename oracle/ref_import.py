"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Makes the UNMODIFIED reference (gpauloski/kfac-pytorch @ /root/reference)
importable in THIS container so that golden vectors can be generated from it
(`oracle/gen_golden.py`) and the CPU restatement (`oracle/kfac_oracle.py`)
can be pinned against it.  `/root/reference` does not exist on the GPU box:
nothing under tests/ marked `gpu`, `smoke()` or `bench.py` may call this.

`kfac/__init__.py:19` reads `importlib.metadata.version('kfac-pytorch')`,
which fails for a source tree that is not pip-installed; we put a stub
`kfac_pytorch-0.4.2.dist-info/METADATA` on sys.path.
"""
from __future__ import annotations

import os
import sys
import tempfile

REFERENCE_ROOT = '/root/reference'


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'kfac'))


def import_reference():
    """Return the reference `kfac` package (imported from /root/reference)."""
    if not reference_available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    scratch = os.path.join(tempfile.gettempdir(), 'kfac_ref_distinfo')
    info = os.path.join(scratch, 'kfac_pytorch-0.4.2.dist-info')
    os.makedirs(info, exist_ok=True)
    meta = os.path.join(info, 'METADATA')
    if not os.path.exists(meta):
        with open(meta, 'w') as f:
            f.write('Metadata-Version: 2.1\nName: kfac-pytorch\nVersion: 0.4.2\n')
    for p in (scratch, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import kfac  # noqa: F401
    import kfac.preconditioner  # noqa: F401
    return kfac

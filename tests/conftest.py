import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build what is missing once, like __graft_entry__.build()."""
    from mapad_b200 import build as b
    need = not os.path.exists(b.LIB) or not os.path.exists(os.path.join(ROOT, "oracle", "libmapad_oracle.so")) or \
        not os.path.exists(os.path.join(ROOT, "tests", "emu", "libmapad_emu.so"))
    if need:
        sys.path.insert(0, ROOT)
        import __graft_entry__ as entry
        entry.build()

"""TEST INFRASTRUCTURE ONLY (see dgl/__init__.py)."""


class DGLError(Exception):
    pass

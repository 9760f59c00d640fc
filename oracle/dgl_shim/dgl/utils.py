"""TEST INFRASTRUCTURE ONLY (see dgl/__init__.py)."""
from .base import DGLError


def expand_as_pair(input_, g=None):
    if isinstance(input_, tuple):
        return input_
    return input_, input_


def check_eq_shape(input_):
    srcdata, dstdata = expand_as_pair(input_)
    if tuple(srcdata.shape[1:]) != tuple(dstdata.shape[1:]):
        raise DGLError("The feature shape of source nodes and destination nodes mismatch")

"""Seed / call-counter bookkeeping for the fused Philox noise.

Every fused forward consumes one ``offset`` (the Philox call counter); the backward
of that call regenerates the same noise from the saved (seed, offset).  This plays
the role torch's global CUDA generator plays for the reference
(``expand([E,K]).rsample()``, stag/layers.py:117-127): ``manual_seed`` makes a run
reproducible, successive calls are independent.
"""
import threading

_state = threading.local()


def _st():
    if not hasattr(_state, "seed"):
        _state.seed = 0x5EED5EED
        _state.offset = 0
    return _state


def manual_seed(seed):
    st = _st()
    st.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    st.offset = 0


def get_state():
    st = _st()
    return st.seed, st.offset


def set_state(seed, offset):
    st = _st()
    st.seed, st.offset = int(seed), int(offset)


def next_offset():
    """Reserve one Philox call counter and return (seed, offset)."""
    st = _st()
    off = st.offset
    st.offset = off + 1
    return st.seed, off

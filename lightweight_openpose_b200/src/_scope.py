"""Variable-scope bookkeeping that reproduces TensorFlow-1 auto-naming.

Every ``tf.layers`` call in the reference passes ``name=''`` (reference
``src/net_factory.py:4,7,14,17,23,27``), so TF names variables per
``tf.variable_scope`` in creation order: ``conv2d``, ``conv2d_1``, ... /
``separable_conv2d``, ... / ``batch_normalization``, ...  Counters are per scope
*name* and restart on each new ``lightweight_openpose()`` trace
(``reuse=tf.AUTO_REUSE`` re-enters the same names)."""
from __future__ import annotations

import contextlib
import threading
from typing import Dict, List

_state = threading.local()


def _stack() -> List[str]:
    if not hasattr(_state, "stack"):
        _state.stack = []
        _state.counts = {}
    return _state.stack


def reset() -> None:
    _stack()
    _state.counts = {}


@contextlib.contextmanager
def variable_scope(name: str):
    st = _stack()
    st.append(name)
    try:
        yield
    finally:
        st.pop()


def next_name(kind: str) -> str:
    st = _stack()
    scope = "/".join(st)
    counts: Dict[str, int] = _state.counts
    key = scope + "|" + kind
    n = counts.get(key, 0)
    counts[key] = n + 1
    leaf = kind + (f"_{n}" if n else "")
    return f"{scope}/{leaf}" if scope else leaf

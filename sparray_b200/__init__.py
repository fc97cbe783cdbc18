"""sparray_b200 -- a B200-native (sm_100a CUDA) implementation of the FlatSparray
hot path of perimosocordiae/sparray, behind the reference's own API
(``from sparray_b200 import FlatSparray, is_sparray`` mirrors
``sparray/__init__.py:1-2``)."""
from .base import is_sparray
from .flat import FlatSparray

__all__ = ["FlatSparray", "is_sparray"]

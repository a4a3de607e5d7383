"""ngacf_b200 -- B200-native (sm_100a) implementation of NGACF's SPUIGACF propagation-and-scoring hot
path behind the reference's own Python API.  See DESIGN.md."""
from ._lib import NgacfError, load as load_library  # noqa: F401

__all__ = ["NgacfError", "load_library"]

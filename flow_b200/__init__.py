"""flow_b200 -- B200-native engine behind the Python API of nschloe/flow's time-stepping
hot path.  Mirrors /root/reference/flow/__init__.py:3-5 (`message`, `navier_stokes`,
`stokes`; `heat` is imported explicitly, as in tests/test_boussinesq.py:12)."""
from . import message  # noqa: F401
from . import navier_stokes  # noqa: F401
from . import stokes  # noqa: F401

__all__ = ["message", "navier_stokes", "stokes"]

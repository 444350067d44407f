from .pressure_correction import Chorin, IPCS, Rotational, last_stats  # noqa: F401

from .pressure_correction import Chorin, IPCS, Rotational, last_stats, set_options, reset_options  # noqa: F401

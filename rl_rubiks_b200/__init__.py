"""rl_rubiks_b200: B200-native (sm_100a) cube dynamics for rl-rubiks, a drop-in for `librubiks.cube`,
the ADI batch generator of `librubiks.train` and the search frontier of `librubiks.solving.agents`.
See DESIGN.md.  Importing requires the built C-ABI library (no CPU fallback)."""
from . import _native  # noqa: F401  (fails loudly when librubiks_b200.so is missing)
from . import cube  # noqa: F401

__all__ = ["cube", "adi", "frontier", "sharding"]

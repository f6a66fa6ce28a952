"""refshim: the legacy (pre-Linen) flax API surface the reference uses (see ../README.md)."""
from . import nn  # noqa: F401
from . import optim  # noqa: F401
from . import serialization  # noqa: F401

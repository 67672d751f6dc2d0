"""`model.from_config(scope, name, **kw)` (recad/model/__init__.py:3-21) for the B200 victims."""
from . import victim
from .attacker import Aush

factories = {"victim": victim.factories, "attacker": {"aush": Aush}}


def from_config(scope, name, **kwargs):
    return factories[scope][name].from_config(**kwargs)

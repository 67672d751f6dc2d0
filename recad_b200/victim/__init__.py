from .base import BaseVictim
from .lightgcn import LightGCN
from .mf import MF
from .ncf import NCF

factories = {"lightgcn": LightGCN, "mf": MF, "ncf": NCF}   # recad/model/__init__.py:3-5

"""tatva_b200 — B200-native implementation of tatva's element-level hot path.

Same public names as the reference package (`from tatva import Mesh, Operator, element`):
    Mesh, Operator, element.{Tri3, Tetrahedron4, Hexahedron8}, sparse, compound, mpi
"""
from . import element  # noqa: F401
from . import materials  # noqa: F401
from .mesh import Mesh  # noqa: F401
from .operator import Operator  # noqa: F401

__version__ = "0.1.0"

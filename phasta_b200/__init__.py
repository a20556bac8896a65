"""phasta_b200 -- B200-native drop-in for PHASTA's compressible implicit-step
hot path (element assembly + EBE GMRES).  See DESIGN.md / INTEGRATION.md."""
from .params import SolverParams, IncompParams  # noqa: F401
from .mesh import MeshPart, make_box, make_state, make_smooth_state, global_node_count, nondimensional  # noqa: F401
from .tables import make_tables  # noqa: F401

__all__ = ["SolverParams", "IncompParams", "MeshPart", "make_box", "make_state", "make_smooth_state", "global_node_count", "make_tables", "nondimensional"]

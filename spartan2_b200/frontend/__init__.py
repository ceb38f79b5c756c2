"""Workload frontends (host side): circuits -> padded R1CS shapes + witnesses, in the layout SplitR1CSShape::new
produces.  Stand-in for the reference's bellpepper frontend (OUT OF SCOPE for the hot path)."""
from .sha256 import Sha256Circuit, build_frontend  # noqa: F401

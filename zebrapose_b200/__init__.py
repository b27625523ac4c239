"""zebrapose_b200 -- B200-native (sm_100a) post-network pose path of ZebraPose: binary-code decode -> 2D-3D
correspondences -> RANSAC-EPnP, behind the reference's Python signatures.  See DESIGN.md / INTEGRATION.md."""
from ._lib import ZpError, STATUS_OK, STATUS_NO_MASK, STATUS_TOO_FEW, STATUS_NO_MODEL  # noqa: F401
from .engine import Engine, Pipeline, default_engine, dict_to_table  # noqa: F401
from .sharding import shard_range, gather_poses  # noqa: F401

__version__ = "0.1.0"

"""pafuse_b200: B200-native (sm_100a) implementation of PAFUSE's denoising inference path."""
from .diffusionpose import D3DP  # noqa: F401
from .mixste import MixSTE2  # noqa: F401
from .utils import aggregate_hypotheses, eval_data_prepare, project_to_2d, wb_pose_from_parts  # noqa: F401
from .h3wb import H3WBSkeleton  # noqa: F401

__version__ = "0.1.0"

"""libcpab_b200 -- B200-native CPAB transformation hot path (drop-in for
SkafteNicki/libcpab with backend='pytorch', device='gpu').

    from libcpab_b200 import Cpab
    T = Cpab([3, 3], backend='pytorch', device='gpu')
    theta = T.sample_transformation(64)
    out = T.transform_data(images, theta, outsize=(256, 256))
"""
from .cpab import Cpab
from .sequential import CpabSequential
from .alignment import CpabAligner

__all__ = ["Cpab", "CpabSequential", "CpabAligner"]
__version__ = "0.1.0"

"""wbc_quadruped_dob_b200 -- B200-native batched whole-body-control cycle.

`api` is the host-side mirror of the reference interface on top of the C ABI (include/wbc_b200.h);
`scenarios` generates the deterministic synthetic inputs that stand in for Gazebo/ROS/TOWR.
"""
from . import scenarios  # noqa: F401

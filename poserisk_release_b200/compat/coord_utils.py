from poserisk_release_b200.coord_utils import *  # noqa: F401,F403
from poserisk_release_b200.coord_utils import axis_angle_to_euler_angle, get_joint_cam, rot_to_angle  # noqa: F401

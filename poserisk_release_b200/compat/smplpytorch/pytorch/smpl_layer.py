from poserisk_release_b200.smpl_layer import SMPL_Layer  # noqa: F401

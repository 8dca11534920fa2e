from poserisk_release_b200.smpl import SMPL  # noqa: F401

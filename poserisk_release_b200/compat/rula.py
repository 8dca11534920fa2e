from poserisk_release_b200.rula import RULA  # noqa: F401

from poserisk_release_b200.reba import REBA  # noqa: F401

"""Moved to workloads/synth.py (input generation is not part of the checker); kept as an alias for the tests."""
from workloads.synth import *  # noqa: F401,F403
from workloads.synth import LM_K, YCBV_K, TLESS_K  # noqa: F401

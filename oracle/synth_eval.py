"""Moved to workloads/synth_eval.py; kept as an alias for the tests."""
from workloads.synth_eval import *  # noqa: F401,F403
from workloads.synth_eval import MODELS, N_PAIRS, N_BOXES, PAD_RATIOS, METHODS  # noqa: F401

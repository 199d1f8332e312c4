"""Shadows ip_adapter/resampler.py: same class name and state-dict keys, engine underneath."""
from followyourclick_b200.ip_adapter import Resampler  # noqa: F401

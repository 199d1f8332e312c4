"""Shadows animatediff/models/unet.py: same class name, engine underneath."""
from followyourclick_b200.unet import UNet3DConditionModel, UNet3DConditionOutput  # noqa: F401

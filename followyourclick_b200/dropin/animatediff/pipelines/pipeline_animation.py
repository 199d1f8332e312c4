"""Shadows animatediff/pipelines/pipeline_animation.py: same class names, engine underneath."""
from followyourclick_b200.pipeline_animation import AnimationPipeline, AnimationPipelineOutput  # noqa: F401

"""CLIP ViT image / text towers used for the cross-modal alignment targets."""
from .model import CLIP, VisionTransformer, build_model, convert_weights, load  # noqa: F401

"""Device-side data layer: see device_pipeline.py."""
from .device_pipeline import DeviceSceneAugmentor, draw_augmentation  # noqa: F401

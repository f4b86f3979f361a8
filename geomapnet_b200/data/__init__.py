"""GPU-side input pipeline (staged)."""
from .preprocess import ImagePipeline, resize_output_size  # noqa: F401

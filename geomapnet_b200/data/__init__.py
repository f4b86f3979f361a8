"""GPU-side input pipeline: Resize -> [ColorJitter] -> ToTensor -> Normalize with the
MF / MFOnline tuple gather, on uint8 frames resident in device memory."""
from .preprocess import ImagePipeline, ColorJitterSampler, resize_output_size  # noqa: F401
from . import tuples  # noqa: F401

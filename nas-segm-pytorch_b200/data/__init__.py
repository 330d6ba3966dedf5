"""Device-side counterpart of the reference's data transforms (src/data/datasets.py): SURVEY 8(f) row f3."""
from .augment import AugmentedLoader, GpuTrainTransform, GpuValTransform, make_even, resized_size  # noqa: F401

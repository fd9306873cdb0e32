"""Import alias: ``import b200face`` -> the package that lives in
``facerecognition-multiarchitecture-pipeline_b200/`` (a directory name Python cannot import
directly because of the hyphens).  This file only redirects ``__path__``."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "facerecognition-multiarchitecture-pipeline_b200")
__path__.append(_PKG_DIR)

from ._lib import lib_path, load_library, LibraryMissingError  # noqa: E402,F401
from .head import (  # noqa: E402,F401
    ArcMarginProduct, ArcFaceNet, arcface_loss, head_schedule, HeadStats, GraphedHeadStep, LazyArcLogits,
)
from .gallery import (  # noqa: E402,F401
    compare_faces, gallery_topk, gallery_topk_batches, cosine_class_match, GalleryIndex, PreparedGallery,
)
from .optim import HeadAdamW  # noqa: E402,F401
from . import parallel  # noqa: E402,F401

__all__ = [
    "ArcMarginProduct", "ArcFaceNet", "arcface_loss", "head_schedule", "HeadStats", "GraphedHeadStep", "LazyArcLogits",
    "compare_faces", "gallery_topk", "gallery_topk_batches", "cosine_class_match", "GalleryIndex", "PreparedGallery",
    "HeadAdamW", "parallel", "lib_path", "load_library", "LibraryMissingError",
]

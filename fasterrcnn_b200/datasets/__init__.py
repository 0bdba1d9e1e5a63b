"""Data path of the reference's training / evaluation loops (pytorch/FasterRCNN/datasets): image loading + preprocessing, the
training-sample records and the PASCAL VOC iterator."""
from . import image, training_sample, voc   # noqa: F401

"""
Image loading and preprocessing (reference: datasets/image.py:34-101).  Decoding is PIL's (the reference decodes through
``imageio.imread(pilmode = "RGB")``, which is PIL's decoder followed by ``convert("RGB")``), resizing is PIL bilinear to a 600-pixel
short side with the reference's truncating size arithmetic, and the per-channel arithmetic keeps the reference's operation order in
fp32 (scale, subtract mean, divide by std) so that the tensors are bit-identical to the reference's for the same file.
"""
import numpy as np
from PIL import Image

from ..backbone import ChannelOrder, PreprocessingParams   # noqa: F401  (re-exported: datasets/image.py:19-32)


def _compute_scale_factor(original_width, original_height, min_dimension_pixels):
  if not min_dimension_pixels:
    return 1.0
  short_side = original_height if original_width > original_height else original_width
  return min_dimension_pixels / short_side


def preprocess(pixels_hwc, preprocessing):
  """(H, W, 3) fp32 RGB -> (3, H, W) fp32 in the backbone's channel order, scaled and standardised."""
  if preprocessing.channel_order == ChannelOrder.BGR:
    pixels_hwc = pixels_hwc[:, :, ::-1]
  elif preprocessing.channel_order != ChannelOrder.RGB:
    raise ValueError("Invalid ChannelOrder value: %s" % str(preprocessing.channel_order))
  out = np.empty((3,) + pixels_hwc.shape[:2], dtype = np.float32)
  scaling = np.float32(preprocessing.scaling)
  for c in range(3):
    plane = pixels_hwc[:, :, c] * scaling                                   # fp32 throughout, one rounding per step as in the reference
    out[c] = (plane - np.float32(preprocessing.means[c])) / np.float32(preprocessing.stds[c])
  return out


def load_image(url, preprocessing, min_dimension_pixels = None, horizontal_flip = False):
  """-> (image_data (3,H,W) fp32, PIL image (scaled), scale_factor, (3, original_height, original_width))."""
  with Image.open(url) as opened:
    image = opened.convert("RGB")
  original_width, original_height = image.width, image.height
  if horizontal_flip:
    image = image.transpose(method = Image.FLIP_LEFT_RIGHT)
  scale_factor = 1.0
  if min_dimension_pixels is not None:
    scale_factor = _compute_scale_factor(image.width, image.height, min_dimension_pixels)
    image = image.resize((int(image.width * scale_factor), int(image.height * scale_factor)), resample = Image.BILINEAR)
  image_data = preprocess(np.array(image).astype(np.float32), preprocessing)
  return image_data, image, scale_factor, (image_data.shape[0], original_height, original_width)

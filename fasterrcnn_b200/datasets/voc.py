"""
PASCAL VOC iterator with the reference's interface and sample order (datasets/voc.py:23-302): same constructor arguments, class
table, annotation rules (1-based corners -> 0-based, ``difficult`` objects skipped unless asked for), shuffle / flip draws from
python's ``random`` in the same sequence, and the same per-file caches.  What differs is where the work runs:

  * anchor maps and the RPN ground-truth map come from the package's device kernels (``anchors.generate_anchor_maps`` /
    ``generate_rpn_map``: frcnn_rpn_decode's anchor generator, frcnn_rpn_targets) -- the reference computes them in NumPy per sample;
    ``anchor_fns`` swaps in other implementations with the same signatures (used by the CPU-only tests);
  * ``prefetch = n`` decodes / resizes / standardises up to n upcoming images on a background thread while the current step runs
    (file decode is the only part of a step that never touches the GPU); the order of samples and of RNG draws is unchanged.
"""
import os
import queue
import random
import threading
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

from . import image
from .training_sample import Box, TrainingSample

_CLASS_NAMES = ("background", "aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow", "diningtable", "dog", "horse",
                "motorbike", "person", "pottedplant", "sheep", "sofa", "train", "tvmonitor")


def _default_anchor_fns():
  from .. import anchors
  return anchors.generate_anchor_maps, anchors.generate_rpn_map


class Dataset:
  num_classes = 21
  class_index_to_name = dict(enumerate(_CLASS_NAMES))

  def __init__(self, split, image_preprocessing_params, compute_feature_map_shape_fn, feature_pixels = 16, dir = "VOCdevkit/VOC2007", augment = True,
               shuffle = True, allow_difficult = False, cache = True, prefetch = 0, anchor_fns = None):
    if not os.path.exists(dir):
      raise FileNotFoundError("Dataset directory does not exist: %s" % dir)
    self.split = split
    self._dir = dir
    self.class_index_to_name = self._get_classes()
    self.class_name_to_index = {name: index for index, name in self.class_index_to_name.items()}
    self.num_classes = len(self.class_index_to_name)
    assert self.num_classes == Dataset.num_classes, "Dataset does not have the expected number of classes (found %d but expected %d)" % (self.num_classes, Dataset.num_classes)
    assert self.class_index_to_name == Dataset.class_index_to_name, "Dataset does not have the expected class mapping"
    self._filepaths = self._get_filepaths()
    self.num_samples = len(self._filepaths)
    self._gt_boxes_by_filepath = self._get_ground_truth_boxes(self._filepaths, allow_difficult)
    self._i = 0
    self._iterable_filepaths = self._filepaths.copy()
    self._image_preprocessing_params = image_preprocessing_params
    self._compute_feature_map_shape_fn = compute_feature_map_shape_fn
    self._feature_pixels = feature_pixels
    self._augment = augment
    self._shuffle = shuffle
    self._cache = cache
    self._unaugmented_cached_sample_by_filepath = {}
    self._augmented_cached_sample_by_filepath = {}
    self._prefetch = int(prefetch)
    self._anchor_fns = anchor_fns
    self._plan = []                      # (filepath, flip) of the current epoch, drawn up front in the reference's RNG order
    self._decoded = None                 # queue of (filepath, flip, load_image result) filled by the prefetch thread
    self._worker = None

  # ---- iteration (voc.py:101-126) ------------------------------------------------------------------
  def __iter__(self):
    self._stop_worker()
    self._i = 0
    if self._shuffle:
      random.shuffle(self._iterable_filepaths)
    # The reference draws one randint per sample inside __next__; nothing else consumes python's RNG between two samples of an
    # epoch EXCEPT the model's own RPN-minibatch sampling (faster_rcnn.py:391-392), so the flips cannot be drawn ahead of time
    # without changing the stream.  Prefetching therefore decodes BOTH orientations' source image once (the decode and resize do
    # not depend on the flip: the flip is applied to the decoded picture) and leaves the draw to __next__.
    if self._prefetch > 0:
      self._start_worker(list(self._iterable_filepaths))
    return self

  def __next__(self):
    if self._i >= len(self._iterable_filepaths):
      self._stop_worker()
      raise StopIteration
    filepath = self._iterable_filepaths[self._i]
    self._i += 1
    flip = random.randint(0, 1) != 0 if self._augment else 0
    cached = self._augmented_cached_sample_by_filepath if flip else self._unaugmented_cached_sample_by_filepath
    predecoded = self._take_prefetched(filepath)
    if filepath in cached:
      sample = cached[filepath]
    else:
      sample = self._generate_training_sample(filepath, flip, predecoded)
    if self._cache:
      cached[filepath] = sample
    return sample

  # ---- sample construction (voc.py:128-164) ------------------------------------------------------------
  def _generate_training_sample(self, filepath, flip, predecoded = None):
    if predecoded is not None:
      scaled_image_data, scaled_image, scale_factor, original_shape = self._finish_decoded(predecoded, flip)
    else:
      scaled_image_data, scaled_image, scale_factor, original_shape = image.load_image(
        url = filepath, preprocessing = self._image_preprocessing_params, min_dimension_pixels = 600, horizontal_flip = flip)
    _, original_height, original_width = original_shape
    scaled_gt_boxes = []
    for box in self._gt_boxes_by_filepath[filepath]:
      corners = box.corners
      if flip:
        corners = np.array([corners[0], original_width - 1 - corners[3], corners[2], original_width - 1 - corners[1]])
      scaled_gt_boxes.append(Box(class_index = box.class_index, class_name = box.class_name, corners = corners * scale_factor))
    generate_anchor_maps, generate_rpn_map = self._anchor_fns or _default_anchor_fns()
    anchor_map, anchor_valid_map = generate_anchor_maps(image_shape = scaled_image_data.shape, feature_map_shape = self._compute_feature_map_shape_fn(scaled_image_data.shape),
                                                        feature_pixels = self._feature_pixels)
    gt_rpn_map, gt_rpn_object_indices, gt_rpn_background_indices = generate_rpn_map(anchor_map = anchor_map, anchor_valid_map = anchor_valid_map, gt_boxes = scaled_gt_boxes)
    return TrainingSample(anchor_map = anchor_map, anchor_valid_map = anchor_valid_map, gt_rpn_map = gt_rpn_map, gt_rpn_object_indices = gt_rpn_object_indices,
                          gt_rpn_background_indices = gt_rpn_background_indices, gt_boxes = scaled_gt_boxes, image_data = scaled_image_data, image = scaled_image,
                          filepath = filepath)

  # ---- prefetch: decode ahead on a thread ----------------------------------------------------------------
  def _start_worker(self, filepaths):
    self._decoded = queue.Queue(maxsize = self._prefetch)
    stop = threading.Event()

    def work():
      from PIL import Image
      for path in filepaths:
        if stop.is_set():
          return
        if path in self._augmented_cached_sample_by_filepath and path in self._unaugmented_cached_sample_by_filepath:
          item = (path, None)                                        # both orientations cached: nothing to decode
        else:
          with Image.open(path) as opened:
            item = (path, opened.convert("RGB"))
        while not stop.is_set():
          try:
            self._decoded.put(item, timeout = 0.1)
            break
          except queue.Full:
            continue

    self._worker = (threading.Thread(target = work, daemon = True), stop)
    self._worker[0].start()

  def _stop_worker(self):
    if self._worker is not None:
      self._worker[1].set()
      self._worker[0].join(timeout = 5)
      self._worker = None
      self._decoded = None

  def _take_prefetched(self, filepath):
    if self._decoded is None:
      return None
    path, picture = self._decoded.get()
    assert path == filepath, "prefetch order diverged from the iteration order"
    return picture

  def _finish_decoded(self, picture, flip):
    """load_image's tail (flip, resize, standardise) on an already decoded RGB picture."""
    from PIL import Image
    original_width, original_height = picture.width, picture.height
    if flip:
      picture = picture.transpose(method = Image.FLIP_LEFT_RIGHT)
    scale_factor = image._compute_scale_factor(picture.width, picture.height, 600)
    picture = picture.resize((int(picture.width * scale_factor), int(picture.height * scale_factor)), resample = Image.BILINEAR)
    data = image.preprocess(np.array(picture).astype(np.float32), self._image_preprocessing_params)
    return data, picture, scale_factor, (data.shape[0], original_height, original_width)

  # ---- directory parsing (voc.py:166-302) ----------------------------------------------------------------
  def _get_classes(self):
    imageset_dir = os.path.join(self._dir, "ImageSets", "Main")
    classes = set(os.path.basename(path).split("_")[0] for path in Path(imageset_dir).glob("*_" + self.split + ".txt"))
    assert len(classes) > 0, "No classes found in ImageSets/Main for '%s' split" % self.split
    class_index_to_name = {1 + i: name for i, name in enumerate(sorted(classes))}
    class_index_to_name[0] = "background"
    return class_index_to_name

  def _get_filepaths(self):
    with open(os.path.join(self._dir, "ImageSets", "Main", self.split + ".txt")) as fp:
      basenames = [line.strip() for line in fp.readlines()]
    return [os.path.join(self._dir, "JPEGImages", basename) + ".jpg" for basename in basenames]

  def _get_ground_truth_boxes(self, filepaths, allow_difficult):
    def only(node, tag):
      found = node.findall(tag)
      assert len(found) == 1, "expected exactly one <%s>" % tag
      return found[0]

    gt_boxes_by_filepath = {}
    for filepath in filepaths:
      basename = os.path.splitext(os.path.basename(filepath))[0]
      root = ET.parse(os.path.join(self._dir, "Annotations", basename) + ".xml").getroot()
      assert int(only(only(root, "size"), "depth").text) == 3
      boxes = []
      for obj in root.findall("object"):
        name, bndbox = only(obj, "name").text, only(obj, "bndbox")
        if int(only(obj, "difficult").text) != 0 and not allow_difficult:
          continue
        x_min, y_min, x_max, y_max = (int(only(bndbox, tag).text) - 1 for tag in ("xmin", "ymin", "xmax", "ymax"))      # 1-based -> 0-based pixels
        boxes.append(Box(class_index = self.class_name_to_index[name], class_name = name, corners = np.array([y_min, x_min, y_max, x_max]).astype(np.float32)))
      assert len(boxes) > 0
      gt_boxes_by_filepath[filepath] = boxes
    return gt_boxes_by_filepath

"""
TEST INFRASTRUCTURE ONLY -- seeded INPUT recipes shared by oracle/make_golden.py (which feeds
them to the unmodified reference) and by tests/ (which feed them to the oracle restatement and to
the CUDA path).  Pure NumPy (bit-reproducible across machines for a fixed NumPy version's
PCG64 stream; the fixtures store digests of reference OUTPUTS, the inputs are regenerated).
"""
import os

import numpy as np

GEOMETRY_CASES = {"tiny": (160, 240), "a600x800": (600, 800), "b600x1000": (600, 1000)}


def gt_boxes_for(h, w):
  """Two VOC-shaped GT boxes scaled from the 600x1000 layout of SURVEY.md section 8d."""
  sy, sx = h / 600.0, w / 1000.0
  return [((100 * sy, 150 * sx, 400 * sy, 600 * sx), 7), ((50 * sy, 650 * sx, 500 * sy, 850 * sx), 15)]


# ---- NMS ------------------------------------------------------------------------------------
NMS_CASES = ("f32_3000_ties_07", "f64_300_03", "f32_12000_07", "f32_degenerate")


def random_boxes(rng, n, h = 600.0, w = 1000.0, dtype = np.float32, min_side = 16.0, max_side = 300.0):
  y1 = rng.uniform(0, h, n)
  x1 = rng.uniform(0, w, n)
  hh = rng.uniform(min_side, max_side, n)
  ww = rng.uniform(min_side, max_side, n)
  b = np.stack([y1, x1, np.minimum(y1 + hh, h), np.minimum(x1 + ww, w)], axis = 1)
  return b.astype(dtype)


def nms_case(tag):
  rng = np.random.default_rng(abs(hash_tag(tag)))
  if tag == "f32_3000_ties_07":
    b = random_boxes(rng, 3000)
    s = (rng.integers(0, 200, 3000) / 200.0).astype(np.float32)          # many exact ties
    return b, s, 0.7
  if tag == "f64_300_03":
    b = random_boxes(rng, 300, dtype = np.float64)
    s = rng.uniform(0, 1, 300).astype(np.float32)
    return b, s, 0.3
  if tag == "f32_12000_07":
    # clustered boxes (jittered copies of 400 seeds): RPN-like heavy overlap
    seeds = random_boxes(rng, 400)
    idx = rng.integers(0, 400, 12000)
    b = seeds[idx] + rng.normal(0, 6.0, (12000, 4)).astype(np.float32)
    b[:, 2:4] = np.maximum(b[:, 2:4], b[:, 0:2] + 1.0)
    s = rng.uniform(0, 1, 12000).astype(np.float32)
    return b.astype(np.float32), s, 0.7
  if tag == "f32_degenerate":
    # zero-area boxes (0/0 -> NaN never suppresses), identical boxes, exact-threshold pairs
    b = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [5, 5, 5, 5], [5, 5, 5, 5], [0, 0, 10, 7], [0, 0, 7, 10],
                  [100, 100, 100, 150], [100, 100, 100, 150], [0, 3, 10, 10], [20, 20, 40, 40]], dtype = np.float32)
    s = np.array([0.9, 0.9, 0.8, 0.8, 0.7, 0.6, 0.5, 0.5, 0.4, 0.3], dtype = np.float32)
    return b, s, 0.7
  raise KeyError(tag)


def hash_tag(tag):
  h = 0
  for ch in tag:
    h = (h * 131 + ord(ch)) % 2147483647
  return h


# ---- RoIPool --------------------------------------------------------------------------------
ROI_CASES = ("small", "vgg_b128")


def roi_case(tag):
  rng = np.random.default_rng(hash_tag(tag))
  if tag == "small":
    C, H, W = 8, 12, 17
    fm = rng.normal(0, 1, (1, C, H, W)).astype(np.float32)
    n = 64
    y1 = rng.uniform(-20, H * 16, n); x1 = rng.uniform(-20, W * 16, n)
    y2 = y1 + rng.uniform(0, 150, n); x2 = x1 + rng.uniform(0, 200, n)
    rois = np.stack([np.zeros(n), x1, y1, x2, y2], axis = 1).astype(np.float32)
    # exact .5 rounding cases (x*1/16 = k + 0.5  <=> x = 16k + 8) and inverted / outside boxes
    rois[0] = [0, 8, 24, 40, 56]
    rois[1] = [0, 24, 8, 24, 8]
    rois[2] = [0, 300, 250, 400, 300]       # entirely outside -> empty bins
    rois[3] = [0, 100, 100, 50, 50]         # inverted -> width/height forced to 1
    rois[4] = [0, -8, -24, 8, 24]           # negative .5 rounding (half away from zero)
    return fm, rois
  if tag == "vgg_b128":
    C, H, W = 512, 37, 62
    fm = np.maximum(rng.normal(0, 1, (1, C, H, W)), 0).astype(np.float32)   # post-ReLU: many exact zeros/ties
    b = random_boxes(rng, 128, 600.0, 1000.0)
    rois = np.stack([np.zeros(128, dtype = np.float32), b[:, 1], b[:, 0], b[:, 3], b[:, 2]], axis = 1).astype(np.float32)
    return fm, rois
  raise KeyError(tag)


def roi_grad(tag, shape):
  rng = np.random.default_rng(hash_tag(tag) + 1)
  return rng.normal(0, 1, shape).astype(np.float32)


# ---- RPN proposal stage ---------------------------------------------------------------------
RPN_CASES = ("b_train", "a_infer", "tiny_train")


def rpn_case(tag):
  rng = np.random.default_rng(hash_tag(tag))
  if tag == "b_train":
    h, w, pre, post = 600, 1000, 12000, 2000
  elif tag == "a_infer":
    h, w, pre, post = 600, 800, 6000, 300
  else:
    h, w, pre, post = 160, 240, 12000, 2000
  fh, fw = h // 16, w // 16
  n = fh * fw * 9
  # distinct scores in (0.01, 0.99): a shuffled arithmetic progression (no ties)
  scores = (0.01 + 0.98 * (rng.permutation(n).astype(np.float64) + 0.5) / n).astype(np.float32)
  assert len(np.unique(scores)) == n
  deltas = rng.normal(0, 0.3, (n, 4)).astype(np.float32)
  return dict(image_shape = (3, h, w), fm_hw = (fh, fw), pre_nms = pre, post_nms = post,
              score_map = scores.reshape(1, fh, fw, 9), delta_map = deltas.reshape(1, fh, fw, 36))


# ---- end-to-end -----------------------------------------------------------------------------
E2E_CASES = {
  "small": dict(hw = (384, 512), weight_seed = 0, sample_seed = 0, heads = "spread", score_threshold = 0.05, backbone = "vgg16"),
  "resnet50_small": dict(hw = (384, 512), weight_seed = 1, sample_seed = 0, heads = "spread", score_threshold = 0.05, backbone = "resnet50"),
}


# ---- evaluation statistics (statistics.py:65-214) ---------------------------------------------
STATS_CASES = {"voc_like": (11, 40), "sparse": (12, 6), "crowded": (13, 25)}     # tag -> (seed, images)


def stats_case(tag):
  """
  Seeded detections: per image a list of ground-truth boxes [(corners f32 (4,), class)] and a dict class -> (k,5) f32 array of
  (y1,x1,y2,x2,score) rows.  Predictions are jittered copies of ground truths (so IoUs straddle 0.5, several predictions compete
  for one object and one prediction can overlap two objects) plus random false positives.
  """
  seed, n_images = STATS_CASES[tag]
  rng = np.random.RandomState(seed)
  images = []
  for _ in range(n_images):
    n_gt = int(rng.randint(0, 7)) if tag != "crowded" else int(rng.randint(6, 14))
    gts = []
    for _ in range(n_gt):
      cy, cx = rng.uniform(80, 520), rng.uniform(80, 920)
      hh, ww = rng.uniform(20, 160), rng.uniform(20, 200)
      if tag == "crowded" and len(gts) > 0 and rng.rand() < 0.6:          # overlapping same-class neighbours
        base = gts[rng.randint(len(gts))]
        corners = base[0] + rng.uniform(-25, 25, size = 4).astype(np.float32)
        gts.append((corners.astype(np.float32), base[1]))
        continue
      gts.append((np.array([cy - hh, cx - ww, cy + hh, cx + ww], dtype = np.float32), int(rng.randint(1, 6 if tag != "voc_like" else 21))))
    preds = {}
    for corners, cls in gts:
      for _ in range(int(rng.randint(0, 4))):
        jitter = rng.normal(0, 0.18, size = 4) * np.array([corners[2] - corners[0], corners[3] - corners[1]] * 2)
        row = np.concatenate([corners + jitter, [rng.uniform(0.05, 1.0)]]).astype(np.float32)
        preds.setdefault(cls, []).append(row)
    for _ in range(int(rng.randint(0, 5))):
      cls = int(rng.randint(1, 6 if tag != "voc_like" else 21))
      cy, cx = rng.uniform(80, 520), rng.uniform(80, 920)
      hh, ww = rng.uniform(20, 160), rng.uniform(20, 200)
      preds.setdefault(cls, []).append(np.array([cy - hh, cx - ww, cy + hh, cx + ww, rng.uniform(0.05, 0.9)], dtype = np.float32))
    images.append((gts, {c: np.stack(r, axis = 0) for c, r in preds.items()}))
  return images


# ---- synthetic PASCAL VOC tree (datasets/voc.py) -----------------------------------------------------
VOC_CLASSES = ("aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow", "diningtable", "dog", "horse", "motorbike", "person",
               "pottedplant", "sheep", "sofa", "train", "tvmonitor")
VOC_IMAGES = (("000005", 500, 375), ("000007", 333, 500), ("000009", 480, 360), ("000012", 500, 333), ("000016", 400, 300))


def make_voc_tree(root, split = "trainval", seed = 0):
  """Writes a 5-image VOC2007-shaped directory under `root` (seeded content; JPEG bytes depend only on the PIL build, which is the
  same in the build container and on the GPU box): ImageSets/Main/<split>.txt + <class>_<split>.txt, JPEGImages, Annotations with
  1-3 objects per image (one of them `difficult`).  Returns the dataset directory."""
  from PIL import Image
  rng = np.random.default_rng(seed)
  d = os.path.join(root, "VOCdevkit", "VOC2007")
  for sub in ("ImageSets/Main", "JPEGImages", "Annotations"):
    os.makedirs(os.path.join(d, sub), exist_ok = True)
  with open(os.path.join(d, "ImageSets", "Main", split + ".txt"), "w") as fp:
    fp.write("".join(name + "\n" for name, _, _ in VOC_IMAGES))
  for cls in VOC_CLASSES:
    with open(os.path.join(d, "ImageSets", "Main", "%s_%s.txt" % (cls, split)), "w") as fp:
      fp.write("".join("%s -1\n" % name for name, _, _ in VOC_IMAGES))
  for i, (name, w, h) in enumerate(VOC_IMAGES):
    # smooth random picture (low-frequency noise upsampled) so that JPEG + bilinear resize exercise real interpolation
    small = rng.integers(0, 256, (h // 8 + 1, w // 8 + 1, 3), dtype = np.uint8)
    Image.fromarray(small, mode = "RGB").resize((w, h), resample = Image.BICUBIC).save(os.path.join(d, "JPEGImages", name + ".jpg"), quality = 92)
    objs = []
    for j in range(1 + i % 3):
      x1 = int(rng.integers(1, w // 2)); y1 = int(rng.integers(1, h // 2))
      x2 = int(rng.integers(x1 + 40, w)); y2 = int(rng.integers(y1 + 40, h))
      cls = VOC_CLASSES[int(rng.integers(0, 20))]
      objs.append((cls, x1, y1, x2, y2, 0))
    if i == 1:
      objs.append(("person", 10, 10, 60, 80, 1))               # a `difficult` object: skipped unless allow_difficult
    xml = ["<annotation><filename>%s.jpg</filename><size><width>%d</width><height>%d</height><depth>3</depth></size>" % (name, w, h)]
    for cls, x1, y1, x2, y2, diff in objs:
      xml.append("<object><name>%s</name><difficult>%d</difficult><bndbox><xmin>%d</xmin><ymin>%d</ymin><xmax>%d</xmax><ymax>%d</ymax></bndbox></object>" % (cls, diff, x1, y1, x2, y2))
    xml.append("</annotation>")
    with open(os.path.join(d, "Annotations", name + ".xml"), "w") as fp:
      fp.write("".join(xml))
  return d

"""Device-side segmentation metrics: mirror of utils/util.py:57-128 (CityscapesMetricTracker).

The reference copies the student's and the teacher's full logit tensors to the host EVERY iteration
(trainer/layerwise_trainer.py:247-250: 2 x 80 MB per 1024^2 image) and bincounts them in numpy -- the largest host
sync of its loop.  Here `update` is one kernel pass (kdcc_confusion_update) into an int64 matrix that
stays on the GPU; `conf` / `get_iou()` copy 19 x 19 numbers when asked.
"""
import numpy as np
import torch

from . import functional as F_kdcc


class ConfusionMatrix:
    """conf[target][pred] counts on the device.  Labels equal to `ignore_index` or outside [0, C) are dropped
    (the reference rewrites them to C in place and masks them out: utils/util.py:109,122; kdcc leaves `labels`
    untouched)."""

    def __init__(self, num_classes=19, ignore_index=255, device="cuda"):
        self.nc, self.ignore = num_classes, ignore_index
        self.mat = torch.zeros(num_classes * num_classes, dtype=torch.long, device=device)

    def reset(self):
        self.mat.zero_()

    def update(self, logits, target):
        F_kdcc.confusion_update(self.mat, logits, target, self.ignore)

    def numpy(self):
        return self.mat.view(self.nc, self.nc).cpu().numpy()

    def iou(self):
        """mean IoU over the classes that occur (union > 0); 0 for an empty matrix"""
        m = self.mat.view(self.nc, self.nc).double()
        inter = m.diag()
        union = m.sum(0) + m.sum(1) - inter
        valid = union > 0
        return float((inter[valid] / union[valid]).mean()) if valid.any() else 0.0


class CityscapesMetricTracker:
    """Same constructor, attributes and methods as the reference class (utils/util.py:57-128)."""
    class_names = ["road", "sidewalk", "building", "wall", "fence", "pole", "traffic_light", "traffic_sight", "vegetation",
                   "terrain", "sky", "person", "rider", "car", "truck", "bus", "train", "motorcycle", "bicycle"]
    num_classes = len(class_names)

    def __init__(self, writer=None, ignore_index=255, device="cuda"):
        self.writer = writer
        self.ignore_index = ignore_index
        self._cm = ConfusionMatrix(self.num_classes, ignore_index, device=device)

    def reset(self):
        self._cm.reset()

    @property
    def conf(self):
        """float64 (C, C) numpy matrix like the reference attribute (device -> host copy of C*C counts)"""
        return self._cm.numpy().astype(np.float64)

    def update(self, outputs, labels):
        self._cm.update(outputs, labels)

    def get_iou(self):
        conf = self.conf
        if not np.any(conf):
            return 1.
        tp = np.diag(conf)
        with np.errstate(divide="ignore", invalid="ignore"):
            iou_pc = tp / (np.sum(conf, 0) + np.sum(conf, 1) - tp)
        return np.nanmean(iou_pc)

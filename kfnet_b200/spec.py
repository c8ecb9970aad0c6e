"""KFNetDataSpec -- configuration record of the reference (KFNet/KFNet.py:6-50), same fields and defaults."""
from __future__ import annotations


class KFNetDataSpec:
    _SCENES = {  # scene -> (image_num, sequence_length, num_sequence)   KFNet.py:26-47
        "chess": (4000, 1000, 4), "fire": (2000, 1000, 2), "heads": (1000, 1000, 1),
        "office": (6000, None, None), "pumpkin": (4000, None, None), "redkitchen": (7000, None, None),
        "stairs": (2000, 500, 4),
    }

    def __init__(self, batch_size=4, image_size=(480, 640), channels=3, crop_size=(480, 640),
                 focal_x=525., focal_y=525., u=320., v=240., scene="stairs"):
        self.batch_size = batch_size
        self.image_size = tuple(image_size)
        self.channels = channels
        self.crop_size = tuple(crop_size)
        self.focal_x, self.focal_y, self.u, self.v = focal_x, focal_y, u, v
        self.scene = scene
        if scene in self._SCENES:
            num, seq, nseq = self._SCENES[scene]
            self.image_num = num
            if seq is not None:
                self.sequence_length, self.num_sequence = seq, nseq
            self.stepvalue = self.image_num * 75 / self.batch_size
            self.max_steps = self.stepvalue * 5

"""Drop-in for the reference's ``augmentation.py``: same class, constructor and method signatures
(augmentation.py:14,121,144,206), same ``random`` draws in the same order, but the per-point
arithmetic runs in the fused pre-pass kernel instead of Python loops / OpenCV calls.  Like the
reference, each method modifies ``sign`` ([T,K,2] tensor) in place and returns it."""
import logging
import math
import random

import numpy as np

from . import preprocess as PP


class augmentation():
    def __init__(self, body_type_identifiers, body_section_dict, device="cuda"):
        self.body_section_dict = body_section_dict
        self.BODY_IDENTIFIERS = (body_type_identifiers['pose'] + body_type_identifiers['left_hand']
                                 + body_type_identifiers['rigth_hand'])
        self.HAND_IDENTIFIERS = body_type_identifiers['left_hand'] + body_type_identifiers['rigth_hand']
        left = ['pose_chest_middle_up', 'pose_left_shoulder', 'pose_left_elbow', 'pose_left_wrist']
        right = ['pose_chest_middle_up', 'pose_right_shoulder', 'pose_right_elbow', 'pose_right_wrist']
        self.ARM_IDENTIFIERS_ORDER = [[body_section_dict[i] for i in left], [body_section_dict[i] for i in right]]
        self.device = device
        self._pp = {}

    def _prepass(self, K_points):
        if K_points not in self._pp:
            self._pp[K_points] = PP.Prepass(K_points, self.device, self.BODY_IDENTIFIERS, self.HAND_IDENTIFIERS,
                                            arm_chains=self.ARM_IDENTIFIERS_ORDER)
        return self._pp[K_points]

    def _apply(self, sign, aug):
        out = self._prepass(sign.shape[1])(sign.unsqueeze(0), augs=[aug], want_inputs=False)["y"][0]
        sign.copy_(out.to(sign.device))
        return sign

    def __random_pass(self, prob):
        return random.random() < prob

    def augment_rotate(self, sign, angle_range):
        """augmentation.py:121-142: one angle per sequence about (0.5, 0.5); hands rotate twice."""
        angle = math.radians(random.uniform(*angle_range))
        return self._apply(sign, PP.aug_rotate(angle))

    def augment_shear(self, sign, type, squeeze_ratio):
        """augmentation.py:144-203."""
        src = np.array(((0, 1), (1, 1), (0, 0), (1, 0)), dtype=np.float32)
        if type == "squeeze":
            move_left = random.uniform(*squeeze_ratio)
            move_right = random.uniform(*squeeze_ratio)
            dest = np.array(((0 + move_left, 1), (1 - move_right, 1), (0 + move_left, 0), (1 - move_right, 0)),
                            dtype=np.float32)
        elif type == "perspective":
            move_ratio = random.uniform(*squeeze_ratio)
            if self.__random_pass(0.5):
                dest = np.array(((0 + move_ratio, 1 - move_ratio), (1, 1), (0 + move_ratio, 0 + move_ratio), (1, 0)),
                                dtype=np.float32)
            else:
                dest = np.array(((0, 1), (1 - move_ratio, 1 - move_ratio), (0, 0), (1 - move_ratio, 0 + move_ratio)),
                                dtype=np.float32)
        else:
            logging.error("Unsupported shear type provided.")
            return {}
        return self._apply(sign, PP.aug_shear(PP.perspective_matrix(src, dest)))

    def augment_arm_joint_rotate(self, sign, probability, angle_range):
        """augmentation.py:206-233."""
        angles = []
        for arm_side_ids in self.ARM_IDENTIFIERS_ORDER:
            row = []
            for _ in arm_side_ids:
                if self.__random_pass(probability):
                    row.append(math.radians(random.uniform(*angle_range)))
                else:
                    row.append(None)
            angles.append(row)
        return self._apply(sign, PP.aug_arm(angles))

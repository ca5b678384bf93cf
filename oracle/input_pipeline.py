"""ORACLE (test infrastructure, not product code): CPU restatement of the tensor half of the reference's sample
construction, data/face_dataset.py:45-90 (`get_train_item`), in numpy exactly as the reference writes it.  Used by tests/ only."""
import numpy as np
import torch


def train_item(img_S, img_F, mask_S, mask_F, lm_S, lm_F, flipped, load_size=128):
    """One sample.  img_* (H,W,3) uint8, mask_* (H,W,1) uint8, lm_* (L,2) int; `flipped` = the reference's
    `index >= len(self.pairs)` (:66).  Returns the dict of torch tensors the reference's Dataset yields."""
    lm_S, lm_F = lm_S.copy(), lm_F.copy()
    if flipped:                                                          # :66-71
        lm_S = np.hstack((127 - lm_S[:, 0:1], lm_S[:, 1:2]))
        lm_F = np.hstack((127 - lm_F[:, 0:1], lm_F[:, 1:2]))
        img_S, img_F = img_S[:, ::-1, :], img_F[:, ::-1, :]
        mask_S, mask_F = mask_S[:, ::-1, :], mask_F[:, ::-1, :]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).transpose((2, 0, 1)).astype('float32')).div(255)   # :77-80
    lm = lambda a: torch.clamp(torch.from_numpy(a).long(), 0, load_size - 1)                                  # :82-85
    return {'img_S': t(img_S), 'img_F': t(img_F), 'mask_S': t(mask_S), 'mask_F': t(mask_F), 'lm_S': lm(lm_S), 'lm_F': lm(lm_F)}

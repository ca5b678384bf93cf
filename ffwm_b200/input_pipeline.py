"""Device-side half of the reference's input pipeline (data/face_dataset.py:45-90).

The reference builds every training sample on the host — optional left-right flip of images, masks and landmarks
(:66-71), HWC -> CHW transpose, astype(float32), div(255) (:77-80), landmark clamp (:82-85) — inside DataLoader workers,
and `set_train_input` then copies fp32 tensors to the GPU (models/ffwm_model.py:61-70).  `BatchIngest` moves that
arithmetic to the GPU: the decoded batch is handed over as uint8 HWC arrays (what `cv2.imread` produces), crosses PCIe
once from pinned memory (4x fewer bytes than fp32), and one kernel per tensor (csrc/ingest.cu) produces exactly the
tensors the reference's loader would have produced, bit for bit.  Decoding files, the random rotation augmentation
(cv2.warpAffine, :112-126) and the landmark dictionaries stay host-side data handling and are out of scope.
"""
import torch

from . import ops


class BatchIngest:
    def __init__(self, device, batch, size=128, load_size=128):
        self.device, self.load_size = torch.device(device), load_size
        self._pin = {k: torch.empty((batch, size, size, c), dtype=torch.uint8).pin_memory()
                     for k, c in (("img_S", 3), ("img_F", 3), ("mask_S", 1), ("mask_F", 1))}
        self._pin_flip = torch.empty(batch, dtype=torch.uint8).pin_memory()

    def __call__(self, batch_u8, flip=None):
        """batch_u8: dict of uint8 arrays / tensors `img_S`, `img_F` (B,H,W,3), `mask_S`, `mask_F` (B,H,W,1) and integer
        landmarks `lm_S` / `lm_F` (B,L,2) as read from disk (unflipped); flip: B booleans (the reference's index >= len(pairs)
        samples).  Returns the dict `set_train_input` expects, on the device."""
        dev = self.device
        b = len(batch_u8["img_S"])
        fl = None
        if flip is not None:
            self._pin_flip[:b].copy_(torch.as_tensor(flip, dtype=torch.uint8))
            fl = self._pin_flip[:b].to(dev, non_blocking=True)
        out = {}
        for k, pin in self._pin.items():
            if k not in batch_u8:
                continue
            src = torch.as_tensor(batch_u8[k], dtype=torch.uint8)
            pin[:b].copy_(src.reshape(pin[:b].shape))
            d = pin[:b].to(dev, non_blocking=True)
            out[k] = torch.empty((b, d.size(3), d.size(1), d.size(2)), dtype=torch.float32, device=dev)
            ops.ingest_u8(d, fl, out[k])
        for k in ("lm_S", "lm_F"):
            if k in batch_u8:
                lm = torch.as_tensor(batch_u8[k]).to(dev, non_blocking=True).long()
                if fl is not None:            # np.hstack((127 - lm[:, 0:1], lm[:, 1:2])) (:66-67)
                    x = torch.where(fl.bool().view(-1, 1), 127 - lm[..., 0], lm[..., 0])
                    lm = torch.stack((x, lm[..., 1]), -1)
                out[k] = lm.clamp(0, self.load_size - 1)
        for k, v in batch_u8.items():
            if k not in out:
                out[k] = v
        return out

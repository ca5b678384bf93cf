"""the device-side input pipeline (csrc/ingest.cu, ffwm_b200/input_pipeline.py) against the CPU restatement
of data/face_dataset.py:45-90 (oracle/input_pipeline.py) — bit-exact, including flipped samples and ragged sizes."""
import numpy as np
import pytest
import torch

DEV = "cuda:0"


def _samples(b, size, seed):
    r = np.random.RandomState(seed)
    return {"img_S": r.randint(0, 256, (b, size, size, 3)).astype(np.uint8), "img_F": r.randint(0, 256, (b, size, size, 3)).astype(np.uint8),
            "mask_S": (r.rand(b, size, size, 1) > 0.4).astype(np.uint8) * 255, "mask_F": (r.rand(b, size, size, 1) > 0.4).astype(np.uint8) * 255,
            "lm_S": r.randint(-5, 140, (b, 50, 2)), "lm_F": r.randint(-5, 140, (b, 50, 2))}


def test_oracle_matches_reference_semantics():
    """CPU: flip is `[:, ::-1, :]` on images and `127 - x` on landmark columns; values are x / 255 in float32."""
    from oracle import input_pipeline as O
    s = _samples(1, 8, 0)
    one = {k: v[0] for k, v in s.items()}
    a, f = O.train_item(flipped=False, **one), O.train_item(flipped=True, **one)
    assert torch.equal(f["img_S"], a["img_S"].flip(-1)) and torch.equal(f["mask_F"], a["mask_F"].flip(-1))
    assert a["img_S"].dtype == torch.float32 and float(a["img_S"].max()) <= 1.0
    assert torch.equal(a["img_S"], torch.from_numpy(s["img_S"][0]).permute(2, 0, 1).float() / 255)
    assert int(a["lm_S"].min()) >= 0 and int(a["lm_S"].max()) <= 127
    assert torch.equal(f["lm_S"][:, 1], a["lm_S"][:, 1])


@pytest.mark.gpu
@pytest.mark.parametrize("b,size", [(8, 128), (3, 37), (1, 1)])
def test_batch_ingest_is_bit_exact(b, size):
    from oracle import input_pipeline as O
    from ffwm_b200.input_pipeline import BatchIngest
    s = _samples(b, size, b + size)
    flips = [(i % 3) == 1 for i in range(b)]
    got = BatchIngest(DEV, b, size=size)(s, flips)
    for i in range(b):
        want = O.train_item(flipped=flips[i], **{k: v[i] for k, v in s.items()})
        for k, w in want.items():
            assert torch.equal(got[k][i].cpu(), w), (i, k)
    plain = BatchIngest(DEV, b, size=size)(s)                      # no flip flags at all
    assert torch.equal(plain["img_F"].cpu(), torch.from_numpy(s["img_F"]).permute(0, 3, 1, 2).float().div(255))


@pytest.mark.gpu
def test_ingest_rejects_bad_arguments():
    from ffwm_b200 import ops
    with pytest.raises(ValueError):
        ops.ingest_u8(torch.zeros(1, 4, 4, 3, device=DEV), None, torch.zeros(1, 3, 4, 4, device=DEV))      # not uint8
    with pytest.raises(RuntimeError):
        ops.ingest_u8(torch.zeros(1, 4, 4, 2, dtype=torch.uint8, device=DEV), None, torch.zeros(1, 2, 4, 4, device=DEV))   # C = 2

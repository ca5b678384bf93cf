"""Shared definition of the golden-vector cases (inputs are rebuilt from
numpy seeds so only outputs need to be stored)."""
import numpy as np


def _rs(seed):
    return np.random.RandomState(seed)


def resample2d_case(name):
    """-> dict(in1, in2, gout, ks, dil) as float64 numpy arrays."""
    spec = {
        # cfg1 of SURVEY 8(d): 1x3x128x128, flow = randn*3 px, sigma const
        "cfg1_ks2": dict(shape=(1, 3, 128, 128), ks=2, dil=1, sigma=5.0, fscale=3.0, seed=0),
        "cfg1_ks4": dict(shape=(1, 3, 128, 128), ks=4, dil=1, sigma=2.0, fscale=3.0, seed=1),
        # small: per-pixel sigma, dilation 2, ragged sizes, input1 larger than the flow grid
        "small_ks2_d2": dict(shape=(2, 5, 13, 11), ks=2, dil=2, sigma=None, fscale=2.0, seed=2),
        "small_ks4_d1": dict(shape=(2, 5, 13, 11), ks=4, dil=1, sigma=None, fscale=2.0, seed=3),
        "small_ks6_d1": dict(shape=(1, 2, 9, 14), ks=6, dil=1, sigma=None, fscale=1.5, seed=4),
        "small_ks3_d1": dict(shape=(1, 4, 8, 8), ks=3, dil=1, sigma=None, fscale=1.0, seed=5),
    }[name]
    b, c, h, w = spec["shape"]
    r = _rs(spec["seed"])
    in1 = r.rand(b, c, h, w) * 2 - 1
    flow = r.standard_normal((b, 2, h, w)) * spec["fscale"]
    if spec["sigma"] is None:
        sigma = r.rand(b, 1, h, w) * 2.5 + 0.5
    else:
        sigma = np.full((b, 1, h, w), spec["sigma"])
    gout = r.standard_normal((b, c, h, w))
    return dict(in1=in1, in2=np.concatenate([flow, sigma], 1), gout=gout, ks=spec["ks"], dil=spec["dil"])


RESAMPLE2D_CASES = ["cfg1_ks2", "cfg1_ks4", "small_ks2_d2", "small_ks4_d1", "small_ks6_d1", "small_ks3_d1"]


def block_extractor_case(name):
    spec = {
        # the reference's own gradcheck recipe (test_block_extractor.py:77-81)
        "ref_gradcheck": dict(src=(4, 6, 14, 10), flow=(4, 2, 14, 10), k=3, mode="rand1.8", seed=10),
        # live use in losses.py:212-217: constant integer flow k//2 on a 1-channel grid
        "unfold_k5": dict(src=(2, 1, 20, 18), flow=(2, 2, 16, 14), k=5, mode="const", seed=11),
        # wild flow: negatives and far out of range -> clamping on all sides
        "wild_k3": dict(src=(2, 3, 9, 12), flow=(2, 2, 7, 5), k=3, mode="randn4", seed=12),
        "even_k2": dict(src=(1, 2, 8, 8), flow=(1, 2, 8, 8), k=2, mode="randn4", seed=13),
    }[name]
    r = _rs(spec["seed"])
    src = r.rand(*spec["src"])
    k = spec["k"]
    if spec["mode"] == "rand1.8":
        flow = r.rand(*spec["flow"]) * 1.8
    elif spec["mode"] == "const":
        flow = np.full(spec["flow"], float(k // 2))
    else:
        flow = r.standard_normal(spec["flow"]) * 4
    b, _, hf, wf = spec["flow"]
    gout = r.standard_normal((b, spec["src"][1], k * hf, k * wf))
    return dict(src=src, flow=flow, gout=gout, k=k)


BLOCK_EXTRACTOR_CASES = ["ref_gradcheck", "unfold_k5", "wild_k3", "even_k2"]


def local_attn_reshape_case(name):
    spec = {
        # the reference's known-answer input (test_local_attn_reshape.py:29-43)
        "kat_0_8": dict(shape=(2, 9, 10, 10), k=3, seed=None),
        "ref_gradcheck": dict(shape=(4, 9, 14, 10), k=3, seed=20),
        "k5": dict(shape=(2, 25, 6, 7), k=5, seed=21),
        "k7": dict(shape=(1, 49, 5, 3), k=7, seed=22),
    }[name]
    b, c, h, w = spec["shape"]
    k = spec["k"]
    if spec["seed"] is None:
        x = np.tile(np.arange(c, dtype=np.float64).reshape(1, c, 1, 1), (b, 1, h, w))
        r = _rs(0)
    else:
        r = _rs(spec["seed"])
        x = r.rand(b, c, h, w)
    gout = r.standard_normal((b, 1, k * h, k * w))
    return dict(x=x, gout=gout, k=k)


LOCAL_ATTN_RESHAPE_CASES = ["kat_0_8", "ref_gradcheck", "k5", "k7"]

# every 7th element of a flattened output is stored for the 128x128 cases
SUBSAMPLE = 7

"""Harness for running the reference's UNMODIFIED orchestrators (test / bench infrastructure).

`baseline/_ref` holds a byte-identical copy of the reference's `models/` and `lightcnn/` packages
(baseline/stage_ref.py).  This module supplies only what SURVEY 8(c) lists as harness-side shims,
never an edit of the reference:

  * `numpy.int = int`                     (models/base_networks.py:366 uses the removed alias)
  * an offline torchvision VGG19 file     (models/losses.py:401 calls vgg19(pretrained=True))
  * state_dict files at `opt.lightcnn / opt.flownetf / opt.flownetb`   (models/ffwm_model.py:30-35)
  * `opt` built as a namespace            (fields read at models/base_model.py:32-38)

Two ways to run the same reference class:

  reference_model(device="cpu")                 the reference's own CPU path: its networks and losses
                                                on PyTorch CPU kernels, `F.grid_sample` warps
  reference_model(device="cuda", product=True)  `ffwm_b200.compat.install()` first, so the reference's
                                                `FFWMModel` / `FlowNetModel` bind to this repo's modules
                                                and sm_100a kernels — the orchestrator code that runs is
                                                still the reference's
"""
import importlib
import os
import sys
import tempfile
import types
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "models", "ffwm_model.py"))


def fill_state(module, dtype=None):
    """Deterministic parameters from the state_dict key alone (same rule as tests/golden/model_cases.py)."""
    import torch
    sd, new = module.state_dict(), {}
    for key, t in sd.items():
        if not t.is_floating_point():
            new[key] = t.clone()
            continue
        g = torch.Generator().manual_seed(zlib.crc32(key.encode()) & 0x7fffffff)
        shape, leaf = tuple(t.shape), key.rsplit('.', 1)[-1]
        if leaf == 'running_var':
            v = torch.rand(shape, generator=g, dtype=torch.float64) + 0.5
        elif leaf == 'running_mean':
            v = torch.randn(shape, generator=g, dtype=torch.float64) * 0.1
        elif leaf in ('weight_u', 'weight_v'):
            v = torch.randn(shape, generator=g, dtype=torch.float64)
            v = v / v.norm()
        elif t.dim() >= 2:
            v = torch.randn(shape, generator=g, dtype=torch.float64) * (1.0 / max(1, t.numel() // t.shape[0])) ** 0.5
        elif leaf == 'weight':
            v = torch.rand(shape, generator=g, dtype=torch.float64) + 0.5
        else:
            v = torch.randn(shape, generator=g, dtype=torch.float64) * 0.05
        new[key] = v.to(dtype or t.dtype)
    module.load_state_dict(new)
    return module


def _purge(prefixes=("models", "lightcnn")):
    for name in list(sys.modules):
        if name in prefixes or name.startswith(tuple(p + "." for p in prefixes)):
            del sys.modules[name]


def import_reference(product=False):
    """Import the staged reference packages (fresh).  product=True binds them to ffwm_b200 first."""
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int
    if not available():
        raise RuntimeError("baseline/_ref is not staged (run python baseline/stage_ref.py where the reference checkout exists)")
    _purge()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    models = importlib.import_module("models")
    if product:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        import ffwm_b200.compat
        ffwm_b200.compat.install()
    return models


def scratch_checkpoints(scratch, product=False):
    """Random-init (seeded by key) state_dicts where the reference expects pretrained files."""
    import torch
    os.environ["TORCH_HOME"] = scratch
    ck = os.path.join(scratch, "hub", "checkpoints")
    os.makedirs(ck, exist_ok=True)
    vgg_file = os.path.join(ck, "vgg19-dcbb9e9d.pth")
    if not os.path.exists(vgg_file):
        import torchvision
        torch.save(fill_state(torchvision.models.vgg19(weights=None)).state_dict(), vgg_file)
    base = importlib.import_module("models.base_networks")
    lc = importlib.import_module("lightcnn.light_cnn")
    paths = {}
    for name, ctor in (("lightcnn", lc.LightCNN_29Layers), ("flownetf", lambda: base.FlowNet(64)), ("flownetb", lambda: base.FlowNet(64))):
        paths[name] = os.path.join(scratch, name + ".pth")
        if not os.path.exists(paths[name]):
            torch.save(fill_state(ctor()).state_dict(), paths[name])
    return paths


def reference_ffwm_model(device="cpu", product=False, scratch=None, fill=True):
    """The reference's FFWMModel (models/ffwm_model.py:10-59), constructed exactly as train_ffwm.py does
    (`create_model(opt)` reduces to `FFWMModel(opt)`), with deterministic weights."""
    import torch
    import_reference(product=product)
    scratch = scratch or tempfile.mkdtemp(prefix="ffwm_ref_")
    paths = scratch_checkpoints(scratch, product=product)
    gpu_ids = [] if str(device) == "cpu" else [torch.device(device).index or 0]
    opt = types.SimpleNamespace(gpu_ids=gpu_ids, isTrain=True, checkpoints_dir=scratch, name="ref", preprocess="none",
                                crop=False, **paths)
    FFWMModel = importlib.import_module("models.ffwm_model").FFWMModel
    model = FFWMModel(opt)
    if product:
        # losses.PerceptualLoss of the product builds its VGG19 with random init (no torchvision download
        # hook): give it the same torchvision-format weights the reference arm loads from TORCH_HOME
        vgg = getattr(model.criterionPerceptual, "vgg", None)
        if vgg is not None and hasattr(vgg, "load_torchvision"):
            vgg.load_torchvision(torch.load(os.path.join(scratch, "hub", "checkpoints", "vgg19-dcbb9e9d.pth")))
            model.criterionPerceptual.to(model.device)
    if fill:
        fill_state(model.netG)
        fill_state(model.netD)
    return model


def synthetic_batch(b, seed, titers=30000):
    """SURVEY 8(d) cfg3 batch, as the reference's data loader would hand it to set_input."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return {
        'img_S': torch.rand(b, 3, 128, 128, generator=g), 'img_F': torch.rand(b, 3, 128, 128, generator=g),
        'mask_F': (torch.rand(b, 1, 128, 128, generator=g) > 0.3).float(),
        'mask_S': (torch.rand(b, 1, 128, 128, generator=g) > 0.3).float(),
        'lm_F': torch.randint(20, 108, (b, 1000, 2), generator=g), 'lm_S': torch.randint(20, 108, (b, 1000, 2), generator=g),
        'gate': (torch.rand(b, 1000, 1, generator=g) > 0.2).float(),
        'titers': titers, 'epoch': 0, 'input_path': [''] * b,
    }


def reference_flownet_model(device="cpu", product=False, scratch=None, fill=True):
    """The reference's FlowNetModel (models/flownet_model.py:8-78; train_flow.py sets `reverse`).

    product=False (CPU): the reference has no CPU implementation of block_extractor / local_attn_reshape
    (NotImplementedError for CPU tensors), so the `extractor` / `reshape` attributes of its
    AffineRegularizationLoss INSTANCES are served by the C oracle (pinned to the outputs of the reference's own
    CUDA kernels), and `criterionLD` gets torch 1.5's integer division — instance-level shims,
    nothing in the reference is edited.  product=True: no shim at all; the reference class runs on ffwm_b200."""
    import torch
    import_reference(product=product)
    scratch = scratch or tempfile.mkdtemp(prefix="ffwm_ref_")
    scratch_checkpoints(scratch, product=product)
    gpu_ids = [] if str(device) == "cpu" else [torch.device(device).index or 0]
    opt = types.SimpleNamespace(gpu_ids=gpu_ids, isTrain=True, checkpoints_dir=scratch, name="ref", preprocess="none")
    model = importlib.import_module("models.flownet_model").FlowNetModel(opt)
    model.reverse = False
    if fill:
        fill_state(model.flowNet)
    if product:
        vgg = getattr(model.Correctness, "vgg", None)
        if vgg is not None and hasattr(vgg, "load_torchvision"):
            vgg.load_torchvision(torch.load(os.path.join(scratch, "hub", "checkpoints", "vgg19-dcbb9e9d.pth")))
            model.Correctness.to(model.device)
        return model
    from oracle import train_cpu

    class _Extract(torch.nn.Module):
        def __init__(self, kz):
            super().__init__()
            self.kz = kz

        def forward(self, s, f):
            return train_cpu._BlockExtractorCPU.apply(s, f, self.kz)

    class _Reshape(torch.nn.Module):
        def forward(self, x, k):
            return train_cpu._LocalAttnReshapeCPU.apply(x, k)

    for reg in model.Regularization.method_dic.values():
        reg.extractor = _Extract(reg.kz)
        reg.reshape = _Reshape()
    ld = model.criterionLD
    inner = ld.criterionLD

    def ld_forward(flows, lm_S, lm_F, gate):
        total = 0
        for i, flow in enumerate(flows):
            scale = ld.img_size // flow.size(3)
            total += ld.weights[i] * inner(flow, torch.div(lm_S, scale, rounding_mode='floor'),
                                           torch.div(lm_F, scale, rounding_mode='floor'), gate)
        return total
    model.criterionLD = ld_forward
    return model

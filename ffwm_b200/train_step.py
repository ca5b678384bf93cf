"""The FFWM and FlowNet training steps as plain callables.

`FFWMTrainer.optimize_parameters()` performs exactly the work of the reference's
`FFWMModel.optimize_parameters` (models/ffwm_model.py:72-160): forward of flowNetF, flowNetB,
netG, guided filter and the 8 facial-part crops; discriminator step; generator + flow step with
the perceptual (VGG19), L1, illumination, identity (LightCNN-29), adversarial and facial-part
losses at the reference's weights; three Adam optimisers at the reference's hard-coded rates
(:46-49).  `FlowNetTrainer` is `FlowNetModel.optimize_parameters` (models/flownet_model.py:57-78),
the only live user of block_extractor / local_attn_reshape.

Under `torch.distributed` (one process per GPU) the step is pure data parallel: every rank holds a
replica, BatchNorm statistics stay per rank (8 images per rank = the reference batch), and the
gradients are averaged with NCCL all-reduce after each backward (`parallel.GradAverager`) before
the optimiser step — discriminator gradients after `backward_D`, generator + both FlowNets after
`backward_G`.  The reference has no multi-GPU path at all.
"""
import itertools
import os

import torch
import torch.nn.functional as F

from . import base_networks, external_function, losses
from .light_cnn import LightCNN_29Layers
from .norm import DeferredCounters
from .parallel import GradAverager


# The perceptual losses of one step through 6 VGG19 passes instead of 14 and the identity losses through 2 LightCNN
# passes instead of 4 (losses.PerceptualLoss.many / IdentityLoss.many: same samples, bigger batches): verified against
# the reference goldens on the CPU and on a B200, default on since round 2 (profiles/r02a_switches.txt).
BATCHED_VGG = os.environ.get("FFWM_BATCHED_VGG", "1") == "1"
FLOW_STREAMS = os.environ.get("FFWM_FLOW_STREAMS", "1") == "1"      # see FFWMTrainer._flownets


def set_requires_grad(nets, flag):
    for net in nets if isinstance(nets, (list, tuple)) else [nets]:
        if net is not None:
            for p in net.parameters():
                p.requires_grad = flag


# torch.optim.Adam's single-kernel ("fused") implementation on CUDA: the same update rule as the reference's default
# (foreach) Adam in one multi-tensor kernel instead of ~12 per parameter chunk — 2.3 ms of foreach kernels per step in
# profiles/r02s_launches_train_summary.txt.  FFWM_FUSED_ADAM=0 restores the default for the A/B.
FUSED_ADAM = os.environ.get("FFWM_FUSED_ADAM", "1") == "1"


def _adam_impl(device):
    return dict(fused=True) if (FUSED_ADAM and torch.device(device).type == "cuda") else {}


class FFWMTrainer:
    loss_names = ['loss_G', 'loss_D', 'loss_l1', 'loss_iden', 'loss_illu', 'loss_adv', 'loss_prc', 'loss_fc']

    def __init__(self, device, crop=False, lightcnn_state=None, flownetf_state=None, flownetb_state=None,
                 vgg_weights=None, distributed=None, graph=False):
        self.device = torch.device(device)
        self._capturable = bool(graph)
        dev = self.device
        with torch.device(dev):        # parameters are created and initialised on the device (122 M of them)
            self.flowNetF = base_networks.FlowNet(64)
            self.flowNetB = base_networks.FlowNet(64)
            self.warpNet = base_networks.WarpNet().eval()
            self.lightCNN = LightCNN_29Layers().eval()
            self.netG = base_networks.FFWM(sn=True)
            self.netD = base_networks.MSDiscriminator(128, sigmoid=False)
        for net, state in ((self.lightCNN, lightcnn_state), (self.flowNetF, flownetf_state), (self.flowNetB, flownetb_state)):
            if state is not None:
                net.load_state_dict(state)
        self.model_names = ['netG', 'netD', 'flowNetF', 'flowNetB']
        # LightCNN is a frozen feature extractor: the reference leaves requires_grad=True on its
        # weights and so computes 14.5 GFLOP/image of weight gradients that no optimiser ever reads
        # (SURVEY 8a a15).  They are not computed here; every loss and every applied gradient is
        # unchanged (the gradient w.r.t. the generated image still flows through the network).
        for p in self.lightCNN.parameters():
            p.requires_grad_(False)

        self.criterionL1 = torch.nn.L1Loss().to(dev)
        self.criterionIllu = losses.MSL1Loss(self.criterionL1).to(dev)
        with torch.device(dev):
            self.criterionPerceptual = losses.PerceptualLoss()
        if vgg_weights is not None:
            self.criterionPerceptual.vgg.load_torchvision(vgg_weights)
        self.criterionIden = losses.IdentityLoss(self.lightCNN, crop=crop).to(dev)
        self.criterionGAN = losses.GANLoss('lsgan').to(dev)

        flow_params = itertools.chain(self.flowNetF.parameters(), self.flowNetB.parameters())
        adam = dict(betas=(0.5, 0.999), capturable=self._capturable, **_adam_impl(dev))
        self.optimizer_F = torch.optim.Adam(flow_params, lr=0.00005, **adam)
        self.optimizer_G = torch.optim.Adam(self.netG.parameters(), lr=0.0004, **adam)
        self.optimizer_D = torch.optim.Adam(self.netD.parameters(), lr=0.0004, **adam)
        self.optimizers = [self.optimizer_G, self.optimizer_F, self.optimizer_D]
        self._counters = DeferredCounters([self.netG, self.netD, self.flowNetF, self.flowNetB])     # BN call counters: one add per step
        self.optimizers_G = [self.optimizer_G, self.optimizer_F]
        self.optimizers_D = [self.optimizer_D]

        self.gf128 = external_function.GuidedFilter(32).to(dev)
        self.gf64 = external_function.GuidedFilter(16).to(dev)
        self.gf32 = external_function.GuidedFilter(8).to(dev)

        self.avg_D = self.avg_G = None
        if distributed is not None:
            distributed.broadcast_module_states([self.netG, self.netD, self.flowNetF, self.flowNetB, self.lightCNN,
                                                 self.criterionPerceptual])
            # eager steps launch the bucket all-reduces from inside the backward pass; CUDA-graph steps reduce one flat
            # buffer per averager between two graph replays (parallel.GradAverager)
            self.avg_D = GradAverager(list(self.netD.parameters()), distributed, in_backward=not graph)
            self.avg_G = GradAverager(list(self.netG.parameters()) + list(self.flowNetF.parameters())
                                      + list(self.flowNetB.parameters()), distributed, in_backward=not graph)

    # ------------------------------------------------------------------ input
    def set_input(self, batch):
        dev = self.device
        self.img_S = batch['img_S'].to(dev, non_blocking=True)
        self.img_F = batch['img_F'].to(dev, non_blocking=True)
        self.lm_F = batch['lm_F'].to(dev, non_blocking=True)
        self.mask_F = batch['mask_F'].to(dev, non_blocking=True).float()
        self.mask_S = batch['mask_S'].to(dev, non_blocking=True).float()
        self.titers = batch['titers']
        self.epoch = batch.get('epoch', 0)

    # ------------------------------------------------------------------ forward
    def _flownets(self):
        """flowNetF and flowNetB on the same input.  They are independent until the losses, and half of their 38
        convolutions each work on maps of 8x8 and below (a few CTAs on 148 SMs), so FFWM_FLOW_STREAMS=1 runs
        flowNetB on a second stream, forked from and joined back into the current one (legal inside CUDA-graph
        capture; autograd runs each net's backward on the stream its forward used).  Same arithmetic, same order
        within each net.  Default on since round 2 (profiles/r02a_switches.txt; FFWM_FLOW_STREAMS=0 for the A/B)."""
        if not (FLOW_STREAMS and self.device.type == "cuda"):
            flows_F = self.flowNetF(self.img_S)
            return flows_F, self.flowNetB(self.img_S)
        cur = torch.cuda.current_stream(self.device)
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        side = self._side_stream
        side.wait_stream(cur)                               # fork: img_S and the weights are ready on `cur`
        with torch.cuda.stream(side):
            flows_B = self.flowNetB(self.img_S)
        flows_F = self.flowNetF(self.img_S)
        cur.wait_stream(side)                               # join before anything consumes flows_B
        if not torch.cuda.is_current_stream_capturing():
            for t in flows_B:
                t.record_stream(cur)                        # allocated on `side`, consumed on `cur`
        return flows_F, flows_B

    def forward(self):
        (flow_F128, flow_F64, flow_F32), (self.flow_B128, self.flow_B64, self.flow_B32) = self._flownets()
        self.img_S_warp = self.warpNet(self.img_S, flow_F128)
        self.img_S_rec = self.warpNet(self.img_F, self.flow_B128)
        self.fake_F32, self.fake_F64, self.fake_F128 = self.netG(self.img_S, flow=[flow_F32, flow_F64, flow_F128])
        self.img_GF128 = self.gf128(self.fake_F128, self.img_F)
        # facial parts (eyes, nose, mouth) cropped from the generated and the real frontal face
        self.parts = []
        for grid in self.get_part_grid():
            self.parts.append((self.warpNet(self.img_GF128, grid), self.warpNet(self.img_F, grid)))

    def backward_D(self):
        dis_fake = self.netD(self.img_GF128.detach() * self.mask_F)
        dis_real = self.netD(self.img_F * self.mask_F)
        self.loss_D = (self.criterionGAN(dis_fake, False, for_dis=True)
                       + self.criterionGAN(dis_real, True, for_dis=True)) * 0.5
        self.loss_D.backward()

    def backward_G(self):
        img_F, mask_F = self.img_F, self.mask_F
        img_F64 = F.interpolate(img_F, (64, 64), mode='bilinear')
        img_F32 = F.interpolate(img_F, (32, 32), mode='bilinear')
        mask_F64 = F.interpolate(mask_F, (64, 64), mode='nearest')
        mask_F32 = F.interpolate(mask_F, (32, 32), mode='nearest')
        if self.titers < 20000:          # warm-up: losses on the raw generator outputs
            gf128, gf64, gf32 = self.fake_F128, self.fake_F64, self.fake_F32
        else:
            gf128 = self.img_GF128
            gf64 = self.gf64(self.fake_F64, img_F64)
            gf32 = self.gf32(self.fake_F32, img_F32)
        scales = ((gf128, img_F, mask_F, 1), (gf64, img_F64, mask_F64, 1), (gf32, img_F32, mask_F32, 1.5))
        if BATCHED_VGG:      # one VGG pass per resolution over the generated and one over the target images
            prc = self.criterionPerceptual.many([(g * m, t * m) for g, t, m, _ in scales] + list(self.parts))
            self.loss_prc = sum(w * l for (_, _, _, w), l in zip(scales, prc))
        else:
            self.loss_prc = sum(w * self.criterionPerceptual(g * m, t * m) for g, t, m, w in scales)
        self.loss_l1 = sum(w * self.criterionL1(g * m, t * m) for g, t, m, w in scales)
        self.loss_illu = self.criterionIllu([self.flow_B128, self.flow_B64, self.flow_B32],
                                            [self.fake_F128, self.fake_F64, self.fake_F32], self.img_S, self.mask_S)
        if BATCHED_VGG:      # likewise the two LightCNN identity losses: 2 passes instead of 4
            self.loss_iden, self.loss_iden_gf = self.criterionIden.many([(self.fake_F128, img_F), (gf128, img_F)])
        else:
            self.loss_iden = self.criterionIden(self.fake_F128, img_F)
            self.loss_iden_gf = self.criterionIden(gf128, img_F)
        self.loss_adv = self.criterionGAN(self.netD(self.img_GF128 * mask_F), True, for_dis=False)
        (eyel, eyer, nose, mouth) = prc[3:] if BATCHED_VGG else [self.criterionPerceptual(g, t) for g, t in self.parts]
        self.loss_fc = 2 * (eyel + eyer) + mouth + nose

        self.loss_l1 = self.loss_l1 * 5
        self.loss_iden = self.loss_iden * 0.5 + self.loss_iden_gf * 1
        self.loss_adv = self.loss_adv * 0.1
        self.loss_illu = self.loss_illu * 15
        self.loss_G = self.loss_iden + self.loss_l1 + self.loss_prc + self.loss_illu + self.loss_fc + self.loss_adv
        self.loss_G.backward()

    def _zero_D(self):
        if self.avg_D is not None:
            self.avg_D.zero()              # data parallel: the gradients live in the averager's persistent buckets
        else:
            self._zero(self.optimizers_D)

    def _zero_G(self):
        if self.avg_G is not None:
            self.avg_G.zero()
        else:
            self._zero(self.optimizers_G)

    def optimize_parameters(self):
        self.forward()
        set_requires_grad(self.netD, True)
        self._zero_D()
        self.backward_D()                  # data parallel: bucket all-reduces start inside the backward pass (hooks)
        if self.avg_D is not None:
            self.avg_D.average()
        self._step(self.optimizers_D)
        set_requires_grad(self.netD, False)
        self._zero_G()
        self.backward_G()
        if self.avg_G is not None:
            self.avg_G.average()
        self._step(self.optimizers_G)
        self._counters.flush()

    # ------------------------------------------------------------------ CUDA graph
    def enable_cuda_graph(self, example_batch, warmup=3, segmented=None):
        """Capture `optimize_parameters()` — ~10^4 kernel launches — into CUDA graph(s) and replay
        from then on (`step` copies the batch into the captured input buffers).  The step is
        launch-bound when issued from Python; replaying removes the CPU from the loop.  Requires the
        optimisers to be built with `capturable=True` (constructor argument `graph=True`).

        Single process: one graph.  Data parallel: three graphs sharing one memory pool,
            [forward, backward_D]  | all-reduce D grads |  [step D, backward_G]  | all-reduce G,F grads |  [step G,F]
        with the NCCL all-reduces issued eagerly between the replays, each ONE collective over the averager's persistent
        flat gradient buffer (the captured backward passes accumulate straight into it: no pack / unpack copies, NCCL
        AVG instead of a scaling pass).  Capturing the collectives into a single graph was measured and rejected
        (parallel.GradAverager)."""
        assert self.device.type == "cuda" and self._capturable, "construct the trainer with graph=True"
        flat_grads = self.avg_G is not None and self.avg_G.overlap      # gradients live in the averagers' flat buffers
        if segmented is None:
            segmented = self.avg_D is not None
        self._static = {k: (v.to(self.device).clone() if torch.is_tensor(v) else v) for k, v in example_batch.items()}
        self._graphs = None
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):              # cudnn autotuning, allocator warm-up, averager plans
                self._bind(self._static)
                self.optimize_parameters()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if not flat_grads:
            self._zero(self.optimizers)
        from . import _lib
        self._bind(self._static)
        n0 = _lib.kernel_launches()
        if not segmented:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.optimize_parameters()
            self._graphs = [(g, None)]
        else:
            g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                self.forward()
                set_requires_grad(self.netD, True)
                self._zero_D()
                self.backward_D()
            pool = g1.pool()
            with torch.cuda.graph(g2, pool=pool):
                self._step(self.optimizers_D)
                set_requires_grad(self.netD, False)
                self._zero_G()
                self.backward_G()
            with torch.cuda.graph(g3, pool=pool):
                self._step(self.optimizers_G)
                self._counters.flush()
            self._graphs = [(g1, self.avg_D), (g2, self.avg_G), (g3, None)]
        self.graph_kernel_nodes = _lib.kernel_launches() - n0     # ffwm_b200 kernels recorded in the graph(s)
        self.graph_replays = 0

    def _mark(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream(self.device))
        return e

    def _bind(self, batch):
        self.img_S, self.img_F, self.lm_F = batch['img_S'], batch['img_F'], batch['lm_F']
        self.mask_F, self.mask_S = batch['mask_F'].float(), batch['mask_S'].float()
        self.titers = batch['titers']
        self.epoch = batch.get('epoch', 0)

    def step(self, batch):
        """set_input + optimize_parameters; replays the captured graph when there is one."""
        if getattr(self, "_graphs", None) is None:
            self.set_input(batch)
            self.optimize_parameters()
            return
        if batch is not self._static:
            assert (batch['titers'] < 20000) == (self._static['titers'] < 20000), "warm-up branch is baked into the graph"
            for k, v in self._static.items():
                if torch.is_tensor(v):
                    v.copy_(batch[k], non_blocking=True)
        timing = getattr(self, "phase_events", None)       # diagnostic: [(name, start, end)] CUDA events per phase
        for i, (graph, averager) in enumerate(self._graphs):
            if timing is not None:
                timing.append(("graph%d" % i, self._mark(), None))
            graph.replay()
            if timing is not None:
                timing.append(("allreduce%d" % i, self._mark(), None))
            if averager is not None:
                averager.average()
        if timing is not None:
            timing.append(("end", self._mark(), None))
        self.graph_replays += 1

    @staticmethod
    def _zero(opts):
        for o in opts:
            o.zero_grad()

    @staticmethod
    def _step(opts):
        for o in opts:
            o.step()

    def get_current_losses(self):
        """float() of each loss: the per-iteration device->host read the reference performs
        (models/base_model.py:164-170)."""
        return {n: float(getattr(self, n).detach()) for n in self.loss_names if hasattr(self, n)}

    # ------------------------------------------------------------------ facial-part grids
    def get_part_grid(self):
        """32x32 crops centred on the eye / nose / mouth landmarks (models/ffwm_model.py:217-246);
        order: left eye, right eye, nose, mouth."""
        lm = self.lm_F
        el, er = lm[:, 63:64], lm[:, 515:516]
        mouth = torch.cat((lm[:, 64:128], lm[:, 516:580]), 1)
        # LongTensor / 2: true division on torch >= 1.6 (x.5 centres), floor division on the torch 1.5 the reference was
        # written for (models/ffwm_model.py:225).  Parity here is pinned to the UNMODIFIED reference as it executes on the
        # installed torch (tests/golden/ref_train_step.json, ref_orchestrators.json), i.e. true division; unlike
        # MultiScaleLDLoss (losses._int_div) nothing fails either way, the mouth crop just sits half a pixel apart.
        mc = (mouth.min(dim=1, keepdim=True)[0] + mouth.max(dim=1, keepdim=True)[0]) / 2
        nc = lm[:, 429:430]
        return [self.build_grid(c, 32) for c in (el, er, nc, mc)]

    def build_grid(self, lm, d):
        """(b,2,d,d) sampling grid of a d x d patch centred on landmark `lm` (b,1,2), pixels -> [-1,1]."""
        r = d // 2
        line = torch.linspace(-r, r, d, device=self.device)
        centre = lm.float().view(-1, 2, 1, 1) - 64
        gx = line.view(1, 1, 1, d).expand(lm.size(0), 1, d, d) + centre[:, 0:1]
        gy = line.view(1, 1, d, 1).expand(lm.size(0), 1, d, d) + centre[:, 1:2]
        return torch.cat((gx, gy), 1) / 64


class FlowNetTrainer:
    """FlowNet pre-training step (models/flownet_model.py:16-78)."""
    loss_names = ['loss', 'loss_reg', 'loss_lm', 'loss_cor']

    def __init__(self, device, reverse=False, vgg_weights=None, distributed=None):
        self.device = torch.device(device)
        self.reverse = reverse
        self.flowNet = base_networks.FlowNet(64).to(self.device)
        self.warpNet = base_networks.WarpNet().to(self.device)
        self.criterionLD = losses.MultiScaleLDLoss().to(self.device)
        self.Correctness = losses.PerceptualCorrectness().to(self.device)
        if vgg_weights is not None:
            self.Correctness.vgg.load_torchvision(vgg_weights)
        self.Regularization = losses.MultiAffineRegularizationLoss(kz_dic={1: 7, 2: 5, 3: 3})
        self.optimizer = torch.optim.Adam(self.flowNet.parameters(), lr=0.0004, betas=(0.5, 0.999), **_adam_impl(self.device))
        self._counters = DeferredCounters([self.flowNet])
        self.avg = None
        if distributed is not None:
            distributed.broadcast_module_states([self.flowNet, self.Correctness])
            self.avg = GradAverager(list(self.flowNet.parameters()), distributed)

    def set_input(self, batch):
        dev = self.device
        s, f = ('F', 'S') if self.reverse else ('S', 'F')
        self.img_S = batch['img_' + s].to(dev).float()
        self.img_F = batch['img_' + f].to(dev).float()
        self.lm_S = batch['lm_' + s].to(dev).long()
        self.lm_F = batch['lm_' + f].to(dev).long()
        self.mask = batch['mask_S' if self.reverse else 'mask_F'].to(dev).float()
        gate = batch['gate'].to(dev).float()
        self.gate = torch.cat((gate, gate), 2)

    def forward(self):
        self.flow, self.flow64, self.flow32 = self.flowNet(self.img_F if self.reverse else self.img_S)
        self.fake_F = self.warpNet(self.img_S, self.flow)

    def backward(self):
        flows = [self.flow, self.flow64, self.flow32]
        self.loss_cor = self.Correctness(self.img_F, self.img_S, flows[::-1], [2, 1, 0], norm_mask=self.mask) * 20
        self.loss_reg = self.Regularization(flows[::-1]) * 0.01
        self.loss_lm = self.criterionLD(flows, self.lm_S, self.lm_F, self.gate)
        self.loss = self.loss_cor + self.loss_lm + self.loss_reg
        self.loss.backward()

    def optimize_parameters(self):
        self.forward()
        self.optimizer.zero_grad()
        self.backward()
        if self.avg is not None:
            self.avg.average()
        self.optimizer.step()
        self._counters.flush()

    def get_current_losses(self):
        return {n: float(getattr(self, n).detach()) for n in self.loss_names if hasattr(self, n)}

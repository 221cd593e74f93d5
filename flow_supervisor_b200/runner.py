"""Opt-in inference driver around an UNMODIFIED reference model.

``RaftRunner(model)`` runs the arithmetic of ``model.forward(image1, image2, iters, test_mode=True)``
(/root/reference/pytorch/core/raft.py:86-144; gma_network.py:74-129) with the model's own
sub-modules (``fnet``, ``cnet``, ``update_block``, ``att``) and weights, restructured for the GPU:

* the correlation block is this package's ``CorrBlock`` (build + one lookup launch per iteration);
* the whole forward -- encoders, volume build, every GRU iteration (lookup, update block,
  ``coords1 += delta``, raft.py:122-132) and the final upsampling -- is captured ONCE per input shape
  in a CUDA graph and replayed: the reference issues ~60 small kernels and 4 host-to-device copies per
  iteration from Python (corr.py:29-50), here an iteration is a fixed chain of launches with no host
  work in between (row f2);
* in test mode only the LAST iteration's upsampled flow is returned (raft.py:141-142), so the convex
  upsampling (raft.py:72-83, a 576-channel softmax per iteration in the reference) runs once, through
  ``fc_upsample_flow`` (row f4; ``fused_upsample=False`` keeps the model's own method).

Graph on / off is bit-identical (same kernels, same order); against the reference forward the result
is within the flow tolerance of the drop-in block (tests/test_gpu_runner.py).  Inference only: the
reference's training path keeps running unchanged through ``patch_reference()``.
"""
from __future__ import annotations

import torch

from . import ops
from .corr import CorrBlock, coords_grid


class RaftRunner:
    def __init__(self, model, iters: int = 12, graph: bool = True, fused_upsample: bool = True,
                 corr_block=CorrBlock, fused_fnet_tail: bool = False, fused_convc1: bool = False):
        self.model = model
        # fused_convc1 (row f1): the lookup and the motion encoder's first 1x1 convolution + ReLU run as one kernel
        # (CorrBlock.lookup_convc1); the rest of the update block runs through the model's own sub-modules.
        self.fused_convc1 = bool(fused_convc1)
        self._packed_convc1 = None
        # fused_fnet_tail (row f3): fnet's 1x1 output convolution runs inside the volume build (CorrBlock.from_fnet_tail)
        # -- no fp32 feature maps, no pack launch.  Needs an encoder shaped like extractor.py's BasicEncoder.
        self.fused_fnet_tail = bool(fused_fnet_tail)
        self._packed_tail = None
        self.iters = int(iters)
        self.use_graph = bool(graph)
        self.fused_upsample = bool(fused_upsample)
        self.corr_block = corr_block
        self._graphs = {}            # (shape, dtype, device, flow_init?) -> (graph, static inputs, static outputs)

    # ------------------------------------------------------------------ the forward itself
    def _autocast(self):
        return torch.autocast("cuda", enabled=bool(getattr(self.model.args, "mixed_precision", False)))

    def forward_eager(self, image1, image2, flow_init=None):
        m = self.model
        hdim, cdim = m.hidden_dim, m.context_dim
        image1 = (2 * (image1 / 255.0) - 1.0).contiguous()                   # raft.py:89-93
        image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
        if self.fused_fnet_tail and self._tail_ok(m.fnet):
            f = m.fnet
            with self._autocast():                                           # extractor.py:166-182 without conv2
                x = torch.cat([image1, image2], dim=0)
                x = f.layer3(f.layer2(f.layer1(f.relu1(f.norm1(f.conv1(x))))))
            if self._packed_tail is None:
                self._packed_tail = ops.fnet_tail_prepare(f.conv2.weight.detach(), f.conv2.bias.detach())
            corr_fn = self.corr_block.from_fnet_tail(x.float(), self._packed_tail, f.conv2.out_channels,
                                                     radius=m.args.corr_radius)
        else:
            with self._autocast():
                fmap1, fmap2 = m.fnet([image1, image2])                      # raft.py:99-100
            corr_fn = self.corr_block(fmap1.float(), fmap2.float(), radius=m.args.corr_radius)
        with self._autocast():
            cnet = m.cnet(image1)                                            # raft.py:110-114
            net, inp = torch.split(cnet, [hdim, cdim], dim=1)
            net, inp = torch.tanh(net), torch.relu(inp)
            attention = m.att(inp) if hasattr(m, "att") else None            # gma_network.py:100
        B, _, H, W = image1.shape
        coords0 = coords_grid(B, H // 8, W // 8, device=image1.device)
        coords1 = coords0.clone()
        if flow_init is not None:
            coords1 = coords1 + flow_init
        up_mask = None
        fuse_c1 = self.fused_convc1 and attention is None and self._convc1_ok(m.update_block, corr_fn)
        if fuse_c1 and self._packed_convc1 is None:
            c1 = m.update_block.encoder.convc1
            self._packed_convc1 = ops.convc1_prepare(c1.weight.detach(), c1.bias.detach())
        for _ in range(self.iters):                                          # raft.py:122-132
            flow = coords1 - coords0
            if fuse_c1:
                cor = corr_fn.lookup_convc1(coords1, self._packed_convc1)    # = relu(convc1(corr_fn(coords1)))
                with self._autocast():
                    net, up_mask, delta = self._update_after_convc1(m.update_block, net, inp, cor, flow)
            else:
                corr = corr_fn(coords1)
                with self._autocast():
                    if attention is None:
                        net, up_mask, delta = m.update_block(net, inp, corr, flow)
                    else:
                        net, up_mask, delta = m.update_block(net, inp, corr, flow, attention)
            coords1 = coords1 + delta
        flow_low = coords1 - coords0
        if up_mask is None:                                                  # small model: bilinear x8 (utils.py:80-82)
            flow_up = 8 * torch.nn.functional.interpolate(flow_low, scale_factor=8, mode="bilinear", align_corners=True)
        elif self.fused_upsample:
            flow_up = ops.upsample_flow(flow_low.float(), up_mask.float())
        else:
            flow_up = m.upsample_flow(flow_low, up_mask)
        return flow_low, flow_up

    @staticmethod
    def _convc1_ok(ub, corr_fn) -> bool:
        enc = getattr(ub, "encoder", None)
        c1 = getattr(enc, "convc1", None)
        return (isinstance(c1, torch.nn.Conv2d) and tuple(c1.weight.shape) == (256, 324, 1, 1) and c1.bias is not None
                and all(hasattr(enc, n) for n in ("convc2", "convf1", "convf2", "conv"))
                and all(hasattr(ub, n) for n in ("gru", "flow_head", "mask")) and hasattr(corr_fn, "lookup_convc1"))

    @staticmethod
    def _update_after_convc1(ub, net, inp, cor, flow):
        """BasicUpdateBlock.forward (update.py:127-136) with BasicMotionEncoder.forward (update.py:89-98) entered after
        its first convolution + ReLU."""
        relu = torch.nn.functional.relu
        enc = ub.encoder
        cor = relu(enc.convc2(cor))
        flo = relu(enc.convf2(relu(enc.convf1(flow))))
        out = relu(enc.conv(torch.cat([cor, flo], dim=1)))
        motion = torch.cat([out, flow], dim=1)
        net = ub.gru(net, torch.cat([inp, motion], dim=1))
        delta = ub.flow_head(net)
        return net, .25 * ub.mask(net), delta

    @staticmethod
    def _tail_ok(fnet) -> bool:
        conv2 = getattr(fnet, "conv2", None)
        return (all(hasattr(fnet, n) for n in ("conv1", "norm1", "relu1", "layer1", "layer2", "layer3"))
                and isinstance(conv2, torch.nn.Conv2d) and conv2.kernel_size == (1, 1) and conv2.bias is not None
                and not (fnet.training and getattr(fnet, "dropout", None) is not None))

    # ------------------------------------------------------------------ graph capture / replay
    def _capture(self, image1, image2, flow_init):
        s_im1, s_im2 = torch.empty_like(image1), torch.empty_like(image2)
        s_init = torch.empty_like(flow_init) if flow_init is not None else None
        s_im1.copy_(image1); s_im2.copy_(image2)
        if s_init is not None:
            s_init.copy_(flow_init)
        side = torch.cuda.Stream(device=image1.device)
        side.wait_stream(torch.cuda.current_stream(image1.device))
        with torch.cuda.stream(side):                                        # warm-up: lazy inits, cuDNN plans, caches
            for _ in range(2):
                self.forward_eager(s_im1, s_im2, s_init)
        torch.cuda.current_stream(image1.device).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.forward_eager(s_im1, s_im2, s_init)
        return g, (s_im1, s_im2, s_init), out

    @torch.no_grad()
    def __call__(self, image1, image2, flow_init=None):
        if not image1.is_cuda:
            raise RuntimeError("RaftRunner needs CUDA images: this package has no CPU fallback")
        if not self.use_graph:
            return self.forward_eager(image1, image2, flow_init)
        key = (tuple(image1.shape), image1.dtype, image1.device, flow_init is not None, self.iters,
               self.fused_upsample, self.fused_fnet_tail, self.fused_convc1)
        if key not in self._graphs:
            self._graphs[key] = self._capture(image1, image2, flow_init)
        g, (s_im1, s_im2, s_init), (flow_low, flow_up) = self._graphs[key]
        s_im1.copy_(image1); s_im2.copy_(image2)
        if s_init is not None:
            s_init.copy_(flow_init)
        g.replay()
        return flow_low.clone(), flow_up.clone()

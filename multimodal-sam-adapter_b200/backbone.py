"""Registry entries of the drop-in boundary: same class names, constructor kwargs,
forward contract and state-dict schema as the reference backbones / head / segmentor."""
import math
import warnings

import numpy as np
import torch
import torch.nn as nn

from . import nn_modules as M
from .engine import EncoderEngine
from .ops.modules import MSDeformAttn
from .registry import BACKBONES, HEADS, SEGMENTORS, build_backbone, build_head


def _trunc_normal(t, std=0.02):
    nn.init.trunc_normal_(t, std=std)


def _init_weights(m):
    """..._new.py:119-134."""
    if isinstance(m, nn.Linear):
        _trunc_normal(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, (nn.LayerNorm, nn.BatchNorm2d)):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
        fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
        m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
        if m.bias is not None:
            m.bias.data.zero_()


@BACKBONES.register_module(force=True)
class ImageEncoderViT(nn.Module):
    """SAM ViT encoder weights (base/image_encoder.py:187-328)."""

    def __init__(self, img_size=1024, patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0,
                 qkv_bias=True, norm_layer=None, act_layer=None, use_abs_pos=True, use_rel_pos=True,
                 rel_pos_zero_init=True, window_size=14, global_attn_indexes=(5, 11, 17, 23), pretrained=None,
                 with_cp=False, pretrained_size=1024, fix=False):
        super().__init__()
        if embed_dim % num_heads or embed_dim // num_heads != 64:
            raise ValueError("the B200 attention kernel is built for head_dim 64 (SAM ViT-B / ViT-L; ViT-H has head_dim 80); "
                             f"got embed_dim={embed_dim}, num_heads={num_heads}")
        self.img_size, self.embed_dim, self.num_heads, self.patch_size = img_size, embed_dim, num_heads, patch_size
        self.patch_embed = M.PatchEmbed((patch_size, patch_size), (patch_size, patch_size), in_chans, embed_dim)
        self.pos_embed = None
        if use_abs_pos:
            self.pos_embed = nn.Parameter(torch.zeros(1, pretrained_size // patch_size, pretrained_size // patch_size, embed_dim))
        self.blocks = nn.ModuleList()
        gidx = list(global_attn_indexes)
        for i in range(depth):
            self.blocks.append(M.Block(embed_dim, num_heads, mlp_ratio, qkv_bias, use_rel_pos,
                                       window_size if i not in gidx else 0,
                                       (pretrained_size // patch_size, pretrained_size // patch_size)))
        if isinstance(pretrained, str):
            self.init_weights(pretrained)

    def init_weights(self, pretrained=None):
        if isinstance(pretrained, str):
            from .checkpoint import load_checkpoint        # mmcv-style: unwrap, strip 'module.', non-strict
            load_checkpoint(self, pretrained, strict=False)


@BACKBONES.register_module(force=True)
class TwinConvNeXt(M.TwinConvNeXt):
    pass


class _AdapterBase(ImageEncoderViT):
    def __init__(self, pretrain_size=1024, num_heads=12, conv_inplane=64, n_points=4,
                 modalities_name=("rgb", "depth", "lidar", "event"), modalities_ch=(3, 3, 3, 1), deform_num_heads=6,
                 init_values=0., gamma_init_values=0., interaction_indexes=None, with_cffn=True, cffn_ratio=0.25,
                 deform_ratio=1.0, add_vit_feature=True, pretrained=None, use_extra_extractor=True, with_cp=True,
                 drop_path_rate=0.4, drop_rate=0., drop_multimodal_path=0.2, arch="base", checkpoint="check",
                 conv_drop_path_rate=None, *args, **kwargs):
        super().__init__(num_heads=num_heads, pretrained=pretrained, with_cp=with_cp, *args, **kwargs)
        modalities_name, modalities_ch = list(modalities_name), list(modalities_ch)
        if "rgb" not in modalities_name or len(modalities_name) != 2:
            raise NotImplementedError("the B200 path implements the bimodal (rgb + one auxiliary modality) adapter "
                                      "(SpatialPriorModuleBimodal); got modalities " + str(modalities_name))
        self.in_ch_im = modalities_ch[modalities_name.index("rgb")]
        img_size = kwargs.get("img_size")
        self.cfg = dict(img_size=img_size, modalities_name=modalities_name, modalities_ch=modalities_ch,
                        interaction_indexes=[list(i) for i in interaction_indexes], add_vit_feature=add_vit_feature,
                        use_extra_extractor=use_extra_extractor, num_heads=num_heads, deform_num_heads=deform_num_heads,
                        n_points=n_points, arch=arch)
        self.interaction_indexes = interaction_indexes
        self.add_vit_feature = add_vit_feature
        E = self.embed_dim
        self.spm = M.SpatialPriorModuleBimodal(conv_inplane, E, img_size, arch)
        self.up = nn.ConvTranspose2d(E, E, 2, 2)
        self.up.apply(_init_weights)
        self.level_embed = nn.Parameter(torch.zeros(3, E))
        n_int = len(interaction_indexes)
        self.interactions = nn.Sequential(*[
            M.InteractionBlock(E, deform_num_heads, n_points, with_cffn, cffn_ratio, init_values, deform_ratio,
                               (i == n_int - 1) and use_extra_extractor, MSDeformAttn) for i in range(n_int)])
        self.norm1 = nn.BatchNorm2d(E)
        self.norm2 = nn.BatchNorm2d(E)
        self.norm3 = nn.BatchNorm2d(E)
        self.norm4 = nn.BatchNorm2d(E)
        self.interactions.apply(_init_weights)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        nn.init.normal_(self.level_embed)
        self._engines = {}
        self._stamp_tensors = {}

    # ------------------------------------------------------------------
    def _weights_stamp(self, head):
        """Changes whenever a parameter / buffer of the backbone or head is written in place (load_state_dict, mmcv's
        load_checkpoint via _load_from_state_dict, optimizer steps ...): the sum of the tensors' version counters. The
        tensor list is cached per head (cleared by invalidate(); module surgery after the first forward needs that call)."""
        ts = self._stamp_tensors.get(id(head))
        if ts is None:
            ts = [t for m in (self, head) if m is not None for t in list(m.parameters()) + list(m.buffers())]
            self._stamp_tensors[id(head)] = ts
        return sum(t._version for t in ts)

    def engine(self, head=None):
        """Packed weights + captured graphs, one per (device, head): `backbone.forward` (no head) and the segmentor's label
        paths (with the decode head) keep separate engines instead of rebuilding each other's. An engine is rebuilt
        when any weight of the backbone / head has changed since it was packed."""
        dev = next(self.parameters()).device
        key = (str(dev), id(head))
        stamp = self._weights_stamp(head)
        ent = self._engines.get(key)
        if ent is None or ent[0] != stamp:
            for k in [k for k in self._engines if k[0] != str(dev)]:
                del self._engines[k]
            ent = (stamp, EncoderEngine(self, head, dev))
            self._engines[key] = ent
        return ent[1]

    def invalidate(self):
        """Drop the packed weights (they are also re-packed automatically when a weight changes)."""
        self._engines = {}
        self._stamp_tensors = {}

    @torch.no_grad()
    def forward(self, x):
        """x [B, sum(modalities_ch), H, W] -> ([f1..f4] NCHW, None)   (..._new.py:161-349)."""
        if self.training:
            raise RuntimeError("the B200 path is inference-only: call .eval() first")
        feats = self.engine().backbone_nhwc(x)
        return [f.permute(0, 3, 1, 2).contiguous().to(x.dtype) for f in feats], None


@BACKBONES.register_module(force=True)
class SAMAdapterbimodalMixModNewInTwinConvNEW(_AdapterBase):
    pass


@BACKBONES.register_module(force=True)
class SAMAdapterbimodalMixModNewInTwinConvNEWwithcp(_AdapterBase):
    pass


@HEADS.register_module(force=True)
class SegformerHead(M.SegformerHeadParams):
    """decode_heads/segformer_head.py:11-66; kwargs of mmseg BaseDecodeHead accepted."""

    def __init__(self, interpolate_mode="bilinear", in_channels=None, in_index=None, channels=None, num_classes=None,
                 dropout_ratio=0.1, norm_cfg=None, act_cfg=None, align_corners=False, loss_decode=None,
                 input_transform="multiple_select", **_ignored):
        if interpolate_mode != "bilinear" or align_corners:
            raise NotImplementedError("SegformerHead (B200): bilinear, align_corners=False only")
        in_index = list(range(len(in_channels))) if in_index is None else list(in_index)
        assert len(in_channels) == len(in_index)
        super().__init__(list(in_channels), channels, num_classes)
        self.in_channels, self.in_index, self.channels, self.num_classes = list(in_channels), in_index, channels, num_classes
        self.align_corners = align_corners


@SEGMENTORS.register_module(force=True)
class EncoderDecoder(nn.Module):
    """Inference side of segmentors/encoder_decoder.py: extract_feat, encode_decode, whole_dim /
    whole_dim_cut / whole / slide inference, simple_test."""

    def __init__(self, backbone, decode_head, neck=None, auxiliary_head=None, train_cfg=None, test_cfg=None,
                 pretrained=None, init_cfg=None):
        super().__init__()
        backbone = dict(backbone)
        if pretrained is not None:
            assert backbone.get("pretrained") is None, "both backbone and segmentor set pretrained weight"
            backbone["pretrained"] = pretrained
        if neck is not None or auxiliary_head is not None:
            raise NotImplementedError("neck / auxiliary_head are not part of the shipped configs")
        self.backbone = build_backbone(backbone)
        self.decode_head = build_head(decode_head)
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes
        self.train_cfg, self.test_cfg = train_cfg, dict(test_cfg or {})

    def extract_feat(self, img):
        x, *_ = self.backbone(img)
        return x

    def _engine(self):
        return self.backbone.engine(self.decode_head)

    use_cuda_graph = True

    @torch.no_grad()
    def encode_decode_labels(self, img, out_hw=None, crop_hw=None):
        """uint8 labels [B, H, W] on the device. With use_cuda_graph the forward is replayed from a per-shape CUDA
        graph and the result lives in the graph's static buffer (copy it before the next call)."""
        if self.use_cuda_graph:
            return self._engine().segment_graphed(img, out_hw, crop_hw)
        return self._engine().segment(img, out_hw, crop_hw)

    def set_input_pipeline(self, mean, std, to_rgb=(False, False), norm_by_max=True, pad_size=None, pad_val=0.0):
        """The test pipeline's Normalize_multimodal + Pad_multimodal arguments (the config's `mod_norm_cfg`, `norm_by_max`,
        `Pad_multimodal(size=..., pad_val=...)`: pipelines/transform.py:2717-2815, 2934): mean / std are the concatenated
        per-channel statistics of [rgb | auxiliary]. Enables the uint8 entry points (stream_labels on (rgb_u8, aux_u8) batches,
        u8_input): normalisation, padding and the HWC -> CHW transpose then run inside the stem / patch-embed kernels."""
        cin = self.backbone.in_ch_im
        pre = 1.0 / 255.0 if norm_by_max else 1.0
        self._pipeline = dict(rgb=(list(mean[:cin]), list(std[:cin]), pre, bool(to_rgb[0])),
                              aux=(list(mean[cin:]), list(std[cin:]), pre, bool(to_rgb[1])),
                              pad_size=None if pad_size is None else tuple(pad_size), pad_val=float(pad_val))

    def u8_input(self, rgb_u8, aux_u8):
        """uint8 HWC device frames [B, H, W, 3] x 2 -> the engine's U8Input (needs set_input_pipeline)."""
        from . import kernels as K
        from .engine import U8Input
        pl = getattr(self, "_pipeline", None)
        if pl is None:
            raise RuntimeError("call set_input_pipeline(mean, std, ...) before feeding uint8 frames")
        hw = pl["pad_size"] or tuple(rgb_u8.shape[1:3])
        mods = [K.U8Modality(d, *pl[k], pad_val=pl["pad_val"]) for k, d in (("rgb", rgb_u8), ("aux", aux_u8))]
        return U8Input(mods[0], mods[1], hw)

    @torch.no_grad()
    def stream_labels(self, batches, out_hw=None, crop_hw=None):
        """Whole-image inference over an iterable of HOST batches (the test loop of mmseg_custom/apis/test_bs.py:91-163 feeds
        one DataLoader batch at a time): yields one uint8 [B, H, W] label tensor in pinned host memory per batch, in order.
        A batch is either the fp32 network input [B, C, H, W] or — after set_input_pipeline — a pair of uint8 HWC frame
        tensors (rgb [B, H, W, 3], auxiliary [B, H, W, 3]) straight from the decoder: 1 byte per value over PCIe instead
        of 4. The host->device copy of batch i+1 runs on a copy stream while batch i is computed, and the labels of batch i
        are read back asynchronously, so a step costs max(copy, compute) instead of their sum. Pass pinned tensors
        (tensor.pin_memory()); pageable ones still work but their copies are synchronous. A yielded tensor is valid
        until the generator has been advanced twice."""
        dev = next(self.parameters()).device
        comp = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
            self._stage = [None, None]
            self._out_host = [None, None]
        copy = self._copy_stream
        free_ev = [None, None]

        def upload(x, slot):
            xs = list(x) if isinstance(x, (tuple, list)) else [x]
            bufs = self._stage[slot]
            if bufs is None or len(bufs) != len(xs) or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(bufs, xs)):
                bufs = self._stage[slot] = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in xs]   # allocated on comp
                copy.wait_stream(comp)
            with torch.cuda.stream(copy):
                if free_ev[slot] is not None:
                    copy.wait_event(free_ev[slot])          # the forward that read this slot has consumed it
                for b, t in zip(bufs, xs):
                    b.copy_(t, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            return bufs, ev

        it = iter(batches)
        try:
            nxt = upload(next(it), 0)
        except StopIteration:
            return
        prev, i = None, 0
        while nxt is not None:
            (bufs, ev), slot = nxt, i % 2
            try:
                nxt = upload(next(it), (i + 1) % 2)
            except StopIteration:
                nxt = None
            comp.wait_event(ev)
            lab = self.encode_decode_labels(bufs[0] if len(bufs) == 1 else self.u8_input(bufs[0], bufs[1]), out_hw, crop_hw)
            free_ev[slot] = torch.cuda.Event()
            free_ev[slot].record(comp)
            out = self._out_host[slot]
            if out is None or out.shape != lab.shape:
                out = self._out_host[slot] = torch.empty(lab.shape, dtype=lab.dtype).pin_memory()
            out.copy_(lab, non_blocking=True)
            done = torch.cuda.Event()
            done.record(comp)
            if prev is not None:
                prev[1].synchronize()
                yield prev[0]
            prev, i = (out, done), i + 1
        prev[1].synchronize()
        yield prev[0]

    # ------------------------------------------------------------------ logits-level API
    @torch.no_grad()
    def _head_logits(self, img):
        """fp32 channels-last head logits [B*h0*w0, npad] + (h0, w0)."""
        eng = self._engine()
        return eng.head_logits(eng.backbone_nhwc(img))

    @torch.no_grad()
    def encode_decode(self, img, img_metas=None):
        """Logits at image size (bilinear, align_corners=False): [B, num_classes, H, W] fp32 (encoder_decoder.py:85-95)."""
        from . import kernels as K
        logits, (h0, w0) = self._head_logits(img)
        B, (H, W) = img.shape[0], img.shape[2:]
        with torch.cuda.device(logits.device):
            lg = K.resize_logits(logits, B, (h0, w0), (H, W))
        return lg.view(B, H, W, -1)[..., : self.num_classes].permute(0, 3, 1, 2)

    def _slide_boxes(self, h_img, w_img):
        """Crop grid of slide_inference (encoder_decoder.py:198-212)."""
        h_stride, w_stride = self.test_cfg["stride"]
        h_crop, w_crop = self.test_cfg["crop_size"]
        h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
        w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
        boxes = []
        for hi in range(h_grids):
            for wi in range(w_grids):
                y1, x1 = hi * h_stride, wi * w_stride
                y2, x2 = min(y1 + h_crop, h_img), min(x1 + w_crop, w_img)
                boxes.append((max(y2 - h_crop, 0), max(x2 - w_crop, 0), y2, x2))
        return boxes

    @torch.no_grad()
    def slide_labels(self, img, crop_batch=8, rescale_to=None):
        """slide_inference (encoder_decoder.py:191-234) -> uint8 labels. The crops of a frame are independent, so they go
        through the network `crop_batch` at a time (one forward for the 6 crops of a MUSES frame); one kernel
        (mmsam_slide_merge_f32) then up-samples every crop's head logits, adds the overlaps, divides by the count and
        takes the argmax in a single pass over the frame. rescale_to=(H, W): the averaged logits are resized first
        (rescale=True with an ori_shape different from the frame, :223-229)."""
        from . import kernels as K
        B, _, h_img, w_img = img.shape
        boxes = self._slide_boxes(h_img, w_img)
        if len({(y2 - y1, x2 - x1) for y1, x1, y2, x2 in boxes}) != 1:
            raise NotImplementedError("frames smaller than the crop size (crops of different sizes)")
        per = max(crop_batch // B, 1)                      # crop positions per forward (each contributes B crops)
        chunks, hw = [], None
        for i in range(0, len(boxes), per):
            crops = torch.cat([img[:, :, y1:y2, x1:x2] for y1, x1, y2, x2 in boxes[i:i + per]], 0).contiguous()
            lg, hw = self._head_logits(crops)              # rows ordered (position, image, y, x)
            chunks.append(lg)
        logits = chunks[0] if len(chunks) == 1 else torch.cat(chunks, 0)
        with torch.cuda.device(logits.device):
            if rescale_to is None or tuple(rescale_to) == (h_img, w_img):
                return K.slide_merge(logits, B, hw, self.num_classes, (h_img, w_img), boxes)
            preds = K.slide_merge(logits, B, hw, self.num_classes, (h_img, w_img), boxes, want_preds=True)
            return K.upsample_argmax(preds, B, (h_img, w_img), self.num_classes, tuple(rescale_to))

    @torch.no_grad()
    def inference_labels(self, img, img_meta=None, rescale=True):
        """inference + argmax of simple_test (encoder_decoder.py:417-508) on the device -> uint8 labels [B, H', W'].
        Every mode of the reference: slide, whole (rescale -> ori_shape), whole_dim (-> test_cfg.dim), whole_dim_cut
        (-> dim when rescale, then the cut_dim window); flip / flip_direction of img_meta flips the result back."""
        from . import kernels as K
        mode = self.test_cfg.get("mode", "whole")
        meta = img_meta[0] if img_meta else {}
        in_hw = tuple(img.shape[2:])
        ori_hw = tuple(meta["ori_shape"][:2]) if "ori_shape" in meta else in_hw
        if "ori_shape" in meta:
            assert all(tuple(m["ori_shape"][:2]) == ori_hw for m in img_meta), "one ori_shape per batch (encoder_decoder.py:432-434)"
        if mode == "slide":
            lab = self.slide_labels(img, rescale_to=ori_hw if rescale else None)
        elif mode in ("whole", "whole_dim", "whole_dim_cut"):
            out_hw, crop = in_hw, None
            if mode == "whole":
                out_hw = ori_hw if rescale else in_hw
            elif mode == "whole_dim":
                if not rescale:
                    raise ValueError("whole_inference_dim returns nothing with rescale=False (encoder_decoder.py:329-362)")
                out_hw = tuple(self.test_cfg["dim"])
            else:
                if rescale:
                    out_hw = tuple(self.test_cfg["dim"])
                cw, ch = self.test_cfg["cut_dim"]
                crop = (min(ch, out_hw[0]), min(cw, out_hw[1]))
            lab = self.encode_decode_labels(img, out_hw, crop)
        else:
            raise ValueError(f"unknown test_cfg.mode {mode!r}")
        if meta.get("flip", False):
            direction = meta.get("flip_direction", "horizontal")
            assert direction in ("horizontal", "vertical")
            lab = lab.flip(dims=(2,) if direction == "horizontal" else (1,))
        return lab

    @torch.no_grad()
    def simple_test(self, img, img_meta=None, rescale=True):
        """-> list of np.int64 [H, W] label maps, one per image (encoder_decoder.py:471-508)."""
        return list(self.inference_labels(img, img_meta, rescale).cpu().numpy().astype(np.int64))

    def forward(self, img, img_metas=None, return_loss=False, rescale=True, **kw):
        if return_loss:
            raise RuntimeError("inference-only build")
        if isinstance(img, (list, tuple)):
            img = img[0]
        return self.simple_test(img, img_metas, rescale)

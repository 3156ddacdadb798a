"""Encoder forward on the tcgen05 convolution kernels (inference / no-grad path).

Takes the weights of an `OdometryModel` (reference parameter names) and runs
`ResNetModified._forward_impl` (reference: src/models/resnet_modified.py:95-120, BasicBlock :159-177)
as a chain of `delora_conv2d_fprop_bf16` launches on NHWC bf16 activations with materialised
circular-W / zero-H padding: stem 3x3 s(1,2) + tanh, 3x3/(1,2) max-pool, 4 stages x 2 BasicBlocks
(second conv fuses residual add + activation, 1x1 strided downsample convs), then average pool,
`fc` and the two MLP heads (tiny GEMMs, left to torch as SURVEY.md §7.6 allows).

Training still differentiates through the torch/cuDNN path (dgrad/wgrad kernels are the next
milestone); this class is what inference (`inference_only`, Tester / ROS node shape, BASELINE
config #5) and the encoder benchmark use.
"""
import torch

from .. import ops


def _prep(weight, cin_pad=None):
    """[Cout,Cin,k,k] fp32 -> [Cout, k*k, Cin_pad] bf16 (tap-major K)."""
    cout, cin, k, _ = weight.shape
    w = weight.detach().permute(0, 2, 3, 1).reshape(cout, k * k, cin)
    if cin_pad is not None and cin_pad > cin:
        w = torch.nn.functional.pad(w, (0, cin_pad - cin))
    return w.contiguous().to(torch.bfloat16)


class TensorCoreEncoder:
    def __init__(self, model):
        self.model = model
        self.act = ops.ACT_RELU if model.config["activation_fct"] == "relu" else ops.ACT_TANH
        r = model.resnet
        if r.conv1.weight.shape[0] % 64 != 0:
            raise Exception("tensor-core encoder needs channel counts that are multiples of 64 "
                            "(factor_fewer_resnet_channels = 1)")
        self.w_stem = _prep(r.conv1.weight, 64)
        self.blocks = []
        for li in range(1, 5):
            for blk in getattr(r, f"layer{li}"):
                stride = blk.stride if isinstance(blk.stride, tuple) else (blk.stride, blk.stride)
                self.blocks.append({
                    "w1": _prep(blk.conv1.weight), "w2": _prep(blk.conv2.weight),
                    "wd": _prep(blk.downsample[0].weight) if blk.downsample is not None else None,
                    "stride": stride, "cout": blk.conv1.weight.shape[0]})
        self._buf = {}

    def _buffer(self, tag, b, h, w, c, device):
        key = (tag, b, h, w, c)
        t = self._buf.get(key)
        if t is None:
            t = ops.padded_nhwc_zeros(b, h, w, c, device)      # halo rows stay zero forever
            self._buf[key] = t
        return t

    @torch.no_grad()
    def features(self, image_1, image_2):
        """-> [x1, x2, x3, x4] as padded NHWC bf16 + their (H, W)."""
        b, _, h, w = image_1.shape
        dev = image_1.device
        x = ops.images_to_nhwc(image_1.float().contiguous(), image_2.float().contiguous(), 64)
        w2 = w // 2
        y = ops.conv2d_fprop(x, self.w_stem, h, w, 3, (1, 2), self.act, None, self._buffer("stem", b, h, w2, 64, dev))
        w4 = w2 // 2
        cur = self._buffer("pool", b, h, w4, 64, dev)
        L = ops._lib.lib()
        ops._lib.check(L.delora_maxpool_w_nhwc_bf16(y.data_ptr(), b, h, w2, 64, cur.data_ptr(), ops._stream()),
                       "delora_maxpool_w_nhwc_bf16")
        ch, cw = h, w4
        feats = []
        for i, blk in enumerate(self.blocks):
            sh, sw = blk["stride"]
            oh, ow, co = ch // sh, cw // sw, blk["cout"]
            t1 = ops.conv2d_fprop(cur, blk["w1"], ch, cw, 3, (sh, sw), self.act, None,
                                  self._buffer(f"b{i}a", b, oh, ow, co, dev))
            if blk["wd"] is not None:
                ident = ops.conv2d_fprop(cur, blk["wd"], ch, cw, 1, (sh, sw), ops.ACT_NONE, None,
                                         self._buffer(f"b{i}d", b, oh, ow, co, dev))
            else:
                ident = cur
            cur = ops.conv2d_fprop(t1, blk["w2"], oh, ow, 3, (1, 1), self.act, ident,
                                   self._buffer(f"b{i}o", b, oh, ow, co, dev))
            ch, cw = oh, ow
            if i % 2 == 1:
                feats.append((cur, ch, cw))
        return feats

    @torch.no_grad()
    def forward(self, image_1, image_2):
        """(translation [B,3], quaternion [B,4]) like OdometryModel.forward (src/models/model.py:103-116)."""
        feats = self.features(image_1, image_2)
        x4, h4, w4 = feats[-1]
        pooled = x4[:, 1:h4 + 1, 1:w4 + 1, :].float().mean(dim=(1, 2))           # AdaptiveAvgPool2d((1,1))
        m = self.model
        out = m.resnet.fc(pooled)
        if m.config["use_single_mlp_at_output"]:
            o = m.fully_connected_rot_trans(out)
            rot, trans = o[:, :4], o[:, 4:]
        else:
            rot, trans = m.fully_connected_rotation(out), m.fully_connected_translation(out)
        return trans, rot / torch.norm(rot)

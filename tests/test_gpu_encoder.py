"""-m gpu: tcgen05 implicit-GEMM convolution and the tensor-core encoder against torch
(fp32 math on bf16-rounded operands for the single layer; the fp32 cuDNN model for the stack)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def to_padded_nhwc(x):
    xp = F.pad(F.pad(x, (1, 1, 0, 0), mode="circular"), (0, 0, 1, 1))
    return xp.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


CASES = [  # B, Cin, Cout, H, W, k, stride, act, residual
    (1, 64, 64, 8, 128, 3, (1, 1), 0, False),
    (2, 64, 64, 16, 256, 3, (1, 1), 2, True),
    (2, 64, 128, 16, 256, 3, (1, 2), 2, False),
    (2, 64, 128, 16, 256, 1, (1, 2), 0, False),
    (2, 128, 128, 16, 128, 3, (1, 1), 1, True),
    (2, 256, 512, 16, 128, 3, (2, 2), 2, False),
    (2, 256, 512, 16, 128, 1, (2, 2), 0, False),
    (1, 512, 512, 32, 64, 3, (1, 1), 2, True),
    # KITTI 64x720 widths down the encoder (ragged tiles, odd sizes): 720 -> 360 -> 180 -> 90 -> 45 -> 23
    (1, 64, 64, 64, 720, 3, (1, 2), 2, False),
    (1, 64, 64, 64, 180, 3, (1, 1), 2, True),
    (1, 64, 128, 64, 180, 3, (1, 2), 2, False),
    (1, 128, 256, 64, 90, 1, (1, 2), 0, False),
    (1, 256, 512, 64, 45, 3, (2, 2), 2, False),
    (1, 256, 512, 64, 45, 1, (2, 2), 0, False),
    (2, 512, 512, 32, 23, 3, (1, 1), 1, True),
    (1, 64, 64, 6, 10, 3, (1, 1), 0, False),
]


@pytest.mark.parametrize("b,cin,cout,h,w,k,stride,act,use_res", CASES)
def test_conv_fprop_matches_torch(b, cin, cout, h, w, k, stride, act, use_res, cuda_lib):
    from delora_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(cin * 131 + cout + k)
    x = torch.randn(b, cin, h, w, device=DEV, generator=g) * 0.5
    wt = torch.randn(cout, cin, k, k, device=DEV, generator=g) / (cin * k * k) ** 0.5
    ho, wo = ops.conv_out_size(h, stride[0]), ops.conv_out_size(w, stride[1])
    res = torch.randn(b, cout, ho, wo, device=DEV, generator=g) * 0.5 if use_res else None
    wn = wt.permute(0, 2, 3, 1).reshape(cout, k * k, cin).contiguous().to(torch.bfloat16)
    y = ops.conv2d_fprop(to_padded_nhwc(x), wn, h, w, k, stride, act, to_padded_nhwc(res) if use_res else None)
    got = ops.nhwc_to_nchw(y, ho, wo)
    xb, wb = x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float()
    if k == 3:       # circular W padding + zero H padding: src/models/resnet_modified.py:162-168, :126-129
        ref = F.conv2d(F.pad(xb, (1, 1, 0, 0), mode="circular"), wb, stride=stride, padding=(1, 0))
    else:            # 1x1 downsample: :132-134
        ref = F.conv2d(xb, wb, stride=stride)
    if use_res:
        ref = ref + res.to(torch.bfloat16).float()
    ref = torch.relu(ref) if act == 1 else torch.tanh(ref) if act == 2 else ref
    # fp32 accumulation of exact bf16 products; the output is rounded to bf16 (rel 2^-8)
    assert (got - ref).abs().max().item() <= 6e-3 * max(1.0, ref.abs().max().item())
    yf = y.float()
    assert torch.equal(yf[:, 1:-1, 0], yf[:, 1:-1, wo]) and torch.equal(yf[:, 1:-1, wo + 1], yf[:, 1:-1, 1])
    assert float(yf[:, 0].abs().max()) == 0.0 and float(yf[:, -1].abs().max()) == 0.0


def test_conv_rejects_unsupported_shapes(cuda_lib):
    from delora_b200 import ops
    x = torch.zeros(1, 10, 130, 32, dtype=torch.bfloat16, device=DEV)
    w = torch.zeros(64, 9, 32, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError, match="multiples of 64"):
        ops.conv2d_fprop(x, w, 8, 128, 3, (1, 1))


@pytest.mark.parametrize("h,w", [(64, 512), (64, 720), (16, 180)])
def test_tensor_core_encoder_matches_torch_model(h, w, cuda_lib):
    from delora_b200 import ops, synthetic
    from delora_b200.models.model import OdometryModel
    from delora_b200.models.tc_encoder import TensorCoreEncoder
    b = 2
    cfg = synthetic.fov_config(h=h, w=w, device=DEV)
    cfg.update({"pre_feature_extraction": False, "resnet_outputs": 1000, "use_dropout": False, "layers": [2, 2, 2, 2],
                "factor_fewer_resnet_channels": 1, "activation_fct": "tanh", "use_single_mlp_at_output": False})
    torch.manual_seed(0)
    cfg["use_tensor_core_encoder"] = False                 # the fp32 torch / cuDNN path is the reference here
    model = OdometryModel(cfg).to(DEV).eval()
    enc = TensorCoreEncoder(model)
    g = torch.Generator(device=DEV).manual_seed(1)
    img1 = torch.randn(b, 4, h, w, device=DEV, generator=g) * 5.0
    img2 = torch.randn(b, 4, h, w, device=DEV, generator=g) * 5.0
    with torch.no_grad():
        ref = model.forward_features(image_1=img1, image_2=img2)
        t_ref, q_ref = model(image_1=img1, image_2=img2)
    feats = enc.features(img1, img2)
    for (x, fh, fw), r in zip(feats, ref[:4]):
        got = ops.nhwc_to_nchw(x, fh, fw)
        assert got.shape == r.shape
        cos = F.cosine_similarity(got.flatten(), r.flatten(), dim=0).item()
        assert cos > 0.999, cos                        # bf16 activations through up to 20 layers
    t, q = enc.forward(img1, img2)
    assert (t - t_ref).abs().max().item() < 2e-2 and (q - q_ref).abs().max().item() < 2e-2
    # model-level switch: no_grad routes through the tensor-core encoder
    model.config["use_tensor_core_encoder"] = True
    with torch.no_grad():
        t2, q2 = model(image_1=img1, image_2=img2)
    assert torch.equal(t2, t) and torch.equal(q2, q)
    # the bf16 filter copies follow the fp32 parameters (optimizer steps, load_state_dict): no stale weights
    with torch.no_grad():
        model.resnet.layer3[0].conv1.weight.mul_(0.5)
        model.resnet.conv1.weight.add_(0.01)
        t3, q3 = model(image_1=img1, image_2=img2)
        model.config["use_tensor_core_encoder"] = False
        t3_ref, q3_ref = model(image_1=img1, image_2=img2)
    assert not torch.equal(t3, t2)
    assert (t3 - t3_ref).abs().max().item() < 2e-2 and (q3 - q3_ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("b,cin,cout,h,w,k,stride", [
    (2, 64, 64, 8, 128, 3, (1, 1)), (2, 64, 128, 16, 256, 3, (1, 2)), (2, 64, 128, 16, 256, 1, (1, 2)),
    (2, 256, 512, 16, 128, 3, (2, 2)), (2, 256, 512, 16, 128, 1, (2, 2)), (1, 512, 512, 32, 64, 3, (1, 1)),
    (1, 512, 512, 8, 16, 3, (1, 1)),
    # ragged / odd sizes (64x720 encoder): K tiles overhang the image, odd widths under stride 2
    (1, 64, 64, 64, 180, 3, (1, 1)), (1, 64, 128, 64, 180, 3, (1, 2)), (1, 128, 256, 64, 90, 1, (1, 2)),
    (1, 256, 512, 64, 45, 3, (2, 2)), (1, 256, 512, 64, 45, 1, (2, 2)), (2, 512, 512, 32, 23, 3, (1, 1)),
    (1, 64, 64, 6, 10, 3, (1, 1))])
def test_conv_dgrad_wgrad_match_autograd(b, cin, cout, h, w, k, stride, cuda_lib):
    """Backward of the convolution: wgrad kernel and dgrad (= fprop kernel on the zero-upsampled output
    gradient with the flipped filter) against torch autograd on bf16-rounded operands."""
    from delora_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(cin + 7 * cout + k)
    x = (torch.randn(b, cin, h, w, device=DEV, generator=g) * 0.5).to(torch.bfloat16).float().requires_grad_(True)
    wt = (torch.randn(cout, cin, k, k, device=DEV, generator=g) / (cin * k * k) ** 0.5).to(torch.bfloat16).float()
    wt.requires_grad_(True)
    ho, wo = ops.conv_out_size(h, stride[0]), ops.conv_out_size(w, stride[1])
    dz = (torch.randn(b, cout, ho, wo, device=DEV, generator=g) * 0.5).to(torch.bfloat16).float()
    if k == 3:
        y = F.conv2d(F.pad(x, (1, 1, 0, 0), mode="circular"), wt, stride=stride, padding=(1, 0))
    else:
        y = F.conv2d(x, wt, stride=stride)
    y.backward(dz)
    dw = ops.conv2d_wgrad(to_padded_nhwc(x.detach()), to_padded_nhwc(dz), h, w, k, stride)
    assert (dw - wt.grad).abs().max().item() <= 1e-3 * wt.grad.abs().max().item()      # fp32 accumulate / output
    wf = wt.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(cin, k * k, cout).contiguous().to(torch.bfloat16)
    dzn = to_padded_nhwc(dz)
    if stride != (1, 1):
        dzn = ops.zero_upsample(dzn, ho, wo, stride, out_hw=(h, w))
    dx = ops.nhwc_to_nchw(ops.conv2d_fprop(dzn, wf, h, w, k, (1, 1), ops.ACT_NONE), h, w)
    assert (dx - x.grad).abs().max().item() <= 8e-3 * x.grad.abs().max().item()        # bf16 output


@pytest.mark.parametrize("h,w", [(64, 512), (64, 720)])
def test_encoder_training_path_gradients(h, w, cuda_lib):
    """Full trunk forward + backward on tcgen05 (autograd Function) against torch fp32 autograd."""
    from delora_b200 import synthetic
    from delora_b200.models.model import OdometryModel
    from delora_b200.models.tc_encoder import TensorCoreEncoder
    b = 2
    cfg = synthetic.fov_config(h=h, w=w, device=DEV)
    cfg.update({"pre_feature_extraction": False, "resnet_outputs": 1000, "use_dropout": False, "layers": [2, 2, 2, 2],
                "factor_fewer_resnet_channels": 1, "activation_fct": "tanh", "use_single_mlp_at_output": False})
    torch.manual_seed(0)
    model = OdometryModel(cfg).to(DEV)
    enc = TensorCoreEncoder(model)
    g = torch.Generator(device=DEV).manual_seed(1)
    img1 = torch.randn(b, 4, h, w, device=DEV, generator=g) * 5.0
    img2 = torch.randn(b, 4, h, w, device=DEV, generator=g) * 5.0
    sel = torch.randn(b, 512, device=DEV, generator=g)
    model.zero_grad()
    (model.forward_features(image_1=img1, image_2=img2)[3].mean(dim=(2, 3)) * sel).sum().backward()
    ref = [p.grad.clone() for p in enc.trunk_parameters()]
    model.zero_grad()
    pooled = enc.pooled_features(img1, img2)
    (pooled * sel).sum().backward()
    cosines = [F.cosine_similarity(p.grad.flatten(), r.flatten(), dim=0).item()
               for p, r in zip(enc.trunk_parameters(), ref)]
    assert len(cosines) == 20 and min(cosines[1:]) > 0.999 and cosines[0] > 0.99, cosines   # stem: see test_gpu_sizes.py
    # model-level: training forward with the flag goes through the differentiable tcgen05 trunk
    model.config["use_tensor_core_encoder"] = True
    model.zero_grad()
    t, q = model(image_1=img1, image_2=img2)
    (t.sum() + q.sum()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_flat_gradient_buffer_matches_plain_autograd(cuda_lib):
    """parallel_grad.BucketedGradAllReduce on one process: the weight-gradient kernels write into the flat buffer,
    `finish()` makes every p.grad a view of it -- same values as the plain autograd path, for two consecutive steps
    and with a second forward before the first backward (separate buffer slots)."""
    from delora_b200 import synthetic
    from delora_b200.models.model import OdometryModel
    from delora_b200.parallel_grad import BucketedGradAllReduce
    h, w, b = 16, 256, 2
    cfg = synthetic.fov_config(h=h, w=w, device=DEV)
    cfg.update({"pre_feature_extraction": False, "resnet_outputs": 1000, "use_dropout": False, "layers": [2, 2, 2, 2],
                "factor_fewer_resnet_channels": 1, "activation_fct": "tanh", "use_single_mlp_at_output": False})
    torch.manual_seed(0)
    model = OdometryModel(cfg).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(1)
    img = [torch.randn(b, 4, h, w, device=DEV, generator=g) * 5.0 for _ in range(4)]

    def loss_of(i1, i2):
        t, q = model(image_1=i1, image_2=i2)
        return (t * 0.7).sum() + (q * 1.3).sum()
    model.zero_grad(set_to_none=True)
    loss_of(img[0], img[1]).backward()
    ref = [p.grad.clone() for p in model.parameters()]
    sync = BucketedGradAllReduce(model, encoder=model._tensor_core_path())
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        loss_of(img[0], img[1]).backward()
        sync.finish()
        for p, r in zip(model.parameters(), ref):
            assert p.grad.data_ptr() == sync.views[id(p)].data_ptr()
            assert torch.equal(p.grad, r)
    # two forwards alive at once: the second must not clobber the activations the first backward needs
    model.zero_grad(set_to_none=True)
    l1 = loss_of(img[0], img[1])
    l2 = loss_of(img[2], img[3])
    l1.backward()
    sync.finish()
    for p, r in zip(model.parameters(), ref):
        assert torch.equal(p.grad, r)
    del l2


@pytest.mark.parametrize("h,w,f16,act", [(5, 64, True, 2), (9, 96, False, 1), (4, 1024, True, 2), (5, 48, True, 2),
                                          (3, 90, False, 2)])
def test_maxpool_forward_argmax_and_backward(h, w, f16, act, cuda_lib):
    """The stem's max pool (3x3, stride (1, 2), circular in W, activation applied to the maximum) and its backward
    (tiled kernel for W % 32 == 0, per-pixel kernel otherwise): pooled values against torch's max_pool2d, the stored
    argmax against the input, and dz against a scatter-add of dy through those argmax codes times act'(z)."""
    from delora_b200 import ops
    L = ops._lib.lib()
    b, c, wo = 2, 64, w // 2
    g = torch.Generator(device=DEV).manual_seed(h * 1000 + w)
    z = (torch.randn((b, c, h, w), generator=g, device=DEV) * 1.5)
    z = z.to(torch.float16 if f16 else torch.bfloat16)
    zp = F.pad(F.pad(z.float(), (1, 1, 0, 0), mode="circular"), (0, 0, 1, 1))           # halo as the conv epilogue leaves it
    y0 = zp.permute(0, 2, 3, 1).contiguous().to(z.dtype)
    pooled = ops.padded_nhwc_zeros(b, h, wo, c, DEV)
    idx = torch.empty((b, h, wo, c), dtype=torch.uint8, device=DEV)
    ops._lib.check(L.delora_maxpool_w_idx_nhwc_bf16(y0.data_ptr(), b, h, w, c, pooled.data_ptr(), idx.data_ptr(), act,
                                                    1 if f16 else 0, ops._stream()), "maxpool idx")
    f = torch.tanh if act == 2 else torch.relu
    zinf = F.pad(F.pad(z.float(), (1, 1, 0, 0), mode="circular"), (0, 0, 1, 1), value=float("-inf"))
    ref = F.max_pool2d(zinf, 3, stride=(1, 2))[..., :wo]
    got = pooled[:, 1:h + 1, 1:wo + 1].permute(0, 3, 1, 2).float()
    assert torch.allclose(got, f(ref), rtol=2 ** -7, atol=0)          # bf16 rounding of the device tanh: <= 1 ulp
    # the argmax code dr * 3 + dq points at an element equal to the window maximum
    code = idx.permute(0, 3, 1, 2).long()
    assert int(code.max()) <= 8
    win = zinf.unfold(2, 3, 1).unfold(3, 3, 2)[:, :, :, :wo].reshape(b, c, h, wo, 9)
    assert torch.equal(win.gather(4, code.unsqueeze(-1)).squeeze(-1), ref)

    dy = torch.randn((b, h, wo, c), generator=g, device=DEV).to(torch.bfloat16)
    dyp = ops.padded_nhwc_zeros(b, h, wo, c, DEV)
    dyp[:, 1:h + 1, 1:wo + 1] = dy
    dyp[:, :, 0] = dyp[:, :, wo]
    dyp[:, :, wo + 1] = dyp[:, :, 1]
    dz = torch.full((b, h + 2, w + 2, c), 7.0, dtype=torch.bfloat16, device=DEV)
    ops._lib.check(L.delora_maxpool_w_bwd_nhwc_bf16(dyp.data_ptr(), idx.data_ptr(), y0.data_ptr(), b, h, w, c, 4 + act,
                                                    dz.data_ptr(), 1 if f16 else 0, ops._stream()), "maxpool bwd")
    acc = torch.zeros((b, h + 2, w + 2, c), dtype=torch.float32, device=DEV)
    for dr in range(3):
        for dq in range(3):
            acc[:, dr:dr + h, dq:dq + 2 * wo:2] += dy.float() * (idx == dr * 3 + dq)
    acc[:, :, w] += acc[:, :, 0]
    acc[:, :, 1] += acc[:, :, w + 1]
    zf = z.float().permute(0, 2, 3, 1)
    dact = (1.0 - torch.tanh(zf) ** 2) if act == 2 else (zf > 0).float()
    want = acc[:, 1:h + 1, 1:w + 1] * dact
    out = dz[:, 1:h + 1, 1:w + 1].float()
    assert torch.allclose(out, want, rtol=2 ** -7, atol=1e-6)
    assert torch.equal(dz[:, 1:h + 1, 0], dz[:, 1:h + 1, w]) and torch.equal(dz[:, 1:h + 1, w + 1], dz[:, 1:h + 1, 1])


def test_weight_prep_multi_matches_per_layer_prep(cuda_lib):
    """One launch for all filters (tiled transposes through shared memory for channel counts that are multiples of 32,
    element-wise otherwise) against the single-layer kernel: both bf16 layouts bit for bit, 3x3 and 1x1, padded Cin."""
    from delora_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(5)
    shapes = [(64, 64, 3, 64), (128, 64, 3, 64), (128, 64, 1, 64), (512, 256, 3, 256), (512, 512, 3, 512), (64, 8, 3, 64),
              (96, 32, 3, 32), (48, 24, 1, 64)]
    rows, keep = [], []
    for (cout, cin, k, cin_pad) in shapes:
        w = torch.randn((cout, cin, k, k), generator=g, device=DEV)
        fwd = torch.full((cout, k * k, cin_pad), 3.0, dtype=torch.bfloat16, device=DEV)
        flip = torch.full((cin, k * k, cout), 3.0, dtype=torch.bfloat16, device=DEV) if k == 3 else None
        ref_fwd, ref_flip = torch.empty_like(fwd), (torch.empty_like(flip) if flip is not None else None)
        ops.conv_weight_prep(w, ref_fwd, ref_flip, cin_pad)
        rows.append([w.data_ptr(), fwd.data_ptr(), flip.data_ptr() if flip is not None else 0, cout, cin, k, cin_pad, 0])
        keep.append((w, fwd, flip, ref_fwd, ref_flip))
    table = torch.tensor(rows, dtype=torch.int64, device=DEV)
    ops.conv_weight_prep_multi(table, len(rows))
    for (w, fwd, flip, ref_fwd, ref_flip) in keep:
        assert torch.equal(fwd, ref_fwd), tuple(w.shape)
        assert torch.equal(fwd[:, :, :w.shape[1]].float(),
                           w.to(torch.bfloat16).float().permute(0, 2, 3, 1).reshape(fwd.shape[0], -1, w.shape[1]))
        if flip is not None:
            assert torch.equal(flip, ref_flip), tuple(w.shape)


@pytest.mark.parametrize("b,h,w,c", [(2, 16, 64, 512), (1, 5, 23, 64), (3, 1, 1, 128)])
def test_avgpool_matches_torch_mean(b, h, w, c, cuda_lib):
    """AdaptiveAvgPool2d((1,1)) on the padded NHWC bf16 map (resnet_modified.py:111): fp32 mean of the bf16 values;
    the halo must not leak into the sum."""
    from delora_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(b * 100 + w)
    x = torch.randn((b, h + 2, w + 2, c), generator=g, device=DEV).to(torch.bfloat16)       # halo filled with junk
    got = ops.avgpool(x, h, w)
    want = x[:, 1:h + 1, 1:w + 1].double().mean(dim=(1, 2))
    assert got.shape == (b, c) and got.dtype == torch.float32
    assert torch.allclose(got.double(), want, rtol=1e-5, atol=1e-6)
    with pytest.raises(RuntimeError):
        ops.avgpool(torch.zeros((1, 3, 3, 40), dtype=torch.bfloat16, device=DEV), 1, 1)

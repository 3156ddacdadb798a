"""Drop-in `models.model.OdometryModel` (reference: src/models/model.py): the two 4-channel range
images are concatenated to 8 channels, encoded by `ResNetModified`, and two small MLP heads
regress the quaternion (x,y,z,w) and the translation.  Parameter names match the reference
(`resnet.*`, `fully_connected_rotation.{1,3}.*`, `fully_connected_translation.{1,3}.*`)."""
import torch

from . import model_parts, resnet_modified


def _act(name):
    return torch.nn.ReLU() if name == "relu" else torch.nn.Tanh()


class OdometryModel(torch.nn.Module):

    def __init__(self, config):
        super().__init__()
        self.device = config["device"]
        self.config = config
        act = config["activation_fct"]
        if act not in ("relu", "tanh"):
            raise Exception('The specified activation function must be either "relu" or "tanh".')
        self.pre_feature_extraction = config["pre_feature_extraction"]
        in_channels, n_fe = 8, 5
        if self.pre_feature_extraction:                                     # src/models/model.py:30-45
            mods = []
            for li in range(n_fe):
                cin = in_channels // 2 if li == 0 else li * in_channels
                mods += [model_parts.CircularPad((1, 1, 0, 0)),
                         torch.nn.Conv2d(cin, (li + 1) * in_channels, kernel_size=3, padding=(1, 0), bias=False),
                         torch.nn.ReLU(inplace=True) if act == "relu" else torch.nn.Tanh()]
            self.feature_extractor = torch.nn.Sequential(*mods)
        self.resnet = resnet_modified.ResNetModified(
            in_channels=in_channels if not self.pre_feature_extraction else 2 * n_fe * in_channels,
            num_outputs=config["resnet_outputs"], use_dropout=config["use_dropout"], layers=config["layers"],
            factor_fewer_resnet_channels=config["factor_fewer_resnet_channels"], activation_fct=act)
        nout = config["resnet_outputs"]
        if config["use_single_mlp_at_output"]:                              # :58-71
            dims = [nout, 512, 512, 256, 64, 7]
            mods = []
            for a, b in zip(dims[:-1], dims[1:]):
                mods += [_act(act), torch.nn.Linear(a, b)]
            self.fully_connected_rot_trans = torch.nn.Sequential(*mods)
        else:                                                               # :73-84
            self.fully_connected_rotation = torch.nn.Sequential(_act(act), torch.nn.Linear(nout, 100), _act(act),
                                                                torch.nn.Linear(100, 4))
            self.fully_connected_translation = torch.nn.Sequential(_act(act), torch.nn.Linear(nout, 100), _act(act),
                                                                   torch.nn.Linear(100, 3))
        self.geometry_handler = model_parts.GeometryHandler(config=config)

    def forward_features(self, image_1, image_2):
        if self.pre_feature_extraction:
            x = torch.cat((self.feature_extractor(image_1), self.feature_extractor(image_2)), dim=1)
        else:
            x = torch.cat((image_1, image_2), dim=1)
        return self.resnet(x)

    def _tensor_core_path(self):
        """The encoder trunk on the tcgen05 kernels (`models/tc_encoder.py`): bf16 NHWC activations,
        forward and (autograd) backward -- fprop, dgrad and wgrad all on the tensor cores.
        DEFAULT whenever the model is eligible: parameters on a CUDA device, the full-width model
        (factor_fewer_resnet_channels == 1, i.e. channel counts that are multiples of 64), no dropout, no
        pre-feature extractor.  config["use_tensor_core_encoder"] = False opts out (torch / cuDNN convolutions,
        the reference's `nn.Conv2d` path)."""
        if not self.config.get("use_tensor_core_encoder", True):
            return None
        if self.pre_feature_extraction or self.config["factor_fewer_resnet_channels"] != 1 or self.config["use_dropout"]:
            return None
        if not self.resnet.conv1.weight.is_cuda:
            return None
        if getattr(self, "_tc_encoder", None) is None:
            from . import tc_encoder
            self._tc_encoder = tc_encoder.TensorCoreEncoder(self)
        return self._tc_encoder

    def forward(self, image_1, image_2):
        tc = self._tensor_core_path()
        if tc is not None and not torch.is_grad_enabled():
            return tc.forward(image_1.contiguous(), image_2.contiguous())
        if tc is not None:
            x = self.resnet.fc(tc.pooled_features(image_1, image_2))        # differentiable tcgen05 trunk
        else:
            x = self.forward_features(image_1=image_1, image_2=image_2)[-1]
        if self.config["use_single_mlp_at_output"]:
            x = self.fully_connected_rot_trans(x)
            x_rotation, x_translation = x[:, :4], x[:, 4:]
        else:
            x_rotation = self.fully_connected_rotation(x)
            x_translation = self.fully_connected_translation(x)
        # the reference divides by the norm of the WHOLE [B,4] tensor (src/models/model.py:114);
        # the per-row normalisation happens inside the quaternion -> R conversion
        x_rotation = x_rotation / torch.norm(x_rotation)
        return (x_translation, x_rotation)

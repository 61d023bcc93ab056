"""Drop-in shell of the reference's stage-1 network (分割/nets/basicUnet.py; byte-identical copy at
分类/nets/basicUnet_new.py).

`UNetTaskAligWeight(n_channels=3, n_classes=1)` keeps the reference's constructor, attribute tree and the 287
state_dict keys (including the tensors the reference forward never reads: fc1/fc2, cca.fc_soft,
cca.deformabel.*, task2...cross_attention_seg.*), so `load_state_dict(torch.load(p)['net'])` works unchanged.
`forward(x)` (basicUnet.py:406-437) is executed by the ugnet engine: fp32 NCHW [B,3,224,224] -> fp32 logits
[B,1,224,224].  There is no PyTorch fallback: a CPU tensor or train mode raises."""
import torch
import torch.nn as nn

from .deform_conv_v2 import DeformConv2d
from .tasks import TransformerDecoder


class ConvBatchNorm(nn.Module):  # basicUnet.py:25-40
    def __init__(self, in_channels, out_channels, activation="ReLU"):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)
        self.norm = nn.BatchNorm2d(out_channels)
        self.activation = nn.ReLU()


def _make_nConv(in_channels, out_channels, nb_Conv, activation="ReLU"):  # basicUnet.py:17-23
    chans = [in_channels] + [out_channels] * nb_Conv
    return nn.Sequential(*[ConvBatchNorm(a, b, activation) for a, b in zip(chans[:-1], chans[1:])])


class DownBlock(nn.Module):  # basicUnet.py:42-52
    def __init__(self, in_channels, out_channels, nb_Conv, activation="ReLU"):
        super().__init__()
        self.maxpool = nn.MaxPool2d(2)
        self.nConvs = _make_nConv(in_channels, out_channels, nb_Conv, activation)


class CoordAtt3(nn.Module):  # basicUnet.py:201-231
    def __init__(self, inp):
        super().__init__()
        self.conv1_e = _make_nConv(inp, inp, 1, "ReLU")
        self.conv2_e = _make_nConv(inp, inp, 1, "ReLU")
        self.avgpool_e = nn.AdaptiveAvgPool2d((1, 1))
        self.maxpool_e = nn.AdaptiveMaxPool2d((1, 1))
        self.fc_avg = nn.Conv2d(inp, inp // 2, kernel_size=1)
        self.fc_max = nn.Conv2d(inp, inp // 2, kernel_size=1)
        self.fc_soft = nn.Conv2d(inp, inp // 2, kernel_size=1)
        self.fc_avg_max_sfot = nn.Conv2d(inp // 2, inp, kernel_size=1)
        self.deformabel = DeformConv2d(in_channels=inp, out_channels=inp, kernel_size=3)


class UpBlockAlig(nn.Module):  # basicUnet.py:115-128
    def __init__(self, in_channels, out_channels, nb_Conv, activation="ReLU"):
        super().__init__()
        self.up = nn.ConvTranspose2d(in_channels // 2, in_channels // 2, (2, 2), 2)
        self.nConvs = _make_nConv(in_channels, out_channels, nb_Conv, activation)
        self.cca = CoordAtt3(in_channels // 2)


class UNetTaskAligWeight(nn.Module):  # basicUnet.py:369-437
    IMG_SIZE = 224  # locked by the learned 14x14 positional embedding (tasks.py:212-217)

    def __init__(self, n_channels=3, n_classes=9):
        super().__init__()
        self.n_channels, self.n_classes = n_channels, n_classes
        c = 64
        self.inc = ConvBatchNorm(n_channels, c)
        self.down1 = DownBlock(c, c * 2, nb_Conv=2)
        self.down2 = DownBlock(c * 2, c * 4, nb_Conv=2)
        self.down3 = DownBlock(c * 4, c * 8, nb_Conv=2)
        self.down4 = DownBlock(c * 8, c * 8, nb_Conv=2)
        self.up4 = UpBlockAlig(c * 16, c * 4, nb_Conv=2)
        self.up3 = UpBlockAlig(c * 8, c * 2, nb_Conv=2)
        self.up2 = UpBlockAlig(c * 4, c, nb_Conv=2)
        self.up1 = UpBlockAlig(c * 2, c, nb_Conv=2)
        self.outc = nn.Conv2d(c, n_classes, kernel_size=(1, 1))
        self.avgpool2 = nn.AdaptiveAvgPool2d((1, 1))
        self.task2 = TransformerDecoder(dim=c * 8, depth=1, heads=8, dim_head=64, mlp_dim=2048, dropout=0,
                                        decoder_pos_size=14, softmax=True)
        self.fc1 = nn.Linear(c * 8, c * 4)
        self.fc2 = nn.Linear(c * 4, 1)
        self._runner = None
        self._register_load_state_dict_pre_hook(self._drop_runner)

    def _drop_runner(self, *args, **kwargs):
        self._runner = None

    def invalidate(self):
        """Forget the packed engine copy of the weights (it is rebuilt on the next forward)."""
        self._runner = None

    def _fingerprint(self):
        # (storage address, in-place version counter) of every parameter / buffer: changes on load_state_dict, .to(),
        # optimizer steps, p.copy_() / p.mul_() under no_grad, and on assigning a new Parameter or submodule.  Writes
        # through `p.data` bypass the version counter: call invalidate() after those.
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_runner"] = None          # the engine handle / device workspaces are per process, never pickled
        state.pop("_runner_key", None)
        return state

    def __deepcopy__(self, memo):
        import copy
        runner, self._runner = self._runner, None
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            new.__dict__ = copy.deepcopy({k: v for k, v in self.__dict__.items() if k != "_runner_key"}, memo)
        finally:
            self._runner = runner
        return new

    def _apply(self, fn, *args, **kwargs):
        self._runner = None  # .to()/.cuda() move parameters: re-pack lazily
        return super()._apply(fn, *args, **kwargs)

    HEAD = "seg"   # "cls" in the classifier-head variant (nets/basicUnet_cls.py)

    def runner(self):
        """The engine-side compiled network (packed weights + per-batch programs)."""
        key = self._fingerprint()
        if self._runner is not None and getattr(self, "_runner_key", None) != key:
            self._runner = None          # weights were edited in place since the last pack
        if self._runner is None:
            from ..lower import UNetRunner
            dev = self.outc.weight.device
            if dev.type != "cuda":
                raise RuntimeError("UNetTaskAligWeight runs on the ugnet CUDA engine only: call .to('cuda') "
                                   "(there is no CPU path)")
            if self.n_channels != 3 or (self.HEAD == "seg" and self.n_classes != 1):
                raise NotImplementedError("the engine lowers UNetTaskAligWeight(3, 1), the configuration every "
                                          "reference entry point instantiates")
            self._runner = UNetRunner(self.state_dict(), dev, head=self.HEAD)
            self._runner_key = key
        return self._runner

    def forward(self, x):
        if self.training:
            raise RuntimeError("the ugnet engine is inference-only: call model.eval() (predict.py:12, roi.py:18)")
        return self.runner().forward(x)

    @torch.no_grad()
    def forward_mask_boxes(self, x, padding=30):
        """Batched stage-1 result: (logits f32 [B,1,224,224], mask u8 [B,224,224], boxes i32 [B,4])."""
        return self.runner().forward(x, with_mask_boxes=True, padding=padding)

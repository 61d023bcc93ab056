"""Shell of the classifier-head `UNetTaskAligWeight` of 分类/nets/basicUnet.py:369-436.

That file defines a class with the SAME name, constructor and state_dict as the segmentation network
(分割/nets/basicUnet.py, mirrored by `nets/basicUnet.py` here) but a different forward: encoder -> TransformerDecoder,
the `x` (cl) token stream -> AdaptiveAvgPool2d(1) -> fc1 -> fc2, returning `cl_out` [B, 1]; the decoder (`up*`) and
`outc` are built and loaded but never executed (:422-435).  No inference script of the reference imports it (they all
import `basicUnet_new`), it is SURVEY.md §8(f) rank 4: the same kernels under a different liveness mask.
"""
import torch

from .basicUnet import UNetTaskAligWeight as _SegShell


class UNetTaskAligWeight(_SegShell):  # 分类/nets/basicUnet.py:369
    HEAD = "cls"   # UNetRunner(head="cls"): encoder + x token stream + avgpool2/fc1/fc2 (lower._emit_cls)

    def forward(self, x):
        """-> cl_out fp32 [B, 1] (分类/nets/basicUnet.py:432-436)."""
        if self.training:
            raise RuntimeError("the ugnet engine is inference-only: call model.eval()")
        return self.runner().forward(x)

    @torch.no_grad()
    def forward_mask_boxes(self, x, padding=30):
        raise RuntimeError("the classifier-head variant produces no mask (its decoder is dead code, :422-430)")

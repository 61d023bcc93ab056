"""The batched two-stage path as one call: `pipeline(imgs) -> (masks, boxes, logits)`.

Replaces the reference's per-image loop (分类/test.py:122-134 calling roi.py:12-51 inside the Dataset, then
test.py:81-86) with one engine program per micro-batch: no host round trip between the UNet and GoogLeNet."""
import torch

from .lower import PipelineRunner


class TwoStagePipeline:
    def __init__(self, unet, classifier, micro_batch=64, padding=30, cls_batch=256):
        """unet: nets.UNetTaskAligWeight (or a reference-format state_dict); classifier: GoogLeNetClassifier
        (or its state_dict).  Both must live on the same CUDA device."""
        usd = unet if isinstance(unet, dict) else unet.state_dict()
        gsd = classifier if isinstance(classifier, dict) else classifier.state_dict()
        dev = next(iter(usd.values())).device if isinstance(unet, dict) else unet.outc.weight.device
        if torch.device(dev).type != "cuda":
            dev = torch.device("cuda", torch.cuda.current_device())
        self.runner = PipelineRunner(usd, gsd, dev, micro_batch=micro_batch, padding=padding, cls_batch=cls_batch)

    @torch.no_grad()
    def __call__(self, imgs, return_logits=False):
        """imgs: float [B,3,224,224] on the pipeline's device, or uint8 HWC sources [B,Hs,Ws,3] of any size (resized
        on the device like the reference's CDDataAugmentation.transform: PIL bilinear + to_tensor) ->
        (masks uint8 [B,224,224], boxes int32 [B,4] = (x0,y0,x1,y1), class logits float32 [B,6])."""
        return self.runner(imgs, return_logits=return_logits)

    def predict(self, imgs):
        """argmax(softmax(logits)) as 分类/test.py:86 (softmax is monotone, so argmax of the logits)."""
        return torch.argmax(self(imgs)[2], dim=1)

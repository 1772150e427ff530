"""Fan-out of the encoder / predictor over frames-per-clip groups and masks -- the call protocol of the reference's
src/utils/wrappers.py (MultiSeqWrapper :9-27, PredictorMultiSeqWrapper :30-43): inputs are lists over groups (and,
inside a group, lists over masks); outputs mirror that nesting.  `backbone` keeps its attribute name because
checkpoints carry it as a key prefix (`module.backbone.*`) and callers read `encoder.backbone.embed_dim`.

(The fused training step does not go through these modules: `JepaTrainStep` hands all masks of a group to the
engine at once so they share every LayerNorm / GEMM launch.  They serve the autograd drop-in path.)"""
import torch.nn as nn


class _GroupFanOut(nn.Module):
    def __init__(self, backbone):
        super().__init__()
        self.backbone = backbone

    @staticmethod
    def _each_group(groups, fn):
        return [fn(i, g) for i, g in enumerate(groups)]


class MultiSeqWrapper(_GroupFanOut):
    def forward(self, x, masks=None):
        if masks is None:
            return self._each_group(x, lambda i, clip: self.backbone(clip))
        return self._each_group(x, lambda i, clip: [self.backbone(clip, masks=m) for m in masks[i]])


class PredictorMultiSeqWrapper(_GroupFanOut):
    def forward(self, x, masks_x, masks_y, has_cls=False):
        def one_group(i, ctx_tokens):
            return [self.backbone(z, mx, my, mask_index=i, has_cls=has_cls)
                    for z, mx, my in zip(ctx_tokens, masks_x[i], masks_y[i])]
        return self._each_group(x, one_group)

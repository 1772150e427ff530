"""vjepa2_b200: B200-native (sm_100a) implementation of the V-JEPA 2 video-encoder pre-training step.

Public API mirrors the reference (weipeilun/vjepa2):
  vision_transformer.VisionTransformer / vit_large / vit_giant_xformers ...   (src/models/vision_transformer.py)
  predictor.VisionTransformerPredictor / vit_predictor                         (src/models/predictor.py)
  masks.apply_masks / MaskCollator                                             (src/masks/*)
  wrappers.MultiSeqWrapper / PredictorMultiSeqWrapper                          (src/utils/wrappers.py)
  train.init_video_model / JepaTrainStep                                       (app/vjepa/utils.py, app/vjepa/train.py)
Every device op is a hand-written CUDA kernel in libvjepa2_b200.so (include/vjepa2_b200.h); there is no
CPU or PyTorch-op fallback.
"""
from . import _cabi  # noqa: F401


def library_path():
    return _cabi.LIB_PATH

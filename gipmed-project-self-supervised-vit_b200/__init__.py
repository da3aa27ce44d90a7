"""b200ssl — B200-native (sm_100a) hot path of the GipMed self-supervised ViT: the DINO-style training
step (ViT encoder, projection head, centred cross-entropy over multi-crop views, teacher EMA, fused
optimiser, bucketed data-parallel all-reduce) behind the reference's own nn.Module / loss / EMA call
conventions. Import as ``import b200ssl`` (shim at the repo root; this directory's name is not a valid
Python identifier).

The compute path is the in-tree C-ABI library ``libb200ssl.so`` (hand-written CUDA, include/b200ssl.h).
There is no CPU, PyTorch-math or Triton fallback: ops raise if the library or an sm_100 GPU is missing.
"""
from . import _lib, ops  # noqa: F401
from .augment import MultiCropAugment  # noqa: F401
from .dino import (DINOLoss, FusedAdamW, GradBucketDataParallel, GraphedDinoStep, ModelEma,  # noqa: F401
                   MultiCropWrapper, apply_schedules, cancel_gradients_last_layer, cosine_momentum,
                   cosine_scheduler, dino_step, param_groups_wd)
from .features import embed_tiles, pack_mil_inference_file, save_slide_features  # noqa: F401
from .vision_transformer import (Attention, Block, DINOHead, DropPath, Mlp, PatchEmbed,  # noqa: F401
                                 VisionTransformer, drop_path, trunc_normal_, vit_base, vit_small, vit_tiny)

__version__ = "0.1.0"

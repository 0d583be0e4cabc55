""" deepcv_b200 — the DeepcvModule conv / BatchNorm / augment hot path of PaulEmmanuelSotir/DeepCV, rebuilt for B200 (sm_100a).

Package layout mirrors the reference's `deepcv` package for the modules on that path only:
  deepcv_b200.meta.base_module / nn_spec / submodule_creators / nn / hyperparams   — YAML architecture -> DeepcvModule (boundary kept verbatim)
  deepcv_b200.meta.data.preprocess                                                 — preprocess-recipe API + the fused uint8 transform
  deepcv_b200.meta.ignite_training                                                 — the training step and its data-parallel wrap
  deepcv_b200.ops / _lib / csrc                                                    — autograd operators over the C ABI (include/deepcv_b200.h) and its CUDA kernels
"""
__version__ = '0.1.0'

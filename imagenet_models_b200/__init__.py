"""imagenet_models_b200: B200-native (sm_100a) hot path of Lab-LVM/imagenet-models.

Drop-in nn.Modules (same timm registry names, attribute names and state_dict keys as GA/ga_convnext.py) whose
arithmetic runs in hand-written CUDA kernels reached through the C ABI of libga_sm100.so (include/ga_sm100.h).
"""
__version__ = '0.1.0'

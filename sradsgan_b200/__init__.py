"""sradsgan_b200 — B200-native implementation of the SRADSGAN training / inference hot path.

    from sradsgan_b200.model.sradsgan import GeneratorResNet, ResGroup, Discriminator, SRADSGAN

mirrors `from model.sradsgan import ...` of the reference.  All arithmetic of the path runs in
libsradsgan_b200.so (hand-written sm_100a CUDA behind the C ABI of include/sradsgan_b200.h).
"""
__version__ = "0.1.0"

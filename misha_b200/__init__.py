"""misha_b200 -- B200-native track extraction and aggregation behind misha's own interfaces.

`misha_b200.gpu`   object layer over the C ABI (include/mishagpu.h -> libmishagpu.so, hand-written sm_100a kernels)
`misha_b200.synth` seeded synthetic hg38-scale inputs in misha's on-disk formats
There is no CPU fallback: importing works anywhere, but every operation needs the CUDA library and a GPU.
"""
__version__ = "0.1.0"

"""histogan_b200 -- B200-native (sm_100a) kernels for the HistoGAN training hot
path behind the reference's own Python class API.

    from histogan_b200 import RGBuvHistBlock          # histogram_classes/RGBuvHistBlock.py
    from histogan_b200 import hellinger_loss          # histoGAN/histoGAN.py:957-960

The arithmetic lives in ``lib/libhistogan_b200.so`` (C ABI: include/histogan_b200.h),
built from ``csrc/*.cu`` by ``python -m histogan_b200.build``.
"""
from .hist import (RGBuvHistBlock, rgChromaHistBlock, LabHistBlock, hellinger_loss, hist_preprocess,  # noqa: F401
                   device_logf)

__all__ = ["RGBuvHistBlock", "rgChromaHistBlock", "LabHistBlock", "hellinger_loss"]

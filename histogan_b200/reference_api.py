"""The names the reference's scripts import from ``histoGAN`` (histoGAN/__init__.py:1-3),
served by the sm_100a implementation.  ``sys.modules["histoGAN"] = histogan_b200.reference_api``
lets the reference's CLI run on top of it unchanged."""
from .gan import (Conv2DMod, Discriminator, DiscriminatorBlock, Generator, GeneratorBlock,  # noqa: F401
                  HistVectorizer, RGBBlock, StyleVectorizer)
from .hist import RGBuvHistBlock  # noqa: F401
from .trainer import HistoGAN, NanException, Trainer  # noqa: F401

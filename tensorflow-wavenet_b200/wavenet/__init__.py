"""B200-native drop-in for the `wavenet` package of jyegerlehner/tensorflow-wavenet
(reference wavenet/__init__.py:1-4 re-exports exactly these names)."""
from .model import WaveNetModel
from .audio_reader import AudioReader
from .ops import (mu_law_encode, mu_law_decode, time_to_batch,
                  batch_to_time, causal_conv, optimizer_factory)
from .train_step import TrainStep
from . import checkpoint, generation
from .audio_reader import Coordinator

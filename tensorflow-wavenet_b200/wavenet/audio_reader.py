"""Placeholder for the reference's AudioReader (wavenet/audio_reader.py:83-193).

The data-input layer is outside the accelerated hot path ("next" row f2: it
needs librosa, which is not available, and the benchmarks use synthetic audio).  The name is
exported so that `from wavenet import AudioReader` keeps working; constructing it says so."""


class AudioReader(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            'AudioReader (wav loading / silence trimming / queueing) is not part of this build: feed '
            'float audio batches [B, T] straight to WaveNetModel.loss() (SURVEY section 8, row f2)')

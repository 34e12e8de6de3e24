r"""Host-side mirror of wavenet/audio_reader.py of the reference: the same module-level
helpers and the same `AudioReader` constructor / `dequeue` / `dequeue_gc` / `start_threads` surface, without TensorFlow
or librosa.

What replaces what:
  * `librosa.load(filename, sr, mono=True)`      -> `load_wav`: scipy.io.wavfile + channel mean + polyphase resampling
  * `librosa.feature.rmse` / `frames_to_samples`  -> `_rms_frames` (frame 2048, hop 512, centred, reflect padding: the
                                                    librosa defaults audio_reader.py:62-69 relies on)
  * `tf.PaddingFIFOQueue` + `sess.run(enqueue)`   -> a bounded `queue.Queue` fed by daemon threads; `dequeue(n)` pads the
                                                    n pieces with zeros to the longest one ([n, T_max, 1]), exactly what
                                                    `dequeue_many` of a padding queue returns (audio_reader.py:103-105,160-162)
  * `tf.train.Coordinator`                        -> `Coordinator` (request_stop / should_stop)

Semantics kept from the reference: files are drawn WITH replacement (randomize_files, :28-31); with `sample_size` set a
file is cut into pieces of that size and the short tail piece is enqueued too, nothing carries over into the next file
(:167-178); the VCTK speaker id is the first group of r'p([0-9]+)_([0-9]+)\.wav' (:13,45); the category cardinality is the
largest id + 1 (:126-134).  Known reference bugs NOT copied: the `categeory_id` typo on the
no-`sample_size` path, the py3 `None` comparisons in get_category_cardinality, `findall(...) is None` never being true.
"""
import fnmatch
import os
import queue
import random
import re
import threading

import numpy as np

_ID_RE = re.compile(r'p([0-9]+)_([0-9]+)\.wav')


def get_category_cardinality(files):
    """(min id, max id) over the file names (audio_reader.py:12-24)."""
    min_id = None
    max_id = None
    for filename in files:
        matches = _ID_RE.findall(filename)[0]
        id_ = int(matches[0])
        if min_id is None or id_ < min_id:
            min_id = id_
        if max_id is None or id_ > max_id:
            max_id = id_
    return min_id, max_id


def randomize_files(files):
    """Yields len(files) file names drawn uniformly WITH replacement (audio_reader.py:27-30)."""
    for _ in files:
        yield files[random.randint(0, len(files) - 1)]


def find_files(directory, pattern='*.wav'):
    '''Recursively finds all files matching the pattern.'''
    files = []
    for root, _, filenames in os.walk(directory):
        for filename in fnmatch.filter(filenames, pattern):
            files.append(os.path.join(root, filename))
    return files


def load_wav(filename, sample_rate):
    """float32 mono waveform in [-1, 1] at `sample_rate` (the librosa.load(..., sr, mono=True) of the reference)."""
    from scipy.io import wavfile
    rate, data = wavfile.read(filename)
    if data.dtype == np.uint8:
        audio = (data.astype(np.float32) - 128.0) / 128.0
    elif np.issubdtype(data.dtype, np.integer):
        audio = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
    else:
        audio = data.astype(np.float32)
    if audio.ndim > 1:
        audio = audio.mean(axis=1)
    if sample_rate is not None and rate != sample_rate:
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(int(sample_rate), int(rate))
        audio = resample_poly(audio, int(sample_rate) // g, int(rate) // g).astype(np.float32)
    return np.ascontiguousarray(audio, dtype=np.float32)


def load_generic_audio(directory, sample_rate):
    '''Generator that yields (audio [T, 1], filename, category id or None) (audio_reader.py:43-60).'''
    files = find_files(directory)
    for filename in randomize_files(files):
        ids = _ID_RE.findall(filename)
        category_id = int(ids[0][0]) if ids else None
        audio = load_wav(filename, sample_rate).reshape(-1, 1)
        yield audio, filename, category_id


def _rms_frames(audio, frame_length=2048, hop_length=512):
    """librosa.feature.rmse(y) with its defaults: centred frames (reflect padding), one RMS value per hop."""
    y = np.asarray(audio, dtype=np.float32).reshape(-1)
    if y.size == 0:
        return np.zeros((0,), np.float32)
    pad = frame_length // 2
    if y.size > 1:
        yp = np.pad(y, pad, mode='reflect') if y.size > pad else np.pad(y, pad, mode='symmetric')
    else:
        yp = np.pad(y, pad, mode='constant')
    n_frames = 1 + (yp.size - frame_length) // hop_length
    sq = np.concatenate([[0.0], np.cumsum(yp.astype(np.float64) ** 2)])
    starts = np.arange(n_frames) * hop_length
    power = (sq[starts + frame_length] - sq[starts]) / frame_length
    return np.sqrt(np.maximum(power, 0.0)).astype(np.float32)


def trim_silence(audio, threshold, frame_length=2048, hop_length=512):
    '''Removes silence at the beginning and end of a sample (audio_reader.py:62-69).'''
    audio = np.asarray(audio)
    energy = _rms_frames(audio, frame_length, hop_length)
    frames = np.nonzero(energy > threshold)[0]
    indices = frames * hop_length                      # librosa.core.frames_to_samples
    # Note: indices can be an empty array, if the whole audio was silence.
    return audio[indices[0]:indices[-1]] if indices.size else audio[0:0]


def not_all_have_id(files):
    '''True iff any of the file names does not carry a category id (audio_reader.py:72-80, with the intended test).'''
    return any(not _ID_RE.findall(f) for f in files)


class Coordinator(object):
    """Minimal stand-in for tf.train.Coordinator (train.py:207, audio_reader.py:158)."""

    def __init__(self):
        self._stop = threading.Event()

    def request_stop(self):
        self._stop.set()

    def should_stop(self):
        return self._stop.is_set()

    def join(self, threads, timeout=5.0):
        for t in threads:
            t.join(timeout)


class AudioReader(object):
    '''Generic background audio reader that preprocesses audio files and enqueues them into a bounded queue
    (audio_reader.py:83-193).'''

    def __init__(self, audio_dir, coord, sample_rate, gc_enabled, sample_size=None, silence_threshold=None,
                 queue_size=32):
        self.audio_dir = audio_dir
        self.sample_rate = sample_rate
        self.coord = coord if coord is not None else Coordinator()
        self.sample_size = sample_size
        self.silence_threshold = silence_threshold
        self.gc_enabled = gc_enabled
        self.threads = []
        self.queue = queue.Queue(maxsize=queue_size)       # (piece [t, 1] float32, category id or None)
        files = find_files(audio_dir)
        if not files:
            raise ValueError("No audio files found in '{}'.".format(audio_dir))
        if self.gc_enabled and not_all_have_id(files):
            raise ValueError("Global conditioning is enabled, but file names "
                             "do not conform to pattern having id.")
        if self.gc_enabled:
            # largest id + 1: ids of the file names index the embedding table directly (audio_reader.py:126-136)
            _, self.gc_category_cardinality = get_category_cardinality(files)
            self.gc_category_cardinality += 1
            print("Detected --gc_cardinality={}".format(self.gc_category_cardinality))
        else:
            self.gc_category_cardinality = None
        self._pending_ids = []

    # ---- consumer side -------------------------------------------------------------------------
    def _get(self, timeout):
        while True:
            try:
                return self.queue.get(timeout=0.1 if timeout is None else timeout)
            except queue.Empty:
                if timeout is not None or (self.coord.should_stop() and self.queue.empty()):
                    raise

    def dequeue(self, num_elements, timeout=None, pin_memory=False):
        """`num_elements` pieces as one float32 array [num_elements, T_max, 1], zero padded like tf.PaddingFIFOQueue
        (zeros encode to class 128).  With global conditioning the matching ids are handed out by the
        next `dequeue_gc` call.  pin_memory: return a pinned torch tensor (the H2D copy of the step is then asynchronous)."""
        pieces, ids = [], []
        for _ in range(num_elements):
            piece, cid = self._get(timeout)
            pieces.append(piece)
            ids.append(cid)
        t_max = max(p.shape[0] for p in pieces)
        out = np.zeros((num_elements, t_max, 1), np.float32)
        for i, p in enumerate(pieces):
            out[i, :p.shape[0]] = p
        self._pending_ids.append(ids)
        if pin_memory:
            import torch
            return torch.from_numpy(out).pin_memory()
        return out

    def dequeue_gc(self, num_elements):
        """int32 [num_elements] category ids of the pieces returned by the oldest unmatched `dequeue` call."""
        if not self.gc_enabled:
            raise ValueError('global conditioning is not enabled for this reader')
        ids = self._pending_ids.pop(0)
        assert len(ids) == num_elements, 'dequeue / dequeue_gc sizes differ'
        return np.asarray(ids, dtype=np.int32)

    # ---- producer side -------------------------------------------------------------------------
    def _put(self, item):
        while not self.coord.should_stop():
            try:
                self.queue.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def thread_main(self, sess=None):
        stop = False
        # Go through the dataset multiple times
        while not stop:
            for audio, filename, category_id in load_generic_audio(self.audio_dir, self.sample_rate):
                if self.coord.should_stop():
                    stop = True
                    break
                if self.silence_threshold is not None:
                    # Remove silence
                    audio = trim_silence(audio[:, 0], self.silence_threshold).reshape(-1, 1)
                    if audio.size == 0:
                        print("Warning: {} was ignored as it contains only "
                              "silence. Consider decreasing trim_silence "
                              "threshold, or adjust volume of the audio.".format(filename))
                if self.sample_size:
                    # Cut samples into fixed size pieces; the short tail is a piece of its own (audio_reader.py:167-178)
                    buffer_ = audio.reshape(-1)
                    while len(buffer_) > 0:
                        piece = np.reshape(buffer_[:self.sample_size], [-1, 1]).astype(np.float32)
                        if not self._put((piece, category_id)):
                            stop = True
                            break
                        buffer_ = buffer_[self.sample_size:]
                elif audio.size:
                    if not self._put((audio.astype(np.float32), category_id)):
                        stop = True
                if stop:
                    break

    def start_threads(self, sess=None, n_threads=1):
        for _ in range(n_threads):
            thread = threading.Thread(target=self.thread_main, args=(sess,))
            thread.daemon = True  # Thread will close when parent quits.
            thread.start()
            self.threads.append(thread)
        return self.threads

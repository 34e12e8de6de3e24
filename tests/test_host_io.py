"""Host-side "next" rows of SURVEY section 8f that need no GPU: AudioReader (f2), the checkpoint files and logdir rules
(f3), wav writing / seed cutting helpers of the generation driver (f1)."""
import os
import time

import numpy as np
import pytest


def _write_wav(path, audio, rate=16000, dtype=np.int16):
    from scipy.io import wavfile
    if dtype == np.int16:
        wavfile.write(path, rate, (np.clip(audio, -1, 1) * 32767).astype(np.int16))
    else:
        wavfile.write(path, rate, audio.astype(np.float32))


def _corpus(tmp_path, n_speakers=3, per_speaker=2, rate=16000):
    rng = np.random.default_rng(0)
    d = tmp_path / 'corpus'
    for s in range(n_speakers):
        sub = d / 'p{}'.format(225 + s)
        sub.mkdir(parents=True)
        for k in range(per_speaker):
            t = np.arange(int(rate * (0.3 + 0.1 * k))) / rate
            tone = 0.5 * np.sin(2 * np.pi * (200 + 40 * s) * t)
            audio = np.concatenate([np.zeros(3000), tone, np.zeros(2500)]) + 1e-4 * rng.standard_normal(t.size + 5500)
            _write_wav(str(sub / 'p{}_{:03d}.wav'.format(225 + s, k + 1)), audio, rate)
    return str(d)


def test_find_files_ids_and_cardinality(tmp_path):
    from wavenet import audio_reader as ar
    d = _corpus(tmp_path)
    files = ar.find_files(d)
    assert len(files) == 6 and all(f.endswith('.wav') for f in files)
    assert ar.get_category_cardinality(files) == (225, 227)
    assert not ar.not_all_have_id(files)
    assert ar.not_all_have_id(files + ['/x/speech.wav'])
    picked = list(ar.randomize_files(files))
    assert len(picked) == len(files) and set(picked) <= set(files)        # drawn with replacement


def test_load_wav_formats_and_resampling(tmp_path):
    from wavenet import audio_reader as ar
    t = np.arange(16000) / 16000.0
    x = 0.6 * np.sin(2 * np.pi * 440 * t)
    p16, pf = str(tmp_path / 'a.wav'), str(tmp_path / 'b.wav')
    _write_wav(p16, x, 16000, np.int16)
    _write_wav(pf, x, 16000, np.float32)
    a, b = ar.load_wav(p16, 16000), ar.load_wav(pf, 16000)
    assert a.dtype == np.float32 and a.shape == (16000,)
    np.testing.assert_allclose(a, x, atol=1e-4)
    np.testing.assert_allclose(b, x, atol=1e-7)
    half = ar.load_wav(pf, 8000)                       # resampled: same tone, half the samples
    assert half.shape == (8000,)
    spec = np.abs(np.fft.rfft(half))
    assert abs(np.argmax(spec) - 440) <= 1
    stereo = str(tmp_path / 'c.wav')
    from scipy.io import wavfile
    wavfile.write(stereo, 16000, np.stack([x, -x * 0.5], axis=1).astype(np.float32))
    np.testing.assert_allclose(ar.load_wav(stereo, 16000), 0.25 * x, atol=1e-6)      # mono = channel mean


def test_trim_silence_matches_frame_definition():
    """audio_reader.py:62-69 with librosa's defaults: frames of 2048 centred every 512 samples, keep [first, last) frame start."""
    from wavenet import audio_reader as ar
    x = np.zeros(20000, np.float32)
    x[6000:12000] = 0.5
    out = ar.trim_silence(x, 0.1)
    energy = ar._rms_frames(x)
    # brute-force definition
    yp = np.pad(x, 1024, mode='reflect')
    ref = np.array([np.sqrt(np.mean(yp[i * 512:i * 512 + 2048] ** 2)) for i in range(1 + (yp.size - 2048) // 512)])
    np.testing.assert_allclose(energy, ref, atol=1e-6)
    fr = np.nonzero(ref > 0.1)[0]
    assert out.size == (fr[-1] - fr[0]) * 512 and out.size > 5000
    start, end = fr[0] * 512, fr[-1] * 512
    np.testing.assert_array_equal(out, x[start:end])
    assert 4000 < start <= 6000 and 11000 <= end < 13500           # the tone [6000, 12000) survives, most of the silence goes
    assert ar.trim_silence(np.zeros(5000, np.float32), 0.1).size == 0      # all silence -> empty
    assert ar.trim_silence(np.zeros(0, np.float32), 0.1).size == 0


def test_audio_reader_chunks_padding_and_ids(tmp_path):
    from wavenet import AudioReader, Coordinator
    d = _corpus(tmp_path)
    coord = Coordinator()
    reader = AudioReader(d, coord, sample_rate=16000, gc_enabled=True, sample_size=3000, silence_threshold=0.05,
                         queue_size=16)
    assert reader.gc_category_cardinality == 228                 # largest id + 1 (audio_reader.py:126-136)
    reader.start_threads(n_threads=2)
    batch = reader.dequeue(8, timeout=30)
    ids = reader.dequeue_gc(8)
    coord.request_stop()
    coord.join(reader.threads)
    assert batch.shape[0] == 8 and batch.shape[2] == 1 and batch.dtype == np.float32
    assert batch.shape[1] == 3000                                 # at least one full piece among 8 -> padded to 3000
    assert set(ids.tolist()) <= {225, 226, 227} and ids.dtype == np.int32
    lengths = [int(np.max(np.nonzero(batch[i, :, 0])[0]) + 1) if np.any(batch[i]) else 0 for i in range(8)]
    assert max(lengths) == 3000 and all(0 < n <= 3000 for n in lengths)
    # whole files when sample_size is None; silence-only files are skipped, not enqueued as empty pieces
    reader2 = AudioReader(d, Coordinator(), 16000, gc_enabled=False, sample_size=None, silence_threshold=None)
    reader2.start_threads()
    whole = reader2.dequeue(2, timeout=30)
    reader2.coord.request_stop()
    assert whole.shape[1] >= 4800 + 5500
    with pytest.raises(ValueError):
        reader2.dequeue_gc(2)
    with pytest.raises(ValueError):
        AudioReader(str(tmp_path / 'nothing_here'), None, 16000, False)


def test_audio_reader_requires_ids_for_gc(tmp_path):
    from wavenet import AudioReader
    d = tmp_path / 'plain'
    d.mkdir()
    _write_wav(str(d / 'speech.wav'), np.zeros(100))
    with pytest.raises(ValueError):
        AudioReader(str(d), None, 16000, gc_enabled=True)


# ------------------------------------------------------------------------------------------------ checkpoints
class _FakeNet(object):
    def __init__(self, sd):
        self.sd = {k: v.copy() for k, v in sd.items()}

    def state_dict(self):
        return {k: v.copy() for k, v in self.sd.items()}

    def load_state_dict(self, sd):
        from wavenet.checkpoint import _BIAS_AUTONAMES
        for name in self.sd:
            src = sd.get(name)
            if src is None:
                scope, leaf = name.rsplit('/', 1)
                src = sd[scope + '/' + _BIAS_AUTONAMES[leaf]]
            self.sd[name] = np.asarray(src, np.float32).copy()


def _variables():
    import wavenet_oracle as O
    specs = O.variable_specs([1, 2, 4], 2, 8, 8, 16, 32, True, False, 32, 4, 5)
    return O.init_variables(specs, seed=3, bias_scale=0.1)


def test_checkpoint_roundtrip_names_and_state_file(tmp_path):
    from wavenet import checkpoint as ck
    sd = _variables()
    net = _FakeNet(sd)
    logdir = str(tmp_path / 'logdir' / 'train' / 'run')
    assert ck.load(_FakeNet(sd), logdir) is None                     # nothing there yet (train.py:133-134)
    for step in (10, 50, 120, 130, 140, 150, 160):
        ck.save(net, logdir, step)
    names = sorted(os.listdir(logdir))
    assert 'checkpoint' in names and 'model.ckpt-160.npz' in names
    assert 'model.ckpt-10.npz' not in names and len([n for n in names if n.endswith('.npz')]) == 5     # max_to_keep
    assert ck.get_checkpoint_state(logdir).endswith('model.ckpt-160')
    assert ck.step_of(ck.get_checkpoint_state(logdir)) == 160
    stored, extra = ck.load_variables(ck.get_checkpoint_state(logdir))
    assert set(stored) == set(sd) and int(extra['global_step']) == 160
    assert 'wavenet/dilated_stack/layer2/slip_bias' in stored and 'wavenet/embeddings/gc_embedding' in stored
    other = _FakeNet({k: np.zeros_like(v) for k, v in sd.items()})
    assert ck.load(other, logdir) == 160
    for k in sd:
        np.testing.assert_array_equal(other.sd[k], sd[k])


def test_checkpoint_accepts_tf_autonames(tmp_path):
    """This snapshot of the reference saves its biases as '.../Variable', 'Variable_1', ... (model.py:28 naming bug)."""
    from wavenet import checkpoint as ck
    sd = _variables()
    auto = ck.to_tf_autonames(sd)
    assert 'wavenet/dilated_stack/layer1/Variable_3' in auto and 'wavenet/postprocessing/Variable_1' in auto
    assert not any(k.endswith('_bias') for k in auto)
    path = str(tmp_path / 'model.ckpt-7')
    ck.save_variables(path, ck.variables_from_mapping({k + ':0': v for k, v in auto.items()}))
    other = _FakeNet({k: np.zeros_like(v) for k, v in sd.items()})
    ck.restore(other, path)
    for k in sd:
        np.testing.assert_array_equal(other.sd[k], sd[k])
    slots = {'wavenet/causal_layer/filter/Adam': np.zeros(3), 'global_step': np.zeros(())}
    assert ck.variables_from_mapping(slots) == {}


def test_validate_directories_rules():
    from wavenet import checkpoint as ck
    with pytest.raises(ValueError):
        ck.validate_directories(logdir='a', logdir_root='b')
    with pytest.raises(ValueError):
        ck.validate_directories(logdir='a', restore_from='c')
    d = ck.validate_directories(logdir_root='/tmp/root')
    assert d['logdir'].startswith('/tmp/root/train/') and d['restore_from'] == d['logdir']
    d = ck.validate_directories(logdir='/tmp/x')
    assert d == {'logdir': '/tmp/x', 'logdir_root': None, 'restore_from': '/tmp/x'}
    d = ck.validate_directories(restore_from='/tmp/old')
    assert d['restore_from'] == '/tmp/old' and d['logdir'].startswith('./logdir/train/')


def test_write_wav_roundtrip(tmp_path):
    from wavenet import audio_reader, generation
    x = (0.3 * np.sin(np.arange(4000) * 0.05)).astype(np.float32)
    path = str(tmp_path / 'out.wav')
    generation.write_wav(x, 16000, path)
    np.testing.assert_array_equal(audio_reader.load_wav(path, 16000), x)

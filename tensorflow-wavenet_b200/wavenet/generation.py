"""The sample loop of the reference's generate.py as a library call: seed from a wav
(`create_seed`, generate.py:125-138), the priming protocol (:195-210), temperature scaling and np.random.choice draws
(:228-241), `--save_every` partial writes (:252-255) and the final wav (:269-272) -- CLI-free, same argument names.

The draws consume the GLOBAL numpy random stream exactly like the reference does (one `random_sample()` per generated
sample, taken by `np.random.choice`): with the same `np.random.seed(...)`, the same weights and the same probabilities
the two produce the same waveform.  Fast generation runs inside the persistent generator kernel in chunks of
`save_every` samples (or all at once); priming is ONE batched launch that fills the delay lines (the reference's own
TODO, generate.py:199-201) instead of one session call per sample.
"""
from __future__ import print_function

import numpy as np

from . import audio_reader
from .ops import mu_law_decode, mu_law_encode

SAMPLES = 16000              # generate.py:15-22
TEMPERATURE = 1.0
WINDOW = 8000
SAVE_EVERY = None
SILENCE_THRESHOLD = 0.1


def write_wav(waveform, sample_rate, filename):
    """generate.py:119-122 (librosa.output.write_wav -> scipy.io.wavfile, float32 samples)."""
    from scipy.io import wavfile
    y = np.asarray(waveform, dtype=np.float32)
    wavfile.write(filename, int(sample_rate), y)
    print('Updated wav file at {}'.format(filename))


def create_seed(filename, sample_rate, quantization_channels, window_size=WINDOW, silence_threshold=SILENCE_THRESHOLD):
    """First `window_size` mu-law ids of the silence-trimmed wav (generate.py:125-138), as a Python list."""
    audio = audio_reader.load_wav(filename, sample_rate)
    audio = audio_reader.trim_silence(audio, silence_threshold)
    quantized = mu_law_encode(audio, quantization_channels).cpu().numpy()
    cut_index = min(quantized.size, window_size)
    return quantized[:cut_index].tolist()


def _scale(prediction, temperature):
    """generate.py:229-233 in float32 (the slow path's temperature scaling)."""
    prediction = np.asarray(prediction, dtype=np.float32)
    with np.errstate(divide='ignore'):
        scaled = (np.log(prediction) / np.float32(temperature)).astype(np.float32)
        scaled = scaled - np.logaddexp.reduce(scaled)
        return np.exp(scaled).astype(np.float32)


def generate(net, samples=SAMPLES, wav_seed=None, temperature=TEMPERATURE, fast_generation=True, window=WINDOW,
             gc_id=None, save_every=SAVE_EVERY, wav_out_path=None, sample_rate=16000, seed_waveform=None,
             silence_threshold=SILENCE_THRESHOLD, verbose=False):
    """Generate `samples` new samples with `net` (a WaveNetModel with batch_size 1); returns (waveform ids incl. the
    seed, decoded float audio).  Arguments mirror generate.py's flags; `seed_waveform` (a list of ids) stands for an
    already-encoded seed."""
    q = net.quantization_channels
    if net.global_condition_channels is not None and gc_id is None:
        raise ValueError("Globally conditioning, but global condition was not specified. Use gc_id to specify global condition.")
    if seed_waveform is not None:
        waveform = [int(v) for v in seed_waveform]
    elif wav_seed:
        waveform = create_seed(wav_seed, sample_rate, q, silence_threshold=silence_threshold)
    else:
        waveform = np.random.randint(q, size=(1,)).tolist()           # generate.py:193
    seeded = bool(wav_seed) or seed_waveform is not None

    def decoded():
        return mu_law_decode(np.asarray(waveform, dtype=np.int32), q).cpu().numpy()

    if fast_generation:
        for op in net.init_ops:
            op(1)
        if seeded:
            # generate.py:195-210: every seed sample but the last (window + 1) is pushed through the generator before
            # the loop starts (for a seed no longer than window + 1 that is nothing at all -- the reference's behaviour)
            prime = waveform[:-(window + 1)]
            if prime:
                if verbose:
                    print('Priming generation with {} samples...'.format(len(prime)))
                net.prime(np.asarray(prime, dtype=np.int32), global_condition=gc_id, reset=False)
        done = 0
        chunk = int(save_every) if (save_every and wav_out_path) else samples
        while done < samples:
            n = min(chunk, samples - done)
            u = np.random.random_sample(n)[None, :]                    # the draws np.random.choice would make
            out = net.generate(n, [waveform[-1]], global_condition=gc_id, temperature=temperature, uniforms=u,
                               reset=False).cpu().numpy()[0]
            waveform.extend(int(v) for v in out)
            done += n
            if wav_out_path and save_every and done % save_every == 0:
                write_wav(decoded(), sample_rate, wav_out_path)       # generate.py:252-255
    else:
        for step in range(samples):
            win = waveform[-window:] if len(waveform) > window else waveform
            prediction = net.predict_proba(np.asarray(win, dtype=np.int32), gc_id).cpu().numpy()
            scaled = _scale(prediction, temperature)
            if temperature == 1.0:
                np.testing.assert_allclose(prediction, scaled, atol=1e-5,
                                           err_msg='Prediction scaling at temperature=1.0 is not working as intended.')
            waveform.append(int(np.random.choice(np.arange(q), p=scaled)))
            if wav_out_path and save_every and (step + 1) % save_every == 0:
                write_wav(decoded(), sample_rate, wav_out_path)
    audio = decoded()
    if wav_out_path:
        write_wav(audio, sample_rate, wav_out_path)                   # generate.py:269-272
    return waveform, audio

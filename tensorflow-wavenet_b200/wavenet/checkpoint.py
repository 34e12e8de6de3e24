"""Checkpoint interchange: the `save` / `load` / logdir rules of the reference's train.py
(:104-180) and the restore of generate.py (:176-182), on files keyed by the reference's TF variable names.

Format: one `.npz` per checkpoint, `<logdir>/model.ckpt-<step>.npz`, whose keys are exactly the names
tf.train.Saver(var_list=tf.trainable_variables()) writes for this graph ('wavenet/causal_layer/filter',
'wavenet/dilated_stack/layer3/gate', ...), biases under their INTENDED names ('.../filter_bias', '.../slip_bias' sic).
Loading also accepts the TF auto-names this snapshot really produces ('.../Variable', 'Variable_1', ...: the
create_bias_variable naming bug, model.py:28).  Next to the checkpoints a text file `checkpoint` in TF's CheckpointState
form names the latest one (`model_checkpoint_path: "model.ckpt-<step>"`), which is what `load` follows -- like
tf.train.get_checkpoint_state; the step is parsed from the suffix after the last '-' (train.py:124-126).
A real TF-V1 bundle (.index / .data) is not readable without TensorFlow; `variables_from_mapping` takes any
name -> array mapping exported from one (e.g. by tf.train.load_checkpoint in a TF environment).
"""
from __future__ import print_function

import os
import re
import sys
from datetime import datetime

import numpy as np

MODEL_NAME = 'model.ckpt'
LOGDIR_ROOT = './logdir'                                             # train.py:25
STARTED_DATESTRING = "{0:%Y-%m-%dT%H-%M-%S}".format(datetime.now())   # train.py:28
MAX_TO_KEEP = 5                                                      # tf.train.Saver default

_BIAS_AUTONAMES = {'filter_bias': 'Variable', 'gate_bias': 'Variable_1', 'dense_bias': 'Variable_2',
                   'slip_bias': 'Variable_3', 'postprocess1_bias': 'Variable', 'postprocess2_bias': 'Variable_1'}


# ------------------------------------------------------------------------------- files (pure NumPy, no device)
def checkpoint_path(logdir, step):
    return os.path.join(logdir, '{}-{}'.format(MODEL_NAME, int(step)))


def save_variables(path, variables, extra=None):
    """Write name -> array (and optional extra arrays under '__extra__/...') to `<path>.npz`."""
    arrays = {k: np.asarray(v, dtype=np.float32) for k, v in variables.items()}
    for k, v in (extra or {}).items():
        arrays['__extra__/' + k] = np.asarray(v)
    tmp = path + '.tmp.npz'
    np.savez(tmp, **arrays)
    os.replace(tmp, path + '.npz')
    return path + '.npz'


def load_variables(path):
    """(variables, extra) of a checkpoint written by save_variables; `path` with or without the .npz suffix."""
    if not path.endswith('.npz'):
        path = path + '.npz'
    with np.load(path) as z:
        variables = {k: z[k] for k in z.files if not k.startswith('__extra__/')}
        extra = {k[len('__extra__/'):]: z[k] for k in z.files if k.startswith('__extra__/')}
    return variables, extra


def to_tf_autonames(variables):
    """The same variables under the names this snapshot of the reference really saves (bias auto-names)."""
    out = {}
    for name, v in variables.items():
        scope, leaf = name.rsplit('/', 1)
        out[scope + '/' + _BIAS_AUTONAMES.get(leaf, leaf)] = v
    return out


def variables_from_mapping(mapping):
    """Normalise a name -> array mapping exported from a TF checkpoint: strips ':0' suffixes and optimizer slots."""
    out = {}
    for name, v in mapping.items():
        name = name[:-2] if name.endswith(':0') else name
        if not name.startswith('wavenet/') or re.search(r'/(Adam(_1)?|Momentum|RMSProp(_1)?)$', name):
            continue
        out[name] = np.asarray(v, dtype=np.float32)
    return out


def _read_state(logdir):
    path = os.path.join(logdir, 'checkpoint')
    if not os.path.exists(path):
        return None, []
    latest, every = None, []
    for line in open(path):
        m = re.match(r'\s*(model_checkpoint_path|all_model_checkpoint_paths):\s*"(.*)"\s*$', line)
        if not m:
            continue
        if m.group(1) == 'model_checkpoint_path':
            latest = m.group(2)
        else:
            every.append(m.group(2))
    return latest, every


def _write_state(logdir, latest, every):
    with open(os.path.join(logdir, 'checkpoint'), 'w') as f:
        f.write('model_checkpoint_path: "{}"\n'.format(latest))
        for p in every:
            f.write('all_model_checkpoint_paths: "{}"\n'.format(p))


def get_checkpoint_state(logdir):
    """Path of the latest checkpoint of `logdir` (like tf.train.get_checkpoint_state(...).model_checkpoint_path) or None."""
    latest, _ = _read_state(logdir)
    if latest is None:
        return None
    return latest if os.path.isabs(latest) else os.path.join(logdir, latest)


def step_of(path):
    """Global step from a checkpoint path: the text after the last '-' (train.py:124-126, generate.py has the same rule)."""
    return int(os.path.basename(path).replace('.npz', '').split('-')[-1])


# ------------------------------------------------------------------------------- train.py:104-134
def save(net, logdir, step, optimizer=None, max_to_keep=MAX_TO_KEEP):
    """Store `net`'s variables as `<logdir>/model.ckpt-<step>` (train.py:104-115)."""
    print('Storing checkpoint to {} ...'.format(logdir), end="")
    sys.stdout.flush()
    if not os.path.exists(logdir):
        os.makedirs(logdir)
    extra = {'global_step': np.int64(step)}
    if optimizer is not None and hasattr(optimizer, 'state_dict'):
        for k, v in optimizer.state_dict().items():
            extra['optimizer/' + k] = v
    path = checkpoint_path(logdir, step)
    save_variables(path, net.state_dict(), extra)
    _, every = _read_state(logdir)
    name = os.path.basename(path)
    every = [p for p in every if p != name] + [name]
    while max_to_keep and len(every) > max_to_keep:
        old = every.pop(0)
        try:
            os.remove(os.path.join(logdir, old + '.npz'))
        except OSError:
            pass
    _write_state(logdir, name, every)
    print(' Done.')
    return path


def load(net, logdir, optimizer=None):
    """Restore the latest checkpoint of `logdir` into `net`; returns its global step, or None when there is none
    (train.py:118-134)."""
    print("Trying to restore saved checkpoints from {} ...".format(logdir), end="")
    path = get_checkpoint_state(logdir)
    if path:
        print("  Checkpoint found: {}".format(path))
        global_step = step_of(path)
        print("  Global step was: {}".format(global_step))
        print("  Restoring...", end="")
        restore(net, path, optimizer)
        print(" Done.")
        return global_step
    print(" No checkpoint found.")
    return None


def restore(net, path, optimizer=None):
    """saver.restore(sess, checkpoint) of generate.py:176-182: one named checkpoint file into `net`."""
    variables, extra = load_variables(path)
    net.load_state_dict(variables)
    if optimizer is not None and hasattr(optimizer, 'load_state_dict'):
        state = {k[len('optimizer/'):]: v for k, v in extra.items() if k.startswith('optimizer/')}
        if state:
            optimizer.load_state_dict(state)
    return extra


# ------------------------------------------------------------------------------- train.py:137-180
def get_default_logdir(logdir_root):
    return os.path.join(logdir_root, 'train', STARTED_DATESTRING)


def validate_directories(logdir=None, logdir_root=None, restore_from=None):
    """Validate and arrange directory related arguments (train.py:142-180): same rules, keyword arguments instead of
    an argparse namespace."""
    if logdir and logdir_root:
        raise ValueError("--logdir and --logdir_root cannot be specified at the same time.")
    if logdir and restore_from:
        raise ValueError(
            "--logdir and --restore_from cannot be specified at the same "
            "time. This is to keep your previous model from unexpected "
            "overwrites.\n"
            "Use --logdir_root to specify the root of the directory which "
            "will be automatically created with current date and time, or use "
            "only --logdir to just continue the training from the last "
            "checkpoint.")
    root = logdir_root if logdir_root is not None else LOGDIR_ROOT
    if logdir is None:
        logdir = get_default_logdir(root)
        print('Using default logdir: {}'.format(logdir))
    if restore_from is None:
        # logdir and restore_from are exclusive, so the logdir here is newly created
        restore_from = logdir
    return {'logdir': logdir, 'logdir_root': logdir_root, 'restore_from': restore_from}

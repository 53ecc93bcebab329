"""Shared test helpers: golden loading, synthetic inputs (oracle side)."""
import hashlib
import json
import os

import numpy as np

from oracle import synth_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SEED = 0x5C6


def sha(a):
  return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_json(name):
  with open(os.path.join(GOLDEN, name)) as f:
    return json.load(f)


def configs():
  return load_json('configs.json')


def effective_config(cfg_name, table=None):
  """Path-relevant params for waveform_NN with the documented legacy defaults
  (tests/golden/make_golden.py: LEGACY_DEFAULTS; waveform_04 per project_log.txt:19-21)."""
  c = dict((table or configs())[cfg_name])
  if cfg_name == 'waveform_04':
    c['min_RHC'], c['use_global_min_max'] = 0, True
  c.setdefault('min_RHC', float('-inf'))
  c.setdefault('use_global_min_max', False)
  return c


def adversarial_windows():
  import importlib.util
  spec = importlib.util.spec_from_file_location('make_golden', os.path.join(GOLDEN, 'make_golden.py'))
  mg = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mg)
  return mg.adversarial_windows()


def _make_golden():
  import importlib.util
  spec = importlib.util.spec_from_file_location('make_golden', os.path.join(GOLDEN, 'make_golden.py'))
  mg = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mg)
  return mg


def short_windows():
  return _make_golden().short_windows()


def predicate_inputs():
  names, adv = adversarial_windows()
  rec = synth_ref.gen_record(SEED, 7, 150000, kinds=(3,))[:, 0].reshape(-1, 750)
  return names, np.concatenate([adv, rec])


def small_record():
  sig = synth_ref.SIG_NAMES_5
  p = synth_ref.gen_record(SEED, 3, 45000, kinds=synth_ref.kinds_for(sig))
  meta = synth_ref.record_meta(100, events={'RA_1': 0.5, 'RV_1': 15.001, 'PA_1': 30, 'PCW_1': 55.25, 'PA_2': 61.7})
  return sig, p, meta


def full_record(r):
  sig = synth_ref.SIG_NAMES_5
  return sig, synth_ref.gen_record(SEED, r, 300000, kinds=synth_ref.kinds_for(sig)), synth_ref.record_meta(600)


def free_port():
  """A TCP port nobody listens on right now (torch.distributed rendezvous of the multi-process tests)."""
  import socket
  with socket.socket() as so:
    so.bind(('127.0.0.1', 0))
    return so.getsockname()[1]

"""TEST INFRASTRUCTURE — runs the UNMODIFIED reference (``/root/reference``) on synthetic records.

Only usable where ``/root/reference`` exists (the build container, never the GPU box).  It is
how the oracle is *pinned*: ``tests/golden/make_golden.py`` calls the reference's own
``get_chamber_intervals`` / ``get_segments`` / ``has_noise`` / ``get_global_minmax_vals`` /
``SCGDataset`` through this harness and commits the outputs as fixtures; ``tests/test_oracle.py``
then checks ``oracle/scgrhc_oracle.py`` and the C restatement against those fixtures.

Mechanism: ``matplotlib`` and ``wfdb`` are absent from this image and the
reference imports them at module scope (`recordutil.py:2,8`, `waveform_noise.py:1`), so empty
stub modules are pre-inserted into ``sys.modules``; ``wfdb.rdrecord`` is replaced by a lookup
into an in-memory dict of fake records; ``recordutil.PROCESSED_DATA_PATH`` (a module global read
at call time, `recordutil.py:87,98,137`) is pointed at a temp dir holding ``{rec}.json`` and an
empty ``{rec}.hea``.  No reference source is edited or copied.
"""
import json
import os
import sys
import tempfile
import types

REFERENCE_PATH = os.environ.get('SCGRHC_REFERENCE_PATH', '/root/reference')
STAGED_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')   # oracle/make_ref.py: byte copies for the GPU box


def reference_available():
  return os.path.isfile(os.path.join(REFERENCE_PATH, 'recordutil.py'))


def staged_available():
  return os.path.isfile(os.path.join(STAGED_PATH, 'recordutil.py'))


class FakeRecord:
  """What `wfdb.rdrecord` returns, reduced to the two attributes the reference reads
  (`recordutil.py:117-118`)."""

  def __init__(self, sig_name, p_signal):
    self.sig_name = list(sig_name)
    self.p_signal = p_signal


class ReferenceHarness:
  def __init__(self, path=None):
    """``path``: where the reference's modules live (default: /root/reference; the GPU box passes STAGED_PATH)."""
    REFERENCE_PATH = path or globals()['REFERENCE_PATH']
    self.path = REFERENCE_PATH
    if not os.path.isfile(os.path.join(REFERENCE_PATH, 'recordutil.py')):
      raise RuntimeError('reference not present at %s' % REFERENCE_PATH)
    self.tmp = tempfile.TemporaryDirectory(prefix='scgrhc_ref_')
    self.records = {}
    self._saved = {}
    for name in ('matplotlib', 'matplotlib.pyplot', 'wfdb'):
      self._saved[name] = sys.modules.get(name)
      if name == 'wfdb' or self._saved[name] is None:
        sys.modules[name] = types.ModuleType(name)
    plt = sys.modules['matplotlib.pyplot']
    for fn in ('plot', 'title', 'xlabel', 'ylabel', 'ylim', 'legend', 'savefig', 'close'):
      if not hasattr(plt, fn):
        setattr(plt, fn, lambda *a, **k: None)
    sys.modules['matplotlib'].pyplot = plt
    sys.modules['wfdb'].rdrecord = lambda path: self.records[os.path.basename(path)]
    # The reference's flat module names must win over this repo's drop-in modules.
    for name in ('recordutil', 'waveform_noise', 'paramutil', 'pathutil', 'timelog'):
      self._saved[name] = sys.modules.pop(name, None)
    sys.path.insert(0, REFERENCE_PATH)
    try:
      import recordutil, waveform_noise, paramutil  # noqa: E401  (the reference's)
    finally:
      sys.path.remove(REFERENCE_PATH)
    assert os.path.dirname(os.path.abspath(recordutil.__file__)) == os.path.abspath(REFERENCE_PATH)
    self.recordutil = recordutil
    self.waveform_noise = waveform_noise
    self.paramutil = paramutil
    recordutil.PROCESSED_DATA_PATH = self.tmp.name

  def close(self):
    for name in ('recordutil', 'waveform_noise', 'paramutil', 'pathutil', 'timelog'):
      sys.modules.pop(name, None)
    for name, mod in self._saved.items():
      if mod is not None:
        sys.modules[name] = mod
      elif name in ('matplotlib', 'matplotlib.pyplot', 'wfdb'):
        sys.modules.pop(name, None)
    self.tmp.cleanup()

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    self.close()

  # -- synthetic data root ---------------------------------------------------------------
  def add_record(self, name, sig_name, p_signal, meta):
    self.records[name] = FakeRecord(sig_name, p_signal)
    with open(os.path.join(self.tmp.name, name + '.json'), 'w') as f:
      json.dump(meta, f)
    open(os.path.join(self.tmp.name, name + '.hea'), 'w').close()

  def params(self, config_dir, **overrides):
    """Reference ``Params`` for ``waveform_NN``; for the 5 legacy configs that the reference's
    own loader rejects the missing keys come from ``overrides``."""
    path = os.path.join(self.path, config_dir, 'params.json')
    try:
      p = self.paramutil.Params(path)
    except KeyError:
      with open(path) as f:
        data = json.load(f)
      p = types.SimpleNamespace(path=path, data=data, **data)
    for k, v in overrides.items():
      setattr(p, k, v)
    return p
